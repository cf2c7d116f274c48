// hop_select_pipe_body.cuh -- software-pipelined FAST horizon selection, one problem per warp.
//
// Same function and the same closed forms as MODE 1 of hop_select_mma_body.cuh (augmented.py:10-87,
// horizon_selection.py:36-86, solver.py:522); what changes is the SCHEDULE.  In the sequential body a
// warp executes, per horizon step, three Gauss-Jordan sweeps and one forward elimination one after the
// other -- 13 dependent pivots each (shuffle -> reciprocal -> multiply -> FMA), which leaves the FP64 /
// DMMA pipe idle most of the time (ncu r1b: issue-active 28 %, dominant stall `wait`).  The three
// inversions that depend only on the prefix state P_k = (Ebar, Fbar, Gbar)_k are independent of one
// another:
//       W_{k+1} = chol_inv(E_{k+1} + Gbar_k)          prefix step k+1   (horizon_selection.py:72)
//       W_t     = chol_inv(X_t + Gbar_k), t = k+1     query of horizon t (:82)
//       pivot_n(X0_{t-1} + eps I)                     cost of horizon t-1 (:84-85), X0 formed one iteration ago
// so iteration k runs them as ONE interleaved sweep (three independent dependency chains in the same
// instruction stream), then issues the DMMA products of the query (-> X0_t) and of prefix step k+1.
// The value J(t) therefore leaves the pipeline one iteration late; an epilogue drains the last one.
//
// Further differences from the sequential FAST body (all value-preserving up to rounding order):
//   * e = wrap(X_{k+1} - xg) serves both Q_aug[k+1] and QT_{k+1} (it is the same vector);
//   * the matvecs Q e | P e and K q | K' p run on the two half-warps concurrently;
//   * K, K' are kept in fragment order in shared memory (conflict-free loads);
//   * X0 is only formed on its lower tiles and its forward elimination reads the pivot row from the
//     pivot column (X0 is symmetric by construction).
// Any non-positive pivot / sigma (or non-finite input) aborts the pipelined sweep for that problem and the
// caller re-runs it through the sequential body, which owns the jitter ladder, the LU fallback and the
// status word.
#pragma once
#include "hop_select_mma_body.cuh"

namespace hop { namespace mma {

template <int D, int M>
struct PipeConst {
    static constexpr int n = D - 1;
    static constexpr int BASE = (FastConst<D, M>::SIZE + 1) & ~1;
    static constexpr int QRAWT = BASE;            // Q^T (raw), so that lane i reads Q[i][j] at [j*n + i]
    static constexpr int KQF = QRAWT + n * n;     // K  = (Qs + eps I)^-1 in fragment order [8][32]
    static constexpr int KPF = KQF + 256;         // K' = (P  + eps I)^-1 in fragment order [8][32]
    static constexpr int SIZE = KPF + 256;
};

// per-warp shared memory (doubles)
struct PipeSlab {
    static constexpr int LU = 0;                  // 512: scratch of the sequential fallback body (its own layout, 576 + stages)
    static constexpr int EV = 0, QE = 16, PE = 32, YQ = 48, YP = 64, DU = 80;   // vectors of the pipelined sweep
    static constexpr int STAGE = 96;              // 2 x kStage staging buffers
    static constexpr int BARS = STAGE + 2 * kStage;
    static constexpr int SIZE = BARS + 2;
};
static_assert(PipeSlab::SIZE <= kWarpScratch, "the pipelined body reuses the sequential body's per-warp slab");

// CTA-cooperative fill of the extra constants (after fast_const_fill_warp).
template <int D, int M>
HOP_DEVICE void pipe_const_fill(double* cst, int tid, int nthr) {
    using FC = FusedConst<D, M>;
    using XC = FastConst<D, M>;
    using PC = PipeConst<D, M>;
    constexpr int n = D - 1;
    for (int i = tid; i < n * n; i += nthr) cst[PC::QRAWT + (i % n) * n + (i / n)] = cst[FC::QRAW + i];
    for (int i = tid; i < 256; i += nthr) {
        const int e = i >> 5, lane = i & 31;
        const int I = e >> 2, J = (e >> 1) & 1, s = e & 1;
        const int g = lane >> 2, t = lane & 3;
        const int R = 8 * I + rho(g), C = 8 * J + t + 4 * s;
        const bool in = (R < n && C < n);
        cst[PC::KQF + i] = in ? cst[XC::KQ + R * n + C] : 0.0;
        cst[PC::KPF + i] = in ? cst[XC::KP + R * n + C] : 0.0;
    }
}

// One pivot of the in-place Gauss-Jordan inversion (same arithmetic as gj_attempt).
template <int D>
HOP_DEVICE void gj_pivot(Mat& a, int j, const LaneGeo& L, bool& ok) {
    const int Ij = j >> 3, gj = rho_inv(j & 7);
    const int Jj = j >> 3, tj = j & 3, sj = (j & 7) >> 2;
    const double p = simt::shfl(a.v[Ij][Jj][sj], (gj << 2) | tj, 32);
    ok = ok && (p > 0.0);
    const double rinv = pivot_rcp(p);
    double pr[2][2], f[2];
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s) pr[J][s] = simt::shfl(a.v[Ij][J][s], (gj << 2) | L.t, 32);
#pragma unroll
    for (int I = 0; I < 2; ++I) f[I] = simt::shfl(a.v[I][Jj][sj], (L.g << 2) | tj, 32) * rinv;
    HOP_FOR_ELEMS(I, J, s) a.v[I][J][s] = fma(-f[I], pr[J][s], a.v[I][J][s]);
    const bool isrow = (L.g == gj), iscol = (L.t == tj);
    if (isrow) {
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) a.v[Ij][J][s] = pr[J][s] * rinv;
    }
    if (iscol) {
        a.v[0][Jj][sj] = -f[0];
        a.v[1][Jj][sj] = -f[1];
        if (isrow) a.v[Ij][Jj][sj] = rinv;
    }
}

// One pivot of the forward elimination of a SYMMETRIC matrix held on its lower tiles (0,0), (1,0), (1,1)
// (tile (0,1) is never read or written).  Row j is taken from column j.  p receives the pivot.
template <int D>
HOP_DEVICE void fe_pivot_lower(Mat& a, int j, const LaneGeo& L, bool& ok, double& p) {
    const int Ij = j >> 3, gj = rho_inv(j & 7);
    const int Jj = j >> 3, tj = j & 3, sj = (j & 7) >> 2;
    p = simt::shfl(a.v[Ij][Jj][sj], (gj << 2) | tj, 32);
    ok = ok && (p > 0.0);
    if (j == D - 1) return;
    const double rinv = pivot_rcp(p);
    double pr[2][2], f[2];
    // rows/cols <= j are dead: once j >= 8 only tile (1,1) is live
#pragma unroll
    for (int J = (j >= 8 ? 1 : 0); J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s)   // M[j][8J+t+4s] = M[8J+t+4s][j]: tile (J, Jj), lane (rho_inv(t+4s), tj)
            pr[J][s] = simt::shfl(a.v[J][Jj][sj], ((2 * L.t + s) << 2) | tj, 32);
#pragma unroll
    for (int I = (j >= 8 ? 1 : 0); I < 2; ++I) f[I] = simt::shfl(a.v[I][Jj][sj], (L.g << 2) | tj, 32) * rinv;
    if (j < 8) {
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[0][0][s] = fma(-f[0], pr[0][s], a.v[0][0][s]);
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[1][0][s] = fma(-f[1], pr[0][s], a.v[1][0][s]);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) a.v[1][1][s] = fma(-f[1], pr[1][s], a.v[1][1][s]);
}
static_assert(rho_inv(0) == 0 && rho_inv(1) == 2 && rho_inv(4) == 1 && rho_inv(7) == 7, "rho_inv(t + 4s) == 2t + s");

// a1 <- a1^-1, a2 <- a2^-1 (Gauss-Jordan), x <- forward elimination; returns the last pivot of x.
template <int D>
HOP_DEVICE double gj3(Mat& a1, Mat& a2, Mat& x, const LaneGeo& L, bool& ok) {
    double p = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        gj_pivot<D>(a1, j, L, ok);
        gj_pivot<D>(a2, j, L, ok);
        fe_pivot_lower<D>(x, j, L, ok, p);
    }
    return p;
}

// D = X * Z^T on the lower tiles (0,0), (1,0), (1,1) only (symmetric result).
template <int KB>
HOP_DEVICE void mma_nt_lower(Mat& Dm, const Mat& X, const Mat& Z) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            Dm.v[I][J][0] = 0.0; Dm.v[I][J][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
                simt::dmma(Dm.v[I][J][0], Dm.v[I][J][1], X.v[I][kb >> 1][kb & 1], Z.v[J][kb >> 1][kb & 1]);
        }
}

// Returns true when the problem was solved by the pipelined sweep; false => caller must run the sequential body.
template <int D, int M>
HOP_DEVICE bool select_fused_pipe_body(const FusedArgs& p, int b, double* scratch, const double* cst) {
    using FC = FusedConst<D, M>;
    using XC = FastConst<D, M>;
    using PC = PipeConst<D, M>;
    using PS = PipeSlab;
    constexpr int n = D - 1;
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4, KBM = (M + 3) / 4, NTM = (M + 7) / 8;
    static_assert(D > 8 && D <= 16 && n <= 16, "one-problem-per-warp mapping: 9 <= d <= 16");
    if (cst[XC::FLAG] != 0.0) return false;      // K / K' needed the ladder: closed forms do not apply
    LaneGeo L;
    L.init();
    double* EV = scratch + PS::EV;   // e = wrap(X_{k+1} - xg)
    double* QE = scratch + PS::QE;   // Q e
    double* PE = scratch + PS::PE;   // P e
    double* YQ = scratch + PS::YQ;   // [K q ; -1 ; 0]
    double* YP = scratch + PS::YP;   // [K' p ; -1 ; 0]
    double* DU = scratch + PS::DU;   // U - u_ref
    bool ok = true;

    // R_inv = chol_inv(sym(R)) (augmented.py:23); its transpose is the Z operand of B R^-1
    Mat RinvT;
    {
        Mat Rs, Ri;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            Rs.v[I][J][s] = ((R < M && C < M) ? cst[FC::RS + R * M + C] : 0.0) + ((R == C && R < M) ? p.jitter : 0.0);
        }
        ok = gj_attempt<M>(Rs, L) && ok;
        mat_copy(Ri, Rs);
        mat_transpose(RinvT, Ri, L);
    }
    // half-warp roles for the vector work: h = 0 -> Q side (Q e, K q), h = 1 -> terminal side (P e, K' p)
    const int h = L.lane >> 4, li = L.lane & 15;
    const bool isx = li < n;
    const double xg_l = isx ? p.xg[(size_t)b * n + li] : 0.0;
    const bool wrap_l = isx && ((p.wrap_mask >> li) & 1u);
    const double w = p.w[b];
    const double* mat1 = cst + (h ? FC::PF : PC::QRAWT);   // symmetric P | Q^T : element [i][j] at [j*n + i]
    const double* mat2 = cst + (h ? XC::KP : XC::KQ);      // symmetric K' | K
    double* V1 = h ? PE : QE;
    double* V2 = h ? YP : YQ;
    if (L.lane < 16) {
        YQ[L.lane] = (L.lane == n) ? -1.0 : 0.0;
        YP[L.lane] = (L.lane == n) ? -1.0 : 0.0;
    }
    const double* KF1 = cst + PC::KQF + L.lane;
    const double* KF2 = cst + PC::KPF + L.lane;

    const size_t baseN = (size_t)b * p.N;
    const size_t baseX = (size_t)b * (p.N + 1);
    static_assert(n * n + n * M + n + M + n <= kStage, "staging buffer too small");
    static_assert((n * n) % 2 == 0 && (n * M) % 2 == 0 && M % 2 == 0 && n % 2 == 0, "bulk copies need 16-byte multiples");
    double* stage0 = scratch + PS::STAGE;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(scratch + PS::BARS);
    constexpr int oA = 0, oB = n * n, oX = oB + n * M, oU = oX + n, oR = oU + M;
    // stage s < T_max: A_s, B_s, X_s, U_s, a_s;  stage T_max: X only
    auto issue = [&](int s) {
        double* st = stage0 + (s & 1) * kStage;
        unsigned long long* bar = bars + (s & 1);
        if (s < p.T_max) {
            const unsigned bytes = 8u * (n * n + n * M + n + M + (p.a_resid ? n : 0));
            simt::mbar_expect_tx(bar, bytes);
            simt::bulk_g2s(st + oA, p.A + (baseN + s) * n * n, 8u * n * n, bar);
            simt::bulk_g2s(st + oB, p.Bm + (baseN + s) * n * M, 8u * n * M, bar);
            simt::bulk_g2s(st + oX, p.X + (baseX + s) * n, 8u * n, bar);
            simt::bulk_g2s(st + oU, p.U + (size_t)b * p.u_stride + (size_t)s * M, 8u * M, bar);
            if (p.a_resid) simt::bulk_g2s(st + oR, p.a_resid + (baseN + s) * n, 8u * n, bar);
        } else {
            simt::mbar_expect_tx(bar, 8u * n);
            simt::bulk_g2s(st + oX, p.X + (baseX + s) * n, 8u * n, bar);
        }
    };
    if (L.lane == 0) {
        simt::mbar_init(bars, 1);
        simt::mbar_init(bars + 1, 1);
        simt::mbar_fence_init();
    }
    simt::sync();
    if (L.lane == 0) { issue(0); issue(1); }
    simt::sync();   // (host emulation: copies complete at issue time, so order the issue before the first read)

    // leave no bulk copy in flight and no live mbarrier behind (the fallback body re-uses the slab)
    auto bail = [&](int pending_stage) {
        if (pending_stage >= 0) simt::mbar_wait(bars + (pending_stage & 1), (unsigned)((pending_stage >> 1) & 1));
        simt::sync();
        if (L.lane == 0) { simt::mbar_inval(bars); simt::mbar_inval(bars + 1); }
        simt::sync();
        return false;
    };
    // Vector stage for index s (inputs in `stg`): e, Q e, P e, y = K q, y' = K' p, sigma, sigma'.
    // On return YQ/YP hold the extended vectors, DU the control deviation (s < T_max only).
    double rsq = 0.0, rsp = 0.0;
    auto vector_stage = [&](const double* stg, bool with_u) {
        double ev = 0.0;
        if (isx) {
            ev = stg[oX + li] - xg_l;                                           // e = wrap(X_s - xg)  (augmented.py:28,80)
            if (wrap_l) ev = wrap_pi(ev);
            if (h == 0) EV[li] = ev;
        }
        if (with_u && L.lane < M) DU[L.lane] = stg[oU + L.lane] - cst[FC::UREF + L.lane];
        simt::sync();
        double v1 = 0.0, qc = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const double ej = EV[j];
                v1 = fma(mat1[j * n + li], ej, v1);                             // (Q e)_i | (P e)_i
                if (h == 0) qc = fma(ej, cst[FC::QRAW + j * n + li], qc);       // (e^T Q)_i
            }
            V1[li] = v1;
        }
        double r0 = (isx && h == 0) ? qc * ev : 0.0;                            // e^T Q e   (augmented.py:37)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) r0 += simt::shfl_xor(r0, o, 32);
        const double eQe = simt::shfl(r0, 0, 32);
        const double corner = eQe + 2.0 * w + p.rho_reg;
        simt::sync();
        double y = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) y = fma(mat2[j * n + li], V1[j], y);    // y = K q | y' = K' p
            V2[li] = y;
        }
        double r1 = isx ? (h ? y * ev : v1 * y) : 0.0;                          // q^T y | y'^T e
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) r1 += simt::shfl_xor(r1, o, 32);
        const double qy = simt::shfl(r1, 0, 32), ye = simt::shfl(r1, 16, 32);
        const double sigq = (corner + p.jitter) - qy;                           // Schur complement of Q_aug + eps I
        const double sigp = (p.rho_reg + p.jitter) + p.jitter * ye;             // ... of QT + eps I, cancellation-free
        ok = ok && (sigq > 0.0) && (sigp > 0.0);
#ifdef HOP_DEBUG_PIPE
        if (L.lane==0) printf("pipe eQe=%.17g corner=%.17g qy=%.17g ye=%.17g sigq=%.17g sigp=%.17g\n", eQe, corner, qy, ye, sigq, sigp);
#endif
        rsq = 1.0 / sigq;
        rsp = 1.0 / sigp;
        simt::sync();
    };
    // closed-form block inverse  Kx + rs * yx yx^T  (yx = [y ; -1 ; 0])
    auto closed_inverse = [&](Mat& E, const double* KF, const double* Y, double rs) {
        double yr[2], yc[2][2];
#pragma unroll
        for (int I = 0; I < 2; ++I) yr[I] = Y[L.row(I)];
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) yc[J][s] = Y[L.col(J, s)];
        HOP_FOR_ELEMS(I, J, s) E.v[I][J][s] = fma(yr[I] * yc[J][s], rs, KF[((I * 2 + J) * 2 + s) * 32]);
    };
    auto load_AB = [&](const double* stg, Mat& A, Mat& Bm) {
        const double* Ak = stg + oA;
        const double* Bk = stg + oB;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            double a = 0.0;
            if (R < n && C < n) {
                a = Ak[R * n + C];
            } else if (R < n && C == n) {
                double sacc = 0.0;
#pragma unroll
                for (int c = 0; c < M; ++c) sacc = fma(Bk[R * M + c], DU[c], sacc);
                a = (p.a_resid ? stg[oR + R] : 0.0) - sacc;                     // a_k - B_k du   (augmented.py:50)
            } else if (R == n && C == n) {
                a = 1.0;
            }
            A.v[I][J][s] = a;
            Bm.v[I][J][s] = (R < n && C < M) ? Bk[R * M + C] : 0.0;
        }
    };
    auto stage_G = [&](Mat& Ft, Mat& G, const Mat& A, const Mat& Bm, const Mat& E) {
        mma_nt<NT, NT, KB, false>(Ft, A, E);                                   // F_k^T = A_k E_k
        mma_nt<NT, NT, KB, false>(G, Ft, A);                                   // (A_k E_k) A_k^T              (:61)
        Mat BR;
        mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);                            // B_k R^-1
        mma_nt<NT, NT, KBM, true>(G, BR, Bm);                                  // + (B_k R^-1) B_k^T
    };

    // ---------------- prologue: prefix step 0 (horizon_selection.py:57-64 with k = 0)
    PrefixL<D> P;
    simt::mbar_wait(bars + 0, 0u);
    vector_stage(stage0, true);
    {
        Mat E, A, Bm, Ft, G;
        closed_inverse(E, KF1, YQ, rsq);
        load_AB(stage0, A, Bm);
        stage_G(Ft, G, A, Bm, E);
        mat_sym(G, L);
        mma_nt<NT, NT, KB, false>(P.fb, E, A);                                 // F_0 = E_0 A_0^T              (:60)
        mat_copy(P.eb, E);
        mat_copy(P.gb, G);
    }
    Mat X0;                                                                    // X0 of the previous horizon (lower tiles)
    HOP_FOR_ELEMS(I, J, s) X0.v[I][J][s] = (L.row(I) == L.col(J, s)) ? 1.0 : 0.0;
    ArgMin am;
    am.init();

    for (int k = 0; k < p.T_max; ++k) {
        const bool last = (k + 1 == p.T_max);
        simt::sync();                                                           // everyone is done with stage k
        if (L.lane == 0 && k + 2 <= p.T_max) issue(k + 2);
        simt::mbar_wait(bars + ((k + 1) & 1), (unsigned)(((k + 1) >> 1) & 1));
        const double* stg = stage0 + ((k + 1) & 1) * kStage;
        vector_stage(stg, !last);
        if (!simt::all(ok)) return bail(k + 2 <= p.T_max ? k + 2 : -1);
        // ---------------- the three independent inversions
        Mat W, Wt;
        {
            closed_inverse(W, KF1, YQ, rsq);                                   // E_{k+1} = chol_inv(Q_aug[k+1])   (:59)
            closed_inverse(Wt, KF2, YP, rsp);                                  // X_t = chol_inv(QT_t), t = k+1     (:79)
            HOP_FOR_ELEMS(I, J, s) {
                const double dg = (I == J && L.row(I) == L.col(J, s) && L.row(I) < D) ? p.jitter : 0.0;
                W.v[I][J][s] = (W.v[I][J][s] + P.gb.v[I][J][s]) + dg;          // E_{k+1} + Gbar_k (+ eps I)        (:72)
                Wt.v[I][J][s] = (Wt.v[I][J][s] + P.gb.v[I][J][s]) + dg;        // X_t + Gbar_k (+ eps I)            (:82)
                if (I >= J) X0.v[I][J][s] += dg;                               // X0_{t-1} + eps I                  (:84)
            }
        }
        const double piv = gj3<D>(W, Wt, X0, L, ok);
        if (!simt::all(ok)) return bail(k + 2 <= p.T_max ? k + 2 : -1);
        if (k > 0 && L.lane == 0) {                                            // J(t-1) = 0.5 / pivot_n  (z0 = e_n, :85)
            const double Jt = 0.5 / piv;
            p.J_out[(size_t)b * p.T_max + (k - 1)] = Jt;
            if (k >= p.T_min) am.push(Jt, k);
        }
        // ---------------- query products of horizon t = k+1 (:83): X0 = Ebar - (Fbar W_t) Fbar^T, lower tiles
        {
            Mat T3, acc;
            mma_nt<NT, NT, KB, false>(T3, P.fb, Wt);
            mma_nt_lower<KB>(acc, T3, P.fb);
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
#pragma unroll
                    for (int s = 0; s < 2; ++s) X0.v[I][J][s] = P.eb.v[I][J][s] - acc.v[I][J][s];
        }
        if (last) break;
        // ---------------- prefix step k+1 (:57-75)
        {
            Mat E, A, Bm, Ft, G;
            closed_inverse(E, KF1, YQ, rsq);
            load_AB(stg, A, Bm);
            stage_G(Ft, G, A, Bm, E);
            Mat T1, acc;
            mma_nt<NT, NT, KB, false>(T1, P.fb, W);                            // Fbar W                       (:73)
            mma_nt<NT, NT, KB, false>(acc, T1, P.fb);                          // (Fbar W) Fbar^T
            mat_sub(P.eb, P.eb, acc);
            mat_sym(P.eb, L);                                                  // Ebar                         (:73)
            mma_nt<NT, NT, KB, false>(acc, T1, Ft);                            // (Fbar W) F_k  -> new Fbar    (:74)
            mma_nt<NT, NT, KB, false>(T1, Ft, W);                              // F_k^T W                      (:75)
            mat_copy(P.fb, acc);
            mma_nt<NT, NT, KB, false>(acc, T1, Ft);                            // (F_k^T W) F_k
            mat_sub(P.gb, G, acc);
            mat_sym(P.gb, L);                                                  // Gbar                         (:75)
        }
    }
    // ---------------- epilogue: cost of the last horizon
    {
        HOP_FOR_ELEMS(I, J, s)
            if (I >= J && I == J && L.row(I) == L.col(J, s) && L.row(I) < D) X0.v[I][J][s] += p.jitter;
        double piv = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) fe_pivot_lower<D>(X0, j, L, ok, piv);
        if (!simt::all(ok)) return bail(-1);
        if (L.lane == 0) {
            const double Jt = 0.5 / piv;
            p.J_out[(size_t)b * p.T_max + (p.T_max - 1)] = Jt;
            if (p.T_max >= p.T_min) am.push(Jt, p.T_max);
            p.T_out[b] = am.idx;
            p.Jstar_out[b] = am.best;
            p.status[b] = 0;
        }
    }
    return true;
}

}}  // namespace hop::mma
