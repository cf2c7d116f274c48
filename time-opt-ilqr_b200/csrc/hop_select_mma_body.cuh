// hop_select_mma_body.cuh -- HOP horizon selection, ONE PROBLEM PER WARP, blocks held as DMMA
// register fragments (hop_mma.cuh).  Same algorithm and operation order as hop_select_core.cuh /
// the reference (horizon_selection.py:36-86, augmented.py:10-87, solver.py:522); only the mapping
// of the d x d algebra onto the machine differs:
//     products      -> DMMA.8x8x4 (D = X Z^T form, no operand movement)
//     chol_inv      -> Gauss-Jordan sweeps with shuffle-exchanged pivot row/column
//     _sym          -> shuffle transpose
// Used for 9 <= d <= 16 (quadrotor d = 13, synthetic d = 12); small d stays on the multi-problem-
// per-warp kernels of hop_select_body.cuh.
#pragma once
#include "hop_mma.cuh"
#include "hop_select_body.cuh"   // SelectArgs, FusedArgs, FusedConst, wrap_pi, ArgMin

namespace hop { namespace mma {

// per-warp shared memory (doubles): LU scratch (512) + EV(16) QE(16) DU(8) YV(16) + pad (8)  = 576,
// then two TMA staging buffers (A_k | B_k | X_k,X_{k+1} | U_k | a_k) of kStage doubles and two mbarriers.
constexpr int kStage = 240;
constexpr int kWarpScratch = 576 + 2 * kStage + 8;

template <int D>
struct PrefixL { Mat eb, fb, gb; };

HOP_DEVICE double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += simt::shfl_xor(v, o, 32);
    return v;
}

// Stage + prefix of step k (horizon_selection.py:57-75).  Qs: symmetrised Q_k; A: A_k; Bm: B_k (d x m);
// RinvT: (R^-1)^T.  All in layout L with zero padding.
template <int D, int M>
HOP_DEVICE void stage_prefix_step(int k, PrefixL<D>& P, const Mat& Qs, const Mat& A, const Mat& Bm, const Mat& RinvT,
                                  const LaneGeo& L, double* scratch, double jitter, int max_tries, int& status) {
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4, KBM = (M + 3) / 4, NTM = (M + 7) / 8;
    Mat E;
    chol_inv<D>(Qs, E, L, scratch, jitter, max_tries, status);                 // E_k = chol_inv(Q_k)          (:59)
    Mat W;
    if (k > 0) {
        Mat S;
        mat_add(S, E, P.gb);
        mat_sym(S, L);
        chol_inv<D>(S, W, L, scratch, jitter, max_tries, status);              // W = chol_inv(E_k + Gbar)     (:72)
    }
    Mat Ft, G;
    mma_nt<NT, NT, KB, false>(Ft, A, E);                                       // F_k^T = A_k E_k
    mma_nt<NT, NT, KB, false>(G, Ft, A);                                       // (A_k E_k) A_k^T              (:61)
    {
        Mat BR;
        mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);                            // B_k R^-1
        mma_nt<NT, NT, KBM, true>(G, BR, Bm);                                  // + (B_k R^-1) B_k^T
    }
    mat_sym(G, L);                                                             // G_k = sym(...)               (:64)
    if (k == 0) {
        mma_nt<NT, NT, KB, false>(P.fb, E, A);                                 // F_0 = E_0 A_0^T              (:60)
        mat_copy(P.eb, E);
        mat_copy(P.gb, G);
    } else {
        Mat T1, acc;
        mma_nt<NT, NT, KB, false>(T1, P.fb, W);                                // Fbar W                       (:73)
        mma_nt<NT, NT, KB, false>(acc, T1, P.fb);                              // (Fbar W) Fbar^T
        mat_sub(P.eb, P.eb, acc);
        mat_sym(P.eb, L);                                                      // Ebar                         (:73)
        mma_nt<NT, NT, KB, false>(acc, T1, Ft);                                // (Fbar W) F_k  -> new Fbar    (:74)
        mma_nt<NT, NT, KB, false>(T1, Ft, W);                                  // F_k^T W                      (:75)
        mat_copy(P.fb, acc);
        mma_nt<NT, NT, KB, false>(acc, T1, Ft);                                // (F_k^T W) F_k
        mat_sub(P.gb, G, acc);
        mat_sym(P.gb, L);                                                      // Gbar                         (:75)
    }
}

// Query for horizon t = k+1 (:77-86).  Returns P0 = chol_inv(X0) (the caller forms 0.5 z0^T P0 z0).
template <int D>
HOP_DEVICE void query_step(const PrefixL<D>& P, const Mat& QTs, Mat& P0, const LaneGeo& L, double* scratch, double jitter,
                           int max_tries, int& status) {
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4;
    Mat Xt, Wt;
    chol_inv<D>(QTs, Xt, L, scratch, jitter, max_tries, status);               // X_t = chol_inv(QT_t)         (:79)
    {
        Mat S;
        mat_add(S, Xt, P.gb);
        mat_sym(S, L);
        chol_inv<D>(S, Wt, L, scratch, jitter, max_tries, status);             // W_t                          (:82)
    }
    Mat T3, X0;
    mma_nt<NT, NT, KB, false>(T3, P.fb, Wt);                                   // Fbar W_t
    mma_nt<NT, NT, KB, false>(X0, T3, P.fb);                                   // (Fbar W_t) Fbar^T
    mat_sub(X0, P.eb, X0);
    mat_sym(X0, L);                                                            // X0                           (:83)
    chol_inv<D>(X0, P0, L, scratch, jitter, max_tries, status);                // P0                           (:84)
}

// ---- LQR-boundary form: drop-in for propagator_all_Jt_aug + argmin ---------------------------------
template <int D, int M>
HOP_DEVICE void select_generic_body(const SelectArgs& p, int b_raw, double* scratch) {
    LaneGeo L;
    L.init();
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    const size_t rinv_inst = (size_t)(p.rinv_step_stride ? p.N : 1) * M * M;
    Mat RinvT;
    mat_load_t(RinvT, p.R_inv + (size_t)b * rinv_inst, M, M, M, L);
    double zr[2], zc[2][2];
#pragma unroll
    for (int I = 0; I < 2; ++I) zr[I] = (L.row(I) < D) ? p.z0[(size_t)b * D + L.row(I)] : 0.0;
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s) zc[J][s] = (L.col(J, s) < D) ? p.z0[(size_t)b * D + L.col(J, s)] : 0.0;
    PrefixL<D> P;
    mat_zero(P.eb); mat_zero(P.fb); mat_zero(P.gb);
    int status = 0;
    ArgMin am;
    am.init();
    const double wexp = p.w_explicit ? p.w_explicit[b] : 0.0;
    const size_t base = (size_t)b * p.N;
    for (int k = 0; k < p.T_max; ++k) {
        {
            Mat A, Bm, Qs;
            mat_load(A, p.A_aug + (base + k) * D * D, D, D, D, L);
            mat_load(Bm, p.B_aug + (base + k) * D * M, D, M, M, L);
            mat_load(Qs, p.Q_aug + (base + k) * D * D, D, D, D, L);
            if (p.rinv_step_stride) mat_load_t(RinvT, p.R_inv + (size_t)b * rinv_inst + (size_t)k * p.rinv_step_stride, M, M, M, L);
            mat_sym(Qs, L);                                                    // chol_inv symmetrises its input (utils.py:74)
            stage_prefix_step<D, M>(k, P, Qs, A, Bm, RinvT, L, scratch, p.jitter, p.max_tries, status);
        }
        Mat P0;
        {
            Mat QTs;
            mat_load(QTs, p.QT + (base + k) * D * D, D, D, D, L);
            mat_sym(QTs, L);
            query_step<D>(P, QTs, P0, L, scratch, p.jitter, p.max_tries, status);
        }
        double part = 0.0;                                                     // 0.5 z0^T P0 z0               (:85)
        HOP_FOR_ELEMS(I, J, s) part = fma(zr[I] * P0.v[I][J][s], zc[J][s], part);
        const double Jt = 0.5 * warp_sum(part);
        if (L.lane == 0 && valid) {
            p.J_out[(size_t)b * p.T_max + k] = Jt;
            const int t = k + 1;
            if (t >= p.T_min) am.push(Jt + wexp * (double)t, t);
        }
    }
    if (L.lane == 0 && valid) {
        p.T_out[b] = am.idx;
        p.Jstar_out[b] = am.best;
        p.status[b] = status;
    }
}

// ---- fused form: augmented embedding built in registers from (A_k, B_k, a_k, X, U) -------------------
//
// MODE 0 (exact): every block of augmented.py is materialised (in registers) and goes through the
//   generic chol_inv, exactly as the reference does.
// MODE 1 (fast): same function, restructured where the embedding makes a block inverse available in
//   closed form (all algebraically exact, the jitter eps = 1e-9 is kept where the reference adds it):
//     * E_k = chol_inv(Q_aug[k]):  Q_aug + eps I = [[Qs + eps I, q], [q^T, c + eps]] with the SAME top-left
//       block for every k, so with K = (Qs + eps I)^-1 (once per launch), y = K q, sigma = c + eps - q^T y:
//           E_k = [[K + y y^T / sigma, -y / sigma], [-y^T / sigma, 1 / sigma]]
//     * X_t = chol_inv(QT_t): same with K' = (P + eps I)^-1, p = P e and the cancellation-free pivot
//           sigma' = rho + eps + eps * (K' p)^T e     (= e^T P e + rho + eps - p^T K' p exactly)
//     * z0 = e_n, so J(t) = 0.5 P0[n][n] = 0.5 / (last LDL^T pivot of X0 + eps I): forward elimination only.
//     * _sym is kept on the carried state (Ebar, Gbar); it is dropped where the operand is symmetric by
//       construction (E_k + Gbar, X_t + Gbar) or only feeds the pivots (G_k, X0).
//   Any non-positive pivot falls back to the generic chol_inv (jitter ladder / LU) of MODE 0.
template <int D, int M>
struct FastConst {
    static constexpr int n = D - 1;
    static constexpr int KQ = FusedConst<D, M>::SIZE, KP = KQ + n * n, FLAG = KP + n * n, SIZE = FLAG + 2;
};

// One warp computes K = (Qs + eps I)^-1 and K' = (P + eps I)^-1 into the CTA constant block.
template <int D, int M>
HOP_DEVICE void fast_const_fill_warp(const FusedArgs& p, double* cst, double* scratch) {
    using FC = FusedConst<D, M>;
    using XC = FastConst<D, M>;
    constexpr int n = D - 1;
    LaneGeo L;
    L.init();
    int st = 0;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        Mat S, Kinv;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            S.v[I][J][s] = (R < n && C < n) ? cst[(which ? FC::PF : FC::QS) + R * n + C] : 0.0;
        }
        chol_inv<n>(S, Kinv, L, scratch, p.jitter, p.max_tries, st);
        mat_sym(Kinv, L);
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            if (R < n && C < n) cst[(which ? XC::KP : XC::KQ) + R * n + C] = Kinv.v[I][J][s];
        }
    }
    if (L.lane == 0) cst[XC::FLAG] = (st != 0) ? 1.0 : 0.0;   // ladder needed: the closed forms do not apply
    // FLAG + 1: K = (Qs + eps I)^-1 is DIAGONAL (a diagonal running weight Q, as in every reference case): the stage product
    // A_k E_k then splits into a column scaling and a rank-1 term (hop_select_pipe_body.cuh)
    simt::sync();
    bool dg = true;
    for (int i = L.lane; i < n * n; i += 32)
        if (i / n != i % n) dg = dg && (cst[XC::KQ + i] == 0.0);
    dg = simt::all(dg);
    if (L.lane == 0) cst[XC::FLAG + 1] = dg ? 1.0 : 0.0;
}

// Last LDL^T pivot of (S + eps I) by forward elimination (no back-substitution).  ok &= all pivots > 0.
template <int D>
HOP_DEVICE double last_pivot(const Mat& S, double eps, const LaneGeo& L, bool& ok) {
    Mat a;
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        a.v[I][J][s] = S.v[I][J][s] + ((R == C && R < D) ? eps : 0.0);
    }
    double p = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int Ij = j >> 3, gj = rho_inv(j & 7), Jj = j >> 3, tj = j & 3, sj = (j & 7) >> 2;
        p = simt::shfl(a.v[Ij][Jj][sj], (gj << 2) | tj, 32);
        ok = ok && (p > 0.0) && (p <= 1.7976931348623157e308);   // +Inf is non-finite input (utils.py:75), not a pivot
        if (j == D - 1) break;
        const double rinv = pivot_rcp(p);
        // rows/cols <= j are dead from here on: once j >= 8 only tile (1,1) is live
        double pr[2][2];
#pragma unroll
        for (int J = (j >= 8 ? 1 : 0); J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) pr[J][s] = simt::shfl(a.v[Ij][J][s], (gj << 2) | L.t, 32);
#pragma unroll
        for (int I = (j >= 8 ? 1 : 0); I < 2; ++I) {
            const double f = simt::shfl(a.v[I][Jj][sj], (L.g << 2) | tj, 32) * rinv;
#pragma unroll
            for (int J = (j >= 8 ? 1 : 0); J < 2; ++J)
#pragma unroll
                for (int s = 0; s < 2; ++s) a.v[I][J][s] = fma(-f, pr[J][s], a.v[I][J][s]);
        }
    }
    return p;
}

template <int D, int M, int MODE>
HOP_DEVICE void select_fused_body(const FusedArgs& p, int b_raw, double* scratch, const double* cst) {
    using FC = FusedConst<D, M>;
    using XC = FastConst<D, M>;
    constexpr int n = D - 1;
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4;
    LaneGeo L;
    L.init();
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    if (!valid || (p.skip && p.skip[b])) return;   // warp-uniform: one problem per warp
    double* EV = scratch + 512;   // e = wrap(X - xg)
    double* QE = EV + 16;         // Q e  /  P e
    double* DU = QE + 16;         // U_k - u_ref
    double* YV = DU + 8;          // K q  /  K' p      (fast mode)
    int status = 0;

    // R_inv = chol_inv(sym(R)) (augmented.py:23), then its transpose as the Z operand of B R^-1
    Mat RinvT;
    {
        Mat Rs, Ri;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            Rs.v[I][J][s] = (R < M && C < M) ? cst[FC::RS + R * M + C] : 0.0;
        }
        chol_inv<M>(Rs, Ri, L, scratch, p.jitter, p.max_tries, status);
        mat_transpose(RinvT, Ri, L);
    }
    const bool isx = L.lane < n;
    const int lx = isx ? L.lane : 0;
    const double xg_l = isx ? p.xg[(size_t)b * n + L.lane] : 0.0;
    const bool wrap_l = isx && ((p.wrap_mask >> L.lane) & 1u);
    const double w = p.w[b];
    const bool closed_ok = (MODE == 1) && (cst[XC::FLAG] == 0.0);
    // the lane that owns element (n, n) of a block: J = 0.5 P0[n][n] because z0 = e_n (augmented.py:59)
    constexpr int own_I = n >> 3, own_J = n >> 3, own_s = (n & 7) >> 2;
    constexpr int own_lane = (rho_inv(n & 7) << 2) | (n & 3);

    PrefixL<D> P;
    mat_zero(P.eb); mat_zero(P.fb); mat_zero(P.gb);
    ArgMin am;
    am.init();
    const size_t baseN = (size_t)b * p.N;
    const size_t baseX = (size_t)b * (p.N + 1);

    // ---- TMA-staged input stream: step k+1's (A, B, X_k+1, X_k+2, U, a) land in shared memory while step k
    // computes.  One elected lane arms the stage's mbarrier with the byte count and issues the bulk copies.
    static_assert(n * n + n * M + 2 * n + M + n <= kStage, "staging buffer too small");
    static_assert((n * n) % 2 == 0 && (n * M) % 2 == 0 && (2 * n) % 2 == 0 && M % 2 == 0 && n % 2 == 0,
                  "bulk copies need 16-byte multiples");
    double* stage0 = scratch + 576;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(scratch + 576 + 2 * kStage);
    constexpr int oA = 0, oB = n * n, oX = oB + n * M, oU = oX + 2 * n, oR = oU + M;
    auto issue = [&](int kk) {
        double* st = stage0 + (kk & 1) * kStage;
        unsigned long long* bar = bars + (kk & 1);
        const unsigned bytes = 8u * (n * n + n * M + 2 * n + M + (p.a_resid ? n : 0));
        simt::mbar_expect_tx(bar, bytes);
        simt::bulk_g2s(st + oA, p.A + (baseN + kk) * n * n, 8u * n * n, bar);
        simt::bulk_g2s(st + oB, p.Bm + (baseN + kk) * n * M, 8u * n * M, bar);
        simt::bulk_g2s(st + oX, p.X + (baseX + kk) * n, 8u * 2 * n, bar);
        simt::bulk_g2s(st + oU, p.U + (size_t)b * p.u_stride + (size_t)kk * M, 8u * M, bar);
        if (p.a_resid) simt::bulk_g2s(st + oR, p.a_resid + (baseN + kk) * n, 8u * n, bar);
    };
    if (L.lane == 0) {
        simt::mbar_init(bars, 1);
        simt::mbar_init(bars + 1, 1);
        simt::mbar_fence_init();
    }
    simt::sync();
    if (L.lane == 0) issue(0);

    for (int k = 0; k < p.T_max; ++k) {
        simt::sync();                                                           // everyone is done with the other stage
        if (L.lane == 0 && k + 1 < p.T_max) issue(k + 1);
        simt::mbar_wait(bars + (k & 1), (unsigned)((k >> 1) & 1));
        const double* stg = stage0 + (k & 1) * kStage;
        const double* Ak = stg + oA;
        const double* Bk = stg + oB;
        const double* Xk = stg + oX;
        double ev = 0.0;
        if (isx) {
            ev = Xk[L.lane] - xg_l;                                             // e = wrap(X_k - xg)  (augmented.py:28)
            if (wrap_l) ev = wrap_pi(ev);
            EV[L.lane] = ev;
        }
        if (L.lane < M) DU[L.lane] = stg[oU + L.lane] - cst[FC::UREF + L.lane];
        simt::sync();
        double qe = 0.0, qc = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                qe = fma(cst[FC::QRAW + L.lane * n + j], EV[j], qe);            // (Q e)_i
                qc = fma(EV[j], cst[FC::QRAW + j * n + L.lane], qc);            // (e^T Q)_i
            }
            QE[L.lane] = qe;
        }
        const double eQe = warp_sum(isx ? qc * ev : 0.0);
        const double corner = eQe + 2.0 * w + p.rho_reg;                        // augmented.py:37
        simt::sync();
        Mat A, Bm;
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
        // ---------------- stage: E_k
        Mat E;
        bool have_E = false;
        if (closed_ok) {
            double y = 0.0;
            if (isx) {
#pragma unroll
                for (int j = 0; j < n; ++j) y = fma(cst[XC::KQ + L.lane * n + j], QE[j], y);   // y = K q
                YV[L.lane] = y;
            }
            const double sigma = (corner + p.jitter) - warp_sum(isx ? qe * y : 0.0);
            simt::sync();
            if (simt::all(sigma > 0.0)) {
                const double rs = 1.0 / sigma;
                HOP_FOR_ELEMS(I, J, s) {
                    const int R = L.row(I), C = L.col(J, s);
                    double e = 0.0;
                    if (R < n && C < n) e = fma(YV[R] * YV[C], rs, cst[XC::KQ + R * n + C]);
                    else if (R < n && C == n) e = -YV[R] * rs;
                    else if (R == n && C < n) e = -YV[C] * rs;
                    else if (R == n && C == n) e = rs;
                    E.v[I][J][s] = e;
                }
                have_E = true;
            }
        }
        if (!have_E) {
            Mat Qs;
            HOP_FOR_ELEMS(I, J, s) {
                const int R = L.row(I), C = L.col(J, s);
                double q = 0.0;
                if (R < n && C < n) q = cst[FC::QS + R * n + C];
                else if (R < n && C == n) q = QE[R];
                else if (R == n && C < n) q = QE[C];
                else if (R == n && C == n) q = corner;
                Qs.v[I][J][s] = q;
            }
            if (MODE == 0) chol_inv<D>(Qs, E, L, scratch, p.jitter, p.max_tries, status);   // E_k = chol_inv(Q_k) (:59)
            else chol_inv_cold<D>(Qs, E, scratch, p.jitter, p.max_tries, status);
        }
        // ---------------- prefix
        {
            constexpr int KBM = (M + 3) / 4, NTM = (M + 7) / 8;
            Mat W;
            if (k > 0) {
                Mat S;
                mat_add(S, E, P.gb);
                if (MODE == 0) mat_sym(S, L);
                chol_inv<D>(S, W, L, scratch, p.jitter, p.max_tries, status);  // W = chol_inv(E_k + Gbar)     (:72)
            }
            Mat Ft, G;
            mma_nt<NT, NT, KB, false>(Ft, A, E);                               // F_k^T = A_k E_k
            mma_nt<NT, NT, KB, false>(G, Ft, A);                               // (A_k E_k) A_k^T              (:61)
            {
                Mat BR;
                mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);                    // B_k R^-1
                mma_nt<NT, NT, KBM, true>(G, BR, Bm);                          // + (B_k R^-1) B_k^T
            }
            if (MODE == 0 || k == 0) mat_sym(G, L);                            // G_k = sym(...)               (:64)
            if (k == 0) {
                mma_nt<NT, NT, KB, false>(P.fb, E, A);                         // F_0 = E_0 A_0^T              (:60)
                mat_copy(P.eb, E);
                mat_copy(P.gb, G);
            } else {
                Mat T1, acc;
                mma_nt<NT, NT, KB, false>(T1, P.fb, W);                        // Fbar W                       (:73)
                mma_nt<NT, NT, KB, false>(acc, T1, P.fb);                      // (Fbar W) Fbar^T
                mat_sub(P.eb, P.eb, acc);
                mat_sym(P.eb, L);                                              // Ebar                         (:73)
                mma_nt<NT, NT, KB, false>(acc, T1, Ft);                        // (Fbar W) F_k  -> new Fbar    (:74)
                mma_nt<NT, NT, KB, false>(T1, Ft, W);                          // F_k^T W                      (:75)
                mat_copy(P.fb, acc);
                mma_nt<NT, NT, KB, false>(acc, T1, Ft);                        // (F_k^T W) F_k
                mat_sub(P.gb, G, acc);
                mat_sym(P.gb, L);                                              // Gbar                         (:75)
            }
        }
        // ---------------- terminal block QT_{k+1} from X[k+1] (augmented.py:78-86) and the query (:77-86)
        simt::sync();
        double et = 0.0;
        if (isx) {
            et = Xk[n + L.lane] - xg_l;
            if (wrap_l) et = wrap_pi(et);
            EV[L.lane] = et;
        }
        simt::sync();
        double px = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) px = fma(cst[FC::PF + L.lane * n + j], EV[j], px);   // (P e)_i
            QE[L.lane] = px;
        }
        const double ePe = warp_sum(isx ? et * px : 0.0);
        simt::sync();
        Mat Xt;
        bool have_Xt = false;
        if (closed_ok) {
            double y = 0.0;
            if (isx) {
#pragma unroll
                for (int j = 0; j < n; ++j) y = fma(cst[XC::KP + L.lane * n + j], QE[j], y);   // y' = K' p
                YV[L.lane] = y;
            }
            const double sigma = (p.rho_reg + p.jitter) + p.jitter * warp_sum(isx ? y * et : 0.0);
            simt::sync();
            if (simt::all(sigma > 0.0)) {
                const double rs = 1.0 / sigma;
                HOP_FOR_ELEMS(I, J, s) {
                    const int R = L.row(I), C = L.col(J, s);
                    double x = 0.0;
                    if (R < n && C < n) x = fma(YV[R] * YV[C], rs, cst[XC::KP + R * n + C]);
                    else if (R < n && C == n) x = -YV[R] * rs;
                    else if (R == n && C < n) x = -YV[C] * rs;
                    else if (R == n && C == n) x = rs;
                    Xt.v[I][J][s] = x;
                }
                have_Xt = true;
            }
        }
        if (!have_Xt) {
            Mat QTs;
            HOP_FOR_ELEMS(I, J, s) {
                const int R = L.row(I), C = L.col(J, s);
                double q = 0.0;
                if (R < n && C < n) q = cst[FC::PF + R * n + C];
                else if (R < n && C == n) q = QE[R];
                else if (R == n && C < n) q = QE[C];
                else if (R == n && C == n) q = 2.0 * (0.5 * ePe) + p.rho_reg;
                QTs.v[I][J][s] = q;
            }
            if (MODE == 0) chol_inv<D>(QTs, Xt, L, scratch, p.jitter, p.max_tries, status);  // X_t = chol_inv(QT_t) (:79)
            else chol_inv_cold<D>(QTs, Xt, scratch, p.jitter, p.max_tries, status);
        }
        double Jt = 0.0;
        {
            Mat Wt;
            {
                Mat S;
                mat_add(S, Xt, P.gb);
                if (MODE == 0) mat_sym(S, L);
                chol_inv<D>(S, Wt, L, scratch, p.jitter, p.max_tries, status); // W_t                          (:82)
            }
            Mat T3, X0;
            mma_nt<NT, NT, KB, false>(T3, P.fb, Wt);                           // Fbar W_t
            mma_nt<NT, NT, KB, false>(X0, T3, P.fb);                           // (Fbar W_t) Fbar^T
            mat_sub(X0, P.eb, X0);
            bool done = false;
            if (MODE == 1) {
                bool ok = true;
                const double piv = last_pivot<D>(X0, p.jitter, L, ok);
                Jt = 0.5 / piv;
                done = simt::all(ok);
            }
            if (!done) {
                Mat P0;
                mat_sym(X0, L);                                                // X0                           (:83)
                if (MODE == 0) chol_inv<D>(X0, P0, L, scratch, p.jitter, p.max_tries, status);   // P0          (:84)
                else chol_inv_cold<D>(X0, P0, scratch, p.jitter, p.max_tries, status);
                Jt = 0.5 * simt::shfl(P0.v[own_I][own_J][own_s], own_lane, 32);
            }
        }
        if (L.lane == 0 && valid) {
            p.J_out[(size_t)b * p.T_max + k] = Jt;
            const int t = k + 1;
            if (t >= p.T_min) am.push(Jt, t);
        }
    }
    (void)lx;
    if (L.lane == 0 && valid) {
        p.T_out[b] = am.idx;
        p.Jstar_out[b] = am.best;
        p.status[b] = status;
    }
}

}}  // namespace hop::mma
