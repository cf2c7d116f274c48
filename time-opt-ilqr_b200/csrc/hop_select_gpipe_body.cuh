// hop_select_gpipe_body.cuh -- software-pipelined horizon selection at the LQR boundary (hop_select_f64), one
// problem per warp: the drop-in for horizon_selection.py:36-86 propagator_all_Jt_aug + argmin on caller-provided
// augmented blocks (A_k, B_k, Q_k, QT_t arbitrary, z0 arbitrary).
//
// Same algorithm, operation order and _sym placement as the sequential body (hop_select_mma_body.cuh
// select_generic_body); what changes is the SCHEDULE, as in hop_select_pipe_body.cuh:
//   sweep A (three interleaved chains, they only depend on the prefix state P_k):
//       W_{k+1} = chol_inv(E_{k+1} + Gbar_k)     W_t = chol_inv(X_t + Gbar_k), t = k+1
//       forward elimination of [X0_{t-1} + eps I | z0]  ->  J(t-1) = 0.5 z0^T (X0 + eps I)^-1 z0 = 0.5 sum_j y_j^2 / p_j
//   products: query (X0_t, lower tiles) and prefix step k+1 (three k-blocks on the tensor pipe + rank-1 last column)
//   sweep B (two interleaved chains, inputs only):  E_{k+2} = chol_inv(Q_{k+2}),  X_{t+1} = chol_inv(QT_{t+1})
// so a step costs two sweep latencies instead of five.  The four input blocks of a step are staged three steps ahead by
// cp.async (8-byte granules: a 13 x 13 block is not 16-byte aligned, so the TMA bulk copy of the fused kernel does not
// apply) into a per-warp triple buffer.  Any non-positive pivot / non-finite value sends the problem to the sequential
// body (jitter ladder, LU fallback, status word).
#pragma once
#include "hop_select_pipe_body.cuh"

namespace hop { namespace mma {

template <int D, int M>
struct GpipeSlab {   // doubles per warp
    static constexpr int STAGE = (3 * D * D + D * M + 1) & ~1;     // A | B | Q | QT of one step
    static constexpr int oA = 0, oB = D * D, oQ = oB + D * M, oT = oQ + D * D;
    static constexpr int SIZE = 3 * STAGE;
};
static_assert(GpipeSlab<13, 4>::SIZE >= kWarpScratch, "the sequential cold path re-uses the slab");

// two interleaved Gauss-Jordan sweeps
template <int D, int GI, int GS>
HOP_DEVICE void gj2_group(Mat& a1, Mat& a2, const LaneGeo& L, int& signs) {
    constexpr int first = 8 * GI + 4 * GS;
    constexpr int cnt = (D - first) < 4 ? (D - first) : 4;
#pragma unroll 1
    for (int tj = 0; tj < cnt; ++tj) {
        gj_pivot<D, GI, GS>(a1, tj, L, signs);
        gj_pivot<D, GI, GS>(a2, tj, L, signs);
    }
}

// forward elimination of a symmetric matrix on its lower tiles WITH a right-hand side z (held at this lane's rows):
// acc += y_j^2 / p_j, z_i -= (a_ij / p_j) y_j for i > j.   0.5 * acc = 0.5 z^T A^-1 z after the last pivot.
template <int D, int GI, int GS>
HOP_DEVICE void fez_pivot_lower(Mat& a, double (&z)[2], double& acc, int tj, const LaneGeo& L, int& signs) {
    const int gj = 2 * tj + GS;
    const int j = 8 * GI + 4 * GS + tj;
    const double p = simt::shfl(a.v[GI][GI][GS], (gj << 2) | tj, 32);
    signs |= hi_word(p);
    const double rinv = pivot_rcp3(p);
    const double yj = simt::shfl(z[GI], gj << 2, 32);
    acc = fma(yj * yj, rinv, acc);
    if (j == D - 1) return;
    double pr[2][2], f[2];
#pragma unroll
    for (int J = GI; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s) pr[J][s] = simt::shfl(a.v[J][GI][GS], ((2 * L.t + s) << 2) | tj, 32);
#pragma unroll
    for (int I = GI; I < 2; ++I) f[I] = simt::shfl(a.v[I][GI][GS], (L.g << 2) | tj, 32) * rinv;
    if (GI == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[0][0][s] = fma(-f[0], pr[0][s], a.v[0][0][s]);
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[1][0][s] = fma(-f[1], pr[0][s], a.v[1][0][s]);
        z[0] = (L.row(0) > j) ? fma(-f[0], yj, z[0]) : z[0];
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) a.v[1][1][s] = fma(-f[1], pr[1][s], a.v[1][1][s]);
    z[1] = (L.row(1) > j) ? fma(-f[1], yj, z[1]) : z[1];
}

template <int D, int GI, int GS>
HOP_DEVICE void gj3z_group(Mat& a1, Mat& a2, Mat& x, double (&z)[2], double& acc, const LaneGeo& L, int& signs) {
    constexpr int first = 8 * GI + 4 * GS;
    constexpr int cnt = (D - first) < 4 ? (D - first) : 4;
#pragma unroll 1
    for (int tj = 0; tj < cnt; ++tj) {
        gj_pivot<D, GI, GS>(a1, tj, L, signs);
        gj_pivot<D, GI, GS>(a2, tj, L, signs);
        fez_pivot_lower<D, GI, GS>(x, z, acc, tj, L, signs);
    }
}

template <int D>
HOP_DEVICE void add_jitter(Mat& a, double eps, const LaneGeo& L) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (L.row(I) == L.col(I, s) && L.row(I) < D) a.v[I][I][s] += eps;
}

// Sweep B only reads input blocks, so it has no place on the sequential critical path: with PRE the two inversions of
// every step come from a fully parallel pre-pass (pre_invert_pair: one warp per (problem, step), the same interleaved
// sweep, hence the same bits) and this body stages E_k / X_t instead of Q_k / QT_t -- the same bytes -- and runs one
// sweep latency per step.  A problem whose pre-pass met a non-positive pivot or a non-finite value goes to the
// sequential body, which owns the jitter ladder.
template <int D>
HOP_DEVICE void pre_invert_pair(const double* Q, const double* QT, double jitter, double* E, double* X, int* bad_out) {
    LaneGeo L;
    L.init();
    Mat a1, a2;
    mat_load(a1, Q, D, D, D, L);
    mat_sym(a1, L);
    mat_load(a2, QT, D, D, D, L);
    mat_sym(a2, L);
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (L.row(I) == L.col(I, s) && L.row(I) < D) { a1.v[I][I][s] += jitter; a2.v[I][I][s] += jitter; }
    int signs = 0;
    gj2_group<D, 0, 0>(a1, a2, L, signs); gj2_group<D, 0, 1>(a1, a2, L, signs);
    gj2_group<D, 1, 0>(a1, a2, L, signs); gj2_group<D, 1, 1>(a1, a2, L, signs);
    const bool fin = mat_all_finite(a1) && mat_all_finite(a2);
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        if (R < D && C < D) { E[R * D + C] = a1.v[I][J][s]; X[R * D + C] = a2.v[I][J][s]; }
    }
    if (L.lane == 0 && (signs < 0 || !fin)) *bad_out = 1;
}

// The same inversion, FOUR LANES PER MATRIX (lane q of a group owns columns 4q .. 4q+3, all rows, in registers): a
// pivot costs one broadcast of the pivot column (D shuffles inside the group) instead of the fragment layout's row AND
// column exchange per lane, and a warp inverts eight matrices at once -- the fragment sweep is shuffle bound in a
// pre-pass (14 double shuffles per pivot and matrix, ~10 ms of shuffles alone for 65 536 x 128 steps).  Every element
// goes through the operations of gj_pivot (rinv = pivot_rcp3(p), f_i = a_ij rinv, a_ic <- fma(-f_i, a_jc, a_ic), pivot
// row rinv a_jc, pivot column -f_i, pivot rinv), so the result has the same bits as pre_invert_pair / sweep B.
// src / dst: one D x D row-major block; bad_out: flag of the owning problem.
template <int D>
HOP_DEVICE void pre_invert_cols(const double* src, double jitter, double* dst, int* bad_out, bool live) {
    static_assert(D > 8 && D <= 16, "four lanes x four columns");
    const int q = simt::lane_id() & 3;
    double a[D][4];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int C = 4 * q + cc;
            double v = 0.0;
            if (C < D) {
                v = 0.5 * (src[i * D + C] + src[C * D + i]);                     // utils.py:35-37
                if (i == C) v += jitter;                                         // utils.py:83
            }
            a[i][cc] = v;
        }
    int signs = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int qj = j >> 2, cj = j & 3;                                       // owner lane of column j, its slot
        double colv[D];
#pragma unroll
        for (int i = 0; i < D; ++i) colv[i] = simt::shfl(a[i][cj], qj, 4);
        const double p = colv[j];
        signs |= hi_word(p);
        const double rinv = pivot_rcp3(p);
        const bool own = (q == qj);
        double prow[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) prow[cc] = a[j][cc];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double f = colv[i] * rinv;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                double v;
                if (i == j) v = fma(rinv, prow[cc], 0.0);
                else v = fma(-f, prow[cc], a[i][cc]);
                if (cc == cj) v = own ? ((i == j) ? rinv : -f) : v;
                a[i][cc] = v;
            }
        }
    }
    bool fin = true;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int C = 4 * q + cc;
            if (C < D) {
                if (live) dst[i * D + C] = a[i][cc];
                fin = fin && isfinite(a[i][cc]);
            }
        }
    if (live && (signs < 0 || !fin)) *bad_out = 1;
}

// ---- diagonal input blocks -------------------------------------------------------------------------------------------
// Sweep B inverts blocks that only depend on the inputs (Q_k, QT_t).  When such a block is DIAGONAL -- every problem of
// the synthetic HOP-LQR family of SURVEY.md s.8d (Q_k = diag, QT_t = 50 I), and any LQR problem with diagonal weights --
// the Gauss-Jordan sweep of sym(block) + eps I degenerates: pivot j is the j-th diagonal entry, every multiplier is an exact
// zero, and the result is the matrix with pivot_rcp3(p_j) on the diagonal.  The same values are produced here element-wise
// (13 reciprocals instead of two 13-round sweeps: one sweep latency per step instead of two), decided per block pair by a
// warp vote on the staged block, so mixed inputs still take the sweep.  Identical bits either way (asserted).
template <int D>
HOP_DEVICE bool is_diagonal(const Mat& S, const LaneGeo& L) {
    bool z = true;
    HOP_FOR_ELEMS(I, J, s) {
        if (L.row(I) != L.col(J, s)) z = z && (S.v[I][J][s] == 0.0);
    }
    return simt::all(z);
}
template <int D>
HOP_DEVICE void diagonal_inverse(Mat& S, const LaneGeo& L, int& signs) {
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I);
        if (R == L.col(J, s) && R < D) {
            const double p = S.v[I][J][s];
            signs |= pivot_bad(p) ? (int)0x80000000 : 0;                         // <= 0, NaN, +inf: the sequential body decides
            S.v[I][J][s] = pivot_rcp3(p);
        }
    }
}

// Returns true when the problem was solved by the pipelined sweep; false => caller must run the sequential body.
template <int D, int M, bool PRE = false>
HOP_DEVICE bool select_generic_pipe_body(const SelectArgs& p, int b, double* slab) {
    using GS_ = GpipeSlab<D, M>;
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4, KBM = (M + 3) / 4, NTM = (M + 7) / 8;
    constexpr bool R1 = LastCol<D>::split;
    constexpr int KD = R1 ? KB - 1 : KB;
    static_assert(D > 8 && D <= 16, "one-problem-per-warp mapping: 9 <= d <= 16");
    if (PRE && p.pre_bad[b] != 0) return false;
    LaneGeo L;
    L.init();
    const size_t base = (size_t)b * p.N;
    const size_t rinv_inst = (size_t)(p.rinv_step_stride ? p.N : 1) * M * M;
    const double* srcQ = PRE ? p.E_pre : p.Q_aug;                               // PRE: the inverses, same layout
    const double* srcT = PRE ? p.X_pre : p.QT;
    // ---- staging: step s -> buffer s % 3, 8-byte cp.async granules
    auto issue = [&](int s) {
        if (s < p.T_max) {
            double* st = slab + (s % 3) * GS_::STAGE;
            const double* gA = p.A_aug + (base + s) * D * D;
            const double* gB = p.B_aug + (base + s) * D * M;
            const double* gQ = srcQ + (base + s) * D * D;
            const double* gT = srcT + (base + s) * D * D;
            for (int i = L.lane; i < D * D; i += 32) {
                simt::cp_async8(st + GS_::oA + i, gA + i);
                simt::cp_async8(st + GS_::oQ + i, gQ + i);
                simt::cp_async8(st + GS_::oT + i, gT + i);
            }
            for (int i = L.lane; i < D * M; i += 32) simt::cp_async8(st + GS_::oB + i, gB + i);
        }
        simt::cp_async_commit();                                                // (an empty group keeps the wait counts uniform)
    };
    // chol_inv input: sym(block) + eps I  (utils.py:74,83)
    auto load_spd = [&](Mat& S, const double* src) {
        mat_load(S, src, D, D, D, L);
        if (!PRE) {
            mat_sym(S, L);
            add_jitter<D>(S, p.jitter, L);
        }
    };
    auto ident = [&](Mat& S) { HOP_FOR_ELEMS(I, J, s) S.v[I][J][s] = (L.row(I) == L.col(J, s)) ? 1.0 : 0.0; };
    int signs = 0;
    const bool diag_ok = (p.no_diag_fastpath == 0);
    // sweep B on a pair of input blocks (each already sym(.) + eps I): element-wise when both are diagonal
    auto sweep_b = [&](Mat& S1, Mat& S2) {
        if (diag_ok && is_diagonal<D>(S1, L) && is_diagonal<D>(S2, L)) {
            diagonal_inverse<D>(S1, L, signs);
            diagonal_inverse<D>(S2, L, signs);
        } else {
            gj2_group<D, 0, 0>(S1, S2, L, signs); gj2_group<D, 0, 1>(S1, S2, L, signs);
            gj2_group<D, 1, 0>(S1, S2, L, signs); gj2_group<D, 1, 1>(S1, S2, L, signs);
        }
    };
    Mat RinvT;
    mat_load_t(RinvT, p.R_inv + (size_t)b * rinv_inst, M, M, M, L);
    double zr[2];
#pragma unroll
    for (int I = 0; I < 2; ++I) zr[I] = (L.row(I) < D) ? p.z0[(size_t)b * D + L.row(I)] : 0.0;
    const double wexp = p.w_explicit ? p.w_explicit[b] : 0.0;
    bool bad = false;

    issue(0); issue(1); issue(2);
    simt::cp_async_wait<0>();
    simt::sync();
    // ---------------- prologue: E_0, E_1, X_1 (terminal block of horizon 1 = QT[0]) and prefix step 0
    Mat En, Xt;                                                                // E_{k+1}, X_t of the current iteration
    PrefixL<D> P;
    {
        Mat E0;
        load_spd(E0, slab + 0 * GS_::STAGE + GS_::oQ);
        load_spd(Xt, slab + 0 * GS_::STAGE + GS_::oT);
        if (!PRE) sweep_b(E0, Xt);
        Mat dummy;
        ident(dummy);
        if (p.T_max > 1) load_spd(En, slab + 1 * GS_::STAGE + GS_::oQ); else ident(En);
        if (!PRE) sweep_b(En, dummy);
        Mat A, Bm, Ft, G, BR;
        mat_load(A, slab + GS_::oA, D, D, D, L);
        mat_load(Bm, slab + GS_::oB, D, M, M, L);
        mma_nt<NT, NT, KB, false>(Ft, A, E0);                                  // F_0^T = A_0 E_0
        mma_nt<NT, NT, KB, false>(G, Ft, A);                                   // (A_0 E_0) A_0^T              (:61)
        mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);
        mma_nt<NT, NT, KBM, true>(G, BR, Bm);
        mat_sym(G, L);                                                         // G_0                          (:64)
        mma_nt<NT, NT, KB, false>(P.fb, E0, A);                                // F_0 = E_0 A_0^T              (:60)
        mat_copy(P.eb, E0);
        mat_copy(P.gb, G);
    }
    Mat X0;
    ident(X0);
    double zq[2] = {0.0, 0.0};                                                 // right-hand side of the elimination in flight
    ArgMin am;
    am.init();

    for (int k = 0; k < p.T_max; ++k) {
        const bool last = (k + 1 == p.T_max);
        simt::sync();                                                           // everyone is done with buffer k % 3
        issue(k + 3);
        // ---------------- sweep A
        Mat W, Wt;
        mat_add(W, En, P.gb);
        mat_sym(W, L);
        add_jitter<D>(W, p.jitter, L);                                          // sym(E_{k+1} + Gbar_k) + eps I     (:72)
        mat_add(Wt, Xt, P.gb);
        mat_sym(Wt, L);
        add_jitter<D>(Wt, p.jitter, L);                                         // sym(X_t + Gbar_k) + eps I         (:82)
        add_jitter<D>(X0, p.jitter, L);                                         // X0_{t-1} + eps I                  (:84)
        double acc = 0.0;
        gj3z_group<D, 0, 0>(W, Wt, X0, zq, acc, L, signs);
        gj3z_group<D, 0, 1>(W, Wt, X0, zq, acc, L, signs);
        gj3z_group<D, 1, 0>(W, Wt, X0, zq, acc, L, signs);
        gj3z_group<D, 1, 1>(W, Wt, X0, zq, acc, L, signs);
        bad = bad || (signs < 0) || !(acc == acc) || !(fabs(acc) < HUGE_VAL);
        if (simt::ballot(bad) != 0u) { simt::cp_async_wait<0>(); simt::sync(); return false; }
        if (k > 0 && L.lane == 0) {
            const double Jt = 0.5 * acc;                                        // J(t-1)                            (:85)
            p.J_out[(size_t)b * p.T_max + (k - 1)] = Jt;
            if (k >= p.T_min) am.push(Jt + wexp * (double)k, k);
        }
        // ---------------- query products of horizon t = k+1 (:83), lower tiles
        double fb_r[2], fb_c[2][2];
        if (R1) { last_col_rows<D>(fb_r, P.fb, L); last_col_cols<D>(fb_c, P.fb, L); }
        {
            Mat T3, accm;
            mma_nt<NT, NT, KD, false>(T3, P.fb, Wt);
            if (R1) { double c[2][2]; last_col_cols<D>(c, Wt, L); rank1_add<false>(T3, fb_r, c); }
            mma_nt_lower<KD, false>(accm, T3, P.fb);
            if (R1) { double r[2]; last_col_rows<D>(r, T3, L); rank1_add<true>(accm, r, fb_c); }
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
#pragma unroll
                    for (int s = 0; s < 2; ++s) X0.v[I][J][s] = P.eb.v[I][J][s] - accm.v[I][J][s];
            // sym() on the lower tiles: the diagonal tiles are averaged in place by the elimination reading column j
            // only; the off-diagonal tile (1,0) stands for both halves
        }
        zq[0] = zr[0]; zq[1] = zr[1];
        // ---------------- prefix step k+1 (:57-75)
        if (!last) {
            const double* stg = slab + ((k + 1) % 3) * GS_::STAGE;
            Mat A, Bm, Ft, G;
            mat_load(A, stg + GS_::oA, D, D, D, L);
            mat_load(Bm, stg + GS_::oB, D, M, M, L);
            if (p.rinv_step_stride) mat_load_t(RinvT, p.R_inv + (size_t)b * rinv_inst + (size_t)(k + 1) * p.rinv_step_stride, M, M, M, L);
            double ft_r[2], ft_c[2][2], w_c[2][2];
            mma_nt<NT, NT, KD, false>(Ft, A, En);                              // F_k^T = A_k E_k
            if (R1) {
                double r[2], c[2][2];
                last_col_rows<D>(r, A, L); last_col_cols<D>(c, En, L);
                rank1_add<false>(Ft, r, c);
            }
            mma_nt<NT, NT, KD, false>(G, Ft, A);                               // (A_k E_k) A_k^T              (:61)
            if (R1) {
                double c[2][2];
                last_col_rows<D>(ft_r, Ft, L); last_col_cols<D>(ft_c, Ft, L); last_col_cols<D>(c, A, L);
                rank1_add<false>(G, ft_r, c);
            }
            {
                Mat BR;
                mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);                    // B_k R^-1
                mma_nt<NT, NT, KBM, true>(G, BR, Bm);                          // + (B_k R^-1) B_k^T
            }
            mat_sym(G, L);                                                     // G_k                          (:64)
            Mat T1, accm;
            mma_nt<NT, NT, KD, false>(T1, P.fb, W);                            // Fbar W                       (:73)
            if (R1) { last_col_cols<D>(w_c, W, L); rank1_add<false>(T1, fb_r, w_c); }
            double t1_r[2];
            if (R1) last_col_rows<D>(t1_r, T1, L);
            mma_nt<NT, NT, KD, false>(accm, T1, P.fb);                         // (Fbar W) Fbar^T
            if (R1) rank1_add<false>(accm, t1_r, fb_c);
            mat_sub(P.eb, P.eb, accm);
            mat_sym(P.eb, L);                                                  // Ebar                         (:73)
            mma_nt<NT, NT, KD, false>(accm, T1, Ft);                           // (Fbar W) F_k  -> new Fbar    (:74)
            if (R1) rank1_add<false>(accm, t1_r, ft_c);
            mma_nt<NT, NT, KD, false>(T1, Ft, W);                              // F_k^T W                      (:75)
            if (R1) rank1_add<false>(T1, ft_r, w_c);
            mat_copy(P.fb, accm);
            mma_nt<NT, NT, KD, false>(accm, T1, Ft);                           // (F_k^T W) F_k
            if (R1) { double r[2]; last_col_rows<D>(r, T1, L); rank1_add<false>(accm, r, ft_c); }
            mat_sub(P.gb, G, accm);
            mat_sym(P.gb, L);                                                  // Gbar                         (:75)
            // ---------------- sweep B: E_{k+2} = chol_inv(Q_{k+2}), X_{t+1} = chol_inv(QT[k+1])
            simt::cp_async_wait<1>();                                           // stage k+2 has landed (k+3 may be in flight)
            simt::sync();
            load_spd(Xt, stg + GS_::oT);
            if (k + 2 < p.T_max) load_spd(En, slab + ((k + 2) % 3) * GS_::STAGE + GS_::oQ); else ident(En);
            if (!PRE) sweep_b(En, Xt);
        }
    }
    // ---------------- epilogue: cost of the last horizon
    {
        add_jitter<D>(X0, p.jitter, L);
        double acc = 0.0;
        Mat d1, d2;
        ident(d1); ident(d2);
        gj3z_group<D, 0, 0>(d1, d2, X0, zq, acc, L, signs);
        gj3z_group<D, 0, 1>(d1, d2, X0, zq, acc, L, signs);
        gj3z_group<D, 1, 0>(d1, d2, X0, zq, acc, L, signs);
        gj3z_group<D, 1, 1>(d1, d2, X0, zq, acc, L, signs);
        bad = bad || (signs < 0) || !(acc == acc) || !(fabs(acc) < HUGE_VAL);
        simt::cp_async_wait<0>();
        simt::sync();
        if (simt::ballot(bad) != 0u) return false;
        if (L.lane == 0) {
            const double Jt = 0.5 * acc;
            p.J_out[(size_t)b * p.T_max + (p.T_max - 1)] = Jt;
            if (p.T_max >= p.T_min) am.push(Jt + wexp * (double)p.T_max, p.T_max);
            p.T_out[b] = am.idx;
            p.Jstar_out[b] = am.best;
            p.status[b] = 0;
        }
    }
    return true;
}

}}  // namespace hop::mma
