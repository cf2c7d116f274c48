// hop_ddp_mma.cuh -- backward_pass_truncated (solver.py:156-230) with the n x n products on the FP64 tensor pipe.
//
// backward_pass_warp (hop_ddp_core.cuh) spreads the elements of every product over the lanes of a warp and reads both
// operands of every multiply from shared memory, one unfused multiply and one add per term, in the reference's summation
// order: bit-identical to the thread-per-problem version, but LSU-bound at ~0.10 of the FP64 peak and a quarter of a
// quadrotor HOP-DDP iteration.  Here the matrix-valued quantities of a step live as DMMA register fragments (layout L of
// hop_mma.cuh, 16 x 16 zero-padded):
//       A^T Vxx,  (A^T Vxx) A,  B^T Vxx,  (B^T Vxx) B,  (B^T Vxx) A              39 DMMAs
//       K^T Qux,  Qux^T K,  K^T Quu,  (K^T Quu) K                                  14 DMMAs (inner dimension m = 4: one k-block)
// Every product has the form X Z^T of mma_nt with X, Z in {A^T, B^T, K^T, Qux^T, Quu^T, Vxx (symmetric)}, so the only
// transposes are the transposing loads of A_k, B_k, K, Qux, Quu from shared memory.  The m x m solves (Cholesky ladder of
// utils.py:96-120), the vector recursions (Qx, Qu, Vx) and the finiteness / positive-definiteness decisions are the
// lane code of backward_pass_warp, on the same shared-memory layout.
//
// Numerics: the products accumulate with FMA in k-blocks of four instead of the reference's unfused left-to-right sums; the
// gains differ from backward_pass_warp by a few ulp (asserted <= 1e-12 relative; <= 1e-9 against the reference goldens).
// HOP_MODE_EXACT solves keep the ordered kernel, so their J_hist stays bit-comparable with the oracle.
#pragma once
#include "hop_ddp_core.cuh"
#include "hop_mma.cuh"

namespace hop { namespace ddp {

template <int n, int m>
HOP_DEVICE int backward_pass_mma(const double* A, const double* Bm, const double* X, const double* U, const CostConst& c,
                                 int T, double lm, double* k_out, double* K_out, int* ok, double* sm, int lane) {
    using S = BwSmem<n, m>;
    using mma::Mat;
    static_assert(n > 8 && n <= 16 && m <= 8 && n + m <= 32 && m * m <= 32, "fragment mapping: two row tiles of the state, one of the control");
    constexpr int NT = 2, KB = (n + 3) / 4, KBM = (m + 3) / 4;
    double *Ak = sm + S::AK, *Bk = sm + S::BK, *Qux = sm + S::QUX, *Kk = sm + S::KK, *KtQuu = sm + S::KTQ, *Quu = sm + S::QUU;
    double *Vx = sm + S::VX, *Qx = sm + S::QX, *Vxn = sm + S::VXN, *e = sm + S::EV, *Qu = sm + S::QU, *kap = sm + S::KAP, *du = sm + S::DU;
    mma::LaneGeo L;
    L.init();
    *ok = 0;
    if (T <= 0) return DDP_OK;
    bool fin = true;
    if (lane < n) {
        double v = sub(X[(size_t)T * n + lane], c.xg[lane]);
        if ((c.wrap_mask >> lane) & 1u) v = wrap_pi(v);
        e[lane] = v;
        fin = isfinite(v);
    }
    if (!simt::all(fin)) return DDP_OK;
    simt::sync();
    if (lane < n) {
        const double s = row_dot<n>(c.Qf, e, lane, c.qf_diag());
        Vx[lane] = s;                                                           // Vx[T] = Qf e_T
    }
    Mat Vxx, Qm;
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        const bool in = (R < n && C < n);
        Vxx.v[I][J][s] = in ? 0.5 * add(c.Qf[R * n + C], c.Qf[C * n + R]) : 0.0;   // Vxx[T] = sym(Qf)
        Qm.v[I][J][s] = in ? c.Q[R * n + C] : 0.0;
    }
    constexpr int PA = (n * n + 31) / 32, PB = (n * m + 31) / 32;
    double pa[PA], pb[PB], pxu = 0.0;
    auto fetch = [&](int k) {
#pragma unroll
        for (int u = 0; u < PA; ++u) { const int q = lane + 32 * u; pa[u] = (q < n * n) ? A[(size_t)k * n * n + q] : 0.0; }
#pragma unroll
        for (int u = 0; u < PB; ++u) { const int q = lane + 32 * u; pb[u] = (q < n * m) ? Bm[(size_t)k * n * m + q] : 0.0; }
        pxu = (lane < n) ? X[(size_t)k * n + lane] : ((lane < n + m) ? U[(size_t)k * m + (lane - n)] : 0.0);
    };
    fetch(T - 1);
    for (int k = T - 1; k >= 0; --k) {
        simt::sync();
#pragma unroll
        for (int u = 0; u < PA; ++u) { const int q = lane + 32 * u; if (q < n * n) Ak[q] = pa[u]; }
#pragma unroll
        for (int u = 0; u < PB; ++u) { const int q = lane + 32 * u; if (q < n * m) Bk[q] = pb[u]; }
        const double xu = pxu;
        if (k > 0) fetch(k - 1);
        fin = true;
        if (lane < n) {
            double v = sub(xu, c.xg[lane]);
            if ((c.wrap_mask >> lane) & 1u) v = wrap_pi(v);
            e[lane] = v;
            fin = isfinite(v);
        } else if (lane < n + m) {
            const double v = sub(xu, c.u_ref[lane - n]);
            du[lane - n] = v;
            fin = isfinite(v);
        }
        if (!simt::all(fin)) return DDP_OK;
        simt::sync();
        // ---- Qx = lx + A^T Vx, Qu = lu + B^T Vx (lanes, as in backward_pass_warp)
        if (lane < n) {
            const int i = lane;
            double s = 0.0;
            const double lx = row_dot<n>(c.Q, e, i, c.q_diag());
            for (int l = 0; l < n; ++l) s = add(s, mul(Ak[l * n + i], Vx[l]));
            Qx[i] = add(lx, s);
        } else if (lane < n + m) {
            const int i = lane - n;
            double s = 0.0;
            const double lu = row_dot<m>(c.R, du, i, c.r_diag());
            for (int l = 0; l < n; ++l) s = add(s, mul(Bk[l * m + i], Vx[l]));
            Qu[i] = add(lu, s);
        }
        // ---- matrix products on the tensor pipe
        Mat Qxx;
        {
            Mat At, Bt, AtV, BtV, Quum, Quxm;
            mma::mat_load_t(At, Ak, n, n, n, L);                                // A_k^T
            mma::mat_load_t(Bt, Bk, m, n, m, L);                                // B_k^T (m x n)
            mma::mma_nt<NT, NT, KB, false>(AtV, At, Vxx);                        // A^T Vxx            (Vxx symmetric)
            mma::mma_nt<NT, NT, KB, false>(Qxx, AtV, At);                        // (A^T Vxx) A
            mma::mma_nt<1, NT, KB, false>(BtV, Bt, Vxx);                         // B^T Vxx
            mma::mma_nt<1, 1, KB, false>(Quum, BtV, Bt);                         // (B^T Vxx) B
            mma::mma_nt<1, NT, KB, false>(Quxm, BtV, At);                        // (B^T Vxx) A
            HOP_FOR_ELEMS(I, J, s) {
                const int R = L.row(I), C = L.col(J, s);
                Qxx.v[I][J][s] = add(Qm.v[I][J][s], Qxx.v[I][J][s]);            // Qxx = Q + A^T Vxx A
                if (I == 0 && R < m) {
                    if (C < m) Quu[R * m + C] = add(c.R[R * m + C], Quum.v[I][J][s]);   // Quu = R + B^T Vxx B
                    if (C < n) Qux[R * n + C] = Quxm.v[I][J][s];
                }
            }
        }
        simt::sync();
        // ---- gains (every lane holds Quu_reg; the right-hand sides are spread over the lanes): backward_pass_warp's code
        double Qreg[m * m], Ltmp[m * m];
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) Qreg[i * m + j] = add(0.5 * add(Quu[i * m + j], Quu[j * m + i]), (i == j) ? lm : 0.0);
        const int gp = gains_one_pass<m, n>(Qreg, Qu, Qux, kap, Kk, lane);
        if (gp == 1) return DDP_OK;                                           // solver.py:213-216
        if (gp < 0) {                                                         // the reference sequence (jitter ladder, error codes)
            if (!cholesky_lower<m>(Qreg, Ltmp)) return DDP_OK;
            int rc = chol_solve_warp<m, 1>(Qreg, Qu, kap, 1e-9, 8, lane);
            if (rc) return rc;
            rc = chol_solve_warp<m, n>(Qreg, Qux, Kk, 1e-9, 8, lane);
            if (rc) return rc;
        }
        simt::sync();
        if (lane < m) kap[lane] = -kap[lane];
        for (int q = lane; q < m * n; q += 32) Kk[q] = -Kk[q];
        simt::sync();
        if (lane < m) k_out[(size_t)k * m + lane] = kap[lane];
        for (int q = lane; q < m * n; q += 32) K_out[(size_t)k * m * n + q] = Kk[q];
        for (int q = lane; q < n * m; q += 32) {                                // K^T Quu for the vector recursion
            const int i = q / m, j = q % m;
            double s = 0.0;
            for (int l = 0; l < m; ++l) s = add(s, mul(Kk[l * n + i], Quu[l * m + j]));
            KtQuu[q] = s;
        }
        simt::sync();
        if (lane < n) {
            const int i = lane;
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int l = 0; l < m; ++l) s1 = add(s1, mul(Kk[l * n + i], Qu[l]));
            for (int l = 0; l < m; ++l) s2 = add(s2, mul(Qux[l * n + i], kap[l]));
            for (int l = 0; l < m; ++l) s3 = add(s3, mul(KtQuu[i * m + l], kap[l]));
            Vxn[i] = add(add(add(Qx[i], s1), s2), s3);                         // solver.py:224
        }
        // ---- Vxx = sym(Qxx + K^T Qux + Qux^T K + (K^T Quu) K)   (solver.py:225), inner dimension m: one k-block each
        {
            Mat Kt, Quxt, QuuT, KtQ, t1, t2, t3;
            mma::mat_load_t(Kt, Kk, n, m, n, L);                                // K^T   (n x m)
            mma::mat_load_t(Quxt, Qux, n, m, n, L);                             // Qux^T (n x m)
            mma::mat_load_t(QuuT, Quu, m, m, m, L);                             // Quu^T
            mma::mma_nt<NT, NT, KBM, false>(t1, Kt, Quxt);                       // K^T Qux
            mma::mma_nt<NT, NT, KBM, false>(t2, Quxt, Kt);                       // Qux^T K
            mma::mma_nt<NT, 1, KBM, false>(KtQ, Kt, QuuT);                       // K^T Quu
            mma::mma_nt<NT, NT, KBM, false>(t3, KtQ, Kt);                        // (K^T Quu) K
            HOP_FOR_ELEMS(I, J, s) Vxx.v[I][J][s] = add(add(add(Qxx.v[I][J][s], t1.v[I][J][s]), t2.v[I][J][s]), t3.v[I][J][s]);
            mma::mat_sym(Vxx, L);
        }
        fin = mma::mat_all_finite(Vxx);
        bool vfin = true;
        if (lane < n) { Vx[lane] = Vxn[lane]; vfin = isfinite(Vxn[lane]); }
        if (!simt::all(vfin) || !fin) return DDP_OK;
    }
    *ok = 1;
    return DDP_OK;
}

}}  // namespace hop::ddp
