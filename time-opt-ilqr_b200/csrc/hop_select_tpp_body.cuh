// hop_select_tpp_body.cuh -- LQR-boundary horizon selection for small blocks (d <= 5), ONE PROBLEM PER THREAD.
//
// Replaces (reference file:line, dmmsjtu-umich/time-opt-ilqr), same as hop_select_core.cuh:
//   utils.py:35-37,69-93          _sym, chol_inv (jitter ladder 1e-9 x10 up to 8 tries, LU fallback)
//   horizon_selection.py:57-86    stage (E_k, F_k, G_k), prefix composition, per-horizon query J(T)
//   solver.py:522,590             argmin over [T_min, T_max]
//
// Why a second mapping for the same function: the group-of-lanes mapping of hop_select_core.cuh (row per
// lane, operands broadcast through shared memory) moves every operand of every small product through
// the LSU -- ncu on k_select_generic<4,2,4>: L1 wavefronts 78 % of peak, FP64 pipe 22 %, 214 warp
// instructions per problem-step, 0.15 of the HBM roof that bounds d <= 5.  With the whole problem in
// one thread every d x d block is a set of registers, a product is d^3 back-to-back DFMAs with no
// data movement at all, and the only shared-memory traffic left is the input feed.  It needs a batch
// that fills the machine with 32 problems per warp, so the launcher picks it for large batches only;
// small batches (the HOP-DDP configurations) keep the group-of-lanes kernel, whose latency is lower.
//
// Every element is produced by the SAME sequence of IEEE operations as in hop_select_core.cuh (same
// fma chains in ascending index order, same divisions, same jitter ladder decisions, same summation
// tree for z0^T P0 z0), so the two kernels are bit-identical; tests assert that.
//
// Input feed (device): the blocks of step k+1 are copied global -> shared by warp-cooperative, fully
// coalesced cp.async (16-byte granules when every block is a multiple of 16 bytes, else 8-byte) into a
// double buffer while step k computes; each thread then reads its own problem's blocks from shared
// memory with a conflict-free row stride.  Per-thread strided global loads would cost one L1 tag lookup
// per thread and instruction (32 lines per warp instruction) and bound the kernel at the L1 instead.
#pragma once
#include "hop_select_body.cuh"

namespace hop { namespace tpp {

// ---- products (fma chains in ascending l, started from +0.0: identical to hop::mm) ---------------------
// C = X * Y           C[r][j] = sum_l X[r][l] Y[l][j]
template <int R, int K, int C>
HOP_DEVICE void mul_nn(double (&c)[R][C], const double (&x)[R][K], const double (&y)[K][C]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < K; ++l) s = fma(x[r][l], y[l][j], s);
            c[r][j] = s;
        }
}
// C (+)= X * Y^T      C[r][j] = sum_l X[r][l] Y[j][l]
template <int R, int K, int C, bool ACC>
HOP_DEVICE void mul_nt(double (&c)[R][C], const double (&x)[R][K], const double (&y)[C][K]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            double s = ACC ? c[r][j] : 0.0;
#pragma unroll
            for (int l = 0; l < K; ++l) s = fma(x[r][l], y[j][l], s);
            c[r][j] = s;
        }
}
// C = X^T * Y         C[r][j] = sum_l X[l][r] Y[l][j]
template <int D>
HOP_DEVICE void mul_tn(double (&c)[D][D], const double (&x)[D][D], const double (&y)[D][D]) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < D; ++l) s = fma(x[l][r], y[l][j], s);
            c[r][j] = s;
        }
}
// utils.py:35-37
template <int D>
HOP_DEVICE void sym(double (&a)[D][D]) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int j = r; j < D; ++j) {   // (the diagonal too: 0.5 * (x + x) overflows where the reference's does)
            const double v = 0.5 * (a[r][j] + a[j][r]);
            a[r][j] = v;
            a[j][r] = v;
        }
}

// One Gauss-Jordan inversion attempt, in place (hop::gj_attempt element for element).
template <int D>
HOP_DEVICE bool gj_attempt(double (&a)[D][D]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double p = a[j][j];
        ok = ok && (p > 0.0) && (p <= 1.7976931348623157e308);   // +Inf is non-finite input (utils.py:75), not a pivot
        const double rinv = simt::rcp_newton(p);
        double rb[D];
#pragma unroll
        for (int c = 0; c < D; ++c) rb[c] = a[j][c];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            if (r == j) {
#pragma unroll
                for (int c = 0; c < D; ++c)
                    if (c != j) a[r][c] = fma(rinv, rb[c], 0.0);
                a[r][j] = rinv;
            } else {
                const double nf = -(a[r][j] * rinv);
#pragma unroll
                for (int c = 0; c < D; ++c)
                    if (c != j) a[r][c] = fma(nf, rb[c], a[r][c]);
                a[r][j] = nf;
            }
        }
    }
    return ok;
}

// Rare tail of chol_inv (utils.py:81-93): the ladder after a failed first attempt and the LU fallback.
// Out of line and through local memory on purpose: it keeps the hot loop free of its registers.
template <int D>
HOP_DEVICE_NOINLINE void chol_inv_cold(const double* s, double* out, double jitter, int max_tries, int* status_io) {
    constexpr int DP = (D + 1) & ~1;
    int status = *status_io;
    bool fin = true;
    for (int i = 0; i < D * D; ++i) fin = fin && isfinite(s[i]);
    if (!fin) {                                      // utils.py:75
        for (int i = 0; i < D * D; ++i) out[i] = nan("");
        *status_io = status | ST_NONFINITE;
        return;
    }
    double eps = jitter;
    for (int tries = 1;; ++tries) {                  // attempt 0 has failed already
        status |= ST_FLAG_RETRY;
        eps *= 10.0;
        if (tries >= max_tries) break;
        double a[D][D];
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) a[r][c] = s[r * D + c] + ((r == c) ? eps : 0.0);
        if (gj_attempt<D>(a)) {
            for (int r = 0; r < D; ++r)
                for (int c = 0; c < D; ++c) out[r * D + c] = a[r][c];
            *status_io = status;
            return;
        }
    }
    double A[D * DP], X[D * DP];
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) A[r * DP + c] = s[r * D + c] + ((r == c) ? eps : 0.0);
    const bool lu_ok = lu_inverse_serial<D, DP>(A, X);
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) out[r * D + c] = X[r * DP + c];
    status |= ST_FLAG_LU;
    if (!lu_ok) status |= ST_LINALG;
    *status_io = status;
}

// chol_inv (utils.py:69-93) of an already symmetrised matrix; s is preserved.
template <int D>
HOP_DEVICE void chol_inv(const double (&s)[D][D], double (&out)[D][D], double jitter, int max_tries, int& status) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) out[r][c] = (r == c) ? s[r][c] + jitter : s[r][c];   // (x + 0.0 only turns -0.0 into +0.0)
    if (gj_attempt<D>(out)) return;
    double sb[D * D], ob[D * D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) sb[r * D + c] = s[r][c];
    int st = status;
    chol_inv_cold<D>(sb, ob, jitter, max_tries, &st);
    status = st;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) out[r][c] = ob[r * D + c];
}

// 0.5 * z0^T P0 z0 with the summation tree of hop::group_sum<G> (G = 4 for d <= 4, 8 for d = 5).
template <int D>
HOP_DEVICE double cost_of(const double (&p0)[D][D], const double (&z0)[D]) {
    constexpr int G = (D <= 4) ? 4 : 8;
    double part[G];
#pragma unroll
    for (int r = 0; r < G; ++r) {
        if (r < D) {
            double dot = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) dot = fma(p0[r][j], z0[j], dot);
            part[r] = simt::mul_rn(z0[r], dot);
        } else {
            part[r] = 0.0;
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < o; ++r) part[r] = simt::add_rn(part[r], part[r + o]);
    return 0.5 * part[0];
}

// Feed concept: acquire(k) makes the blocks of step k readable (and may start fetching step k+1);
// a(i), bm(i), q(i), qt(i) return element i (row-major) of A_aug[k], B_aug[k], Q_aug[k], QT[k].
template <int D, int M>
struct GlobalFeed {   // straight from global memory (host emulation; also the device fallback for odd layouts)
    const double *A, *Bm, *Q, *QT;
    const double *ak, *bk, *qk, *tk;
    HOP_DEVICE void init(const SelectArgs& p, int b) {
        const size_t base = (size_t)b * p.N;
        A = p.A_aug + base * D * D; Q = p.Q_aug + base * D * D; QT = p.QT + base * D * D; Bm = p.B_aug + base * D * M;
    }
    HOP_DEVICE void acquire(int k) { ak = A + (size_t)k * D * D; qk = Q + (size_t)k * D * D; tk = QT + (size_t)k * D * D; bk = Bm + (size_t)k * D * M; }
    HOP_DEVICE double a(int i) const { return ak[i]; }
    HOP_DEVICE double bm(int i) const { return bk[i]; }
    HOP_DEVICE double q(int i) const { return qk[i]; }
    HOP_DEVICE double qt(int i) const { return tk[i]; }
};

#ifndef HOP_HOST_EMUL
// Shared-memory stage of one warp: [2 buffers][32 problems][S doubles]; S is the per-problem row stride,
// chosen so that the per-thread reads are conflict-free (16-byte granules: S = 2 mod 4; 8-byte: S odd).
template <int D, int M>
struct StageGeo {
    static constexpr int E = 3 * D * D + D * M;
    static constexpr bool V16 = (D * D) % 2 == 0 && (D * M) % 2 == 0;
    static constexpr int S = V16 ? E + ((2 - E % 4 + 4) % 4) : (E | 1);
    static constexpr int oA = 0, oQ = D * D, oT = 2 * D * D, oB = 3 * D * D;
    static constexpr int WARP_DOUBLES = 2 * 32 * S;
};
template <int D, int M>
struct SmemFeed {
    using SG = StageGeo<D, M>;
    const SelectArgs* p;
    double* stage;
    const double* mine;
    int b0, lane;
    HOP_DEVICE void init(const SelectArgs& args, int first_problem, int lane_, double* warp_stage) {
        p = &args; b0 = first_problem; lane = lane_; stage = warp_stage; mine = warp_stage;
        issue(0);
    }
    // C doubles per problem and step, contiguous in global memory; consecutive lanes take consecutive granules
    template <int C>
    HOP_DEVICE void copy_array(double* dst, const double* src, int k) const {
        constexpr int GR = SG::V16 ? 2 : 1;          // doubles per granule
        constexpr int CH = C / GR;                   // granules per problem
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int idx = i * 32 + lane;
            const int prob = idx / CH, ch = idx % CH;
            int b = b0 + prob;
            b = b < p->B ? b : p->B - 1;
            const double* g = src + ((size_t)b * p->N + k) * C + GR * ch;
            double* d = dst + prob * SG::S + GR * ch;
            if (SG::V16) simt::cp_async16(d, g); else simt::cp_async8(d, g);
        }
    }
    HOP_DEVICE void issue(int k) const {
        double* buf = stage + (k & 1) * 32 * SG::S;
        copy_array<D * D>(buf + SG::oA, p->A_aug, k);
        copy_array<D * D>(buf + SG::oQ, p->Q_aug, k);
        copy_array<D * D>(buf + SG::oT, p->QT, k);
        copy_array<D * M>(buf + SG::oB, p->B_aug, k);
        simt::cp_async_commit();
    }
    HOP_DEVICE void acquire(int k) {
        simt::cp_async_wait<0>();
        simt::sync();                                // step k has landed for every lane; everyone is done with step k-1
        if (k + 1 < p->T_max) issue(k + 1);
        mine = stage + ((k & 1) * 32 + lane) * SG::S;
    }
    HOP_DEVICE double a(int i) const { return mine[SG::oA + i]; }
    HOP_DEVICE double bm(int i) const { return mine[SG::oB + i]; }
    HOP_DEVICE double q(int i) const { return mine[SG::oQ + i]; }
    HOP_DEVICE double qt(int i) const { return mine[SG::oT + i]; }
};
#endif

template <int D, int M, class Feed>
HOP_DEVICE void select_generic_tpp_body(const SelectArgs& p, int b, bool valid, Feed& feed) {
    int status = 0;
    double z0[D], rinv[M][M];
#pragma unroll
    for (int j = 0; j < D; ++j) z0[j] = p.z0[(size_t)b * D + j];
#pragma unroll
    for (int i = 0; i < M * M; ++i) rinv[i / M][i % M] = p.R_inv[(size_t)b * M * M + i];
    const double wexp = p.w_explicit ? p.w_explicit[b] : 0.0;
    double eb[D][D], fb[D][D], gb[D][D];
    ArgMin am;
    am.init();

    for (int k = 0; k < p.T_max; ++k) {
        feed.acquire(k);
        // ---- stage (horizon_selection.py:57-64): E_k = chol_inv(Q_k), F_k = E_k A_k^T, G_k = sym((A E) A^T + (B R^-1) B^T)
        double e[D][D];
        {
            double q[D][D];
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int j = 0; j < D; ++j) q[r][j] = 0.5 * (feed.q(r * D + j) + feed.q(j * D + r));
            chol_inv<D>(q, e, p.jitter, p.max_tries, status);
        }
        double f[D][D], g[D][D];
        {
            double a[D][D], bm[D][M];
#pragma unroll
            for (int i = 0; i < D * D; ++i) a[i / D][i % D] = feed.a(i);
#pragma unroll
            for (int i = 0; i < D * M; ++i) bm[i / M][i % M] = feed.bm(i);
            mul_nt<D, D, D, false>(f, e, a);
            double t[D][D], br[D][M];
            mul_nn<D, D, D>(t, a, e);
            mul_nt<D, D, D, false>(g, t, a);
            mul_nn<D, M, M>(br, bm, rinv);
            mul_nt<D, M, D, true>(g, br, bm);
            sym<D>(g);
        }
        if (k == 0) {
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int j = 0; j < D; ++j) { eb[r][j] = e[r][j]; fb[r][j] = f[r][j]; gb[r][j] = g[r][j]; }
        } else {
            // ---- prefix composition (:66-75); every right-hand side uses the OLD (Ebar, Fbar, Gbar)
            double w[D][D];
            {
                double s[D][D];
#pragma unroll
                for (int r = 0; r < D; ++r)
#pragma unroll
                    for (int j = 0; j < D; ++j) s[r][j] = e[r][j] + gb[r][j];
                sym<D>(s);
                chol_inv<D>(s, w, p.jitter, p.max_tries, status);
            }
            double t1[D][D], acc[D][D];
            mul_nn<D, D, D>(t1, fb, w);                    // Fbar W
            mul_nt<D, D, D, false>(acc, t1, fb);           // (Fbar W) Fbar^T
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int j = 0; j < D; ++j) eb[r][j] = eb[r][j] - acc[r][j];
            mul_nn<D, D, D>(fb, t1, f);                    // Fbar <- (Fbar W) F_k
            mul_tn<D>(t1, f, w);                           // F_k^T W
            mul_nn<D, D, D>(acc, t1, f);                   // (F_k^T W) F_k
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int j = 0; j < D; ++j) gb[r][j] = g[r][j] - acc[r][j];
            sym<D>(eb);
            sym<D>(gb);
        }
        // ---- query of horizon t = k+1 (:77-86)
        double J;
        {
            double wt[D][D];
            {
                double qt[D][D], xt[D][D];
#pragma unroll
                for (int r = 0; r < D; ++r)
#pragma unroll
                    for (int j = 0; j < D; ++j) qt[r][j] = 0.5 * (feed.qt(r * D + j) + feed.qt(j * D + r));
                chol_inv<D>(qt, xt, p.jitter, p.max_tries, status);
#pragma unroll
                for (int r = 0; r < D; ++r)
#pragma unroll
                    for (int j = 0; j < D; ++j) xt[r][j] = xt[r][j] + gb[r][j];
                sym<D>(xt);
                chol_inv<D>(xt, wt, p.jitter, p.max_tries, status);
            }
            double t3[D][D], x0[D][D];
            mul_nn<D, D, D>(t3, fb, wt);                   // Fbar W_t
            mul_nt<D, D, D, false>(x0, t3, fb);            // (Fbar W_t) Fbar^T
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int j = 0; j < D; ++j) x0[r][j] = eb[r][j] - x0[r][j];
            sym<D>(x0);
            double p0[D][D];
            chol_inv<D>(x0, p0, p.jitter, p.max_tries, status);
            J = cost_of<D>(p0, z0);
        }
        if (valid) {
            p.J_out[(size_t)b * p.T_max + k] = J;
            const int t = k + 1;
            if (t >= p.T_min) am.push(simt::add_rn(J, simt::mul_rn(wexp, (double)t)), t);
        }
    }
    if (valid) {
        p.T_out[b] = am.idx;
        p.Jstar_out[b] = am.best;
        p.status[b] = status;
    }
}

}}  // namespace hop::tpp
