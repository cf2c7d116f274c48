// hop_select_core.cuh -- HOP horizon selection, one problem per G-lane group of a warp.
//
// Replaces (reference file:line, dmmsjtu-umich/time-opt-ilqr):
//   utils.py:35-37,69-93          _sym, chol_inv (jitter ladder 1e-9 x10 up to 8 tries, LU fallback)
//   horizon_selection.py:57-64    stage  (E_k, F_k, G_k)
//   horizon_selection.py:66-75    prefix (Ebar, Fbar, Gbar) composition
//   horizon_selection.py:77-86    per-horizon query J(T)
//   solver.py:522,590             argmin over [T_min, T_max]
//
// Mapping: lane r (< D) of a group owns ROW r of every d x d block, held in registers; the operand
// that is broadcast to all rows of a product lives in the group's shared-memory slab (row-major,
// row stride DP = D rounded up to even so rows are 16-byte aligned for LDS.128 broadcast loads).
// The three passes of the reference are fused into ONE sweep over k (stage k -> prefix k -> query
// k+1 -> running argmin), so no (Ebar, Fbar, Gbar) history is ever stored.
//
// Inverses: the reference forms (sym(A)+eps I)^-1 through Cholesky; here the same inverse is formed
// by in-place Gauss-Jordan sweeps without pivoting.  The sweep pivots are the squared Cholesky
// pivots, so "Cholesky fails" <=> "a sweep pivot is <= 0": the jitter ladder takes the same
// decisions.  After max_tries failures the LU (partial pivoting) fallback of utils.py:90-93 runs on
// lane 0 of the group (rare path: non-PD X0 on cartpole-like problems).
#pragma once
#include <math.h>

#include "hop_simt.cuh"

namespace hop {

enum : int {
    ST_OK = 0,
    ST_NONFINITE = 1,        // reference: FloatingPointError (utils.py:40-42)
    ST_LINALG = 2,           // reference: LinAlgError (utils.py:93)
    ST_ERRMASK = 0xff,
    ST_FLAG_RETRY = 0x100,   // info: some Cholesky attempt failed and the ladder was climbed
    ST_FLAG_LU = 0x200       // info: the LU fallback branch was taken
};

template <int D, int M, int G>
struct Geo {
    static_assert(D <= G && G <= 32 && (G & (G - 1)) == 0, "group must cover the rows");
    static constexpr int DP = (D + 1) & ~1;
    static constexpr int MAT = D * DP;
    static constexpr int SA = 0;             // A^T          [l][j] = A[j][l]
    static constexpr int SX = SA + MAT;      // scratch
    static constexpr int SY = SX + MAT;      // scratch
    static constexpr int SZ = SY + MAT;      // scratch
    static constexpr int SW = SZ + MAT;      // scratch
    static constexpr int SQ = SW + MAT;      // raw QT_{k+1} staged at the top of the step
    static constexpr int SB = SQ + MAT;      // B^T          [l][j] = B[j][l]   (M x DP)
    static constexpr int SR = SB + M * DP;   // R^-1         (M x MP)
    static constexpr int MP = (M + 1) & ~1;
    static constexpr int ROW = SR + M * MP;  // 2 pivot-row buffers
    static constexpr int Z0 = ROW + 2 * DP;  // z0
    static constexpr int VEC = Z0 + DP;      // 4 scratch vectors (fused builders)
    static constexpr int RAW = VEC + 4 * DP;
    // slab size: == 2 (mod 16) doubles, so consecutive groups start 16 B apart modulo the 128-B bank
    // window and their LDS.128 broadcasts hit disjoint banks.
    static constexpr int SLAB = ((RAW + 13) / 16) * 16 + 2;
};

// ---------------------------------------------------------------------------------------------
// row/column movement between registers and the shared slab
// ---------------------------------------------------------------------------------------------
template <int C, int DP>
HOP_DEVICE void st_row(double* S, int r, const double (&v)[C]) {
#pragma unroll
    for (int j = 0; j < C; ++j) S[r * DP + j] = v[j];
}
template <int C, int DP>
HOP_DEVICE void ld_row(const double* S, int r, double (&v)[C]) {
#pragma unroll
    for (int j = 0; j < C; ++j) v[j] = S[r * DP + j];
}
template <int C, int DP>
HOP_DEVICE void st_col(double* S, int r, const double (&v)[C]) {
#pragma unroll
    for (int j = 0; j < C; ++j) S[j * DP + r] = v[j];
}
template <int C, int DP>
HOP_DEVICE void ld_col(const double* S, int r, double (&v)[C]) {
#pragma unroll
    for (int j = 0; j < C; ++j) v[j] = S[j * DP + r];
}

// out[j] (+)= sum_l v[l] * S[l][j]   (row-vector times the broadcast matrix S, K x C, row stride DP)
template <int K, int C, int DP, bool ACC>
HOP_DEVICE void mm(const double (&v)[K], const double* S, double (&out)[C]) {
    if (!ACC) {
#pragma unroll
        for (int j = 0; j < C; ++j) out[j] = 0.0;
    }
#pragma unroll
    for (int l = 0; l < K; ++l) {
        const double a = v[l];
#pragma unroll
        for (int j = 0; j < C; ++j) out[j] = fma(a, S[l * DP + j], out[j]);
    }
}

// utils.py:35-37 on a matrix held one row per lane: v <- 0.5 (v + v^T), via buffer S.
template <int D, int DP>
HOP_DEVICE void sym_rows(double (&v)[D], double* S, int r, bool act) {
    simt::sync();
    if (act) st_row<D, DP>(S, r, v);
    simt::sync();
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = 0.5 * (v[j] + S[j * DP + r]);
}

template <int G>
HOP_DEVICE unsigned group_mask(int lane) {
    return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
}
template <int G>
HOP_DEVICE bool group_all(bool p, int lane) {
    const unsigned m = group_mask<G>(lane);
    return (simt::ballot(p) & m) == m;
}
template <int G>
HOP_DEVICE double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += simt::shfl_xor(v, o, G);
    return v;
}

// ---------------------------------------------------------------------------------------------
// chol_inv (utils.py:69-93)
// ---------------------------------------------------------------------------------------------
// One Gauss-Jordan inversion attempt on rows a[] (in place).  Returns whether every pivot was > 0,
// i.e. whether np.linalg.cholesky would have succeeded on the same matrix.
template <int D, int G, int DP>
HOP_DEVICE bool gj_attempt(double (&a)[D], int r, double* rowbuf) {
    bool ok = true;
    simt::sync();   // the previous user of rowbuf (an earlier inversion) may still be reading it
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double p = simt::shfl(a[j], j, G);
        ok = ok && (p > 0.0) && (p <= 1.7976931348623157e308);   // +Inf is non-finite input (utils.py:75), not a pivot
        const double rinv = simt::rcp_newton(p);
        double* rb = rowbuf + (j & 1) * DP;
        const bool piv = (r == j);
        if (piv) st_row<D, DP>(rb, 0, a);
        simt::sync();
        // lane j:   a[c] <- a[c] * rinv          (= fma(rinv, rb[c], 0))
        // others:   a[c] <- a[c] - (a[j] rinv) rb[c]
        const double nf = piv ? rinv : -(a[j] * rinv);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            if (c == j) continue;
            const double base = piv ? 0.0 : a[c];
            a[c] = fma(nf, rb[c], base);
        }
        a[j] = nf;   // lane j: 1/p ; others: -a[j]/p
    }
    return ok;
}

// LU with partial pivoting on lane 0 of the group (utils.py:90-93): X = (S)^-1, S in buffer A (D x DP,
// destroyed), result in buffer Xb.  Returns false on an exactly singular pivot.
template <int D, int DP>
HOP_DEVICE_NOINLINE bool lu_inverse_serial(double* A, double* Xb) {
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Xb[i * DP + j] = (i == j) ? 1.0 : 0.0;
    for (int j = 0; j < D; ++j) {
        int piv = j;
        double best = fabs(A[j * DP + j]);
        for (int i = j + 1; i < D; ++i) {
            const double v = fabs(A[i * DP + j]);
            if (v > best) { best = v; piv = i; }
        }
        if (A[piv * DP + j] == 0.0) return false;
        if (piv != j)
            for (int q = 0; q < D; ++q) {
                double t = A[j * DP + q]; A[j * DP + q] = A[piv * DP + q]; A[piv * DP + q] = t;
                t = Xb[j * DP + q]; Xb[j * DP + q] = Xb[piv * DP + q]; Xb[piv * DP + q] = t;
            }
        const double rinv = 1.0 / A[j * DP + j];
        for (int i = j + 1; i < D; ++i) {
            const double f = A[i * DP + j] * rinv;
            for (int q = j + 1; q < D; ++q) A[i * DP + q] -= f * A[j * DP + q];
            for (int q = 0; q < D; ++q) Xb[i * DP + q] -= f * Xb[j * DP + q];
        }
    }
    for (int c = 0; c < D; ++c)
        for (int i = D - 1; i >= 0; --i) {
            double s = Xb[i * DP + c];
            for (int p = i + 1; p < D; ++p) s -= A[i * DP + p] * Xb[p * DP + c];
            Xb[i * DP + c] = s / A[i * DP + i];
        }
    return true;
}

// chol_inv on a SYMMETRISED matrix given one row per lane in s[]; result rows in out[].
// f1/f2: two slab buffers that are free at the call site (used by the rare LU fallback only).
template <int D, int G, int DP>
HOP_DEVICE void chol_inv_rows(const double (&s)[D], double (&out)[D], int r, bool act, int lane, double* rowbuf,
                              double* f1, double* f2, double jitter, int max_tries, int& status) {
    double eps = jitter;
    int tries = 0;
    bool done = false;
    for (;;) {
        double a[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = s[j] + ((j == r) ? eps : 0.0);
        const bool ok = gj_attempt<D, G, DP>(a, r, rowbuf);
        // utils.py:75 -- a non-finite input raises before any attempt.  Checked lazily (only when the
        // first attempt failed, which a NaN always forces); the test is kept warp-uniform.
        const bool fail_first = (!done) && (!ok) && (tries == 0);
        if (!simt::all(!fail_first)) {
            bool fin = true;
#pragma unroll
            for (int j = 0; j < D; ++j) fin = fin && isfinite(s[j]);
            const unsigned bal = simt::ballot(fin || !act);
            const unsigned gm = group_mask<G>(lane);
            if (fail_first && (bal & gm) != gm) {
                status |= ST_NONFINITE;
#pragma unroll
                for (int j = 0; j < D; ++j) out[j] = nan("");
                done = true;
            }
        }
        if (!done) {
            if (ok) {
#pragma unroll
                for (int j = 0; j < D; ++j) out[j] = a[j];
                done = true;
            } else {
                status |= ST_FLAG_RETRY;
                eps *= 10.0;
                ++tries;
            }
        }
        const bool need_lu = (!done) && (tries >= max_tries);
        if (!simt::all(!need_lu)) {  // warp-uniform: somebody in this warp needs the LU branch
            simt::sync();
            if (need_lu && act) {
                double a2[D];
#pragma unroll
                for (int j = 0; j < D; ++j) a2[j] = s[j] + ((j == r) ? eps : 0.0);
                st_row<D, DP>(f1, r, a2);
            }
            simt::sync();
            bool lu_ok = true;
            if (need_lu && r == 0) lu_ok = lu_inverse_serial<D, DP>(f1, f2);
            simt::sync();
            lu_ok = simt::shfl(lu_ok ? 1.0 : 0.0, 0, G) != 0.0;   // executed by every lane of the warp
            if (need_lu) {
                ld_row<D, DP>(f2, act ? r : 0, out);
                status |= ST_FLAG_LU;
                if (!lu_ok) status |= ST_LINALG;
                done = true;
            }
            simt::sync();
        }
        if (simt::all(done)) break;
    }
}

// ---------------------------------------------------------------------------------------------
// per-problem sweep state
// ---------------------------------------------------------------------------------------------
template <int D>
struct Prefix {
    double eb[D], fb[D], gb[D];   // row r of Ebar, Fbar, Gbar
};

struct ArgMin {
    double best;
    int idx;       // 1-based horizon; 0 = nothing seen yet
    bool nan_hit;  // np.argmin semantics: the first NaN wins
    HOP_DEVICE void init() { best = 0.0; idx = 0; nan_hit = false; }
    HOP_DEVICE void push(double v, int t) {
        if (nan_hit) return;
        if (v != v) { nan_hit = true; best = v; idx = t; return; }
        if (idx == 0 || v < best) { best = v; idx = t; }
    }
};

// Stage + prefix part of step k.  On entry the slab holds SA = A_k^T, SB = B_k^T, SR = R^-1 and the
// caller provides the already symmetrised row q of Q_k.  Updates P to (Ebar, Fbar, Gbar)_k.
template <int D, int M, int G>
HOP_DEVICE void stage_prefix_step(int k, Prefix<D>& P, const double (&q)[D], double* sm, int r, bool act, int lane,
                                  double jitter, int max_tries, int& status) {
    using Ge = Geo<D, M, G>;
    constexpr int DP = Ge::DP;
    double* SA = sm + Ge::SA; double* SX = sm + Ge::SX; double* SY = sm + Ge::SY;
    double* SZ = sm + Ge::SZ; double* SW = sm + Ge::SW; double* SB = sm + Ge::SB;
    double* SR = sm + Ge::SR; double* ROWB = sm + Ge::ROW;
    const int rr = act ? r : 0;   // safe row for loads on idle lanes

    // ---- stage (horizon_selection.py:57-64): E_k = chol_inv(Q_k)
    double e[D];
    chol_inv_rows<D, G, DP>(q, e, r, act, lane, ROWB, SZ, SW, jitter, max_tries, status);
    simt::sync();
    if (act) st_row<D, DP>(SX, r, e);     // SX = E_k (rows)
    simt::sync();

    double w[D];
    if (k > 0) {
        // ---- W = chol_inv(E_k + Gbar)  (:72)
        double s[D];
#pragma unroll
        for (int j = 0; j < D; ++j) s[j] = e[j] + P.gb[j];
        sym_rows<D, DP>(s, SY, rr, act);
        chol_inv_rows<D, G, DP>(s, w, r, act, lane, ROWB, SZ, SW, jitter, max_tries, status);
        simt::sync();
        if (act) st_row<D, DP>(SY, r, w); // SY = W (rows)
    }

    // ---- F_k = E_k A_k^T ; G_k = sym((A_k E_k) A_k^T + (B_k R^-1) B_k^T)
    double f[D], g[D];
    mm<D, D, DP, false>(e, SA, f);
    {
        double arow[D], t[D];
        ld_col<D, DP>(SA, rr, arow);                 // row r of A_k
        mm<D, D, DP, false>(arow, SX, t);            // (A E) row
        mm<D, D, DP, false>(t, SA, g);               // (A E) A^T
        double brow[M], br[M];
        ld_col<M, DP>(SB, rr, brow);                 // row r of B_k
        mm<M, M, Ge::MP, false>(brow, SR, br);       // (B R^-1) row
        mm<M, D, DP, true>(br, SB, g);               // + (B R^-1) B^T
    }
    sym_rows<D, DP>(g, SW, rr, act);                 // G_k rows (SW scratch)

    if (k == 0) {
#pragma unroll
        for (int j = 0; j < D; ++j) { P.eb[j] = e[j]; P.fb[j] = f[j]; P.gb[j] = g[j]; }
    } else {
        // ---- prefix composition (:70-75), all right-hand sides use the OLD (Ebar, Fbar, Gbar)
        simt::sync();
        if (act) {
            st_row<D, DP>(SZ, r, f);                 // SZ = F_k (rows)
            st_col<D, DP>(SX, r, P.fb);              // SX = Fbar^T
        }
        simt::sync();
        double t1[D], acc[D];
        mm<D, D, DP, false>(P.fb, SY, t1);           // Fbar W
        mm<D, D, DP, false>(t1, SX, acc);            // (Fbar W) Fbar^T
#pragma unroll
        for (int j = 0; j < D; ++j) P.eb[j] = P.eb[j] - acc[j];
        mm<D, D, DP, false>(t1, SZ, P.fb);           // Fbar <- (Fbar W) F_k
        double fc[D];
        ld_col<D, DP>(SZ, rr, fc);                   // row r of F_k^T
        mm<D, D, DP, false>(fc, SY, t1);             // F_k^T W
        mm<D, D, DP, false>(t1, SZ, acc);            // (F_k^T W) F_k
#pragma unroll
        for (int j = 0; j < D; ++j) P.gb[j] = g[j] - acc[j];
        sym_rows<D, DP>(P.eb, SX, rr, act);
        sym_rows<D, DP>(P.gb, SY, rr, act);
    }

}

// Query for horizon t = k+1 (:77-86) given the symmetrised row qt of QT_t and Z0 = z0 in the slab.
// Returns J(t) (valid on every lane of the group).
template <int D, int M, int G>
HOP_DEVICE double query_step(const Prefix<D>& P, const double (&qt)[D], double* sm, int r, bool act, int lane,
                             double jitter, int max_tries, int& status) {
    using Ge = Geo<D, M, G>;
    constexpr int DP = Ge::DP;
    double* SX = sm + Ge::SX; double* SY = sm + Ge::SY; double* SZ = sm + Ge::SZ; double* SW = sm + Ge::SW;
    double* ROWB = sm + Ge::ROW;
    const int rr = act ? r : 0;
    double xt[D], w[D];
    simt::sync();
    chol_inv_rows<D, G, DP>(qt, xt, r, act, lane, ROWB, SZ, SW, jitter, max_tries, status);
    {
        double s[D];
#pragma unroll
        for (int j = 0; j < D; ++j) s[j] = xt[j] + P.gb[j];
        sym_rows<D, DP>(s, SY, rr, act);
        chol_inv_rows<D, G, DP>(s, w, r, act, lane, ROWB, SZ, SW, jitter, max_tries, status);
    }
    simt::sync();
    if (act) {
        st_row<D, DP>(SY, r, w);                     // SY = W_t
        st_col<D, DP>(SX, r, P.fb);                  // SX = Fbar^T
    }
    simt::sync();
    double x0[D];
    {
        double t3[D], acc[D];
        mm<D, D, DP, false>(P.fb, SY, t3);           // Fbar W_t
        mm<D, D, DP, false>(t3, SX, acc);            // (Fbar W_t) Fbar^T
#pragma unroll
        for (int j = 0; j < D; ++j) x0[j] = P.eb[j] - acc[j];
    }
    sym_rows<D, DP>(x0, SZ, rr, act);
    double p0[D];
    simt::sync();
    chol_inv_rows<D, G, DP>(x0, p0, r, act, lane, ROWB, SZ, SW, jitter, max_tries, status);
    // J = 0.5 z0^T P0 z0
    const double* z0 = sm + Ge::Z0;
    double dot = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) dot = fma(p0[j], z0[j], dot);
    const double part = act ? simt::mul_rn(z0[r], dot) : 0.0;   // (never fused with the first add of the tree)
    return 0.5 * group_sum<G>(part);
}

}  // namespace hop
