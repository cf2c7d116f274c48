// hop_ddp_core.cuh -- per-problem device functions of the HOP-DDP iteration that surround the
// horizon selection (one thread per problem; small dense blocks in thread-local arrays).
//
// Replaces (reference file:line):
//   utils.py:96-120        chol_solve (jitter ladder, NO fallback -> LinAlgError)
//   solver.py:65-105       cost_timeopt_true
//   solver.py:156-230      backward_pass_truncated
//   solver.py:233-286      forward_linesearch_fixedT
// Operation order follows the numpy expressions (left-to-right products), with unfused mul/add so the
// rounding matches the reference's scalar arithmetic as closely as a different BLAS allows.
#pragma once
#include <math.h>

#include "hop_dynamics.cuh"

namespace hop { namespace ddp {

enum : int { DDP_OK = 0, DDP_NONFINITE = 1, DDP_LINALG = 2 };

template <int N_>
HOP_DEVICE bool all_finite(const double* x) {
    bool f = true;
#pragma unroll
    for (int i = 0; i < N_; ++i) f = f && isfinite(x[i]);
    return f;
}

// e = wrap(x - xg) on the indices of wrap_mask (utils.py:131-137)
template <int n>
HOP_DEVICE void wrapped_error(const double* x, const double* xg, unsigned wrap_mask, double* e) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
        double v = sub(x[i], xg[i]);
        if ((wrap_mask >> i) & 1u) v = wrap_pi(v);
        e[i] = v;
    }
}

// (M v)_i as numpy's left-to-right sum.  When M is DIAGONAL (CostConst::diag; every reference case) only the term j = i is not
// an exact zero: the skipped terms are +-0 products of finite v (the callers test finiteness first, or discard the value), the
// running sum starts at +0.0 and +0 + (+-0) = +0 in round-to-nearest, so add(+0.0, mul(M_ii, v_i)) is the same bits as the full
// sum -- n times fewer dependent operations on the chain of every roll-out and backward step.
template <int n>
HOP_DEVICE double row_dot(const double* M, const double* v, int i, bool diag) {
    if (diag) return add(0.0, mul(M[i * n + i], v[i]));
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = add(s, mul(M[i * n + j], v[j]));
    return s;
}
// v^T (M v) as `v @ (M @ v)`
template <int n>
HOP_DEVICE double quad_form(const double* M, const double* v, bool diag) {
    double acc = 0.0;
    if (diag) {
        for (int i = 0; i < n; ++i) acc = add(acc, mul(v[i], add(0.0, mul(M[i * n + i], v[i]))));
        return acc;
    }
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = add(s, mul(M[i * n + j], v[j]));
        acc = add(acc, mul(v[i], s));
    }
    return acc;
}
// all off-diagonal entries of the row-major d x d block are +-0 (CostConst::diag; the kernels compute it per CTA, this serial form
// serves the host emulation)
template <int d>
HOP_DEVICE bool is_diagonal(const double* M) {
    bool dg = true;
    for (int i = 0; i < d * d; ++i)
        if (i / d != i % d) dg = dg && (M[i] == 0.0);
    return dg;
}

// LAPACK dpotf2-style lower Cholesky of M (d x d); returns false on a non-positive pivot.
template <int d>
HOP_DEVICE bool cholesky_lower(const double* M, double* Lo) {
    for (int i = 0; i < d * d; ++i) Lo[i] = 0.0;
    for (int j = 0; j < d; ++j) {
        double ajj = M[j * d + j];
        for (int p = 0; p < j; ++p) ajj = sub(ajj, mul(Lo[j * d + p], Lo[j * d + p]));
        if (!(ajj > 0.0)) return false;
        ajj = sqrt(ajj);
        Lo[j * d + j] = ajj;
        const double rinv = 1.0 / ajj;
        for (int i = j + 1; i < d; ++i) {
            double s = M[i * d + j];
            for (int p = 0; p < j; ++p) s = sub(s, mul(Lo[i * d + p], Lo[j * d + p]));
            Lo[i * d + j] = mul(s, rinv);
        }
    }
    return true;
}

// utils.py:96-120: solve sym(A) X = B (d x c) with the jitter ladder; DDP_LINALG when it is exhausted.
template <int d, int c>
HOP_DEVICE int chol_solve(const double* A, const double* Bm, double* X, double jitter, int max_tries) {
    double S[d * d], M[d * d], Lo[d * d], Y[d * c];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) S[i * d + j] = 0.5 * add(A[i * d + j], A[j * d + i]);
    if (!all_finite<d * d>(S) || !all_finite<d * c>(Bm)) return DDP_NONFINITE;
    double eps = jitter;
    for (int t = 0; t < max_tries; ++t) {
        for (int i = 0; i < d * d; ++i) M[i] = S[i];
        for (int i = 0; i < d; ++i) M[i * d + i] = add(S[i * d + i], eps);
        if (cholesky_lower<d>(M, Lo)) {
            for (int col = 0; col < c; ++col) {
                for (int i = 0; i < d; ++i) {
                    double s = Bm[i * c + col];
                    for (int p = 0; p < i; ++p) s = sub(s, mul(Lo[i * d + p], Y[p * c + col]));
                    Y[i * c + col] = s / Lo[i * d + i];
                }
                for (int i = d - 1; i >= 0; --i) {
                    double s = Y[i * c + col];
                    for (int p = i + 1; p < d; ++p) s = sub(s, mul(Lo[p * d + i], X[p * c + col]));
                    X[i * c + col] = s / Lo[i * d + i];
                }
            }
            if (all_finite<d * c>(X)) return DDP_OK;
        }
        eps *= 10.0;
    }
    return DDP_LINALG;
}

struct CostConst {   // shared case constants, row-major
    const double *xg, *u_ref, *Q, *R, *Qf;
    double w;
    unsigned wrap_mask;
    unsigned diag;   // bit 0: Q, bit 1: R, bit 2: Qf is diagonal (is_diagonal): row_dot / quad_form take the one-term path
    HOP_DEVICE bool q_diag() const { return diag & 1u; }
    HOP_DEVICE bool r_diag() const { return (diag >> 1) & 1u; }
    HOP_DEVICE bool qf_diag() const { return (diag >> 2) & 1u; }
};

// solver.py:65-105.  X [N+1][n], U [N][m] of ONE problem.
template <int n, int m>
HOP_DEVICE double cost_timeopt_true(const double* X, const double* U, const CostConst& c, int T) {
    const double inf = HUGE_VAL;
    if (T <= 0) return inf;
    for (int k = 0; k <= T; ++k)
        if (!all_finite<n>(X + (size_t)k * n)) return inf;
    for (int k = 0; k < T; ++k)
        if (!all_finite<m>(U + (size_t)k * m)) return inf;
    double acc = 0.0, e[n], du[m];
    for (int k = 0; k < T; ++k) {
        wrapped_error<n>(X + (size_t)k * n, c.xg, c.wrap_mask, e);
#pragma unroll
        for (int i = 0; i < m; ++i) du[i] = sub(U[(size_t)k * m + i], c.u_ref[i]);
        if (!all_finite<n>(e) || !all_finite<m>(du)) return inf;
        acc = add(acc, add(add(mul(0.5, quad_form<n>(c.Q, e, c.q_diag())), mul(0.5, quad_form<m>(c.R, du, c.r_diag()))), c.w));
    }
    wrapped_error<n>(X + (size_t)T * n, c.xg, c.wrap_mask, e);
    if (!all_finite<n>(e)) return inf;
    return add(acc, mul(0.5, quad_form<n>(c.Qf, e, c.qf_diag())));
}

// small dense helpers on thread-local row-major arrays
template <int r, int k, int c>
HOP_DEVICE void mm_nn(const double* A, const double* Bm, double* C) {   // C = A B
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s = add(s, mul(A[i * k + l], Bm[l * c + j]));
            C[i * c + j] = s;
        }
}
template <int r, int k, int c>
HOP_DEVICE void mm_tn(const double* A, const double* Bm, double* C) {   // C = A^T B, A is k x r
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s = add(s, mul(A[l * r + i], Bm[l * c + j]));
            C[i * c + j] = s;
        }
}

// solver.py:156-230.  A [N][n][n], Bm [N][n][m], X, U of ONE problem; writes k_out [T][m], K_out [T][m][n].
// *ok = 0 reproduces `return None, None, False`; a non-zero return reproduces an escaping exception.
template <int n, int m>
HOP_DEVICE int backward_pass(const double* A, const double* Bm, const double* X, const double* U, const CostConst& c,
                             int T, double lm, double* k_out, double* K_out, int* ok) {
    *ok = 0;
    if (T <= 0) return DDP_OK;
    double e[n], du[m], Vx[n], Vxx[n * n];
    wrapped_error<n>(X + (size_t)T * n, c.xg, c.wrap_mask, e);
    if (!all_finite<n>(e)) return DDP_OK;
    for (int i = 0; i < n; ++i) {
        Vx[i] = row_dot<n>(c.Qf, e, i, c.qf_diag());
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Vxx[i * n + j] = 0.5 * add(c.Qf[i * n + j], c.Qf[j * n + i]);
    for (int k = T - 1; k >= 0; --k) {
        const double* Ak = A + (size_t)k * n * n;
        const double* Bk = Bm + (size_t)k * n * m;
        wrapped_error<n>(X + (size_t)k * n, c.xg, c.wrap_mask, e);
        for (int i = 0; i < m; ++i) du[i] = sub(U[(size_t)k * m + i], c.u_ref[i]);
        if (!all_finite<n>(e) || !all_finite<m>(du)) return DDP_OK;
        double Qx[n], Qu[m], Qxx[n * n], Quu[m * m], Qux[m * n], AtV[n * n], BtV[m * n], t[n * n];
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            const double lx = row_dot<n>(c.Q, e, i, c.q_diag());
            for (int l = 0; l < n; ++l) s = add(s, mul(Ak[l * n + i], Vx[l]));
            Qx[i] = add(lx, s);
        }
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            const double lu = row_dot<m>(c.R, du, i, c.r_diag());
            for (int l = 0; l < n; ++l) s = add(s, mul(Bk[l * m + i], Vx[l]));
            Qu[i] = add(lu, s);
        }
        mm_tn<n, n, n>(Ak, Vxx, AtV);
        mm_nn<n, n, n>(AtV, Ak, t);
        for (int i = 0; i < n * n; ++i) Qxx[i] = add(c.Q[i], t[i]);
        mm_tn<m, n, n>(Bk, Vxx, BtV);
        mm_nn<m, n, m>(BtV, Bk, t);
        for (int i = 0; i < m * m; ++i) Quu[i] = add(c.R[i], t[i]);
        mm_nn<m, n, n>(BtV, Ak, Qux);
        double Qreg[m * m], Ltmp[m * m];
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) Qreg[i * m + j] = add(0.5 * add(Quu[i * m + j], Quu[j * m + i]), (i == j) ? lm : 0.0);
        if (!cholesky_lower<m>(Qreg, Ltmp)) return DDP_OK;                    // solver.py:213-216
        double kap[m], Kk[m * n];
        int rc = chol_solve<m, 1>(Qreg, Qu, kap, 1e-9, 8);
        if (rc) return rc;
        rc = chol_solve<m, n>(Qreg, Qux, Kk, 1e-9, 8);
        if (rc) return rc;
        for (int i = 0; i < m; ++i) kap[i] = -kap[i];
        for (int i = 0; i < m * n; ++i) Kk[i] = -Kk[i];
        for (int i = 0; i < m; ++i) k_out[(size_t)k * m + i] = kap[i];
        for (int i = 0; i < m * n; ++i) K_out[(size_t)k * m * n + i] = Kk[i];
        double KtQuu[n * m];
        mm_tn<n, m, m>(Kk, Quu, KtQuu);
        double Vxn[n];
        for (int i = 0; i < n; ++i) {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int l = 0; l < m; ++l) s1 = add(s1, mul(Kk[l * n + i], Qu[l]));
            for (int l = 0; l < m; ++l) s2 = add(s2, mul(Qux[l * n + i], kap[l]));
            for (int l = 0; l < m; ++l) s3 = add(s3, mul(KtQuu[i * m + l], kap[l]));
            Vxn[i] = add(add(add(Qx[i], s1), s2), s3);                         // solver.py:224
        }
        double t1[n * n], t2[n * n], t3[n * n];
        mm_tn<n, m, n>(Kk, Qux, t1);
        mm_tn<n, m, n>(Qux, Kk, t2);
        mm_nn<n, m, n>(KtQuu, Kk, t3);
        for (int i = 0; i < n * n; ++i) t[i] = add(add(add(Qxx[i], t1[i]), t2[i]), t3[i]);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) Vxx[i * n + j] = 0.5 * add(t[i * n + j], t[j * n + i]);   // solver.py:225
        for (int i = 0; i < n; ++i) Vx[i] = Vxn[i];
        if (!all_finite<n>(Vx) || !all_finite<n * n>(Vxx)) return DDP_OK;
    }
    *ok = 1;
    return DDP_OK;
}

// ---- warp-cooperative backward pass ------------------------------------------------------------------------
// Same recursion as backward_pass above, ONE WARP per problem: every matrix lives in per-warp shared memory and each
// output ELEMENT is produced by one lane with exactly the operation sequence of the thread-per-problem version
// (ordered 12-term sums, unfused mul/add), so the gains, Vx/Vxx and the ok/err flags are bit-identical -- only the
// elements of a product are spread over the lanes.  The thread-per-problem kernel needs >= 3e5 problems to fill a
// B200 (11 KB of local memory per thread); at 16 384 problems it ran 3.5 warps per SM (13 ms per pass).
template <int n, int m>
struct BwSmem {   // doubles per warp
    static constexpr int VXX = 0, AK = VXX + n * n, BK = AK + n * n, ATV = BK + n * m, TT = ATV + n * n, QXX = TT + n * n,
                         BTV = QXX + n * n, QUX = BTV + m * n, KK = QUX + m * n, KTQ = KK + m * n, QUU = KTQ + n * m,
                         VX = QUU + m * m, QX = VX + n, VXN = QX + n, EV = VXN + n, QU = EV + n, KAP = QU + m, DU = KAP + m,
                         LXU = DU + m,                          // lx [n], lu [m] of the brute-force sweep
                         SIZE = (LXU + n + m + 1) & ~1;
};

// chol_solve (utils.py:96-120) with the right-hand-side columns spread over the lanes: every lane factors the d x d
// matrix redundantly (same bits everywhere), lane j < c substitutes column j; the ladder decision (all X finite) is a vote.
template <int d, int c>
HOP_DEVICE int chol_solve_warp(const double* A, const double* Bm, double* X, double jitter, int max_tries, int lane) {
    double S[d * d], M[d * d], Lo[d * d];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) S[i * d + j] = 0.5 * add(A[i * d + j], A[j * d + i]);
    bool fin = all_finite<d * d>(S);
    for (int i = lane; i < d * c; i += 32) fin = fin && isfinite(Bm[i]);
    if (!simt::all(fin)) return DDP_NONFINITE;
    double eps = jitter;
    for (int t = 0; t < max_tries; ++t) {
        for (int i = 0; i < d * d; ++i) M[i] = S[i];
        for (int i = 0; i < d; ++i) M[i * d + i] = add(S[i * d + i], eps);
        if (cholesky_lower<d>(M, Lo)) {
            bool okc = true;
            double Y[d], Xc[d];
            if (lane < c) {
                const int col = lane;
                for (int i = 0; i < d; ++i) {
                    double s = Bm[i * c + col];
                    for (int p = 0; p < i; ++p) s = sub(s, mul(Lo[i * d + p], Y[p]));
                    Y[i] = s / Lo[i * d + i];
                }
                for (int i = d - 1; i >= 0; --i) {
                    double s = Y[i];
                    for (int p = i + 1; p < d; ++p) s = sub(s, mul(Lo[p * d + i], Xc[p]));
                    Xc[i] = s / Lo[i * d + i];
                }
                okc = all_finite<d>(Xc);
            }
            if (simt::all(okc)) {
                if (lane < c)
                    for (int i = 0; i < d; ++i) X[i * c + lane] = Xc[i];
                return DDP_OK;
            }
        }
        eps *= 10.0;
    }
    return DDP_LINALG;
}

// The gains of one backward step in ONE pass (solver.py:213-221): the positive-definiteness test of Quu_reg and the two
// chol_solve calls (k = -Quu_reg^-1 Qu, K = -Quu_reg^-1 Qux) as written above are three Cholesky factorisations one after the
// other -- the test on Quu_reg, then twice the SAME jittered factor of sym(Quu_reg) + 1e-9 I -- and two substitution passes; fp64
// sqrt and divide are ~250-clock dependent sequences, so they, not the matrix products, bound a backward step.  Here the two
// DIFFERENT factorisations run column by column in lockstep (two independent chains, no early exit), the jittered factor is formed
// once, and the n + 1 right-hand-side columns (Qux on lanes 0 .. n-1, Qu on lane n) are substituted together.  Every
// value is produced by the operations of cholesky_lower / chol_solve_warp in their order, so the results are bit-identical; any
// case the straight-line path does not cover (non-finite input, a failed first try, a non-finite solution: the jitter ladder)
// returns -1 and the caller runs the reference sequence.  Returns 0: kap, Kk written (not yet negated); 1: Quu_reg is not PD.
template <int m, int n>
HOP_DEVICE int gains_one_pass(const double* Qreg, const double* Qu, const double* Qux, double* kap, double* Kk, int lane) {
    static_assert(n + 1 <= 32, "one right-hand-side column per lane");
    double S[m * m], M[m * m], L0[m * m], L1[m * m];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) S[i * m + j] = 0.5 * add(Qreg[i * m + j], Qreg[j * m + i]);
    for (int i = 0; i < m * m; ++i) M[i] = S[i];
    for (int i = 0; i < m; ++i) M[i * m + i] = add(S[i * m + i], 1e-9);
    for (int i = 0; i < m * m; ++i) { L0[i] = 0.0; L1[i] = 0.0; }
    bool ok0 = true, ok1 = true;
    for (int j = 0; j < m; ++j) {                                               // cholesky_lower<m> on Qreg and on M, interleaved
        double a0 = Qreg[j * m + j], a1 = M[j * m + j];
        for (int q = 0; q < j; ++q) {
            a0 = sub(a0, mul(L0[j * m + q], L0[j * m + q]));
            a1 = sub(a1, mul(L1[j * m + q], L1[j * m + q]));
        }
        ok0 = ok0 && (a0 > 0.0);
        ok1 = ok1 && (a1 > 0.0);
        a0 = sqrt(a0); a1 = sqrt(a1);                                           // (NaN after a failed pivot: the flag is already down)
        L0[j * m + j] = a0; L1[j * m + j] = a1;
        const double r0 = 1.0 / a0, r1 = 1.0 / a1;
        for (int i = j + 1; i < m; ++i) {
            double s0 = Qreg[i * m + j], s1 = M[i * m + j];
            for (int q = 0; q < j; ++q) {
                s0 = sub(s0, mul(L0[i * m + q], L0[j * m + q]));
                s1 = sub(s1, mul(L1[i * m + q], L1[j * m + q]));
            }
            L0[i * m + j] = mul(s0, r0);
            L1[i * m + j] = mul(s1, r1);
        }
    }
    if (!ok0) return 1;                                                         // solver.py:213-216
    bool fin = all_finite<m * m>(S);
    for (int i = lane; i < m; i += 32) fin = fin && isfinite(Qu[i]);
    for (int i = lane; i < m * n; i += 32) fin = fin && isfinite(Qux[i]);
    if (!simt::all(fin) || !ok1) return -1;
    bool okc = true;
    double Y[m], Xc[m];
    if (lane <= n) {
        const double* rhs = (lane < n) ? Qux + lane : Qu;                       // column `lane` of Qux (stride n) | Qu (stride 1)
        const int ld = (lane < n) ? n : 1;
        for (int i = 0; i < m; ++i) {
            double sacc = rhs[i * ld];
            for (int q = 0; q < i; ++q) sacc = sub(sacc, mul(L1[i * m + q], Y[q]));
            Y[i] = sacc / L1[i * m + i];
        }
        for (int i = m - 1; i >= 0; --i) {
            double sacc = Y[i];
            for (int q = i + 1; q < m; ++q) sacc = sub(sacc, mul(L1[q * m + i], Xc[q]));
            Xc[i] = sacc / L1[i * m + i];
        }
        okc = all_finite<m>(Xc);
    }
    if (!simt::all(okc)) return -1;
    if (lane < n) {
        for (int i = 0; i < m; ++i) Kk[i * n + lane] = Xc[i];
    } else if (lane == n) {
        for (int i = 0; i < m; ++i) kap[i] = Xc[i];
    }
    return 0;
}

template <int n, int m>
HOP_DEVICE int backward_pass_warp(const double* A, const double* Bm, const double* X, const double* U, const CostConst& c,
                                  int T, double lm, double* k_out, double* K_out, int* ok, double* sm, int lane) {
    using S = BwSmem<n, m>;
    static_assert(n + m <= 32 && m * m <= 32, "lane roles below assume n + m <= 32");
    double *Vxx = sm + S::VXX, *Ak = sm + S::AK, *Bk = sm + S::BK, *AtV = sm + S::ATV, *t = sm + S::TT, *Qxx = sm + S::QXX;
    double *BtV = sm + S::BTV, *Qux = sm + S::QUX, *Kk = sm + S::KK, *KtQuu = sm + S::KTQ, *Quu = sm + S::QUU;
    double *Vx = sm + S::VX, *Qx = sm + S::QX, *Vxn = sm + S::VXN, *e = sm + S::EV, *Qu = sm + S::QU, *kap = sm + S::KAP, *du = sm + S::DU;
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
        Vx[lane] = s;
    }
    for (int q = lane; q < n * n; q += 32) {
        const int i = q / n, j = q % n;
        Vxx[q] = 0.5 * add(c.Qf[i * n + j], c.Qf[j * n + i]);
    }
    // The inputs of step k-1 (lane owns the elements lane, lane + 32, ... of A and B, and one component of X / U) are
    // loaded into registers while step k computes and moved to shared memory at the top of step k-1: the sweep is one
    // dependent chain per warp and these loads were exposed once per step.
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
        // ---- Qx, Qu, A^T Vxx, B^T Vxx
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
        for (int q = lane; q < n * n; q += 32) {                                // mm_tn<n, n, n>(Ak, Vxx, AtV)
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(Ak[l * n + i], Vxx[l * n + j]));
            AtV[q] = s;
        }
        for (int q = lane; q < m * n; q += 32) {                                // mm_tn<m, n, n>(Bk, Vxx, BtV)
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(Bk[l * m + i], Vxx[l * n + j]));
            BtV[q] = s;
        }
        simt::sync();
        // ---- Qxx = Q + (A^T Vxx) A, Quu = R + (B^T Vxx) B, Qux = (B^T Vxx) A
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(AtV[i * n + l], Ak[l * n + j]));
            Qxx[q] = add(c.Q[q], s);
        }
        for (int q = lane; q < m * m; q += 32) {
            const int i = q / m, j = q % m;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(BtV[i * n + l], Bk[l * m + j]));
            Quu[q] = add(c.R[q], s);
        }
        for (int q = lane; q < m * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(BtV[i * n + l], Ak[l * n + j]));
            Qux[q] = s;
        }
        simt::sync();
        // ---- gains (every lane holds Quu_reg; the right-hand sides are spread over the lanes)
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
        for (int q = lane; q < n * m; q += 32) {                                // mm_tn<n, m, m>(Kk, Quu, KtQuu)
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
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            double t1 = 0.0, t2 = 0.0, t3 = 0.0;
            for (int l = 0; l < m; ++l) t1 = add(t1, mul(Kk[l * n + i], Qux[l * n + j]));
            for (int l = 0; l < m; ++l) t2 = add(t2, mul(Qux[l * n + i], Kk[l * n + j]));
            for (int l = 0; l < m; ++l) t3 = add(t3, mul(KtQuu[i * m + l], Kk[l * n + j]));
            t[q] = add(add(add(Qxx[q], t1), t2), t3);
        }
        simt::sync();
        fin = true;
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            const double v = 0.5 * add(t[i * n + j], t[j * n + i]);             // solver.py:225
            Vxx[q] = v;
            fin = fin && isfinite(v);
        }
        if (lane < n) { Vx[lane] = Vxn[lane]; fin = fin && isfinite(Vxn[lane]); }
        if (!simt::all(fin)) return DDP_OK;
    }
    *ok = 1;
    return DDP_OK;
}

// ---- brute-force J(T) comparator ---------------------------------------------------------------------------
// solver.py:293-358 bruteforce_all_Jt_backward_expansion for ONE horizon T: a full Riccati sweep T -> 0 that also
// carries the constant term V0; J(T) = V0[0] (it already contains w T).  One warp per (problem, T), same shared-memory
// layout and element-per-lane products as backward_pass_warp.  It is the reference's baseline-1 *method*, kept here as
// an independent on-device check of the propagator curve (O(T^2 n^3) per problem against the propagator's O(T n^3)).
template <int n, int m>
HOP_DEVICE int bruteforce_one_T_warp(const double* A, const double* Bm, const double* X, const double* U, const CostConst& c,
                                     int T, double lm, double* V0_out, double* sm, int lane) {
    using S = BwSmem<n, m>;
    double *Vxx = sm + S::VXX, *Ak = sm + S::AK, *Bk = sm + S::BK, *AtV = sm + S::ATV, *t = sm + S::TT, *Qxx = sm + S::QXX;
    double *BtV = sm + S::BTV, *Qux = sm + S::QUX, *iQux = sm + S::KK, *Quu = sm + S::QUU;
    double *Vx = sm + S::VX, *Qx = sm + S::QX, *Vxn = sm + S::VXN, *e = sm + S::EV, *Qu = sm + S::QU, *iQu = sm + S::KAP, *du = sm + S::DU;
    double *lx = sm + S::LXU, *lu = sm + S::LXU + n;   // (own slots: n + m doubles do not fit the n m slots of KTQ when m = 1)
    if (lane < n) {
        double v = sub(X[(size_t)T * n + lane], c.xg[lane]);
        if ((c.wrap_mask >> lane) & 1u) v = wrap_pi(v);
        e[lane] = v;
    }
    simt::sync();
    if (lane < n) {
        const double s = row_dot<n>(c.Qf, e, lane, c.qf_diag());
        Vx[lane] = s;                                                           // Vx[T] = Qf e_T
    }
    for (int q = lane; q < n * n; q += 32) {
        const int i = q / n, j = q % n;
        Vxx[q] = 0.5 * add(c.Qf[i * n + j], c.Qf[j * n + i]);                   // Vxx[T] = sym(Qf)
    }
    simt::sync();
    double V0 = 0.0;
    if (lane == 0) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = add(acc, mul(e[i], Vx[i]));           // e_T . (Qf e_T)
        V0 = mul(0.5, acc);
    }
    for (int k = T - 1; k >= 0; --k) {
        simt::sync();
        for (int q = lane; q < n * n; q += 32) Ak[q] = A[(size_t)k * n * n + q];
        for (int q = lane; q < n * m; q += 32) Bk[q] = Bm[(size_t)k * n * m + q];
        if (lane < n) {
            double v = sub(X[(size_t)k * n + lane], c.xg[lane]);
            if ((c.wrap_mask >> lane) & 1u) v = wrap_pi(v);
            e[lane] = v;
        } else if (lane < n + m) {
            du[lane - n] = sub(U[(size_t)k * m + (lane - n)], c.u_ref[lane - n]);
        }
        simt::sync();
        if (lane < n) {
            const int i = lane;
            double s = 0.0;
            const double a = row_dot<n>(c.Q, e, i, c.q_diag());                 // lx = Q e
            for (int l = 0; l < n; ++l) s = add(s, mul(Ak[l * n + i], Vx[l]));
            lx[i] = a;
            Qx[i] = add(a, s);
        } else if (lane < n + m) {
            const int i = lane - n;
            double s = 0.0;
            const double a = row_dot<m>(c.R, du, i, c.r_diag());                // lu = R du
            for (int l = 0; l < n; ++l) s = add(s, mul(Bk[l * m + i], Vx[l]));
            lu[i] = a;
            Qu[i] = add(a, s);
        }
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(Ak[l * n + i], Vxx[l * n + j]));
            AtV[q] = s;
        }
        for (int q = lane; q < m * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(Bk[l * m + i], Vxx[l * n + j]));
            BtV[q] = s;
        }
        simt::sync();
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(AtV[i * n + l], Ak[l * n + j]));
            Qxx[q] = add(c.Q[q], s);
        }
        for (int q = lane; q < m * m; q += 32) {
            const int i = q / m, j = q % m;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(BtV[i * n + l], Bk[l * m + j]));
            Quu[q] = add(c.R[q], s);
        }
        for (int q = lane; q < m * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = add(s, mul(BtV[i * n + l], Ak[l * n + j]));
            Qux[q] = s;
        }
        double l0 = 0.0;
        if (lane == 0) {                                                        // 0.5 e.(Q e) + 0.5 du.(R du) + w
            double a = 0.0, b = 0.0;
            for (int i = 0; i < n; ++i) a = add(a, mul(e[i], lx[i]));
            for (int i = 0; i < m; ++i) b = add(b, mul(du[i], lu[i]));
            l0 = add(add(mul(0.5, a), mul(0.5, b)), c.w);
        }
        simt::sync();
        double Qreg[m * m];
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) Qreg[i * m + j] = add(0.5 * add(Quu[i * m + j], Quu[j * m + i]), (i == j) ? lm : 0.0);
        int rc = chol_solve_warp<m, 1>(Qreg, Qu, iQu, 1e-9, 8, lane);
        if (rc) return rc;
        rc = chol_solve_warp<m, n>(Qreg, Qux, iQux, 1e-9, 8, lane);
        if (rc) return rc;
        simt::sync();
        if (lane < n) {
            const int i = lane;
            double s = 0.0;
            for (int l = 0; l < m; ++l) s = add(s, mul(Qux[l * n + i], iQu[l]));
            Vxn[i] = sub(Qx[i], s);                                             // Vx = Qx - Qux^T invQuuQu
        }
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            double s = 0.0;
            for (int l = 0; l < m; ++l) s = add(s, mul(Qux[l * n + i], iQux[l * n + j]));
            t[q] = sub(Qxx[q], s);
        }
        if (lane == 0) {
            double qq = 0.0;
            for (int l = 0; l < m; ++l) qq = add(qq, mul(Qu[l], iQu[l]));
            V0 = sub(add(l0, V0), mul(0.5, qq));                                // V0 = l0 + V0 - 0.5 Qu.invQuuQu
        }
        simt::sync();
        for (int q = lane; q < n * n; q += 32) {
            const int i = q / n, j = q % n;
            Vxx[q] = 0.5 * add(t[i * n + j], t[j * n + i]);
        }
        if (lane < n) Vx[lane] = Vxn[lane];
    }
    if (lane == 0) *V0_out = V0;
    return DDP_OK;
}

// solver.py:233-286, alphas = (1, .5, .25, .1, .05).  Writes the accepted candidate (or a copy of the
// nominal when nothing improved) into X_new / U_new.
template <int SYS>
HOP_DEVICE void forward_linesearch(const double* prm, int N, const double* X, const double* U, const CostConst& c, int T,
                                   const double* k_list, const double* K_list, double* X_new, double* U_new,
                                   double* J_out, int* accepted) {
    constexpr int n = SysDims<SYS>::n, m = SysDims<SYS>::m;
    const double alphas[5] = {1.0, 0.5, 0.25, 0.1, 0.05};
    const double J_old = cost_timeopt_true<n, m>(X, U, c, T);
    for (int ai = 0; ai < 5; ++ai) {
        const double a = alphas[ai];
        double x[n], xn[n], u[m];
        for (int i = 0; i < n; ++i) { x[i] = X[i]; X_new[i] = x[i]; }
        bool ok = true;
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < m; ++i) u[i] = U[(size_t)k * m + i];
            if (k < T) {
                double dx[n];
                for (int i = 0; i < n; ++i) {
                    double v = sub(x[i], X[(size_t)k * n + i]);
                    if ((c.wrap_mask >> i) & 1u) v = wrap_pi(v);
                    dx[i] = v;
                }
                for (int i = 0; i < m; ++i) {
                    double s = 0.0;
                    for (int j = 0; j < n; ++j) s = add(s, mul(K_list[(size_t)k * m * n + i * n + j], dx[j]));
                    u[i] = add(u[i], add(s, mul(a, k_list[(size_t)k * m + i])));
                }
            }
            for (int i = 0; i < m; ++i) U_new[(size_t)k * m + i] = u[i];
            dynamics<SYS>(prm, x, u, xn);
            for (int i = 0; i < n; ++i) { x[i] = xn[i]; X_new[(size_t)(k + 1) * n + i] = xn[i]; }
            if (!all_finite<n>(xn)) { ok = false; break; }
        }
        if (!ok) continue;
        const double J_new = cost_timeopt_true<n, m>(X_new, U_new, c, T);
        if (J_new < J_old) { *J_out = J_new; *accepted = 1; return; }
    }
    for (size_t i = 0; i < (size_t)(N + 1) * n; ++i) X_new[i] = X[i];
    for (size_t i = 0; i < (size_t)N * m; ++i) U_new[i] = U[i];
    *J_out = J_old;
    *accepted = 0;
}

// ---- the five step sizes of solver.py:233-286 evaluated SIDE BY SIDE ----------------------------------
// One candidate of the line search: roll the closed-loop policy du = K dx + alpha k out to N and evaluate
// cost_timeopt_true of the candidate on the fly (same terms, same accumulation order as the function above applied to
// the stored trajectory).  STORE: also write the candidate into X_new / U_new.  Returns false when the candidate has
// to be skipped (non-finite state somewhere on [1, N], solver.py:273-276); J is +inf where cost_timeopt_true returns inf.
template <int SYS, bool STORE>
HOP_DEVICE bool linesearch_candidate(const double* prm, int N, const double* X, const double* U, const CostConst& c, int T,
                                     const double* k_list, const double* K_list, double a, double* X_new, double* U_new,
                                     double* J_out) {
    constexpr int n = SysDims<SYS>::n, m = SysDims<SYS>::m;
    // Small systems: the policy of step k+1 (U, X, K, k: 10 doubles for n = 4, m = 1) is loaded while step k computes --
    // the roll-out is one dependent chain per thread and these loads were its top stall (ncu: long_scoreboard 3.7 per issue).
    constexpr bool PF = (n * m <= 8);
    double x[n], xn[n], u[m], e[n], du[m];
    double pu[m], pX[n], pK[m * n], pk[m];
    auto fetch = [&](int k) {
        for (int i = 0; i < m; ++i) pu[i] = U[(size_t)k * m + i];
        if (k < T) {
            for (int i = 0; i < n; ++i) pX[i] = X[(size_t)k * n + i];
            for (int i = 0; i < m * n; ++i) pK[i] = K_list[(size_t)k * m * n + i];
            for (int i = 0; i < m; ++i) pk[i] = k_list[(size_t)k * m + i];
        }
    };
    for (int i = 0; i < n; ++i) { x[i] = X[i]; if (STORE) X_new[i] = x[i]; }
    double acc = 0.0, J = HUGE_VAL;
    bool inf = (T <= 0);
    if (PF && N > 0) fetch(0);
    for (int k = 0; k < N; ++k) {
        double cX[n], cK[m * n], ck[m];
        if (PF) {
            for (int i = 0; i < m; ++i) u[i] = pu[i];
            for (int i = 0; i < n; ++i) cX[i] = pX[i];
            for (int i = 0; i < m * n; ++i) cK[i] = pK[i];
            for (int i = 0; i < m; ++i) ck[i] = pk[i];
            if (k + 1 < N) fetch(k + 1);
        } else {
            for (int i = 0; i < m; ++i) u[i] = U[(size_t)k * m + i];
        }
        if (k < T) {
            double dx[n];
            for (int i = 0; i < n; ++i) {
                double v = sub(x[i], PF ? cX[i] : X[(size_t)k * n + i]);
                if ((c.wrap_mask >> i) & 1u) v = wrap_pi(v);
                dx[i] = v;
            }
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
                for (int j = 0; j < n; ++j) s = add(s, mul(PF ? cK[i * n + j] : K_list[(size_t)k * m * n + i * n + j], dx[j]));
                u[i] = add(u[i], add(s, mul(a, PF ? ck[i] : k_list[(size_t)k * m + i])));
            }
            // stage term of the candidate (solver.py:92-101)
            wrapped_error<n>(x, c.xg, c.wrap_mask, e);
#pragma unroll
            for (int i = 0; i < m; ++i) du[i] = sub(u[i], c.u_ref[i]);
            if (!all_finite<n>(x) || !all_finite<m>(u) || !all_finite<n>(e) || !all_finite<m>(du)) inf = true;
            acc = add(acc, add(add(mul(0.5, quad_form<n>(c.Q, e, c.q_diag())), mul(0.5, quad_form<m>(c.R, du, c.r_diag()))), c.w));
        } else if (k == T) {
            wrapped_error<n>(x, c.xg, c.wrap_mask, e);                          // terminal term at X_new[T] (solver.py:102-105)
            if (!all_finite<n>(x) || !all_finite<n>(e)) inf = true;
            J = add(acc, mul(0.5, quad_form<n>(c.Qf, e, c.qf_diag())));
        }
        if (STORE)
            for (int i = 0; i < m; ++i) U_new[(size_t)k * m + i] = u[i];
        dynamics<SYS>(prm, x, u, xn);
        for (int i = 0; i < n; ++i) { x[i] = xn[i]; if (STORE) X_new[(size_t)(k + 1) * n + i] = xn[i]; }
        if (!all_finite<n>(xn)) return false;
    }
    if (T >= N) {
        wrapped_error<n>(x, c.xg, c.wrap_mask, e);
        if (!all_finite<n>(e)) inf = true;
        J = add(acc, mul(0.5, quad_form<n>(c.Qf, e, c.qf_diag())));
    }
    *J_out = inf ? HUGE_VAL : J;
    return true;
}

}}  // namespace hop::ddp
