// hop_mma.cuh -- d x d (d <= 16) fp64 blocks held as DMMA register fragments: one problem per warp.
//
// Why: on B200 the broadcast-through-shared-memory mapping of hop_select_core.cuh is bound by the
// LSU issue rate (measured 0.43 LDS.128/clk/SM, profiles/r1_probes_lds_dmma_shfl.txt), i.e. ~40 % of
// the FP64 pipe at best.  DMMA.8x8x4 delivers the same 64 FMA/clk/SM with both operands coming from
// registers, so the 13x13x13 products of the LFT composition (10 per horizon step) run with no
// shared-memory traffic and ~8x fewer instructions.
//
// Layout "L" of a 16x16 (zero-padded) matrix over the 32 lanes, g = lane>>2, t = lane&3:
//     v[I][J][s]  holds  M[ 8I + rho(g) ][ 8J + t + 4s ],   rho(g) = (g>>1) + 4(g&1)
// With this choice the accumulator fragment of mma.m8n8k4.f64 *is* the A-operand fragment of the
// same matrix and the B-operand fragment of its transpose, so
//     D = X * Z^T      (X, Z, D all in layout L)
// is 16 DMMAs and zero data movement.  Every product of the sweep is of that form (the symmetric
// factors E, W, W_t serve as their own transposes; F_k enters only through F_k^T = A_k E_k).
// Transposes (needed by _sym only) and the Gauss-Jordan pivot row/column exchange use warp shuffles.
#pragma once
#include <math.h>

#include "hop_simt.cuh"

namespace hop { namespace mma {

struct Mat { double v[2][2][2]; };   // [I][J][s]

HOP_DEVICE int rho(int g) { return (g >> 1) + 4 * (g & 1); }
// inverse of rho on 0..7: x = a + 4b  ->  g = 2a + b
HOP_DEVICE constexpr int rho_inv(int x) { return 2 * (x & 3) + (x >> 2); }

struct LaneGeo {
    int lane, g, t, r0;   // r0 = rho(g): row inside a tile
    HOP_DEVICE void init() { lane = simt::lane_id(); g = lane >> 2; t = lane & 3; r0 = rho(g); }
    HOP_DEVICE int row(int I) const { return 8 * I + r0; }
    HOP_DEVICE int col(int J, int s) const { return 8 * J + t + 4 * s; }
};

#define HOP_FOR_ELEMS(I, J, s) \
    _Pragma("unroll") for (int I = 0; I < 2; ++I) \
    _Pragma("unroll") for (int J = 0; J < 2; ++J) \
    _Pragma("unroll") for (int s = 0; s < 2; ++s)

HOP_DEVICE void mat_zero(Mat& m) { HOP_FOR_ELEMS(I, J, s) m.v[I][J][s] = 0.0; }
HOP_DEVICE void mat_copy(Mat& d, const Mat& a) { HOP_FOR_ELEMS(I, J, s) d.v[I][J][s] = a.v[I][J][s]; }
HOP_DEVICE void mat_add(Mat& d, const Mat& a, const Mat& b) { HOP_FOR_ELEMS(I, J, s) d.v[I][J][s] = a.v[I][J][s] + b.v[I][J][s]; }
HOP_DEVICE void mat_sub(Mat& d, const Mat& a, const Mat& b) { HOP_FOR_ELEMS(I, J, s) d.v[I][J][s] = a.v[I][J][s] - b.v[I][J][s]; }

// D (+)= X * Z^T.  NI/NJ: row tiles of X / Z that hold data, KB: number of 4-wide k blocks.
template <int NI, int NJ, int KB, bool ACC>
HOP_DEVICE void mma_nt(Mat& D, const Mat& X, const Mat& Z) {
    if (!ACC) { HOP_FOR_ELEMS(I, J, s) D.v[I][J][s] = 0.0; }
    // k-block outermost: consecutive DMMAs go to different accumulator tiles, so a dependent pair is NI*NJ
    // instructions apart (tile-major order serialises on the DMMA latency: ncu r1d, `wait` 83 % in the products)
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int I = 0; I < NI; ++I)
#pragma unroll
            for (int J = 0; J < NJ; ++J)
                simt::dmma(D.v[I][J][0], D.v[I][J][1], X.v[I][kb >> 1][kb & 1], Z.v[J][kb >> 1][kb & 1]);
}

// T = M^T (layout L -> layout L): 16 double shuffles.
HOP_DEVICE void mat_transpose(Mat& T, const Mat& M, const LaneGeo& L) {
    const bool odd = (L.g & 1) != 0;
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                // T[8I+rho(g)][8J+t+4s] = M[8J+t+4s][8I+rho(g)] : lane (g'=2t+s, t'=g>>1), tile (J,I), slot g&1
                const int src = ((2 * L.t + s) << 2) | (L.g >> 1);
                const double v0 = simt::shfl(M.v[J][I][0], src, 32);
                const double v1 = simt::shfl(M.v[J][I][1], src, 32);
                T.v[I][J][s] = odd ? v1 : v0;
            }
}

// utils.py:35-37  M <- 0.5 (M + M^T)
HOP_DEVICE void mat_sym(Mat& M, const LaneGeo& L) {
    Mat T;
    mat_transpose(T, M, L);
    HOP_FOR_ELEMS(I, J, s) M.v[I][J][s] = 0.5 * (M.v[I][J][s] + T.v[I][J][s]);
}

HOP_DEVICE bool mat_all_finite(const Mat& M) {
    bool fin = true;
    HOP_FOR_ELEMS(I, J, s) fin = fin && isfinite(M.v[I][J][s]);
    return simt::all(fin);
}

// 1/p for a pivot.  Device: MUFU.RCP64H seed + two Newton steps (<= 1 ulp; the IEEE-correct `1.0/p`
// costs ~3x the instructions and sits on the critical path of every elimination step).
HOP_DEVICE double pivot_rcp(double p) { return simt::rcp_newton(p); }

// One in-place Gauss-Jordan inversion attempt of the leading D x D block (no pivoting, natural pivot
// order).  The pivots met are those of LDL^T, so "some pivot <= 0" <=> np.linalg.cholesky fails; the
// function returns (warp-uniform) whether all D pivots were > 0.
//
// NOTE (numerics): a blocked (8, D-8) elimination with DMMA Schur updates was tried and rejected: it
// needs the explicit inverse of the leading tile, and for the query matrices X_t + Gbar (one ~1e8
// direction from the rank-deficient terminal block, augmented.py:85) that loses the last pivot
// completely (-3.6 instead of +1e-3).  The sequential sweep below forms every pivot as a running
// Schur complement, like Cholesky, and keeps it to ~1e-8 absolute.
template <int D>
HOP_DEVICE bool gj_attempt(Mat& a, const LaneGeo& L) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int Ij = j >> 3, gj = rho_inv(j & 7);             // pivot row: tile row Ij, lanes (gj, *)
        const int Jj = j >> 3, tj = j & 3, sj = (j & 7) >> 2;   // pivot column: tile col Jj, lanes (*, tj), slot sj
        const double p = simt::shfl(a.v[Ij][Jj][sj], (gj << 2) | tj, 32);
        ok = ok && (p > 0.0) && (p <= 1.7976931348623157e308);   // +Inf is non-finite input (utils.py:75), not a pivot
        const double rinv = pivot_rcp(p);
        double pr[2][2], f[2];
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) pr[J][s] = simt::shfl(a.v[Ij][J][s], (gj << 2) | L.t, 32);   // M[j][my cols]
#pragma unroll
        for (int I = 0; I < 2; ++I) f[I] = simt::shfl(a.v[I][Jj][sj], (L.g << 2) | tj, 32) * rinv;   // M[my rows][j] / p
        HOP_FOR_ELEMS(I, J, s) a.v[I][J][s] = fma(-f[I], pr[J][s], a.v[I][J][s]);
        const bool isrow = (L.g == gj), iscol = (L.t == tj);
        if (isrow) {                                             // pivot row: M[j][c] / p
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int s = 0; s < 2; ++s) a.v[Ij][J][s] = pr[J][s] * rinv;
        }
        if (iscol) {                                             // pivot column: -M[i][j] / p ; pivot: 1 / p
            a.v[0][Jj][sj] = -f[0];
            a.v[1][Jj][sj] = -f[1];
            if (isrow) a.v[Ij][Jj][sj] = rinv;
        }
    }
    return ok;
}

// LU with partial pivoting (utils.py:90-93) on lane 0, through a [16][16] shared scratch pair.
template <int D>
HOP_DEVICE_NOINLINE bool lu_inverse_serial16(double* A, double* Xb) {
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Xb[i * 16 + j] = (i == j) ? 1.0 : 0.0;
    for (int j = 0; j < D; ++j) {
        int piv = j;
        double best = fabs(A[j * 16 + j]);
        for (int i = j + 1; i < D; ++i) {
            const double v = fabs(A[i * 16 + j]);
            if (v > best) { best = v; piv = i; }
        }
        if (A[piv * 16 + j] == 0.0) return false;
        if (piv != j)
            for (int q = 0; q < D; ++q) {
                double tmp = A[j * 16 + q]; A[j * 16 + q] = A[piv * 16 + q]; A[piv * 16 + q] = tmp;
                tmp = Xb[j * 16 + q]; Xb[j * 16 + q] = Xb[piv * 16 + q]; Xb[piv * 16 + q] = tmp;
            }
        const double rinv = 1.0 / A[j * 16 + j];
        for (int i = j + 1; i < D; ++i) {
            const double f = A[i * 16 + j] * rinv;
            for (int q = j + 1; q < D; ++q) A[i * 16 + q] -= f * A[j * 16 + q];
            for (int q = 0; q < D; ++q) Xb[i * 16 + q] -= f * Xb[j * 16 + q];
        }
    }
    for (int c = 0; c < D; ++c)
        for (int i = D - 1; i >= 0; --i) {
            double sacc = Xb[i * 16 + c];
            for (int q = i + 1; q < D; ++q) sacc -= A[i * 16 + q] * Xb[q * 16 + c];
            Xb[i * 16 + c] = sacc / A[i * 16 + i];
        }
    return true;
}

enum : int { ST_NONFINITE = 1, ST_LINALG = 2, ST_FLAG_RETRY = 0x100, ST_FLAG_LU = 0x200 };

// utils.py:69-93 chol_inv of a SYMMETRISED matrix S (layout L, zero padding outside D x D).
// scratch: 512 doubles of per-warp shared memory (rare LU fallback only).
template <int D>
HOP_DEVICE void chol_inv(const Mat& S, Mat& out, const LaneGeo& L, double* scratch, double jitter, int max_tries,
                         int& status) {
    double eps = jitter;
    for (int tries = 0;; ++tries) {
        if (tries == max_tries) {   // LU fallback on (S + eps I), eps = jitter * 10^max_tries
            simt::sync();
            HOP_FOR_ELEMS(I, J, s) {
                const int R = L.row(I), C = L.col(J, s);
                scratch[R * 16 + C] = S.v[I][J][s] + ((R == C && R < D) ? eps : 0.0);
            }
            simt::sync();
            bool lu_ok = true;
            if (L.lane == 0) lu_ok = lu_inverse_serial16<D>(scratch, scratch + 256);
            simt::sync();
            lu_ok = simt::all(lu_ok);
            HOP_FOR_ELEMS(I, J, s) {
                const int R = L.row(I), C = L.col(J, s);
                out.v[I][J][s] = (R < D && C < D) ? scratch[256 + R * 16 + C] : 0.0;
            }
            simt::sync();
            status |= ST_FLAG_LU;
            if (!lu_ok) status |= ST_LINALG;
            return;
        }
        Mat a;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            a.v[I][J][s] = S.v[I][J][s] + ((R == C && R < D) ? eps : 0.0);
        }
        if (gj_attempt<D>(a, L)) {
            mat_copy(out, a);
            return;
        }
        if (tries == 0 && !mat_all_finite(S)) {   // utils.py:75: non-finite input raises before any attempt
            status |= ST_NONFINITE;
            HOP_FOR_ELEMS(I, J, s) out.v[I][J][s] = nan("");
            return;
        }
        status |= ST_FLAG_RETRY;
        eps *= 10.0;
    }
}

// Out-of-line copy for cold call sites (fallbacks of the closed-form paths): keeps the hot loop small.
template <int D>
HOP_DEVICE_NOINLINE void chol_inv_cold(const Mat& S, Mat& out, double* scratch, double jitter, int max_tries, int& status) {
    LaneGeo L;
    L.init();
    chol_inv<D>(S, out, L, scratch, jitter, max_tries, status);
}

// Element-wise loaders: every lane fetches its 8 entries of a rows x cols row-major block
// (leading dimension ld); entries outside the block are zero.
HOP_DEVICE void mat_load(Mat& M, const double* __restrict__ src, int rows, int cols, int ld, const LaneGeo& L) {
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        M.v[I][J][s] = (R < rows && C < cols) ? src[R * ld + C] : 0.0;
    }
}
// M = src^T restricted to rows x cols of the RESULT (src is cols x rows, leading dimension ld)
HOP_DEVICE void mat_load_t(Mat& M, const double* __restrict__ src, int rows, int cols, int ld, const LaneGeo& L) {
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        M.v[I][J][s] = (R < rows && C < cols) ? src[C * ld + R] : 0.0;
    }
}

}}  // namespace hop::mma
