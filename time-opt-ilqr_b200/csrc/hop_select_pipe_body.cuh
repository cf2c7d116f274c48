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
//   * the vector stage of step k+2 (e, Q e, P e, K q, K' p, Schur complements) is spread over the pivot
//     sweep of iteration k, and X / U are read straight from global two iterations ahead;
//   * pivot row / column fix-ups are folded into the rank-1 update; positivity is tested on the integer pipe;
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
    // vectors of the pipelined sweep; YQ, YP (2 x 16) and DU (2 x 8) are double-buffered by step parity
    static constexpr int EV = 0, QE = 16, PE = 32, YQ = 48, YP = 80, DU = 112;
    static constexpr int STAGE = 128;             // 2 x kStage staging buffers
    static constexpr int BARS = STAGE + 2 * kStage;
    static constexpr int ROWB = (BARS + 2 + 1) & ~1;   // pivot row / column exchange: [2 parities][3 matrices][16]
    static constexpr int SIZE = ROWB + 2 * 3 * 16;
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

// ---- small helpers --------------------------------------------------------------------------------
// Pivot positivity on the integer pipe (the FP64 pipe is the bottleneck).  Inside the sweeps only the SIGN of
// every pivot is accumulated (one LOP3: signs |= hi word); zero, NaN and +inf pivots turn the reciprocal into
// NaN, which reaches the last pivot of a later X0 and is caught by the once-per-iteration test on that value.
// Any hit sends the problem to the sequential body, which owns the jitter ladder and the status word.
HOP_DEVICE int hi_word(double p) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(p);
#else
    long long bits;
    static_assert(sizeof(bits) == sizeof(p), "");
    __builtin_memcpy(&bits, &p, 8);
    return (int)(bits >> 32);
#endif
}
// true for p <= 0, NaN, +inf and p < 2^-1042
HOP_DEVICE bool pivot_bad(double p) { return (unsigned)(hi_word(p) - 1) >= 0x7fefffffu; }

// 1/p: MUFU.RCP64H seed r0 (rel. error e <= 2^-23) and r0 (1 + e + e^2): three dependent DFMAs, error ~ e^3.
HOP_DEVICE double pivot_rcp3(double p) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    return (p - p == 0.0) ? 1.0 / p : (p - p);   // the device sequence turns +-inf and NaN into NaN (fma(-inf, 0, 1))
#endif
}

// utils.py:127-128 with a branch-free exact fast path: for s = a + pi in [-2pi, 4pi) the floored modulo is
// s, s + 2pi or s - 2pi (the subtraction is exact by Sterbenz), which is what fmod + sign fix-up returns.
HOP_DEVICE double wrap_pi_fast(double a) {
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    const double s = a + pi;
    if (s >= 0.0 && s < two_pi) return s - pi;
    if (s >= two_pi && s < 2.0 * two_pi) return (s - two_pi) - pi;
    return wrap_pi(a);
}

// One pivot of the in-place Gauss-Jordan inversion.  Same values as gj_attempt; the pivot row / column
// fix-ups are folded into the rank-1 update by zeroing the target and patching f / pr:
//   row j:  0 - (-1/p) M[j][c] = M[j][c]/p;   column j:  0 - f_i * 1 = -M[i][j]/p;   (j,j):  0 - (-1/p) * 1 = 1/p
// The pivot index is j = 8 GI + 4 GS + tj: (GI, GS) select REGISTERS and must be compile-time, tj only enters
// lane numbers and predicates and may be a run-time loop variable (looped variant, smaller code).
template <int D, int GI, int GS>
HOP_DEVICE void gj_pivot(Mat& a, int tj, const LaneGeo& L, int& signs) {
    const int gj = 2 * tj + GS;                                  // rho_inv(4 GS + tj)
    const double p = simt::shfl(a.v[GI][GI][GS], (gj << 2) | tj, 32);
    signs |= hi_word(p);
    const double rinv = pivot_rcp3(p);
    double pr[2][2], f[2];
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s) pr[J][s] = simt::shfl(a.v[GI][J][s], (gj << 2) | L.t, 32);
#pragma unroll
    for (int I = 0; I < 2; ++I) f[I] = simt::shfl(a.v[I][GI][GS], (L.g << 2) | tj, 32) * rinv;
    const bool isrow = (L.g == gj), iscol = (L.t == tj);
    if (iscol) {
        a.v[0][GI][GS] = 0.0;
        a.v[1][GI][GS] = 0.0;
        pr[GI][GS] = 1.0;
    }
    if (isrow) {
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) a.v[GI][J][s] = 0.0;
        f[GI] = -rinv;
    }
    HOP_FOR_ELEMS(I, J, s) a.v[I][J][s] = fma(-f[I], pr[J][s], a.v[I][J][s]);
}

// One pivot of the forward elimination of a SYMMETRIC matrix held on its lower tiles (0,0), (1,0), (1,1)
// (tile (0,1) is never read or written).  Row j is taken from column j.  p receives the pivot.
template <int D, int GI, int GS>
HOP_DEVICE void fe_pivot_lower(Mat& a, int tj, const LaneGeo& L, int& signs, double& p) {
    const int gj = 2 * tj + GS;
    p = simt::shfl(a.v[GI][GI][GS], (gj << 2) | tj, 32);
    signs |= hi_word(p);
    if (8 * GI + 4 * GS + tj == D - 1) return;
    const double rinv = pivot_rcp3(p);
    double pr[2][2], f[2];
    // rows/cols <= j are dead: once j >= 8 only tile (1,1) is live
#pragma unroll
    for (int J = GI; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s)   // M[j][8J+t+4s] = M[8J+t+4s][j]: tile (J, GI), lane (rho_inv(t+4s), tj)
            pr[J][s] = simt::shfl(a.v[J][GI][GS], ((2 * L.t + s) << 2) | tj, 32);
#pragma unroll
    for (int I = GI; I < 2; ++I) f[I] = simt::shfl(a.v[I][GI][GS], (L.g << 2) | tj, 32) * rinv;
    if (GI == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[0][0][s] = fma(-f[0], pr[0][s], a.v[0][0][s]);
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[1][0][s] = fma(-f[1], pr[0][s], a.v[1][0][s]);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) a.v[1][1][s] = fma(-f[1], pr[1][s], a.v[1][1][s]);
}
// ---- pivot exchange through shared memory -----------------------------------------------------------------
// The shuffle exchange costs 14 SHFL.32 per pivot and matrix (7 doubles); the sweeps are issue bound (ncu r1e), so the
// exchange goes through a 16-double row buffer instead: the 4 lanes that own pivot row j store it (2 x STS.128), one
// __syncwarp per pivot ROUND (three matrices), and every lane fetches its 4 column values (2 x LDS.128), its 2 row
// values (1 x LDS.128) and the pivot (1 x LDS.64): 18 shared-memory instructions per round instead of 39 shuffles.
// Buffer position of column c:  pos(c) = 4 (c & 3) + 2 ((c >> 2) & 1) + (c >> 3), so that a lane's own four columns
// {t, t+4, t+8, t+12} and its own two rows {rho(g), rho(g)+8} are contiguous, 16-byte aligned runs (4t.., 2g..).
// To take the column from the ROW the sweep must keep the matrix symmetric: this is the symmetric sweep operator
// (row AND column scaled by +1/p, pivot -> -1/p), which turns A into -A^-1; it is applied to -S, whose pivots are
// negative, and returns S^-1.  Same rank-1 update and same folded fix-ups as gj_pivot (pr' = -1 instead of +1).
HOP_DEVICE void st2(double* p, double a, double b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
HOP_DEVICE void ld2(const double* p, double& a, double& b) {
#if defined(__CUDA_ARCH__)
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
// a, or a value that acts as zero, when z: only the HIGH word is cleared (one SEL instead of two).  What is left is
// the low word read as a subnormal (< 2^-1042); the only consumer is the addend of  fma(x, y, .)  with |x y| > 1e-290
// (or x y == 0 exactly, and then a == 0 and its low word is zero too), so the rounded result is the one a true zero gives
// (a product that is an exact rounding tie, probability 2^-53, aside).
HOP_DEVICE double zero_hi_if(double a, bool z) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(z ? 0 : __double2hiint(a), __double2loint(a));
#else
    unsigned long long bits;
    __builtin_memcpy(&bits, &a, 8);
    if (z) bits &= 0xffffffffull;
    __builtin_memcpy(&a, &bits, 8);
    return a;
#endif
}
template <int GI>
HOP_DEVICE void gjs_store(const Mat& a, double* buf, bool isrow, int t) {      // pivot row: lanes g == gj
    if (isrow) {
        st2(buf + 4 * t, a.v[GI][0][0], a.v[GI][1][0]);
        st2(buf + 4 * t + 2, a.v[GI][0][1], a.v[GI][1][1]);
    }
}
template <int GI, int GS>
HOP_DEVICE void fes_store(const Mat& a, double* buf, bool iscol, int g) {      // pivot COLUMN (lower tiles): lanes t == tj
    if (iscol) st2(buf + 2 * g, a.v[0][GI][GS], a.v[1][GI][GS]);
}
template <int D, int GI, int GS, bool HIZ = false>
HOP_DEVICE void gjs_apply(Mat& a, int tj, const LaneGeo& L, int& signs, const double* buf, bool isrow, bool iscol) {
    double pr[2][2], f[2];
    ld2(buf + 4 * L.t, pr[0][0], pr[1][0]);
    ld2(buf + 4 * L.t + 2, pr[0][1], pr[1][1]);
    ld2(buf + 2 * L.g, f[0], f[1]);
    const double p = buf[4 * tj + 2 * GS + GI];
    signs |= ~hi_word(p);                                        // pivots of -S must be negative
    const double rinv = pivot_rcp3(p);
    f[0] *= rinv;
    f[1] *= rinv;
    if (iscol) {
        a.v[0][GI][GS] = 0.0;
        a.v[1][GI][GS] = 0.0;
        pr[GI][GS] = -1.0;
    }
    if (HIZ) {
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) a.v[GI][J][s] = zero_hi_if(a.v[GI][J][s], isrow);
        if (isrow) f[GI] = -rinv;
    } else if (isrow) {
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) a.v[GI][J][s] = 0.0;
        f[GI] = -rinv;
    }
    HOP_FOR_ELEMS(I, J, s) a.v[I][J][s] = fma(-f[I], pr[J][s], a.v[I][J][s]);
}
template <int D, int GI, int GS>
HOP_DEVICE void fes_apply(Mat& a, int tj, const LaneGeo& L, int& signs, double& p, const double* buf) {
    p = buf[4 * tj + 2 * GS + GI];
    signs |= hi_word(p);
    if (8 * GI + 4 * GS + tj == D - 1) return;
    const double rinv = pivot_rcp3(p);
    double pr[2][2], f[2];
    ld2(buf + 4 * L.t, pr[0][0], pr[1][0]);
    ld2(buf + 4 * L.t + 2, pr[0][1], pr[1][1]);
    ld2(buf + 2 * L.g, f[0], f[1]);
    f[0] *= rinv;
    f[1] *= rinv;
    if (GI == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[0][0][s] = fma(-f[0], pr[0][s], a.v[0][0][s]);
#pragma unroll
        for (int s = 0; s < 2; ++s) a.v[1][0][s] = fma(-f[1], pr[0][s], a.v[1][0][s]);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) a.v[1][1][s] = fma(-f[1], pr[1][s], a.v[1][1][s]);
}
// one pivot round of the three sweeps; a1, a2 hold NEGATED matrices (see above), x the lower tiles of X0 + eps I.
// PAR: parity of the exchange buffer, -1 = tj & 1 at run time; a compile-time parity turns every shared-memory address
// of the round into base register + immediate.
template <int D, int GI, int GS, int PAR = -1, bool HIZ = false>
HOP_DEVICE void gj3s_round(Mat& a1, Mat& a2, Mat& x, int tj, const LaneGeo& L, int& signs, double& p, double* rowb) {
    const int gj = 2 * tj + GS;
    const bool isrow = (L.g == gj), iscol = (L.t == tj);
    double* buf = rowb + (PAR < 0 ? (tj & 1) : PAR) * 48;
    gjs_store<GI>(a1, buf, isrow, L.t);
    gjs_store<GI>(a2, buf + 16, isrow, L.t);
    fes_store<GI, GS>(x, buf + 32, iscol, L.g);
    simt::sync();
    gjs_apply<D, GI, GS, HIZ>(a1, tj, L, signs, buf, isrow, iscol);
    gjs_apply<D, GI, GS, HIZ>(a2, tj, L, signs, buf + 16, isrow, iscol);
    fes_apply<D, GI, GS>(x, tj, L, signs, p, buf + 32);
}
// U2: the run-time loop advances two pivots per trip (buffer parity known at compile time, half the loop overhead,
// twice the code of the plain loop) and the pivot-row zeroing is a high-word select (zero_hi_if)
template <int D, int GI, int GS, bool UNROLL, bool U2 = false>
HOP_DEVICE void gj3s_group(Mat& a1, Mat& a2, Mat& x, const LaneGeo& L, int& signs, double& p, double* rowb) {
    constexpr int first = 8 * GI + 4 * GS;
    constexpr int cnt = (D - first) < 4 ? (D - first) : 4;
    if (UNROLL) {
#pragma unroll
        for (int tj = 0; tj < cnt; ++tj) gj3s_round<D, GI, GS>(a1, a2, x, tj, L, signs, p, rowb);
    } else if (U2) {
#pragma unroll 1
        for (int tj = 0; tj + 1 < cnt; tj += 2) {
            gj3s_round<D, GI, GS, 0, true>(a1, a2, x, tj, L, signs, p, rowb);
            gj3s_round<D, GI, GS, 1, true>(a1, a2, x, tj + 1, L, signs, p, rowb);
        }
        if (cnt & 1) gj3s_round<D, GI, GS, 0, true>(a1, a2, x, cnt - 1, L, signs, p, rowb);
    } else {
#pragma unroll 1
        for (int tj = 0; tj < cnt; ++tj) gj3s_round<D, GI, GS>(a1, a2, x, tj, L, signs, p, rowb);
    }
}

// the three interleaved sweeps over the pivots of group (GI, GS); UNROLL = false keeps tj a run-time loop
template <int D, int GI, int GS, bool UNROLL>
HOP_DEVICE void gj3_group(Mat& a1, Mat& a2, Mat& x, const LaneGeo& L, int& signs, double& p) {
    constexpr int first = 8 * GI + 4 * GS;
    constexpr int cnt = (D - first) < 4 ? (D - first) : 4;
    if (UNROLL) {
#pragma unroll
        for (int tj = 0; tj < cnt; ++tj) {
            gj_pivot<D, GI, GS>(a1, tj, L, signs);
            gj_pivot<D, GI, GS>(a2, tj, L, signs);
            fe_pivot_lower<D, GI, GS>(x, tj, L, signs, p);
        }
    } else {
#pragma unroll 1
        for (int tj = 0; tj < cnt; ++tj) {
            gj_pivot<D, GI, GS>(a1, tj, L, signs);
            gj_pivot<D, GI, GS>(a2, tj, L, signs);
            fe_pivot_lower<D, GI, GS>(x, tj, L, signs, p);
        }
    }
}
static_assert(rho_inv(0) == 0 && rho_inv(1) == 2 && rho_inv(4) == 1 && rho_inv(7) == 7, "rho_inv(t + 4s) == 2t + s");

// D = X * Z^T on the lower tiles (0,0), (1,0), (1,1) only (symmetric result).
template <int KB, bool ACC>
HOP_DEVICE void mma_nt_lower(Mat& Dm, const Mat& X, const Mat& Z) {
    if (!ACC) {
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) { Dm.v[I][J][0] = 0.0; Dm.v[I][J][1] = 0.0; }
    }
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J)
                simt::dmma(Dm.v[I][J][0], Dm.v[I][J][1], X.v[I][kb >> 1][kb & 1], Z.v[J][kb >> 1][kb & 1]);
}

// ---- the k = D-1 column as a rank-1 DFMA update ----------------------------------------------------------
// For d = 13 the fourth k-block of every product holds ONE non-zero column (k = 12): four DMMAs (64 pipe clocks)
// that do the work of a rank-1 update.  The products of the pipelined body therefore run three k-blocks on the
// tensor pipe and add  x_12 z_12^T  with DFMAs (8 per product, 16 pipe clocks); the two vectors come from the
// operands' fragments by shuffles (or from shared memory where the column is known in closed form).
template <int D>
struct LastCol {
    static constexpr int k = D - 1, Jk = k >> 3, sk = (k & 7) >> 2, tk = k & 3;   // register [.][Jk][sk], lanes t == tk
    static constexpr bool split = (D % 4) == 1;                                    // exactly one column in the last k-block
};
// M[row(I)][D-1] for the two row tiles of this lane (the X-operand role)
template <int D>
HOP_DEVICE void last_col_rows(double (&r)[2], const Mat& M, const LaneGeo& L) {
    using LC = LastCol<D>;
#pragma unroll
    for (int I = 0; I < 2; ++I) r[I] = simt::shfl(M.v[I][LC::Jk][LC::sk], (L.g << 2) | LC::tk, 32);
}
// M[col(J,s)][D-1] for the four column slots of this lane (the Z-operand role)
template <int D>
HOP_DEVICE void last_col_cols(double (&c)[2][2], const Mat& M, const LaneGeo& L) {
    using LC = LastCol<D>;
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int s = 0; s < 2; ++s) c[J][s] = simt::shfl(M.v[J][LC::Jk][LC::sk], ((2 * L.t + s) << 2) | LC::tk, 32);
}
template <bool LOWER>
HOP_DEVICE void rank1_add(Mat& Dm, const double (&r)[2], const double (&c)[2][2]) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
            if (!LOWER || J <= I) {
#pragma unroll
                for (int s = 0; s < 2; ++s) Dm.v[I][J][s] = fma(r[I], c[J][s], Dm.v[I][J][s]);
            }
}

// NOTE (numerics): forming Ebar / Gbar from the lower tiles of their products and mirroring was tried and
// rejected.  W comes out of the Gauss-Jordan sweep with an antisymmetric rounding component Omega, and
// Fbar Omega Fbar^T (|Fbar| ~ 1e8) is exactly antisymmetric: 0.5 (M + M^T) removes it, mirroring the lower
// tiles keeps it (error at T* 8e-8 instead of 4e-10 on the S1 goldens).  X0 only feeds pivots and is fine.

// Returns true when the problem was solved by the pipelined sweep; false => caller must run the sequential body.
// SCHED 0: unrolled sweep, shuffle exchange; 1: looped sweep, shuffle exchange; 2: looped sweep, shared-memory exchange;
// 3: as 2 with two pivots per loop trip and high-word zeroing of the pivot row
template <int D, int M, int SCHED>
HOP_DEVICE bool select_fused_pipe_body(const FusedArgs& p, int b, double* scratch, const double* cst) {
    using FC = FusedConst<D, M>;
    using XC = FastConst<D, M>;
    using PC = PipeConst<D, M>;
    using PS = PipeSlab;
    constexpr int n = D - 1;
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4, KBM = (M + 3) / 4, NTM = (M + 7) / 8;
    constexpr bool LOOPED = (SCHED != 0), SMX = (SCHED >= 2), U2 = (SCHED == 3);
    double* rowb = scratch + PipeSlab::ROWB;
    constexpr bool R1 = LastCol<D>::split;          // last k-block as a rank-1 DFMA update
    constexpr int KD = R1 ? KB - 1 : KB;            // k-blocks left on the tensor pipe
    static_assert(D > 8 && D <= 16 && n <= 16 && M <= 4, "one-problem-per-warp mapping: 9 <= d <= 16, m <= 4");
    if (cst[XC::FLAG] != 0.0) return false;      // K / K' needed the ladder: closed forms do not apply
    LaneGeo L;
    L.init();
    double* EV = scratch + PS::EV;   // e = wrap(X_s - xg)
    double* QE = scratch + PS::QE;   // Q e
    double* PE = scratch + PS::PE;   // P e
    bool bad = false;

    // R_inv = chol_inv(sym(R)) (augmented.py:23); its transpose is the Z operand of B R^-1
    Mat RinvT;
    {
        Mat Rs;
        HOP_FOR_ELEMS(I, J, s) {
            const int R = L.row(I), C = L.col(J, s);
            Rs.v[I][J][s] = ((R < M && C < M) ? cst[FC::RS + R * M + C] : 0.0) + ((R == C && R < M) ? p.jitter : 0.0);
        }
        bad = !gj_attempt<M>(Rs, L);
        mat_transpose(RinvT, Rs, L);
    }
    // half-warp roles for the vector work: h = 0 -> Q side (Q e, K q), h = 1 -> terminal side (P e, K' p)
    const int h = L.lane >> 4, li = L.lane & 15;
    const bool isx = li < n;
    const double xg_l = isx ? p.xg[(size_t)b * n + li] : 0.0;
    const bool wrap_l = isx && ((p.wrap_mask >> li) & 1u);
    const double uref_l = (L.lane < M) ? cst[FC::UREF + L.lane] : 0.0;
    const double w = p.w[b];
    const double* mat1 = cst + (h ? FC::PF : PC::QRAWT) + li;   // symmetric P | Q^T : element [i][j] at [j*n + i]
    const double* mat2 = cst + (h ? XC::KP : XC::KQ) + li;      // symmetric K' | K
    const double* matc = cst + FC::QRAW + li;                   // column i of Q
    double* V1 = h ? PE : QE;
    // extended vectors [y ; -1 ; 0], control deviations: double-buffered by the parity of the step index
    if (L.lane < 16) {
        const double v = (L.lane == n) ? -1.0 : 0.0;
        scratch[PS::YQ + L.lane] = v; scratch[PS::YQ + 16 + L.lane] = v;
        scratch[PS::YP + L.lane] = v; scratch[PS::YP + 16 + L.lane] = v;
    }
    const double* KF1 = cst + PC::KQF + L.lane;
    const double* KF2 = cst + PC::KPF + L.lane;

    const size_t baseN = (size_t)b * p.N;
    const double* Xb = p.X + (size_t)b * (p.N + 1) * n;
    const double* Ub = p.U + (size_t)b * p.u_stride;
    static_assert(n * n + n * M + n <= kStage, "staging buffer too small");
    static_assert((n * n) % 2 == 0 && (n * M) % 2 == 0 && n % 2 == 0, "bulk copies need 16-byte multiples");
    double* stage0 = scratch + PS::STAGE;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(scratch + PS::BARS);
    constexpr int oA = 0, oB = n * n, oR = oB + n * M;
    // stage s (< T_max): A_s, B_s, a_s through TMA bulk copies; X and U are read straight from global
    auto issue = [&](int s) {
        double* st = stage0 + (s & 1) * kStage;
        unsigned long long* bar = bars + (s & 1);
        const unsigned bytes = 8u * (n * n + n * M + (p.a_resid ? n : 0));
        simt::mbar_expect_tx(bar, bytes);
        simt::bulk_g2s(st + oA, p.A + (baseN + s) * n * n, 8u * n * n, bar);
        simt::bulk_g2s(st + oB, p.Bm + (baseN + s) * n * M, 8u * n * M, bar);
        if (p.a_resid) simt::bulk_g2s(st + oR, p.a_resid + (baseN + s) * n, 8u * n, bar);
    };
    if (L.lane == 0) {
        simt::mbar_init(bars, 1);
        simt::mbar_init(bars + 1, 1);
        simt::mbar_fence_init();
    }
    simt::sync();
    if (L.lane == 0) { issue(0); if (1 < p.T_max) issue(1); }
    simt::sync();   // (host emulation: copies complete at issue time, so order the issue before the first read)
    // leave no bulk copy in flight and no live mbarrier behind (the fallback body re-uses the slab)
    auto bail = [&](int pending_stage) {
        if (pending_stage >= 0) simt::mbar_wait(bars + (pending_stage & 1), (unsigned)((pending_stage >> 1) & 1));
        simt::sync();
        if (L.lane == 0) { simt::mbar_inval(bars); simt::mbar_inval(bars + 1); }
        simt::sync();
        return false;
    };
    auto load_x = [&](int s) { return (isx && s <= p.T_max) ? Xb[(size_t)s * n + li] : 0.0; };
    auto load_u = [&](int s) { return (L.lane < M && s < p.T_max) ? Ub[(size_t)s * M + L.lane] : 0.0; };

    // ---- vector stage of step index s, in four phases (A: e, du; B: Q e | P e, e^T Q e; C: K q | K' p and the
    // two dot products; D: Schur complements and their reciprocals).  In the main loop the phases of step k+2
    // are spread over the Gauss-Jordan sweep of iteration k so that their latency hides behind the pivots.
    struct Vec { double ev, v1, corner, qy, ye, rsq, rsp; };
    auto vecA = [&](Vec& V, int s, double xs, double us) {
        double ev = 0.0;
        if (isx) {
            ev = xs - xg_l;                                                     // e = wrap(X_s - xg)  (augmented.py:28,80)
            if (wrap_l) ev = wrap_pi_fast(ev);
            if (h == 0) EV[li] = ev;
        }
        V.ev = ev;
        if (L.lane < M) scratch[PS::DU + (s & 1) * 8 + L.lane] = us - uref_l;   // du = U_s - u_ref   (augmented.py:29)
    };
    auto vecB = [&](Vec& V) {
        simt::sync();
        double v1 = 0.0, qc = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const double ej = EV[j];
                v1 = fma(mat1[j * n], ej, v1);                                  // (Q e)_i | (P e)_i
                qc = fma(ej, matc[j * n], qc);                                  // (e^T Q)_i
            }
            V1[li] = v1;
        }
        V.v1 = v1;
        double r0 = (isx && h == 0) ? qc * V.ev : 0.0;                          // e^T Q e   (augmented.py:37)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) r0 += simt::shfl_xor(r0, o, 32);
        V.corner = simt::shfl(r0, 0, 32) + 2.0 * w + p.rho_reg;
    };
    auto vecC = [&](Vec& V, int s) {
        simt::sync();
        double y = 0.0;
        if (isx) {
#pragma unroll
            for (int j = 0; j < n; ++j) y = fma(mat2[j * n], V1[j], y);         // y = K q | y' = K' p
            scratch[(h ? PS::YP : PS::YQ) + (s & 1) * 16 + li] = y;
        }
        double r1 = isx ? (h ? y * V.ev : V.v1 * y) : 0.0;                      // q^T y | y'^T e
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) r1 += simt::shfl_xor(r1, o, 32);
        V.qy = simt::shfl(r1, 0, 32);
        V.ye = simt::shfl(r1, 16, 32);
    };
    auto vecD = [&](Vec& V, bool live) {
        const double sigq = (V.corner + p.jitter) - V.qy;                       // Schur complement of Q_aug + eps I
        const double sigp = (p.rho_reg + p.jitter) + p.jitter * V.ye;           // ... of QT + eps I, cancellation-free
        bad = bad || (live && (pivot_bad(sigq) || pivot_bad(sigp)));
        V.rsq = pivot_rcp(sigq);
        V.rsp = pivot_rcp(sigp);
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
    // F_k^T = A_k E_k without the tensor pipe when K is diagonal:  E = diag(K, 0) + rs yx yx^T  (yx = [y ; -1 ; 0])  gives
    //   A E = A diag(K, 0) + rs (A yx) yx^T :  a column scaling and a rank-1 term -- 30 FP64 instructions (60 pipe clocks)
    // instead of 12 DMMAs + the rank-1 last column (208 clocks), and E_k itself is never formed.
    const bool kdiag = (cst[XC::FLAG + 1] != 0.0);
    auto ft_closed = [&](Mat& Ft, const Mat& A, const double* Y, double rs) {
        double yc[2][2], kd[2][2], u[2];
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int c = L.col(J, s);
                yc[J][s] = Y[c];
                kd[J][s] = (c < n) ? cst[XC::KQ + c * n + c] : 0.0;
            }
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            double a = 0.0;
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int s = 0; s < 2; ++s) a = fma(A.v[I][J][s], yc[J][s], a);
            a += simt::shfl_xor(a, 1, 32);
            a += simt::shfl_xor(a, 2, 32);                                      // (A yx)_row: the four lanes of a row hold 4 columns each
            u[I] = a * rs;
        }
        HOP_FOR_ELEMS(I, J, s) Ft.v[I][J][s] = fma(u[I], yc[J][s], A.v[I][J][s] * kd[J][s]);
    };
    auto load_AB = [&](const double* stg, const double* DU, Mat& A, Mat& Bm) {
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

    // ---------------- prologue: vector stages 0 and 1, prefix step 0 (horizon_selection.py:57-64 with k = 0)
    Vec V0, V1s;
    vecA(V0, 0, load_x(0), load_u(0)); vecB(V0); vecC(V0, 0); vecD(V0, true);
    simt::sync();
    vecA(V1s, 1, load_x(1), load_u(1)); vecB(V1s); vecC(V1s, 1); vecD(V1s, true);
    simt::sync();
    double rsq = V1s.rsq, rsp = V1s.rsp;       // reciprocal Schur complements of step k+1 (current iteration)
    double xcur = load_x(2), ucur = load_u(2);  // inputs of the vector stage that runs inside iteration k: step k+2
    PrefixL<D> P;
    simt::mbar_wait(bars + 0, 0u);
    {
        Mat E, A, Bm, Ft, G;
        closed_inverse(E, KF1, scratch + PS::YQ, V0.rsq);
        load_AB(stage0, scratch + PS::DU, A, Bm);
        mma_nt<NT, NT, KB, false>(Ft, A, E);                                   // F_0^T = A_0 E_0
        mma_nt<NT, NT, KB, false>(G, Ft, A);                                   // (A_0 E_0) A_0^T              (:61)
        Mat BR;
        mma_nt<NT, NTM, KBM, false>(BR, Bm, RinvT);                            // B_0 R^-1
        mma_nt<NT, NT, KBM, true>(G, BR, Bm);                                  // + (B_0 R^-1) B_0^T
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
        const int q1 = (k + 1) & 1;                                             // buffers of step k+1
        const bool do_vec = (k + 2 <= p.T_max);
        simt::sync();                                                           // everyone is done with stage k and vectors k
        if (L.lane == 0 && k + 2 < p.T_max) issue(k + 2);
        const double xnext = load_x(k + 3), unext = load_u(k + 3);              // consumed one iteration from now
        // ---------------- the three independent inversions
        Mat W, Wt;
        {
            closed_inverse(W, KF1, scratch + PS::YQ + q1 * 16, rsq);            // E_{k+1} = chol_inv(Q_aug[k+1])   (:59)
            closed_inverse(Wt, KF2, scratch + PS::YP + q1 * 16, rsp);           // X_t = chol_inv(QT_t), t = k+1     (:79)
            HOP_FOR_ELEMS(I, J, s) {
                const double dg = (I == J && L.row(I) == L.col(J, s) && L.row(I) < D) ? p.jitter : 0.0;
                const double s1 = (W.v[I][J][s] + P.gb.v[I][J][s]) + dg;       // E_{k+1} + Gbar_k (+ eps I)        (:72)
                const double s2 = (Wt.v[I][J][s] + P.gb.v[I][J][s]) + dg;      // X_t + Gbar_k (+ eps I)            (:82)
                W.v[I][J][s] = SMX ? -s1 : s1;                                 // (the symmetric sweep inverts -S)
                Wt.v[I][J][s] = SMX ? -s2 : s2;
                if (I == J) X0.v[I][J][s] += dg;                               // X0_{t-1} + eps I                  (:84)
            }
        }
        Vec V;                                                                  // (past the horizon the inputs are zeros: harmless)
        double piv = 0.0;
        static_assert(D > 12, "the phase placement below assumes four pivot groups");
        vecA(V, k + 2, xcur, ucur);
        int signs = 0;
        if (SMX) {
            gj3s_group<D, 0, 0, false, U2>(W, Wt, X0, L, signs, piv, rowb);
            vecB(V);
            gj3s_group<D, 0, 1, false, U2>(W, Wt, X0, L, signs, piv, rowb);
            vecC(V, k + 2);
            gj3s_group<D, 1, 0, false, U2>(W, Wt, X0, L, signs, piv, rowb);
            vecD(V, do_vec);
            gj3s_group<D, 1, 1, false, U2>(W, Wt, X0, L, signs, piv, rowb);
        } else {
            gj3_group<D, 0, 0, !LOOPED>(W, Wt, X0, L, signs, piv);
            vecB(V);
            gj3_group<D, 0, 1, !LOOPED>(W, Wt, X0, L, signs, piv);
            vecC(V, k + 2);
            gj3_group<D, 1, 0, !LOOPED>(W, Wt, X0, L, signs, piv);
            vecD(V, do_vec);
            gj3_group<D, 1, 1, !LOOPED>(W, Wt, X0, L, signs, piv);
        }
        bad = bad || (signs < 0) || pivot_bad(piv);                             // piv: last pivot of X0_{t-1} (or of I at k = 0)
        if (simt::ballot(bad) != 0u) return bail(k + 2 < p.T_max ? k + 2 : -1);
        if (k > 0 && L.lane == 0) {                                            // J(t-1) = 0.5 / pivot_n  (z0 = e_n, :85)
            const double Jt = 0.5 / piv;
            p.J_out[(size_t)b * p.T_max + (k - 1)] = Jt;
            if (k >= p.T_min) am.push(Jt, k);
        }
        // ---------------- query products of horizon t = k+1 (:83): X0 = Ebar - (Fbar W_t) Fbar^T, lower tiles
        double fb_r[2], fb_c[2][2];                                            // column d-1 of Fbar in both operand roles
        if (R1) { last_col_rows<D>(fb_r, P.fb, L); last_col_cols<D>(fb_c, P.fb, L); }
        {
            Mat T3, acc;
            mma_nt<NT, NT, KD, false>(T3, P.fb, Wt);
            if (R1) { double c[2][2]; last_col_cols<D>(c, Wt, L); rank1_add<false>(T3, fb_r, c); }
            mma_nt_lower<KD, false>(acc, T3, P.fb);
            if (R1) { double r[2]; last_col_rows<D>(r, T3, L); rank1_add<true>(acc, r, fb_c); }
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
#pragma unroll
                    for (int s = 0; s < 2; ++s) X0.v[I][J][s] = P.eb.v[I][J][s] - acc.v[I][J][s];
        }
        if (last) break;
        // ---------------- prefix step k+1 (:57-75)
        simt::mbar_wait(bars + q1, (unsigned)(((k + 1) >> 1) & 1));
        {
            Mat A, Bm, Ft, G;
            load_AB(stage0 + q1 * kStage, scratch + PS::DU + q1 * 8, A, Bm);
            double ft_r[2], ft_c[2][2], w_c[2][2];
            if (kdiag) {
                ft_closed(Ft, A, scratch + PS::YQ + q1 * 16, rsq);             // F_k^T = A_k E_k, closed form
            } else {
                Mat E;
                closed_inverse(E, KF1, scratch + PS::YQ + q1 * 16, rsq);
                mma_nt<NT, NT, KD, false>(Ft, A, E);                           // F_k^T = A_k E_k
                if (R1) {
                    double r[2], c[2][2];
                    last_col_rows<D>(r, A, L); last_col_cols<D>(c, E, L);
                    rank1_add<false>(Ft, r, c);
                }
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
            Mat T1, acc;
            mma_nt<NT, NT, KD, false>(T1, P.fb, W);                            // Fbar W                       (:73)
            if (R1) { last_col_cols<D>(w_c, W, L); rank1_add<false>(T1, fb_r, w_c); }
            double t1_r[2];
            if (R1) last_col_rows<D>(t1_r, T1, L);
            mma_nt<NT, NT, KD, false>(acc, T1, P.fb);                          // (Fbar W) Fbar^T
            if (R1) rank1_add<false>(acc, t1_r, fb_c);
            mat_sub(P.eb, P.eb, acc);
            mat_sym(P.eb, L);                                                  // Ebar                         (:73)
            mma_nt<NT, NT, KD, false>(acc, T1, Ft);                            // (Fbar W) F_k  -> new Fbar    (:74)
            if (R1) rank1_add<false>(acc, t1_r, ft_c);
            mma_nt<NT, NT, KD, false>(T1, Ft, W);                              // F_k^T W                      (:75)
            if (R1) rank1_add<false>(T1, ft_r, w_c);
            mat_copy(P.fb, acc);
            mma_nt<NT, NT, KD, false>(acc, T1, Ft);                            // (F_k^T W) F_k
            if (R1) { double r[2]; last_col_rows<D>(r, T1, L); rank1_add<false>(acc, r, ft_c); }
            mat_sub(P.gb, G, acc);
            mat_sym(P.gb, L);                                                  // Gbar                         (:75)
        }
        rsq = V.rsq; rsp = V.rsp;
        xcur = xnext; ucur = unext;
    }
    // ---------------- epilogue: cost of the last horizon
    {
        HOP_FOR_ELEMS(I, J, s)
            if (I == J && L.row(I) == L.col(J, s) && L.row(I) < D) X0.v[I][J][s] += p.jitter;
        double piv = 0.0;
        Mat d1, d2;                                                            // dummies: the sweep is shared with the main loop
        HOP_FOR_ELEMS(I, J, s) d1.v[I][J][s] = d2.v[I][J][s] = (L.row(I) == L.col(J, s)) ? (SMX ? -1.0 : 1.0) : 0.0;
        int signs = 0;
        if (SMX) {
            simt::sync();
            gj3s_group<D, 0, 0, false>(d1, d2, X0, L, signs, piv, rowb);
            gj3s_group<D, 0, 1, false>(d1, d2, X0, L, signs, piv, rowb);
            gj3s_group<D, 1, 0, false>(d1, d2, X0, L, signs, piv, rowb);
            gj3s_group<D, 1, 1, false>(d1, d2, X0, L, signs, piv, rowb);
        } else {
            gj3_group<D, 0, 0, false>(d1, d2, X0, L, signs, piv);
            gj3_group<D, 0, 1, false>(d1, d2, X0, L, signs, piv);
            gj3_group<D, 1, 0, false>(d1, d2, X0, L, signs, piv);
            gj3_group<D, 1, 1, false>(d1, d2, X0, L, signs, piv);
        }
        bad = bad || (signs < 0) || pivot_bad(piv);
        if (simt::ballot(bad) != 0u) return bail(-1);
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
