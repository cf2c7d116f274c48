// hop_select_scan_body.cuh -- HOP_MODE_SCAN: the prefix composition as a chunked parallel scan over the horizon.
//
// The LFT maps g_k = (E_k, F_k, G_k) compose associatively (horizon_selection.py:66-75 is the left fold
// gbar_{0:k} = gbar_{0:k-1} o g_k), so the horizon can be cut into C contiguous chunks, one warp each:
//   phase 1  every warp folds ITS chunk from scratch (local aggregate); warp 0's local prefixes already are the
//            true prefixes, so it also answers its queries;
//   phase 2  warp c composes the aggregates of chunks 0..c-1 (c-1 general compositions) -> its incoming prefix;
//   phase 3  warp c folds its chunk again, now from the incoming prefix, answering the query of every horizon.
// Latency per problem: 2 T/C + C - 1 step-equivalents instead of T (3.3x shorter at T = 128, C = 8) for ~2x the
// prefix work -- the right trade for SMALL batches (the single-instance drop-in call, a handful of problems),
// where the sequential kernels leave the machine empty.  With >= 10^4 independent problems the batch already
// fills the GPU and the sequential modes are the efficient ones.
//
// Numerics: re-association changes the rounding (<= 1e-9 on the synthetic S2 problems and on the
// quadrotor at iteration 0, 4e-8 on a converged quadrotor trajectory, UNSAFE on cartpole where it flips T*), so the
// scan is opt-in and only offered at the LQR boundary (hop_select_f64); chunk 0 is bit-identical to the
// sequential sweep.
#pragma once
#include "hop_select_mma_body.cuh"

namespace hop { namespace mma {

constexpr int kScanWarps = 8;

// per-CTA shared memory (doubles): C warp slabs, C aggregates (3 matrices in fragment order), C argmin records
struct ScanSmem {
    static constexpr int AGG = 3 * 8 * 32;                        // one aggregate
    HOP_HD static constexpr int slabs(int C) { return C * kWarpScratch; }
    HOP_HD static constexpr int aggs(int C) { return slabs(C); }
    HOP_HD static constexpr int recs(int C) { return aggs(C) + C * AGG; }   // per warp: best, idx, nan_hit, status
    HOP_HD static constexpr int size(int C) { return recs(C) + C * 4; }
};

HOP_DEVICE void agg_store(double* dst, const Mat& E, const Mat& F, const Mat& G, int lane) {
    HOP_FOR_ELEMS(I, J, s) {
        const int e = (I * 2 + J) * 2 + s;
        dst[(0 * 8 + e) * 32 + lane] = E.v[I][J][s];
        dst[(1 * 8 + e) * 32 + lane] = F.v[I][J][s];
        dst[(2 * 8 + e) * 32 + lane] = G.v[I][J][s];
    }
}
HOP_DEVICE void agg_load(const double* src, Mat& E, Mat& F, Mat& G, int lane) {
    HOP_FOR_ELEMS(I, J, s) {
        const int e = (I * 2 + J) * 2 + s;
        E.v[I][J][s] = src[(0 * 8 + e) * 32 + lane];
        F.v[I][J][s] = src[(1 * 8 + e) * 32 + lane];
        G.v[I][J][s] = src[(2 * 8 + e) * 32 + lane];
    }
}

// P <- P o (E2, F2, G2): W = chol_inv(E2 + Gbar); Ebar <- sym(Ebar - (Fbar W) Fbar^T); Fbar <- (Fbar W) F2;
// Gbar <- sym(G2 - (F2^T W) F2)   (the general form of horizon_selection.py:72-75)
template <int D>
HOP_DEVICE void compose(PrefixL<D>& P, const Mat& E2, const Mat& F2, const Mat& G2, const LaneGeo& L, double* scratch,
                        double jitter, int max_tries, int& status) {
    constexpr int NT = (D + 7) / 8, KB = (D + 3) / 4;
    Mat S, W, F2T, T1, acc;
    mat_add(S, E2, P.gb);
    mat_sym(S, L);
    chol_inv<D>(S, W, L, scratch, jitter, max_tries, status);
    mat_transpose(F2T, F2, L);
    mma_nt<NT, NT, KB, false>(T1, P.fb, W);                                    // Fbar W
    mma_nt<NT, NT, KB, false>(acc, T1, P.fb);                                  // (Fbar W) Fbar^T
    mat_sub(P.eb, P.eb, acc);
    mat_sym(P.eb, L);
    mma_nt<NT, NT, KB, false>(acc, T1, F2T);                                   // (Fbar W) F2
    mma_nt<NT, NT, KB, false>(T1, F2T, W);                                     // F2^T W
    mat_copy(P.fb, acc);
    mma_nt<NT, NT, KB, false>(acc, T1, F2T);                                   // (F2^T W) F2
    mat_sub(P.gb, G2, acc);
    mat_sym(P.gb, L);
}

struct ScanChunk {
    int k0, k1;
    HOP_DEVICE void set(int T_max, int C, int c) {
        const int Lc = (T_max + C - 1) / C;
        k0 = c * Lc < T_max ? c * Lc : T_max;
        k1 = (c + 1) * Lc < T_max ? (c + 1) * Lc : T_max;
    }
    HOP_DEVICE bool empty() const { return k1 <= k0; }
};

// One pass of a warp over its chunk.  from_scratch: fold the chunk alone (phase 1); otherwise continue from P.
// queries: evaluate J(t) for every step of the chunk (always, except for the from-scratch pass of warps c > 0).
template <int D, int M>
HOP_DEVICE void scan_chunk_pass(const SelectArgs& p, int b, const ScanChunk& ch, PrefixL<D>& P, bool from_scratch, bool queries,
                                const LaneGeo& L, double* scratch, ArgMin& am, int& status) {
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
    const double wexp = p.w_explicit ? p.w_explicit[b] : 0.0;
    const size_t base = (size_t)b * p.N;
    for (int k = ch.k0; k < ch.k1; ++k) {
        {
            Mat A, Bm, Qs;
            mat_load(A, p.A_aug + (base + k) * D * D, D, D, D, L);
            mat_load(Bm, p.B_aug + (base + k) * D * M, D, M, M, L);
            mat_load(Qs, p.Q_aug + (base + k) * D * D, D, D, D, L);
            if (p.rinv_step_stride) mat_load_t(RinvT, p.R_inv + (size_t)b * rinv_inst + (size_t)k * p.rinv_step_stride, M, M, M, L);
            mat_sym(Qs, L);
            // step index 0 makes stage_prefix_step START a fold (P <- g_k); any other value continues it (P <- P o g_k)
            stage_prefix_step<D, M>((from_scratch && k == ch.k0) ? 0 : 1, P, Qs, A, Bm, RinvT, L, scratch, p.jitter, p.max_tries, status);
        }
        if (!queries) continue;
        Mat P0, QTs;
        mat_load(QTs, p.QT + (base + k) * D * D, D, D, D, L);
        mat_sym(QTs, L);
        query_step<D>(P, QTs, P0, L, scratch, p.jitter, p.max_tries, status);
        double part = 0.0;                                                     // 0.5 z0^T P0 z0   (horizon_selection.py:85)
        HOP_FOR_ELEMS(I, J, s) part = fma(zr[I] * P0.v[I][J][s], zc[J][s], part);
        const double Jt = 0.5 * warp_sum(part);
        if (L.lane == 0) {
            p.J_out[(size_t)b * p.T_max + k] = Jt;
            const int t = k + 1;
            if (t >= p.T_min) am.push(Jt + wexp * (double)t, t);
        }
    }
}

// phase 1: warp c of the CTA that owns problem b
template <int D, int M>
HOP_DEVICE void scan_phase1(const SelectArgs& p, int b, int c, int C, double* smem) {
    LaneGeo L;
    L.init();
    ScanChunk ch;
    ch.set(p.T_max, C, c);
    double* scratch = smem + (size_t)c * kWarpScratch;
    double* rec = smem + ScanSmem::recs(C) + c * 4;
    PrefixL<D> P;
    mat_zero(P.eb); mat_zero(P.fb); mat_zero(P.gb);
    ArgMin am;
    am.init();
    int status = 0;
    if (!ch.empty()) scan_chunk_pass<D, M>(p, b, ch, P, true, c == 0, L, scratch, am, status);
    agg_store(smem + ScanSmem::aggs(C) + (size_t)c * ScanSmem::AGG, P.eb, P.fb, P.gb, L.lane);
    if (L.lane == 0) { rec[0] = am.best; rec[1] = (double)am.idx; rec[2] = am.nan_hit ? 1.0 : 0.0; rec[3] = (double)status; }
}

// phases 2 + 3: warps c > 0
template <int D, int M>
HOP_DEVICE void scan_phase23(const SelectArgs& p, int b, int c, int C, double* smem) {
    if (c == 0) return;
    LaneGeo L;
    L.init();
    ScanChunk ch;
    ch.set(p.T_max, C, c);
    if (ch.empty()) return;
    double* scratch = smem + (size_t)c * kWarpScratch;
    double* rec = smem + ScanSmem::recs(C) + c * 4;
    const double* aggs = smem + ScanSmem::aggs(C);
    int status = (int)rec[3];
    PrefixL<D> P;
    agg_load(aggs, P.eb, P.fb, P.gb, L.lane);                                  // chunk 0: the true prefix up to its end
    for (int j = 1; j < c; ++j) {
        Mat E2, F2, G2;
        agg_load(aggs + (size_t)j * ScanSmem::AGG, E2, F2, G2, L.lane);
        compose<D>(P, E2, F2, G2, L, scratch, p.jitter, p.max_tries, status);
    }
    ArgMin am;
    am.init();
    scan_chunk_pass<D, M>(p, b, ch, P, false, true, L, scratch, am, status);
    if (L.lane == 0) { rec[0] = am.best; rec[1] = (double)am.idx; rec[2] = am.nan_hit ? 1.0 : 0.0; rec[3] = (double)status; }
}

// merge of the per-chunk argmin records in horizon order (np.argmin: first minimum, the first NaN wins)
HOP_DEVICE void scan_finish(const SelectArgs& p, int b, int C, const double* smem) {
    ArgMin am;
    am.init();
    int status = 0;
    for (int c = 0; c < C; ++c) {
        const double* rec = smem + ScanSmem::recs(C) + c * 4;
        status |= (int)rec[3];
        if ((int)rec[1] != 0) am.push(rec[0], (int)rec[1]);
    }
    p.T_out[b] = am.idx;
    p.Jstar_out[b] = am.best;
    p.status[b] = status;
}

}}  // namespace hop::mma
