// emul_select.cpp -- host entry points that run the HOP select kernel bodies under the SIMT
// emulator (test infrastructure; mirrors the grid/slab arithmetic of hop_select.cu).
#include <cstdlib>
#include <vector>

#include "hop_select_body.cuh"
#include "hop_select_mma_body.cuh"
#include "hop_select_pipe_body.cuh"
#include "hop_select_scan_body.cuh"
#include "hop_select_gpipe_body.cuh"
#include "hop_select_tpp_body.cuh"
#include "hop_select_epl_body.cuh"

namespace {
template <int D, int M, int G>
struct GenericJob { const hop::SelectArgs* p; int b0; double* smem; };
template <int D, int M, int G>
void generic_lane(void* a) {
    auto* j = (GenericJob<D, M, G>*)a;
    const int lane = hop::simt::lane_id();
    const int slot = lane / G;
    hop::select_generic_body<D, M, G>(*j->p, j->b0 + slot, j->smem + (size_t)slot * hop::Geo<D, M, G>::SLAB);
}
template <int D, int M, int G>
int run_generic(const hop::SelectArgs& p) {
    constexpr int GPW = 32 / G;
    std::vector<double> smem((size_t)GPW * hop::Geo<D, M, G>::SLAB, -7.0);
    for (int b0 = 0; b0 < p.B; b0 += GPW) {
        GenericJob<D, M, G> j{&p, b0, smem.data()};
        if (hop::simt::run_warp(generic_lane<D, M, G>, &j)) return -1;
    }
    return 0;
}
template <int D, int M, int G>
struct FusedJob { const hop::FusedArgs* p; int b0; double* smem; const double* cst; };
template <int D, int M, int G>
void fused_lane(void* a) {
    auto* j = (FusedJob<D, M, G>*)a;
    const int lane = hop::simt::lane_id();
    const int slot = lane / G;
    hop::select_fused_body<D, M, G>(*j->p, j->b0 + slot, j->smem + (size_t)slot * hop::Geo<D, M, G>::SLAB, j->cst);
}
template <int D, int M, int G>
int run_fused(const hop::FusedArgs& p) {
    constexpr int GPW = 32 / G;
    std::vector<double> smem((size_t)GPW * hop::Geo<D, M, G>::SLAB, -7.0);
    std::vector<double> cst(hop::FusedConst<D, M>::SIZE, 0.0);
    hop::fused_const_fill<D, M>(p, cst.data(), 0, 1);
    for (int b0 = 0; b0 < p.B; b0 += GPW) {
        FusedJob<D, M, G> j{&p, b0, smem.data(), cst.data()};
        if (hop::simt::run_warp(fused_lane<D, M, G>, &j)) return -1;
    }
    return 0;
}
// ---- one-problem-per-warp DMMA bodies
template <int D, int M>
struct MmaGenericJob { const hop::SelectArgs* p; int b; double* scratch; };
template <int D, int M>
void mma_generic_lane(void* a) {
    auto* j = (MmaGenericJob<D, M>*)a;
    hop::mma::select_generic_body<D, M>(*j->p, j->b, j->scratch);
}
template <int D, int M>
int run_generic_mma(const hop::SelectArgs& p) {
    std::vector<double> scratch(hop::mma::kWarpScratch, -7.0);
    for (int b = 0; b < p.B; ++b) {
        MmaGenericJob<D, M> j{&p, b, scratch.data()};
        if (hop::simt::run_warp(mma_generic_lane<D, M>, &j)) return -1;
    }
    return 0;
}
template <int D, int M>
struct MmaFusedJob { const hop::FusedArgs* p; int b; double* scratch; double* cst; };
template <int D, int M, int MODE>
void mma_fused_lane(void* a) {
    auto* j = (MmaFusedJob<D, M>*)a;
    hop::mma::select_fused_body<D, M, MODE>(*j->p, j->b, j->scratch, j->cst);
}
template <int D, int M, int LOOPED>
void mma_pipe_lane(void* a) {   // mirrors k_select_fused_mma<.., MODE = 2 | 3 | 4> (SCHED 0 | 1 | 2)
    auto* j = (MmaFusedJob<D, M>*)a;
    if (hop::mma::select_fused_pipe_body<D, M, LOOPED>(*j->p, j->b, j->scratch, j->cst)) return;
    hop::mma::select_fused_body<D, M, 1>(*j->p, j->b, j->scratch, j->cst);
}
template <int D, int M>
void mma_fast_const_lane(void* a) {
    auto* j = (MmaFusedJob<D, M>*)a;
    hop::mma::fast_const_fill_warp<D, M>(*j->p, j->cst, j->scratch);
}
template <int D, int M>
int run_fused_mma(const hop::FusedArgs& p) {
    std::vector<double> scratch(hop::mma::kWarpScratch, -7.0);
    std::vector<double> cst(hop::mma::PipeConst<D, M>::SIZE, 0.0);
    hop::fused_const_fill<D, M>(p, cst.data(), 0, 1);
    if (p.mode >= 1) {
        MmaFusedJob<D, M> j{&p, 0, scratch.data(), cst.data()};
        if (hop::simt::run_warp(mma_fast_const_lane<D, M>, &j)) return -1;
    }
    if (p.mode >= 2) hop::mma::pipe_const_fill<D, M>(cst.data(), 0, 1);
    for (int b = 0; b < p.B; ++b) {
        MmaFusedJob<D, M> j{&p, b, scratch.data(), cst.data()};
        auto fn = p.mode == 5 ? mma_pipe_lane<D, M, 3> : p.mode == 4 ? mma_pipe_lane<D, M, 2> : p.mode == 3 ? mma_pipe_lane<D, M, 1> : p.mode == 2 ? mma_pipe_lane<D, M, 0>
                                : (p.mode == 1 ? mma_fused_lane<D, M, 1> : mma_fused_lane<D, M, 0>);
        if (hop::simt::run_warp(fn, &j)) return -1;
    }
    return 0;
}
}  // namespace

namespace {
struct InvJob { const double* A; double* X; int* status; int* ok_first; double* scratch; };
template <int D>
void inv_lane(void* a) {
    auto* j = (InvJob*)a;
    hop::mma::LaneGeo L;
    L.init();
    hop::mma::Mat S, out;
    hop::mma::mat_load(S, j->A, D, D, D, L);
    hop::mma::mat_sym(S, L);
    int st = 0;
    hop::mma::chol_inv<D>(S, out, L, j->scratch, 1e-9, 8, st);
    HOP_FOR_ELEMS(I, J, s) {
        const int R = L.row(I), C = L.col(J, s);
        if (R < D && C < D) j->X[R * D + C] = out.v[I][J][s];
    }
    if (L.lane == 0) *j->status = st;
}
}  // namespace
extern "C" int emul_chol_inv_mma(int d, const double* A, double* X, int* status) {
    std::vector<double> scratch(hop::mma::kWarpScratch, 0.0);
    InvJob j{A, X, status, nullptr, scratch.data()};
    if (d == 13) return hop::simt::run_warp(inv_lane<13>, &j);
    if (d == 12) return hop::simt::run_warp(inv_lane<12>, &j);
    if (d == 4) return hop::simt::run_warp(inv_lane<4>, &j);
    return -2;
}

// ---- HOP_MODE_SCAN: the CTA's warps run one after the other; the two __syncthreads() are the phase boundaries
namespace {
struct ScanJob { const hop::SelectArgs* p; int b, c; double* smem; };
template <int D, int M>
void scan_p1_lane(void* a) { auto* j = (ScanJob*)a; hop::mma::scan_phase1<D, M>(*j->p, j->b, j->c, hop::mma::kScanWarps, j->smem); }
template <int D, int M>
void scan_p23_lane(void* a) { auto* j = (ScanJob*)a; hop::mma::scan_phase23<D, M>(*j->p, j->b, j->c, hop::mma::kScanWarps, j->smem); }
template <int D, int M>
int run_generic_scan(const hop::SelectArgs& p) {
    constexpr int C = hop::mma::kScanWarps;
    std::vector<double> smem(hop::mma::ScanSmem::size(C), -7.0);
    for (int b = 0; b < p.B; ++b) {
        for (int c = 0; c < C; ++c) { ScanJob j{&p, b, c, smem.data()}; if (hop::simt::run_warp(scan_p1_lane<D, M>, &j)) return -1; }
        for (int c = 0; c < C; ++c) { ScanJob j{&p, b, c, smem.data()}; if (hop::simt::run_warp(scan_p23_lane<D, M>, &j)) return -1; }
        hop::mma::scan_finish(p, b, C, smem.data());
    }
    return 0;
}
}  // namespace
extern "C" int emul_select_generic_scan(int d, int m, const hop::SelectArgs* p) {
    if (d == 12 && m == 4) return run_generic_scan<12, 4>(*p);
    if (d == 13 && m == 4) return run_generic_scan<13, 4>(*p);
    return -2;
}

// ---- pipelined LQR-boundary body (mirrors k_select_generic_pipe: pipelined sweep, sequential body as cold path)
namespace {
template <int D, int M>
void gpipe_lane(void* a) {
    auto* j = (MmaGenericJob<D, M>*)a;
    if (hop::mma::select_generic_pipe_body<D, M>(*j->p, j->b, j->scratch)) return;
    hop::mma::select_generic_body<D, M>(*j->p, j->b, j->scratch);
}
template <int D, int M>
int run_generic_pipe(const hop::SelectArgs& p) {
    std::vector<double> slab(hop::mma::GpipeSlab<D, M>::SIZE, -7.0);
    for (int b = 0; b < p.B; ++b) {
        MmaGenericJob<D, M> j{&p, b, slab.data()};
        if (hop::simt::run_warp(gpipe_lane<D, M>, &j)) return -1;
    }
    return 0;
}
}  // namespace
extern "C" int emul_select_generic_pipe(int d, int m, const hop::SelectArgs* p) {
    if (d == 12 && m == 4) return run_generic_pipe<12, 4>(*p);
    if (d == 13 && m == 4) return run_generic_pipe<13, 4>(*p);
    return -2;
}

extern "C" int emul_select_generic_mma(int d, int m, const hop::SelectArgs* p) {
    if (d == 12 && m == 4) return run_generic_mma<12, 4>(*p);
    if (d == 13 && m == 4) return run_generic_mma<13, 4>(*p);
    return -2;
}
extern "C" int emul_select_fused_mma(int n, int m, const hop::FusedArgs* p) {
    if (n == 12 && m == 4) return run_fused_mma<13, 4>(*p);
    return -2;
}

extern "C" int emul_select_generic(int d, int m, const hop::SelectArgs* p) {
    if (d == 3 && m == 1) return run_generic<3, 1, 4>(*p);
    if (d == 4 && m == 2) return run_generic<4, 2, 4>(*p);
    if (d == 5 && m == 1) return run_generic<5, 1, 8>(*p);
    if (d == 12 && m == 4) return run_generic<12, 4, 16>(*p);
    if (d == 13 && m == 4) return run_generic<13, 4, 16>(*p);
    return -2;
}
extern "C" int emul_select_fused(int n, int m, const hop::FusedArgs* p) {
    if (n == 2 && m == 1) return run_fused<3, 1, 4>(*p);
    if (n == 4 && m == 1) return run_fused<5, 1, 8>(*p);
    if (n == 12 && m == 4) return run_fused<13, 4, 16>(*p);
    return -2;
}

// ---- thread-per-problem LQR-boundary body (hop_select_tpp_body.cuh): no cross-lane traffic, so a plain loop
namespace {
template <int D, int M>
int run_generic_tpp(const hop::SelectArgs& p) {
    for (int b = 0; b < p.B; ++b) {
        hop::tpp::GlobalFeed<D, M> feed;
        feed.init(p, b);
        hop::tpp::select_generic_tpp_body<D, M>(p, b, true, feed);
    }
    return 0;
}
}  // namespace
extern "C" int emul_select_generic_tpp(int d, int m, const hop::SelectArgs* p) {
    if (d == 3 && m == 1) return run_generic_tpp<3, 1>(*p);
    if (d == 4 && m == 2) return run_generic_tpp<4, 2>(*p);
    if (d == 5 && m == 1) return run_generic_tpp<5, 1>(*p);
    return -2;
}

// ---- element-per-lane fused body of the small systems (hop_select_epl_body.cuh): one warp per problem
namespace {
template <int D, int M>
struct EplJob { const hop::FusedArgs* p; int b; double* scratch; const double* cst; };
template <int D, int M>
void epl_lane(void* a) {
    auto* j = (EplJob<D, M>*)a;
    hop::epl::select_fused_epl_body<D, M>(*j->p, j->b, j->scratch, j->cst);
}
template <int D, int M>
int run_fused_epl(const hop::FusedArgs& p) {
    std::vector<double> scratch(2 * D * ((D + 1) & ~1), -7.0);
    std::vector<double> cst(hop::FusedConst<D, M>::SIZE, 0.0);
    hop::fused_const_fill<D, M>(p, cst.data(), 0, 1);
    for (int b = 0; b < p.B; ++b) {
        EplJob<D, M> j{&p, b, scratch.data(), cst.data()};
        if (hop::simt::run_warp(epl_lane<D, M>, &j)) return -1;
    }
    return 0;
}
}  // namespace
extern "C" int emul_select_fused_epl(int n, int m, const hop::FusedArgs* p) {
    if (n == 2 && m == 1) return run_fused_epl<3, 1>(*p);
    if (n == 4 && m == 1) return run_fused_epl<5, 1>(*p);
    return -2;
}

// ---- HOP_MODE_EXACT / HOP_MODE_FP32: reference-operation-order body (hop_select_ref_body.cuh), one warp per problem
#include "hop_select_ref_body.cuh"
namespace {
template <typename R>
struct RefGenericJob { const hop::SelectArgs* p; int d, m, b; R* slab; };
template <typename R>
void ref_generic_lane(void* a) {
    auto* j = (RefGenericJob<R>*)a;
    hop::ref::select_generic_body<R>(*j->p, j->d, j->m, j->b, j->slab);
}
template <typename R>
int run_generic_ref(int d, int m, const hop::SelectArgs& p) {
    std::vector<R> slab(hop::ref::Layout::make(d, m).size, (R)-7.0);
    for (int b = 0; b < p.B; ++b) {
        RefGenericJob<R> j{&p, d, m, b, slab.data()};
        if (hop::simt::run_warp(ref_generic_lane<R>, &j)) return -1;
    }
    return 0;
}
struct RefFusedJob { const hop::FusedArgs* p; int n, m, b; double* slab; const double* cst; };
void ref_fused_lane(void* a) {
    auto* j = (RefFusedJob*)a;
    hop::ref::select_fused_body(*j->p, j->n, j->m, j->b, j->slab, j->cst);
}
}  // namespace
extern "C" int emul_select_generic_ref(int d, int m, const hop::SelectArgs* p, int fp32) {
    if (d < 1 || d > hop::ref::kMaxD || m < 1 || m > hop::ref::kMaxD) return -2;
    return fp32 ? run_generic_ref<float>(d, m, *p) : run_generic_ref<double>(d, m, *p);
}
extern "C" int emul_select_fused_ref(int n, int m, const hop::FusedArgs* p) {
    if (n < 1 || n + 1 > hop::ref::kMaxD || m < 1 || m > n + 1) return -2;
    std::vector<double> slab(hop::ref::Layout::make(n + 1, m).size, -7.0);
    std::vector<double> cst(hop::ref::FusedCst::make(n, m).size, 0.0);
    hop::ref::fused_cst_fill(*p, n, m, cst.data(), 0, 1);
    for (int b = 0; b < p->B; ++b) {
        if (p->skip && p->skip[b]) continue;
        RefFusedJob j{p, n, m, b, slab.data(), cst.data()};
        if (hop::simt::run_warp(ref_fused_lane, &j)) return -1;
    }
    return 0;
}
