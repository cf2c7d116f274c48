// hop_select_epl.cu -- __global__ wrapper + launcher of the element-per-lane fused selection kernel for small systems
// (hop_select_epl_body.cuh).  Grid: one warp per problem, kEplWarps warps per CTA (small CTAs spread a small batch
// over many SMs: these launches are latency bound).  Dynamic shared memory = per-warp LU scratch + the CTA constants.
#include <cstdlib>

#include "hop_common.cuh"
#include "hop_select_epl_body.cuh"
#include "../../include/hop_b200.h"

namespace hop {

constexpr int kEplWarps = 2;
template <int D>
struct EplScratch { static constexpr int SIZE = 2 * D * ((D + 1) & ~1); };   // LU fallback: two D x DP buffers per warp

template <int D, int M>
__global__ void __launch_bounds__(kEplWarps * 32) k_select_fused_epl(const FusedArgs p) {
    extern __shared__ __align__(16) double smem[];
    double* cst = smem + kEplWarps * EplScratch<D>::SIZE;
    fused_const_fill<D, M>(p, cst, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kEplWarps + warp;
    if (b >= p.B || (p.skip && p.skip[b])) return;
    epl::select_fused_epl_body<D, M>(p, b, smem + warp * EplScratch<D>::SIZE, cst);
}

template <int D, int M>
static int launch_fused_epl(const FusedArgs& p, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)kEplWarps * EplScratch<D>::SIZE + FusedConst<D, M>::SIZE);
    k_select_fused_epl<D, M><<<(p.B + kEplWarps - 1) / kEplWarps, kEplWarps * 32, smem, st>>>(p);
    return check_launch("k_select_fused_epl");
}

// fused selection kernel of the small systems: 0 element per lane (default), 1 lane group per problem (test / A-B hook;
// identical bits)
int g_fused_small_variant = getenv("HOP_FUSED_LANES") ? atoi(getenv("HOP_FUSED_LANES")) : 0;

// A warp per problem only pays while the batch leaves the machine empty (measured on B200, x0 -> T* pipeline, element per
// lane vs lane group: Segway B = 25 0.86 vs 1.30 ms, Cartpole B = 25 1.42 vs 2.13 ms, DI B = 25 0.28 vs 0.40 ms; at
// B = 4096 the lane-group kernel, which packs 4-8 problems into a warp, is ahead: 2.15 vs 1.88 ms).  $HOP_EPL_MAX_BATCH overrides.
int dispatch_select_fused_epl(int n, int m, const FusedArgs& p, cudaStream_t st) {
    static const long max_batch = getenv("HOP_EPL_MAX_BATCH") ? atol(getenv("HOP_EPL_MAX_BATCH")) : 1024;
    if (g_fused_small_variant != 0 || p.B > max_batch) return HOP_E_UNSUPPORTED_DIMS;
    if (n == 2 && m == 1) return launch_fused_epl<3, 1>(p, st);
    if (n == 4 && m == 1) return launch_fused_epl<5, 1>(p, st);
    return HOP_E_UNSUPPORTED_DIMS;
}

}  // namespace hop
