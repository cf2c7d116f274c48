// hop_select_ref.cu -- __global__ wrappers + launchers of HOP_MODE_EXACT (reference operation order,
// hop_select_ref_body.cuh) and HOP_MODE_FP32 (the same sweep in single precision, LQR-boundary entry only).
//
// Grid: one warp per problem, no CTA barrier inside the sweep.  Dynamic shared memory = one slab per warp (ref::Layout,
// 19.3 KB at d = 13 in fp64) [+ the CTA-wide case constants of the fused form].  Warps per CTA are chosen at launch: as
// many as the 227 KB of an SM hold (11 at d = 13) when the batch fills the machine, fewer for small batches so that the
// problems spread over the SMs.
#include "hop_common.cuh"
#include "hop_select_ref_body.cuh"
#include "../../include/hop_b200.h"

namespace hop {

constexpr int kRefMaxWarps = 16;
constexpr size_t kSmemPerSm = 227 * 1024;

// warps per CTA: fill one SM's shared memory when there is at least one such CTA per SM, else spread the batch
static int ref_warps(int B, size_t slab_bytes, size_t cst_bytes) {
    int fit = (int)((kSmemPerSm - cst_bytes - 1024) / slab_bytes);
    fit = fit < 1 ? 1 : (fit > kRefMaxWarps ? kRefMaxWarps : fit);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((long)B >= (long)fit * sms) return fit;
    const int spread = (B + sms - 1) / sms;
    return spread < 1 ? 1 : (spread > fit ? fit : spread);
}

template <typename R>
__global__ void __launch_bounds__(kRefMaxWarps * 32) k_select_ref_generic(const SelectArgs p, int d, int m, int slab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= p.B) return;
    ref::select_generic_body<R>(p, d, m, b, reinterpret_cast<R*>(smem_raw) + (size_t)warp * slab);
}

__global__ void __launch_bounds__(kRefMaxWarps * 32) k_select_ref_fused(const FusedArgs p, int n, int m, int slab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* cst = smem + (size_t)(blockDim.x >> 5) * slab;
    ref::fused_cst_fill(p, n, m, cst, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= p.B || (p.skip && p.skip[b])) return;
    ref::select_fused_body(p, n, m, b, smem + (size_t)warp * slab, cst);
}

template <typename R>
static int launch_ref_generic(int d, int m, const SelectArgs& p, cudaStream_t st) {
    const int slab = (ref::Layout::make(d, m).size + 1) & ~1;
    const int warps = ref_warps(p.B, sizeof(R) * (size_t)slab, 0);
    const size_t smem = sizeof(R) * (size_t)warps * slab;
    cudaError_t e = cudaFuncSetAttribute(k_select_ref_generic<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return report_cuda(e, "cudaFuncSetAttribute(k_select_ref_generic)");
    k_select_ref_generic<R><<<(p.B + warps - 1) / warps, warps * 32, smem, st>>>(p, d, m, slab);
    return check_launch("k_select_ref_generic");
}

int dispatch_select_ref_generic(int d, int m, bool fp32, const SelectArgs& p, cudaStream_t st) {
    if (d < 1 || d > ref::kMaxD || m < 1 || m > ref::kMaxD) {
        set_last_error("hop_select_f64: HOP_MODE_EXACT / HOP_MODE_FP32 need 1 <= d, m <= 16");
        return HOP_E_UNSUPPORTED_DIMS;
    }
    return fp32 ? launch_ref_generic<float>(d, m, p, st) : launch_ref_generic<double>(d, m, p, st);
}

int dispatch_select_ref_fused(int n, int m, const FusedArgs& p, cudaStream_t st) {
    if (n < 1 || n + 1 > ref::kMaxD || m < 1 || m > n + 1) {
        set_last_error("hop_select_fused_f64: HOP_MODE_EXACT needs 1 <= n <= 15 and 1 <= m <= n + 1");
        return HOP_E_UNSUPPORTED_DIMS;
    }
    const int slab = (ref::Layout::make(n + 1, m).size + 1) & ~1;
    const size_t cst_bytes = sizeof(double) * (size_t)ref::FusedCst::make(n, m).size;
    const int warps = ref_warps(p.B, sizeof(double) * (size_t)slab, cst_bytes);
    const size_t smem = sizeof(double) * (size_t)warps * slab + cst_bytes;
    cudaError_t e = cudaFuncSetAttribute(k_select_ref_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return report_cuda(e, "cudaFuncSetAttribute(k_select_ref_fused)");
    k_select_ref_fused<<<(p.B + warps - 1) / warps, warps * 32, smem, st>>>(p, n, m, slab);
    return check_launch("k_select_ref_fused");
}

}  // namespace hop
