// hop_simt.cuh -- the handful of warp-level primitives the HOP kernels use.
//
// Device build (nvcc, sm_100a): thin wrappers over the CUDA warp intrinsics.
// Host-emulation build (-DHOP_HOST_EMUL, g++ only, used by tests/emul/): the same names are
// provided by a fiber scheduler (tests/emul/simt_emul.h) that runs the 32 lanes of a warp as
// cooperative fibers, so the *same kernel source* can be exercised in a container without a GPU.
// The emulation is test infrastructure: it is never compiled into libhop_b200.so.
#pragma once

#ifdef HOP_HOST_EMUL
#include "simt_emul.h"   // defines hop::simt::{lane_id, sync, shfl, shfl_xor, ballot, all} + HOP_DEVICE
#else
#include <cuda_runtime.h>
#define HOP_DEVICE __device__ __forceinline__
#define HOP_DEVICE_NOINLINE __device__ __noinline__
namespace hop { namespace simt {
HOP_DEVICE int lane_id() { return (int)(threadIdx.x & 31u); }
HOP_DEVICE void sync() { __syncwarp(); }
HOP_DEVICE double shfl(double v, int src_in_group, int width) { return __shfl_sync(0xffffffffu, v, src_in_group, width); }
HOP_DEVICE double shfl_xor(double v, int mask, int width) { return __shfl_xor_sync(0xffffffffu, v, mask, width); }
HOP_DEVICE unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
HOP_DEVICE bool all(bool p) { return __all_sync(0xffffffffu, p) != 0; }
// D(8x8) += A(8x4) * B(4x8) in fp64 on the tensor pipe (SASS: DMMA.8x8x4).  Fragments (g = lane>>2,
// t = lane&3): a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
HOP_DEVICE void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
}}  // namespace hop::simt
#endif
