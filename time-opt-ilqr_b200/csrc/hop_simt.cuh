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
namespace hop { namespace simt {
// (g++ on x86-64 without -mfma never contracts a * b + c)
inline double mul_rn(double a, double b) { return a * b; }
inline double add_rn(double a, double b) { return a + b; }
inline double rcp_newton(double p) { return 1.0 / p; }
}}  // namespace hop::simt
#else
#include <cuda_runtime.h>
#define HOP_DEVICE __device__ __forceinline__
#define HOP_DEVICE_NOINLINE __device__ __noinline__
#define HOP_HD __host__ __device__
namespace hop { namespace simt {
HOP_DEVICE int lane_id() { return (int)(threadIdx.x & 31u); }
HOP_DEVICE void sync() { __syncwarp(); }
HOP_DEVICE double shfl(double v, int src_in_group, int width) { return __shfl_sync(0xffffffffu, v, src_in_group, width); }
HOP_DEVICE double shfl_xor(double v, int mask, int width) { return __shfl_xor_sync(0xffffffffu, v, mask, width); }
HOP_DEVICE unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
HOP_DEVICE bool all(bool p) { return __all_sync(0xffffffffu, p) != 0; }
// a product / a sum that must NOT be contracted into an FMA with its neighbours (numpy evaluates them separately)
HOP_DEVICE double mul_rn(double a, double b) { return __dmul_rn(a, b); }
HOP_DEVICE double add_rn(double a, double b) { return __dadd_rn(a, b); }
// 1/p for a Gauss-Jordan pivot: MUFU.RCP64H seed (rel. error <= 2^-23) + two Newton steps (<= 1 ulp).  The IEEE-correct
// `1.0 / p` is a ~25-instruction dependent sequence with a slow-path test and sits on the critical path of every
// elimination step (25 of them per horizon step in the d <= 5 kernels); the host emulation uses the exact quotient.
HOP_DEVICE double rcp_newton(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    double e = fma(-p, r, 1.0);
    r = fma(r, e, r);
    e = fma(-p, r, 1.0);
    r = fma(r, e, r);
    return r;
}
// D(8x8) += A(8x4) * B(4x8) in fp64 on the tensor pipe (SASS: DMMA.8x8x4).  Fragments (g = lane>>2,
// t = lane&3): a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
HOP_DEVICE void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared, completion on an mbarrier --------
HOP_DEVICE unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
HOP_DEVICE void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
HOP_DEVICE void mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
HOP_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
HOP_DEVICE void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned
HOP_DEVICE void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// ---- cp.async (LDGSTS) in 8-byte granules: global -> shared without a register round trip, any 8-byte alignment
HOP_DEVICE void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
HOP_DEVICE void cp_async16(void* dst, const void* src) {   // dst and src 16-byte aligned
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
HOP_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
HOP_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
HOP_DEVICE void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
}}  // namespace hop::simt
#endif
