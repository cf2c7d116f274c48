// simt_emul.h -- TEST INFRASTRUCTURE: run the HOP kernel bodies on the host.
//
// The 32 lanes of a warp are cooperative fibers (ucontext); sync()/shfl()/ballot()/dmma() are barriers
// among them.  This lets `pytest -m "not gpu"` execute the very same kernel source
// (time-opt-ilqr_b200/csrc/hop_select_body.cuh, compiled with -DHOP_HOST_EMUL by g++) in a container
// without a GPU.  It is never linked into the product library and is not a CPU fallback: the
// product's loader refuses to run without CUDA.
#pragma once
#include <cmath>
#include <cstddef>

#define HOP_DEVICE inline
#define HOP_DEVICE_NOINLINE inline
#define HOP_HD

using std::isfinite;

namespace hop { namespace simt {
int lane_id();
void sync();
double shfl(double v, int src_in_group, int width);
double shfl_xor(double v, int mask, int width);
unsigned ballot(bool p);
bool all(bool p);
void dmma(double& c0, double& c1, double a, double b);   // mma.m8n8k4.f64 semantics
// Run fn(arg) once per lane of one emulated warp.  Returns 0, or -1 if lanes exited non-uniformly
// (some lane still waiting at a barrier when another finished), which on a GPU would be a hang.
int run_warp(void (*fn)(void*), void* arg);
// bulk-copy / mbarrier stand-ins: the copy completes at issue time, waits are no-ops
inline void mbar_init(unsigned long long*, unsigned) {}
inline void mbar_fence_init() {}
inline void mbar_inval(unsigned long long*) {}
inline void mbar_expect_tx(unsigned long long*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) {
    const char* s = (const char*)src; char* d = (char*)dst;
    for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
}
inline void mbar_wait(unsigned long long*, unsigned) {}
// cp.async stand-ins: the copy completes at issue time
inline void cp_async8(void* dst, const void* src) { *(double*)dst = *(const double*)src; }
inline void cp_async_commit() {}
template <int N_>
inline void cp_async_wait() {}
}}  // namespace hop::simt
