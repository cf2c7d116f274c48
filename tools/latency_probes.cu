// latency_probes.cu -- dependent-chain latencies on one SM sub-partition (clock64 around a chain in ONE warp), and the
// interference of a DMMA-streaming warp on another warp's dependent DFMA chain (same sub-partition vs another).
// Not part of the product library; results: profiles/r1_latency_probes.txt.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: dependent DFMA chain; 1: dependent DMMA chain (accumulator); 2: dependent SHFL(double) chain;
// 3: MUFU.RCP64H + 3 dependent DFMA (pivot_rcp3) chained; 4: 8 independent DFMA chains (throughput of one warp)
template <int MODE>
__global__ void k_chain(int iters, long long* cycles, double* sink) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - 1e-9, c = 1e-9, c1 = 0.5;
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = a + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) a = fma(a, b, c);
            if (MODE == 1) dmma(a, c1, b, c);
            if (MODE == 2) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
            if (MODE == 3) {
                double r;
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
                const double e = fma(-a, r, 1.0);
                const double t = fma(e, e, e);
                a = fma(r, t, r) + 1.5;
            }
            if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma(v[i], b, c);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    double s = a + c1;
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 1.2345) sink[0] = s;
}

// warp 0: dependent DFMA chain (timed).  Warps 1..: stream independent DMMAs (STREAM = 1) or idle.  With 4 warps per
// CTA warp w sits on sub-partition w % 4: `same` puts the streaming warp on warp 0's sub-partition (warp 4) or not (warp 1).
__global__ void k_interfere(int iters, int stream_warp, long long* cycles, double* sink) {
    const int warp = threadIdx.x >> 5;
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - 1e-9, c = 1e-9;
    if (warp == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) a = fma(a, b, c);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) cycles[0] = t1 - t0;
    } else if (warp == stream_warp) {
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0, d7 = 0;
        for (int it = 0; it < iters * 4; ++it) {
            dmma(d0, d1, a, b); dmma(d2, d3, a, b); dmma(d4, d5, a, b); dmma(d6, d7, a, b);
        }
        a = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    }
    if (a == 1.2345) sink[0] = a;
}

int main() {
    long long* cyc; double* sink;
    cudaMallocManaged(&cyc, 64 * sizeof(long long));
    cudaMalloc(&sink, 8);
    const int iters = 2000;
    const char* names[] = {"dependent DFMA", "dependent DMMA.8x8x4 (accumulator chain)", "dependent SHFL (64-bit = 2 x SHFL.32)",
                           "MUFU.RCP64H + 3 DFMA + DADD (pivot_rcp3 chain)", "8 independent DFMA chains, one warp (per DFMA)"};
    for (int m = 0; m < 5; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) k_chain<0><<<1, 32>>>(iters, cyc, sink);
            if (m == 1) k_chain<1><<<1, 32>>>(iters, cyc, sink);
            if (m == 2) k_chain<2><<<1, 32>>>(iters, cyc, sink);
            if (m == 3) k_chain<3><<<1, 32>>>(iters, cyc, sink);
            if (m == 4) k_chain<4><<<1, 32>>>(iters, cyc, sink);
            cudaDeviceSynchronize();
        }
        const double per = (double)cyc[0] / (iters * 16.0) / (m == 4 ? 8.0 : 1.0);
        printf("%-55s %7.2f clk per op\n", names[m], per);
    }
    for (int sw : {-1, 1, 4}) {
        for (int rep = 0; rep < 2; ++rep) { k_interfere<<<1, 256>>>(iters, sw, cyc, sink); cudaDeviceSynchronize(); }
        printf("dependent DFMA chain, %-42s %7.2f clk per DFMA\n",
               sw < 0 ? "alone" : sw == 1 ? "DMMA stream on ANOTHER sub-partition" : "DMMA stream on the SAME sub-partition",
               (double)cyc[0] / (iters * 16.0));
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
