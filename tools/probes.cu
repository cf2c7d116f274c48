// probes.cu -- B200 micro-benchmarks that decide the kernel mapping (not part of the product library).
//   lds128b / lds64b / lds32b : shared-memory loads where each half-warp reads ONE address (broadcast)
//   lds128d                   : LDS.128 with 32 distinct addresses (conflict-free)
//   dmma                      : mma.sync.aligned.m8n8k4.f64 throughput
//   shfl                      : __shfl_sync throughput
//   dfma                      : DFMA throughput
// Each kernel is timed with CUDA events at 1..16 warps per SM-sub-partition-equivalent occupancy.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_lds(int iters, double* sink) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = lane >> 4;
    double acc0 = 0, acc1 = 0;
    int off = (threadIdx.x >> 5) * 64;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int base = (off + u * 32) & 2047;
            if (MODE == 0) {  // LDS.128 broadcast per half-warp (two addresses 16 B apart-of-bank-window)
                const double2 v = *reinterpret_cast<const double2*>(&sm[base + grp * 2]);
                acc0 += v.x; acc1 += v.y;
            } else if (MODE == 1) {  // LDS.64 broadcast per half-warp
                acc0 += sm[base + grp * 2];
            } else if (MODE == 2) {  // LDS.32 broadcast per half-warp
                acc0 += (double)reinterpret_cast<const float*>(sm)[2 * base + grp * 2];
            } else if (MODE == 3) {  // LDS.128 distinct (32 lanes x 16 B = 512 B)
                const double2 v = *reinterpret_cast<const double2*>(&sm[base + lane * 2]);
                acc0 += v.x; acc1 += v.y;
            } else {  // LDS.64 distinct (256 B)
                acc0 += sm[base + lane];
            }
        }
        off += 16;
    }
    if (acc0 + acc1 == 1.2345) sink[0] = acc0;
}

__global__ void k_dmma(int iters, double* sink) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
        }
    }
    if (c0[0] + c1[0] + c2[1] + c3[1] == 1.2345) sink[0] = c0[0];
}

__global__ void k_shfl(int iters, double* sink) {
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v0 = __shfl_sync(0xffffffffu, v0, (u + 1) & 15, 16);
            v1 = __shfl_sync(0xffffffffu, v1, (u + 3) & 15, 16);
            v2 = __shfl_sync(0xffffffffu, v2, (u + 5) & 15, 16);
            v3 = __shfl_sync(0xffffffffu, v3, (u + 7) & 15, 16);
        }
    }
    if (v0 + v1 + v2 + v3 == 1.2345) sink[0] = v0;
}

__global__ void k_dfma(int iters, double* sink) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) sink[0] = a0;
}

template <typename F>
static void timeit(const char* name, F launch, double ops_per_thread_iter, int iters, int blocks, int threads, int nsm, double clk_ghz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(iters / 4 + 1);
    cudaEventRecord(e0);
    launch(iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = ops_per_thread_iter * iters * (double)blocks * threads / 32.0;
    const double cyc = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-10s warps/SM=%3d  ms=%8.3f  warp-instr/clk/SM=%7.3f  (clk/warp-instr/SM=%6.3f)\n", name,
           blocks * threads / 32 / nsm, ms, warp_instr / cyc / nsm, cyc * nsm / warp_instr);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s  SMs=%d  clock=%.3f GHz (max; rates below assume it)\n", p.name, nsm, ghz);
    double* sink;
    cudaMalloc(&sink, 8);
    for (int wps : {4, 8, 16, 32}) {
        const int threads = 32 * wps / 1, blocks = nsm;   // one CTA per SM with wps warps
        const int th = threads > 1024 ? 1024 : threads, bl = threads > 1024 ? blocks * (threads / 1024) : blocks;
        const int it = 2000;
        timeit("lds128b", [&](int n) { k_lds<0><<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("lds64b", [&](int n) { k_lds<1><<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("lds32b", [&](int n) { k_lds<2><<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("lds128d", [&](int n) { k_lds<3><<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("lds64d", [&](int n) { k_lds<4><<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("dmma884", [&](int n) { k_dmma<<<bl, th>>>(n, sink); }, 16, it, bl, th, nsm, ghz);
        timeit("shfl64", [&](int n) { k_shfl<<<bl, th>>>(n, sink); }, 32, it, bl, th, nsm, ghz);
        timeit("dfma", [&](int n) { k_dfma<<<bl, th>>>(n, sink); }, 32, it, bl, th, nsm, ghz);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
