// Microbenchmark: fp64 pipe behaviour on B200 (latency, per-warp / per-SM throughput, lane occupancy).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool F32>
__global__ void k(double* out, long long* cyc, int iters, int active_lanes, int active_warp_stride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool on = lane < active_lanes && (warp % active_warp_stride) == 0;
    double a[CHAINS]; float f[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { a[c] = 1.0 + c + threadIdx.x; f[c] = 1.f + c + threadIdx.x; }
    const double m = 1.0000001, ad = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (on) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (F32) f[c] = __fmaf_rn(f[c], 1.0000001f, 1e-9f);
                else a[c] = __fma_rn(a[c], m, ad);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += a[c] + f[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS, bool F32>
void run(const char* name, int threads, int lanes, int wstride, double* out, long long* cyc) {
    const int iters = 2000;
    k<CHAINS, F32><<<1, threads>>>(out, cyc, iters, lanes, wstride);
    cudaDeviceSynchronize();
    k<CHAINS, F32><<<1, threads>>>(out, cyc, iters, lanes, wstride);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const int warps = (threads / 32 + wstride - 1) / wstride;
    printf("%-6s threads %4d active_lanes %2d warp_stride %d chains %d: %8lld cycles, %.2f cyc per warp-FMA (per warp), %.2f cyc per warp-FMA (SM aggregate)\n",
           name, threads, lanes, wstride, CHAINS, c, (double)c / (iters * CHAINS), (double)c / (iters * CHAINS * warps));
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 12);
    run<1, false>("f64", 32, 32, 1, out, cyc);      // dependent latency
    run<8, false>("f64", 32, 32, 1, out, cyc);      // one warp, 8 independent chains
    run<8, false>("f64", 32, 1, 1, out, cyc);       // one lane
    run<8, false>("f64", 64, 32, 1, out, cyc);      // two warps (different SMSPs)
    run<8, false>("f64", 128, 32, 1, out, cyc);     // four warps, one per SMSP
    run<8, false>("f64", 128, 1, 1, out, cyc);      // four warps, 1 lane each
    run<8, false>("f64", 256, 32, 4, out, cyc);     // warps 0,4 (same SMSP)
    run<8, false>("f64", 512, 32, 1, out, cyc);     // 16 warps
    run<1, true>("f32", 32, 32, 1, out, cyc);
    run<8, true>("f32", 128, 32, 1, out, cyc);
    return 0;
}
