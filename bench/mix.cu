// mix.cu -- issue-model probe: per-iteration NL LOP3 + NP POPC + NI IMAD + NM VIMNMX on independent registers,
// 2 CTAs x 256 threads per SM.  Prints SM cycles per iteration per SM-wide warp set and the implied lanes/clk.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
constexpr int ITERS = 2048;
template <int NL, int NP, int NI, int NM>
__global__ void __launch_bounds__(256, 2) mix_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t l[16], p[8], m[8], v[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) l[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = l[i] ^ 0x1234567u; m[i] = l[i] + 77u * i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = l[i] * 3u;
    const uint32_t b = seed ^ 0xabcdefu, c = seed + 77u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {   // 4 repetitions per loop trip to amortise the branch
#pragma unroll
            for (int i = 0; i < NL; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i % 16]) : "r"(b), "r"(c));
#pragma unroll
            for (int i = 0; i < NP; ++i) asm volatile("popc.b32 %0, %0;" : "+r"(p[i % 8]));
#pragma unroll
            for (int i = 0; i < NI; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[i % 8]) : "r"(b | 3u), "r"(c));
#pragma unroll
            for (int i = 0; i < NM; ++i) asm volatile("min.u32 %0, %0, %1;" : "+r"(v[i % 4]) : "r"(b + i));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= l[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= p[i] ^ m[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int NL, int NP, int NI, int NM>
static void run(int sms, uint32_t* out, long long* clk) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    mix_kernel<NL, NP, NI, NM><<<2 * sms, 256>>>(out, 123u, clk);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); mix_kernel<NL, NP, NI, NM><<<2 * sms, 256>>>(out, 123u, clk); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    long long h[1024]; cudaMemcpy(h, clk, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 2 * sms; ++i) cyc += (double)h[i]; cyc /= 2 * sms;
    const double iters = (double)ITERS * 4;                 // per thread
    const double lane_iters_sm = iters * 512;               // 2 CTAs x 256 threads per SM
    // wall-clock cycles at 1.965 GHz per (lane-iteration per SM)
    const double wall_clk = best * 1e-3 * 1.965e9 / lane_iters_sm;
    printf("{\"NL\": %d, \"NP\": %d, \"NI\": %d, \"NM\": %d, \"clk64_per_lane_iter\": %.4f, \"wall_clk_per_lane_iter\": %.4f, \"ms\": %.4f}\n",
           NL, NP, NI, NM, cyc / lane_iters_sm, wall_clk, best);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 4 * 512 * sms); cudaMalloc(&clk, 8 * 2 * sms);
    run<16, 0, 0, 0>(sms, out, clk);
    run<0, 4, 0, 0>(sms, out, clk);
    run<0, 0, 6, 0>(sms, out, clk);
    run<16, 4, 0, 0>(sms, out, clk);
    run<16, 0, 6, 0>(sms, out, clk);
    run<16, 4, 6, 0>(sms, out, clk);
    run<16, 4, 6, 2>(sms, out, clk);
    run<14, 5, 6, 2>(sms, out, clk);
    run<12, 6, 6, 2>(sms, out, clk);
    run<8, 8, 6, 2>(sms, out, clk);
    run<16, 2, 6, 2>(sms, out, clk);
    run<0, 4, 6, 0>(sms, out, clk);
    run<16, 4, 12, 2>(sms, out, clk);
    return 0;
}
