// peaks.cu -- measured pipe rates that set the roofline denominators of the Hamming kernels
// (SURVEY 7 step 0): POPC, LOP3, IMAD, VIMNMX issue rates per SM per clock (register-only loops), the
// register-only CSA Hamming inner loop, and an HBM copy.  Prints one JSON object.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;

// explicit variants (8 independent chains per thread)
__global__ void __launch_bounds__(1024) popc_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024) lop3_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t b = seed ^ 0xabcdef, c = seed + 77;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(b), "r"(c));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024) imad_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t b = seed | 3, c = seed + 77;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024) min_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t b = seed ^ 0xabcdef;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b + i + it));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// register-only Hamming inner loops: NPOPC = 8 (plain) or 4 (CSA), 4 queries per thread
__device__ __forceinline__ uint32_t x3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t mj(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256, 2) ham_kernel(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t q[4][8], best[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        best[j] = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 8; ++i) q[j][i] = seed * (threadIdx.x + 1 + j * 977) + i * 0x9e3779b9u;
    }
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = seed + i * 0x85ebca6bu;
    long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = t[i] * 0x01000193u + it;   // uniform "next train descriptor"
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = q[j][i] ^ t[i];
            uint32_t r;
            if (MODE == 13) {   // re-encoded rows (common.cuh ham256_key_enc): 13 LOP3 + 4 POPC
                auto ms = [](uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; };
                const uint32_t c0 = ms(x[0], x[1], x[2]), c1 = ms(x[3], x[4], x[5]), c2 = ms(x[2], x[5], x[6]);
                const uint32_t s3 = x3(c0, c1, c2), c3 = mj(c0, c1, c2);
                r = it;
                r = __popc(x[6]) * 65536u + r; r = __popc(x[7]) * 65536u + r;
                r = __popc(s3) * 131072u + r; r = __popc(c3) * 262144u + r;
            } else if (MODE == 8) {
                r = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) r += __popc(x[i]);
                r = r * 65536u + it;
            } else {
                uint32_t s0 = x3(x[0], x[1], x[2]), c0 = mj(x[0], x[1], x[2]);
                uint32_t s1 = x3(x[3], x[4], x[5]), c1 = mj(x[3], x[4], x[5]);
                uint32_t s2 = x3(s0, s1, x[6]), c2 = mj(s0, s1, x[6]);
                uint32_t s3 = x3(c0, c1, c2), c3 = mj(c0, c1, c2);
                r = it;
                r = __popc(s2) * 65536u + r; r = __popc(x[7]) * 65536u + r;
                r = __popc(s3) * 131072u + r; r = __popc(c3) * 262144u + r;
            }
            best[j] = min(best[j], r);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = best[0] ^ best[1] ^ best[2] ^ best[3];
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <typename F>
static double time_ms(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* out; long long* clk;
    CK(cudaMalloc(&out, sizeof(uint32_t) * 1024 * sms * 4));
    CK(cudaMalloc(&clk, sizeof(long long) * sms * 4));
    long long h_clk[1024];
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", prop.name, sms, prop.clockRate);
    struct { const char* name; void (*k)(uint32_t*, uint32_t, long long*); } pipes[] = {
        {"popc", popc_kernel}, {"lop3", lop3_kernel}, {"imad", imad_kernel}, {"min_u32", min_kernel}};
    // Rates are derived from CUDA-event time; "per clk" divides by the device's maximum SM clock (cudaDeviceProp::clockRate).
    // clock64() differences are reported as *_clk64_* only: on this part they advance slower than the SM clock the event
    // time implies (round 1 printed them as clk/pair, which contradicted the Gcmp/s next to them).
    const double max_hz = (double)prop.clockRate * 1e3;
    for (auto& p : pipes) {
        double ms = time_ms([&] { p.k<<<sms, 1024>>>(out, 12345u, clk); }, 5);
        CK(cudaMemcpy(h_clk, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)h_clk[i]; cyc /= sms;
        double ops_per_sm = 1024.0 * 8 * ITERS;
        printf(", \"%s_per_clk_per_sm\": %.2f, \"%s_gops\": %.1f, \"%s_clk64_over_event_clk\": %.3f", p.name,
               ops_per_sm / (ms * 1e-3 * max_hz), p.name, ops_per_sm * sms / (ms * 1e6), p.name, cyc / (ms * 1e-3 * max_hz));
    }
    {
        struct { const char* name; void (*k)(uint32_t*, uint32_t, long long*); } hams[] = {
            {"ham_plain8", ham_kernel<8>}, {"ham_csa4", ham_kernel<4>}, {"ham_enc13", ham_kernel<13>}};
        double pairs_sm = 2.0 * 256 * 4 * ITERS;   // 2 CTAs per SM
        for (auto& h : hams) {
            double ms = time_ms([&] { h.k<<<sms * 2, 256>>>(out, 12345u, clk); }, 5);
            printf(", \"%s_clk_per_pair_per_sm\": %.4f, \"%s_gcmps\": %.1f", h.name, ms * 1e-3 * max_hz / pairs_sm, h.name,
                   pairs_sm * sms / (ms * 1e6));
        }
    }
    {
        size_t n = (size_t)1 << 26;  // 1 GiB of uint4
        uint4 *a, *b;
        CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16));
        double ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
        printf(", \"hbm_copy_gbs\": %.1f", 2.0 * n * 16 / (ms * 1e6));
        cudaFree(a); cudaFree(b);
    }
    printf("}\n");
    return 0;
}
