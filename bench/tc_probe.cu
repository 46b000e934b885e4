// tc_probe.cu -- can the loop-closure sweep run on the 5th-generation tensor cores, bit-exactly, and how fast?
//
// The sweep is a 1000 x 1e7 x 256-bit binary contraction.  With rows expanded to signed bytes, q_k -> +-16 and t_k -> +-8
// (bit 0 -> +, bit 1 -> -), one tcgen05.mma kind::i8 accumulates  sum_k q_k t_k = 128 (256 - 2 Ham) = 256 (128 - Ham)
// in int32, exactly.  Four extra K slots (of the 32 that pad K from 256 to 288) add the index fields, so that the raw
// accumulator IS the comparison key of the popcount kernel:   acc = 256 (128 - Ham) + (255 - t_local) + (255 - q_local)
// -- a maximum over a row (fixed q) picks the lowest Hamming distance and, among equals, the lowest t; a maximum over a
// column (fixed t) the lowest q: OpenCV's cross-check tie-break without a single ALU instruction per pair.
//
// Sections (each prints one JSON line; every wait has a clock time-out so that a wrong descriptor cannot hang the GPU):
//   layout   one 128 x 256 x 288 tile through expansion -> smem descriptors -> MMA -> TMEM -> tcgen05.ld, compared with the CPU
//   mma      back-to-back MMA issue rate (no epilogue)        -> int8 ops/s per SM and per chip
//   ldtm     tcgen05.ld rate with 4 and 8 warps               -> TMEM read bytes/clk/SM (the epilogue's ceiling)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s\", \"line\": %d}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

#include "../putslam_b200/csrc/lc_tc.cuh"

namespace tc {
using namespace pslam::tc;

__device__ __forceinline__ bool probe_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 400000000ll) return false;
    return true;
}
__device__ __forceinline__ void load_row(const uint32_t* p, uint32_t (&w)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = p[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// section "layout": one tile, every accumulator back to the host
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layout_kernel(const uint32_t* __restrict__ q /* 128 x 8 */, const uint32_t* __restrict__ t /* 256 x 8 */,
                                                     int* __restrict__ out /* 128 x 256 */, int* __restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sq = smem;                         // 128 rows
    uint8_t* st = smem + 128 * kRowBytes;       // 256 rows
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    uint32_t w[8];
    if (tid < 128) { load_row(q + tid * 8, w); expand_row(w, true, sq, tid, 255 - tid, true); }
    load_row(t + tid * 8, w);
    expand_row(w, true, st, tid, 255 - tid, false);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        const uint64_t ad = make_desc(smem_u32(sq)), bd = make_desc(smem_u32(st));
        const uint32_t idesc = make_idesc(128, 256);
#pragma unroll 1
        for (int k = 0; k < kRowBytes / 32; ++k)
            mma_i8(tm, ad + (uint64_t)((k * 2 * kLBO) >> 4), bd + (uint64_t)((k * 2 * kLBO) >> 4), idesc, k > 0);
        mma_commit(smem_u32(&s_bar));
    }
    const bool ok = probe_wait(smem_u32(&s_bar), 0);
    fence_after();
    if (!ok) { if (tid == 0) status[0] = 1; }
    else if (warp < 4) {
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
            int v[32];
            tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 256 + c + j] = v[j];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// section "mma": issue rate.  One thread issues `groups` accumulation groups of 9 MMAs (M = 128, N = n) on whatever the
// shared memory holds; a commit per group, the barrier waited two groups behind (as a double-buffered pipeline would)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_rate_kernel(int groups, int n, long long* __restrict__ clk, int* __restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * kRowBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    if (tid == 0) { mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem + 128 * kRowBytes));
        const uint32_t idesc = make_idesc(128, n);
        bool ok = true;
        const long long t0 = clock64();
#pragma unroll 1
        for (int g = 0; g < groups && ok; ++g) {
            const int b = g & 1;
            if (g >= 2) ok = probe_wait(smem_u32(&s_bar[b]), ((g >> 1) - 1) & 1);
#pragma unroll 1
            for (int k = 0; k < kRowBytes / 32; ++k)
                mma_i8(tm + (uint32_t)(b * 256), ad + (uint64_t)((k * 2 * kLBO) >> 4), bd + (uint64_t)((k * 2 * kLBO) >> 4), idesc, k > 0);
            mma_commit(smem_u32(&s_bar[b]));
        }
        for (int g = groups - 2; g < groups && ok; ++g)
            if (g >= 0) ok = probe_wait(smem_u32(&s_bar[g & 1]), (g >> 1) & 1);
        const long long t1 = clock64();
        clk[blockIdx.x] = t1 - t0;
        if (!ok) status[0] = 2;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// section "ldtm": TMEM read rate.  `warps` warps (4 or 8) each read x32 chunks of their lane quarter in a loop and fold
// them with 3-input maxima (what the sweep's epilogue does with them)
// ---------------------------------------------------------------------------------------------------------------------
template <int fold>
__global__ void __launch_bounds__(256) ldtm_rate_kernel(int iters, long long* __restrict__ clk, int* __restrict__ sink) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    int m = -0x7fffffff;
    int colmax[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) colmax[i] = -0x7fffffff;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 256; c += 32) {
            int v[32];
            tmem_ld32(base + (uint32_t)c, v);
            tmem_ld_wait();
            if (fold == 2) {          // both reductions of a single-orientation epilogue: row maximum + warp-wide column maxima
                int keep = -0x7fffffff;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    m = max(m, max(v[j], v[j + 1]));
                    const int r0 = __reduce_max_sync(0xffffffffu, v[j]), r1 = __reduce_max_sync(0xffffffffu, v[j + 1]);
                    if ((tid & 31) == j) keep = r0;
                    if ((tid & 31) == j + 1) keep = r1;
                }
                colmax[c >> 5] = max(colmax[c >> 5], keep);
            } else if (fold) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) m = max(m, max(v[j], v[j + 1]));
            } else {
                m ^= v[0] + v[31];
            }
        }
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
#pragma unroll
    for (int i = 0; i < 8; ++i) m ^= colmax[i];
    sink[blockIdx.x * blockDim.x + tid] = m;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// section "sweep": the whole tensor-core sweep (lc_tc.cuh) on a synthetic map against a plain popcount kernel
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(const long long* __restrict__ kf_off, int nq, long long n_desc, const uint32_t* __restrict__ row_best,
                                                       const uint32_t* __restrict__ col_best, int tau, int* __restrict__ scores) {
    __shared__ uint32_t s_col[4096];
    __shared__ int s_cnt;
    finalize_keyframe((int)blockIdx.x, kf_off, nq, n_desc, row_best, col_best, tau, scores, s_col, &s_cnt);
}
__global__ void __launch_bounds__(1024) ref_scores_kernel(const uint32_t* __restrict__ db, const long long* __restrict__ kf_off,
                                                          const uint32_t* __restrict__ query, int nq, int tau, int* __restrict__ scores) {
    __shared__ uint32_t s_row[1024], s_col[4096];
    __shared__ int s_cnt;
    const int kf = blockIdx.x;
    const long long r0 = kf_off[kf];
    const int n_t = (int)(kf_off[kf + 1] - r0);
    if (threadIdx.x == 0) s_cnt = 0;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        uint32_t best = 0xffffffffu;
        for (int t = 0; t < n_t; ++t) {
            int d = 0;
            for (int w = 0; w < 8; ++w) d += __popc(query[q * 8 + w] ^ db[(r0 + t) * 8 + w]);
            const uint32_t key = ((uint32_t)d << 16) | (uint32_t)t;
            best = min(best, key);
        }
        s_row[q] = best;
    }
    for (int t = threadIdx.x; t < n_t; t += blockDim.x) {
        uint32_t best = 0xffffffffu;
        for (int q = 0; q < nq; ++q) {
            int d = 0;
            for (int w = 0; w < 8; ++w) d += __popc(query[q * 8 + w] ^ db[(r0 + t) * 8 + w]);
            best = min(best, ((uint32_t)d << 16) | (uint32_t)q);
        }
        s_col[t] = best;
    }
    __syncthreads();
    int cnt = 0;
    if (n_t > 0)
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            const uint32_t rb = s_row[q];
            if ((int)(rb >> 16) <= tau && (int)(s_col[rb & 0xffffu] & 0xffffu) == q) ++cnt;
        }
    if (cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) scores[kf] = s_cnt;
}

}  // namespace tc

static int popc32(uint32_t x) { return __builtin_popcount(x); }

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    int* d_status;
    CK(cudaMalloc(&d_status, 16));
    CK(cudaMemset(d_status, 0, 16));
    int h_status = 0;

    const bool only_sweep = argc > 5 && atoi(argv[5]);
    // ---- layout ----
    if (!only_sweep) {
        std::vector<uint32_t> q(128 * 8), t(256 * 8);
        uint64_t s = 0x9e3779b97f4a7c15ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
        for (auto& x : q) x = rnd();
        for (auto& x : t) x = rnd();
        for (int w = 0; w < 8; ++w) { t[5 * 8 + w] = q[3 * 8 + w]; t[200 * 8 + w] = ~q[100 * 8 + w]; }   // distance 0 and 256
        uint32_t *dq, *dt; int* dout;
        CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dt, t.size() * 4)); CK(cudaMalloc(&dout, 128 * 256 * 4));
        CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dt, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dout, 0xff, 128 * 256 * 4));
        const int smem = (128 + 256) * tc::kRowBytes;
        CK(cudaFuncSetAttribute(tc::layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tc::layout_kernel<<<1, 256, smem>>>(dq, dt, dout, d_status);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h_status, d_status, 4, cudaMemcpyDeviceToHost));
        std::vector<int> out(128 * 256);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0; int first_q = -1, first_t = -1, first_got = 0, first_want = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 256; ++j) {
                int ham = 0;
                for (int w = 0; w < 8; ++w) ham += popc32(q[i * 8 + w] ^ t[j * 8 + w]);
                const int want = 512 * (128 - ham) + (255 - j) + (255 - i);
                if (out[i * 256 + j] != want) { if (!bad) { first_q = i; first_t = j; first_got = out[i * 256 + j]; first_want = want; } ++bad; }
            }
        printf("{\"section\": \"layout\", \"status\": %d, \"mismatches\": %lld, \"of\": %d, \"first\": [%d, %d, %d, %d], \"sample\": [%d, %d, %d, %d]}\n",
               h_status, bad, 128 * 256, first_q, first_t, first_got, first_want, out[0], out[1], out[256], out[3 * 256 + 5]);
        fflush(stdout);
        if (h_status || bad) return 0;      // do not time a path that is not right
    }
    // ---- mma rate ----
    long long* d_clk; CK(cudaMalloc(&d_clk, sizeof(long long) * 1024));
    for (int n : {256, 128}) {
        if (only_sweep) break;
        const int groups = 2000, smem = (128 + 256) * tc::kRowBytes;
        CK(cudaFuncSetAttribute(tc::mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        tc::mma_rate_kernel<<<sms, 128, smem>>>(200, n, d_clk, d_status);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tc::mma_rate_kernel<<<sms, 128, smem>>>(groups, n, d_clk, d_status);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> clk(sms);
        CK(cudaMemcpy(clk.data(), d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&h_status, d_status, 4, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto c : clk) mx = c > mx ? c : mx;
        const double macs = (double)groups * 9 * 128.0 * n * 32.0;
        printf("{\"section\": \"mma\", \"n\": %d, \"status\": %d, \"clk_per_group\": %.1f, \"macs_per_clk_per_sm\": %.1f, \"chip_tops_event_time\": %.1f, \"ms\": %.3f}\n",
               n, h_status, (double)mx / groups, macs / (double)mx, 2.0 * macs * sms / (ms * 1e-3) / 1e12, ms);
        fflush(stdout);
    }
    // ---- ldtm rate ----
    int* d_sink; CK(cudaMalloc(&d_sink, 4 * 256 * sms));
    for (int warps : {4, 8})
        for (int fold : {0, 1, 2}) {
            if (only_sweep) break;
            const int iters = 2000;
            auto launch = [&](int n) {
                if (fold == 0) tc::ldtm_rate_kernel<0><<<sms, warps * 32>>>(n, d_clk, d_sink);
                else if (fold == 1) tc::ldtm_rate_kernel<1><<<sms, warps * 32>>>(n, d_clk, d_sink);
                else tc::ldtm_rate_kernel<2><<<sms, warps * 32>>>(n, d_clk, d_sink);
            };
            launch(100);
            CK(cudaDeviceSynchronize());
            launch(iters);
            CK(cudaDeviceSynchronize());
            std::vector<long long> clk(sms);
            CK(cudaMemcpy(clk.data(), d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            long long mx = 0; for (auto c : clk) mx = c > mx ? c : mx;
            const double bytes = (double)iters * 8 * 32 * 32 * 4 * warps;
            printf("{\"section\": \"ldtm\", \"warps\": %d, \"fold\": %d, \"bytes_per_clk_per_sm\": %.1f, \"values_per_clk_per_sm\": %.1f}\n",
                   warps, fold, bytes / (double)mx, bytes / 4 / (double)mx);
            fflush(stdout);
        }

    // ---- sweep ----
    {
        const int nq = argc > 1 ? atoi(argv[1]) : 1000;
        const int n_kf = argc > 2 ? atoi(argv[2]) : 2048;
        const int check_kf = argc > 3 ? atoi(argv[3]) : 296;
        const int tau = 64;
        uint64_t s = 0x243f6a8885a308d3ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
        std::vector<long long> off(n_kf + 1, 0);
        for (int k = 0; k < n_kf; ++k) {
            int n = 1000;
            if (k % 7 == 3) n = 700 + (int)(rnd() % 600);       // ragged
            if (k == 5) n = 1;
            if (k == 6) n = 256;
            if (k == 9) n = 0;
            off[k + 1] = off[k] + n;
        }
        const long long n_desc = off[n_kf];
        std::vector<uint32_t> q((size_t)nq * 8), db((size_t)n_desc * 8);
        for (auto& x : q) x = rnd();
        for (auto& x : db) x = rnd();
        // planted near-duplicates (and exact duplicates -> ties) so that the scores are not all zero
        for (int k = 0; k < n_kf; ++k) {
            const long long n = off[k + 1] - off[k];
            if (!n) continue;
            const int plant = (int)(rnd() % 200);
            for (int j = 0; j < plant; ++j) {
                const long long t = off[k] + (long long)(rnd() % n);
                const int qq = (int)(rnd() % nq);
                for (int w = 0; w < 8; ++w) db[t * 8 + w] = q[(size_t)qq * 8 + w];
                const int flips = (int)(rnd() % 90);
                for (int f = 0; f < flips; ++f) { const int bit = (int)(rnd() % 256); db[t * 8 + bit / 32] ^= 1u << (bit % 32); }
            }
        }
        uint32_t *dq, *ddb, *drow, *dcol; long long* doff; int *dsc, *dref;
        CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&ddb, db.size() * 4)); CK(cudaMalloc(&doff, off.size() * 8));
        CK(cudaMalloc(&drow, (size_t)n_kf * tc::kMaxQueries * 4)); CK(cudaMalloc(&dcol, (size_t)n_desc * tc::kSplits * 4));
        CK(cudaMalloc(&dsc, n_kf * 4)); CK(cudaMalloc(&dref, n_kf * 4)); int* dref2; CK(cudaMalloc(&dref2, n_kf * 4));
        CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ddb, db.data(), db.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(doff, off.data(), off.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_status, 0, 16));
        tc::SweepArgs A;
        A.db = ddb; A.kf_off = doff; A.n_kf = n_kf; A.db_encoded = 0; A.query = dq; A.nq = nq; A.n_desc = n_desc;
        A.row_best = drow; A.col_best = dcol; A.status = d_status;
        A.n_splits = (nq + tc::kQRows - 1) / tc::kQRows; A.qflag = nullptr; A.qepoch = 0;
        unsigned long long* dst; CK(cudaMalloc(&dst, 64)); CK(cudaMemset(dst, 0, 64)); A.stamps = dst;
        int* dkfd; CK(cudaMalloc(&dkfd, n_kf * 4)); CK(cudaMemset(dkfd, 0, n_kf * 4));
        const int fused = argc > 4 ? atoi(argv[4]) : 1;
        A.tau = tau; A.scores = fused ? dsc : nullptr; A.kf_done = dkfd;     // fused finalize (the separate kernel below then only re-derives the same scores)
        CK(cudaFuncSetAttribute(tc::lc_tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        const int grid = (sms / A.n_splits) * A.n_splits;
        cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        float ms_sweep = 0, ms_fin = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            tc::lc_tc_sweep_kernel<<<grid, tc::kThreads, tc::kSmemBytes>>>(A);
            cudaEventRecord(e1);
            tc::finalize_kernel<<<n_kf, 256>>>(doff, nq, n_desc, drow, dcol, tau, fused ? dref2 : dsc);
            if (!fused) cudaMemcpyAsync(dref2, dsc, n_kf * 4, cudaMemcpyDeviceToDevice);
            cudaEventRecord(e2);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms_sweep, e0, e1); cudaEventElapsedTime(&ms_fin, e1, e2);
        }
        CK(cudaMemcpy(&h_status, d_status, 4, cudaMemcpyDeviceToHost));
        unsigned long long stp[8]; CK(cudaMemcpy(stp, dst, 64, cudaMemcpyDeviceToHost));
        printf("{\"section\": \"sweep_phases_cta0_us\", \"n_kf\": %d, \"prologue_producers\": %.2f, \"prologue_rest\": %.2f, \"first_group_read\": %.2f, \"last_keyframe_read\": %.2f, \"last_finalize\": %.2f, \"exit\": %.2f}\n",
               n_kf, (stp[1] - stp[0]) * 1e-3, (stp[2] - stp[0]) * 1e-3, (stp[3] - stp[0]) * 1e-3, (stp[4] - stp[0]) * 1e-3, (stp[5] - stp[0]) * 1e-3, (stp[6] - stp[0]) * 1e-3);
        const int nchk = check_kf < n_kf ? check_kf : n_kf;
        tc::ref_scores_kernel<<<nchk, 1024>>>(ddb, doff, dq, nq, tau, dref);
        CK(cudaDeviceSynchronize());
        std::vector<int> sc(n_kf), rf(nchk), sc2(n_kf);
        CK(cudaMemcpy(sc.data(), dsc, n_kf * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(sc2.data(), dref2, n_kf * 4, cudaMemcpyDeviceToHost));
        int fused_vs_separate = 0;
        for (int k = 0; k < n_kf; ++k) fused_vs_separate += sc[k] != sc2[k];
        CK(cudaMemcpy(rf.data(), dref, nchk * 4, cudaMemcpyDeviceToHost));
        int bad = 0, first = -1; long long sum = 0;
        for (int k = 0; k < nchk; ++k) { if (sc[k] != rf[k]) { if (first < 0) first = k; ++bad; } sum += rf[k]; }
        const double pairs = (double)nq * (double)n_desc;
        printf("{\"section\": \"sweep\", \"fused_finalize\": %d, \"status\": %d, \"nq\": %d, \"n_kf\": %d, \"n_desc\": %lld, \"checked_kf\": %d, \"fused_vs_separate_finalize_mismatches\": %d, \"score_mismatches\": %d, \"first_bad\": %d, "
               "\"got\": %d, \"want\": %d, \"mean_score\": %.2f, \"sweep_ms\": %.4f, \"finalize_ms\": %.4f, \"gcmp_per_s\": %.1f, \"gcmp_per_s_sweep_only\": %.1f}\n",
               fused, h_status, nq, n_kf, n_desc, nchk, fused_vs_separate, bad, first, first >= 0 ? sc[first] : 0, first >= 0 ? rf[first] : 0, (double)sum / nchk,
               ms_sweep, ms_fin, pairs / ((ms_sweep + ms_fin) * 1e-3) / 1e9, pairs / (ms_sweep * 1e-3) / 1e9);
        fflush(stdout);
    }
    printf("{\"section\": \"device\", \"sms\": %d, \"clock_khz\": %d}\n", sms, khz);
    return 0;
}
