// guided.cu -- K3: guided frame-to-map matching, the inner loops of Matcher::matchXYZ
// (reference src/Matcher/matcher.cpp:670-748).
//
// One warp per map feature j; lanes stride over the current keypoints i.  Gate = 3-D distance below
// the sphere radius AND predicted pyramid levels within one of each other (:699-711); distance =
// popcount of the per-byte saturating difference (the reference's cv::Mat subtraction quirk, :719-721)
// or XOR Hamming; best = first minimum (:714-726); every candidate with ratio*value <= best is emitted
// (:734-747) in (j, i) order.  Three launches: count -> scan -> emit (gates are recomputed, they are a
// handful of float ops; only gated candidates touch descriptors).
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

namespace pslam {

constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;

__device__ __forceinline__ uint32_t desc_distance(const uint4* __restrict__ a, const uint4* __restrict__ b, int mode) {
    const uint4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
    if (mode == 0) {
        return satsub_popc32(a0.x, b0.x) + satsub_popc32(a0.y, b0.y) + satsub_popc32(a0.z, b0.z) +
               satsub_popc32(a0.w, b0.w) + satsub_popc32(a1.x, b1.x) + satsub_popc32(a1.y, b1.y) +
               satsub_popc32(a1.z, b1.z) + satsub_popc32(a1.w, b1.w);
    }
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

struct GuidedArgs {
    const float* map_xyz;
    const uint4* map_desc;
    const int* map_level;
    int M;
    const float* cur_xyz;
    const uint4* cur_desc;
    const int* cur_level;
    int N;
    float radius_f;       // smallest float >= radius: (double)norm < radius  <=>  norm < radius_f
    double accept_ratio;
    int mode;
};

__device__ __forceinline__ bool gate(const GuidedArgs& A, float px, float py, float pz, int lvl, const float* sxyz,
                                     const int* slvl, int i) {
    const float dx = px - sxyz[3 * i], dy = py - sxyz[3 * i + 1], dz = pz - sxyz[3 * i + 2];
    const float nrm = norm3(dx, dy, dz);
    const int li = slvl[i];
    return (nrm < A.radius_f) && (li - 1 <= lvl) && (lvl <= li + 1);
}

// EMIT = false: writes best[j] (value<<16 | i, 0xffffffff if no candidate) and count[j].
// EMIT = true : writes the matches of feature j at offsets[j].
template <bool EMIT>
__global__ void __launch_bounds__(kGThreads)
guided_kernel(GuidedArgs A, uint32_t* __restrict__ best, int* __restrict__ count, const int* __restrict__ offsets,
              int cap, int* __restrict__ out_q, int* __restrict__ out_t, float* __restrict__ out_d) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* sxyz = reinterpret_cast<float*>(smem_raw);
    int* slvl = reinterpret_cast<int*>(sxyz + 3 * (size_t)A.N);
    for (int i = threadIdx.x; i < 3 * A.N; i += kGThreads) sxyz[i] = A.cur_xyz[i];
    for (int i = threadIdx.x; i < A.N; i += kGThreads) slvl[i] = A.cur_level[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kGWarps + (threadIdx.x >> 5);
    for (int j = gw; j < A.M; j += gridDim.x * kGWarps) {
        const float px = A.map_xyz[3 * j], py = A.map_xyz[3 * j + 1], pz = A.map_xyz[3 * j + 2];
        const int lvl = A.map_level[j];
        const uint4* dj = A.map_desc + 2 * (size_t)j;
        uint32_t bestp;
        if (!EMIT) {
            bestp = 0xffffffffu;
            for (int i = lane; i < A.N; i += 32)
                if (gate(A, px, py, pz, lvl, sxyz, slvl, i))
                    bestp = min(bestp, (desc_distance(dj, A.cur_desc + 2 * (size_t)i, A.mode) << 16) | (uint32_t)i);
            bestp = warp_min_u32(bestp);
            if (lane == 0) best[j] = bestp;
        } else {
            bestp = best[j];
        }
        if (bestp == 0xffffffffu) {
            if (!EMIT && lane == 0) count[j] = 0;
            continue;
        }
        const double bestVal = (double)(float)(bestp >> 16);
        int run = EMIT ? offsets[j] : 0;
        for (int base = 0; base < A.N; base += 32) {
            const int i = base + lane;
            bool emit = false;
            uint32_t v = 0;
            if (i < A.N && gate(A, px, py, pz, lvl, sxyz, slvl, i)) {
                v = desc_distance(dj, A.cur_desc + 2 * (size_t)i, A.mode);
                emit = __dmul_rn(A.accept_ratio, (double)(float)v) <= bestVal;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, emit);
            if (EMIT && emit) {
                const int pos = run + __popc(bal & ((1u << lane) - 1u));
                if (pos < cap) { out_q[pos] = j; out_t[pos] = i; out_d[pos] = (float)v; }
            }
            run += __popc(bal);
        }
        if (!EMIT && lane == 0) count[j] = run;
    }
}

// Exclusive scan of count[0..M) into offsets[0..M]; header[0] = total, header[1] = perfect matches
// (best value < 0.1, i.e. 0; reference matcher.cpp:729-731).  Single CTA.
__global__ void __launch_bounds__(1024, 1)
guided_scan_kernel(const int* __restrict__ count, const uint32_t* __restrict__ best, int M, int* __restrict__ offsets,
                   int* __restrict__ header) {
    __shared__ int warp_tot[32];
    __shared__ int carry, perfect;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; perfect = 0; }
    __syncthreads();
    int my_perfect = 0;
    for (int base = 0; base < M; base += 1024) {
        const int j = base + tid;
        const int c = (j < M) ? count[j] : 0;
        if (j < M && (best[j] >> 16) == 0u) ++my_perfect;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int cw = warp_tot[w];
            if (w < warp) woff += cw;
            tot += cw;
        }
        if (j < M) offsets[j] = carry + woff + incl - c;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    my_perfect = (int)warp_add_u32((uint32_t)my_perfect);
    if (lane == 0 && my_perfect) atomicAdd(&perfect, my_perfect);
    __syncthreads();
    if (tid == 0) { offsets[M] = carry; header[0] = carry; header[1] = perfect; }
}

cudaError_t launch_guided_match(const float* d_map_xyz, const uint8_t* d_map_desc, const int* d_map_level, int M,
                                const float* d_cur_xyz, const uint8_t* d_cur_desc, const int* d_cur_level, int N,
                                float radius_f, double accept_ratio, int mode, int* d_count, int* d_best, int* d_out,
                                int cap, cudaStream_t st, int* launches) {
    GuidedArgs A;
    A.map_xyz = d_map_xyz; A.map_desc = reinterpret_cast<const uint4*>(d_map_desc); A.map_level = d_map_level; A.M = M;
    A.cur_xyz = d_cur_xyz; A.cur_desc = reinterpret_cast<const uint4*>(d_cur_desc); A.cur_level = d_cur_level; A.N = N;
    A.radius_f = radius_f; A.accept_ratio = accept_ratio; A.mode = mode;
    // d_count: M counts followed by M+1 offsets
    int* d_offsets = d_count + M;
    int* out_q = d_out + 2;
    int* out_t = d_out + 2 + cap;
    float* out_d = reinterpret_cast<float*>(d_out + 2 + 2 * cap);
    const size_t smem = sizeof(float) * 4 * (size_t)N;
    int grid = (M + kGWarps - 1) / kGWarps;
    if (grid > 4 * PSLAM_SM_COUNT_HINT) grid = 4 * PSLAM_SM_COUNT_HINT;
    if (grid < 1) grid = 1;
    cudaError_t e;
    if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(guided_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(guided_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    }
    guided_kernel<false><<<grid, kGThreads, smem, st>>>(A, reinterpret_cast<uint32_t*>(d_best), d_count, nullptr, cap,
                                                        nullptr, nullptr, nullptr);
    guided_scan_kernel<<<1, 1024, 0, st>>>(d_count, reinterpret_cast<const uint32_t*>(d_best), M, d_offsets, d_out);
    guided_kernel<true><<<grid, kGThreads, smem, st>>>(A, reinterpret_cast<uint32_t*>(d_best), d_count, d_offsets, cap,
                                                       out_q, out_t, out_d);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace pslam
