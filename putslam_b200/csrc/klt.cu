// klt.cu -- K11/K12: the KLT tracking seam, MatcherOpenCV::performTracking (reference src/Matcher/matcherOpenCV.cpp:209-300;
// virtual, include/putslam/Matcher/matcher.h:415-422; called by Matcher::trackKLT, src/Matcher/matcher.cpp:133-160).
//   klt_pyrdown_kernel   level l+1 of both frames' pyramids from level l (cv::pyrDown, integer): one launch per level
//   klt_track_kernel     cv::calcOpticalFlowPyrLK: ONE launch for all points and all levels -- a point's levels depend
//                        on each other, the points do not, so a warp takes a point from the coarsest level to the base
//                        (OpenCV runs one parallel_for per level); Scharr gradients are computed on the fly for the
//                        8 x 8 pixels a window touches instead of for whole images (klt_point.cuh)
//   klt_prune_kernel     the rest of performTracking: err threshold, and of every pair of tracked positions closer than
//                        minimalReprojDistanceNewTrackingFeatures the one with the larger err is dropped (:247-266) --
//                        N^2 pair tests, one warp per feature
// Frames are at most a few MB and stay in L2 between the kernels; all three are latency-bound.
#include "common.cuh"
#include "kernels.h"

namespace pslam {

__global__ void __launch_bounds__(256)
klt_pyrdown_kernel(const uint8_t* __restrict__ src0, uint8_t* __restrict__ dst0, const uint8_t* __restrict__ src1,
                   uint8_t* __restrict__ dst1, int w, int h, int cn, int ow, int oh) {
    chain_begin();
    const uint8_t* src = blockIdx.z ? src1 : src0;
    uint8_t* dst = blockIdx.z ? dst1 : dst0;
    const int e = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;   // e = ox * cn + c
    if (e >= ow * cn) return;
    const int ox = e / cn, c = e - ox * cn;
    dst[(size_t)oy * ow * cn + e] = klt_pyrdown_px(src, w, h, cn, ox, oy, c);
}

cudaError_t launch_klt_pyramid(uint8_t* d_pyrI, uint8_t* d_pyrJ, const KltPlan& P, int cn, cudaStream_t st, int* launches) {
    // d_pyrI == nullptr: the previous frame's pyramid is already there (it was the current frame of the last call)
    for (int l = 1; l < P.n_levels; ++l) {
        const int w = P.w[l - 1], h = P.h[l - 1], ow = P.w[l], oh = P.h[l];
        dim3 grid((ow * cn + 255) / 256, oh, d_pyrI ? 2 : 1);
        cudaError_t e = launch_chained(klt_pyrdown_kernel, grid, dim3(256), 0, st, (const uint8_t*)(d_pyrJ + P.off[l - 1]),
                                       d_pyrJ + P.off[l], (const uint8_t*)(d_pyrI ? d_pyrI + P.off[l - 1] : nullptr),
                                       d_pyrI ? d_pyrI + P.off[l] : nullptr, w, h, cn, ow, oh);
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    return cudaSuccess;
}

template <int WIN_T, int CN_T>
__global__ void __launch_bounds__(128)
klt_track_kernel(const __grid_constant__ KltParams P, const float* __restrict__ prev_xy, float* __restrict__ cur_xy, int n,
                 uint8_t* __restrict__ status, float* __restrict__ err, int work_bytes) {
    extern __shared__ __align__(16) uint8_t klt_smem[];
    chain_begin();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= n) return;                                  // whole warps leave together
    const KltWork W = klt_carve(klt_smem + (size_t)warp * work_bytes, P.win, P.cn);
    float nx = 0.f, ny = 0.f, e;
    uint8_t st;
    if (P.use_initial_flow) { nx = cur_xy[2 * i]; ny = cur_xy[2 * i + 1]; }
    klt_track_point<WIN_T, CN_T>(P, W, prev_xy[2 * i], prev_xy[2 * i + 1], nx, ny, st, e);
    if (lane == 0) {
        cur_xy[2 * i] = nx; cur_xy[2 * i + 1] = ny;
        status[i] = st; err[i] = e;
    }
}

cudaError_t launch_klt_track(const KltParams& P, const float* d_prev_xy, float* d_cur_xy, int n, uint8_t* d_status,
                             float* d_err, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    const int work = (int)klt_work_bytes(P.win, P.cn);
    int warps = 4;
    while (warps > 1 && (size_t)warps * work > 48 * 1024) warps >>= 1;
    if ((size_t)warps * work > 48 * 1024) return cudaErrorInvalidValue;
    void (*k)(const KltParams, const float*, float*, int, uint8_t*, float*, int) = klt_track_kernel<0, 0>;
    if (P.win == 7 && P.cn == 3) k = klt_track_kernel<7, 3>;        // the reference's configuration
    else if (P.win == 7 && P.cn == 1) k = klt_track_kernel<7, 1>;
    cudaError_t e = launch_chained(k, dim3((n + warps - 1) / warps), dim3(32 * warps), (size_t)warps * work, st, P, d_prev_xy,
                                   d_cur_xy, n, d_status, d_err, work);
    if (e == cudaSuccess) ++*launches;
    return e;
}

// keep[i] = status[i] && !(err[i] > err_thr) && no closer-than-threshold neighbour wins against i (klt_pair_removes).
// One warp per feature: the lanes share the n - 1 pair tests (the positions are 12 KB per 1000 features and stay in L1),
// a float pre-test discards the pairs that are far apart, a warp vote collects the verdict.
__global__ void __launch_bounds__(128)
klt_prune_kernel(const float* __restrict__ xy, const float* __restrict__ err, const uint8_t* __restrict__ status, int n,
                 double err_thr, double sq_thr, float lim, uint8_t* __restrict__ keep) {
    chain_begin();
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;                                  // whole warps leave together
    const bool removed = __any_sync(0xffffffffu, klt_prune_lane(i, lane, n, xy, err, sq_thr, lim));
    if (lane == 0) keep[i] = (uint8_t)(status[i] != 0 && !((double)err[i] > err_thr) && !removed);
}

cudaError_t launch_klt_prune(const float* d_xy, const float* d_err, const uint8_t* d_status, int n, double err_thr,
                             double sq_thr, float lim, uint8_t* d_keep, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    cudaError_t e = launch_chained(klt_prune_kernel, dim3((n + 3) / 4), dim3(128), 0, st, d_xy, d_err, d_status, n, err_thr,
                                   sq_thr, lim, d_keep);
    if (e == cudaSuccess) ++*launches;
    return e;
}

// Ordered compaction of the survivors for the fused tracking frame (pslam_klt_frame): survivor j is feature kept[j].
// mout = {m, queryIdx[cap] = kept[j], trainIdx[cap] = j, distance[cap] = 0}: the DMatch(i, j, 0) list performTracking
// returns (matcherOpenCV.cpp:269-277), in the layout the RANSAC launcher reads match lists from; cxy[j] = xy[kept[j]]
// (the compacted `features`), zero beyond m so that the back-projection launched over all cap slots reads defined values.
// One CTA: n is a frame's feature count (a few thousand at most) and the order is the point.
__global__ void __launch_bounds__(1024)
klt_compact_kernel(const uint8_t* __restrict__ keep, const float2* __restrict__ xy, int n, int cap, int* __restrict__ mout,
                   float2* __restrict__ cxy) {
    __shared__ int warp_count[32];
    __shared__ int running;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const bool k = i < n && keep[i] != 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_count[warp] = __popc(ballot);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_count[w];
        if (k) {
            const int j = before + __popc(ballot & ((1u << lane) - 1u));
            mout[1 + j] = i; mout[1 + cap + j] = j; mout[1 + 2 * cap + j] = 0;
            cxy[j] = xy[i];
        }
        __syncthreads();
        if (tid == 0) {
            int t = running;
            for (int w = 0; w < 32; ++w) t += warp_count[w];
            running = t;
        }
        __syncthreads();
    }
    const int m = running;
    if (tid == 0) mout[0] = m;
    for (int j = m + tid; j < cap; j += 1024) cxy[j] = make_float2(0.f, 0.f);
}

cudaError_t launch_klt_compact(const uint8_t* d_keep, const float* d_xy, int n, int cap, int* d_mout, float* d_cxy,
                               cudaStream_t st, int* launches) {
    cudaError_t e = launch_chained(klt_compact_kernel, dim3(1), dim3(1024), 0, st, d_keep, (const float2*)d_xy, n, cap, d_mout,
                                   (float2*)d_cxy);
    if (e == cudaSuccess) ++*launches;
    return e;
}

}  // namespace pslam
