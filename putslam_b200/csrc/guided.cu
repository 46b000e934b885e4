// guided.cu -- K3: guided frame-to-map matching, the inner loops of Matcher::matchXYZ
// (reference src/Matcher/matcher.cpp:670-748).
//
// One warp per map feature j; lanes stride over the current keypoints i.  Gate = 3-D distance below
// the sphere radius AND predicted pyramid levels within one of each other (:699-711); distance =
// popcount of the per-byte saturating difference (the reference's cv::Mat subtraction quirk, :719-721)
// or XOR Hamming; best = first minimum (:714-726); every candidate with ratio*value <= best is emitted
// (:734-747) in (j, i) order.  Two launches: collect (gate, distances, best, ratio filter into a per-feature
// candidate cache) -> scan + emit (one CTA).  Only gated candidates touch descriptors.
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
    int M;                // number of map features, or their capacity when M_dev is set
    const int* M_dev;     // device-resident count (map filtered on the device, mapprep.cu); nullptr = use M
    const float* cur_xyz;
    const uint4* cur_desc;
    const int* cur_level;
    int N;
    float sq_radius_f;    // squared-norm form of the gate: (double)sqrtf(s) < radius  <=>  s < sq_radius_f (exact)
    float z_reach;        // >= sqrt(sq_radius_f) with margin: |dz| of any keypoint that can pass the gate
    double accept_ratio;
    int mode;
};

constexpr int kCacheCap = 16;   // gated candidates cached per map feature; more than that -> recompute path

__device__ __forceinline__ bool gate(const GuidedArgs& A, float px, float py, float pz, int lvl, const float* sxyz,
                                     const int* slvl, int i) {
    const float dx = px - sxyz[3 * i], dy = py - sxyz[3 * i + 1], dz = pz - sxyz[3 * i + 2];
    const float yy = dy * dy, zz = dz * dz;
    const float sq = dx * dx + (yy + zz);           // Eigen squaredNorm association
    const int li = slvl[i];
    return (sq < A.sq_radius_f) && (li - 1 <= lvl) && (lvl <= li + 1);
}

// Pass 1 (grid): one warp per map feature.  Every CTA first sorts the current keypoints into 64 depth bins of
// 0.125 m in shared memory (counting sort: histogram, scan, scatter; positions + (level << 16 | index) per slot), so
// that a feature only visits the bins its sphere can reach in z -- a keypoint inside the sphere has
// |dz| <= sqrt(sq_radius), and the bin index is a monotone function of z, so no candidate is missed -- instead of all
// N keypoints (6 % of them at the default 0.12 m gate on a 0.8-5 m scene).  The gated candidates' descriptor
// distances are evaluated, (i, value) of the first kCacheCap are cached per feature, the best (first minimum: the key
// (value << 16 | i) does not depend on the visiting order) is found, then the cached candidates are put back into
// ascending i -- the reference's loop order -- and filtered by the accept ratio in place.  count[j] = matches of
// feature j; bit 31 set = more than kCacheCap candidates, the emit pass recomputes that feature in index order.
constexpr int kZBins = 64;
constexpr float kZBinsPerMetre = 8.f;   // power of two: z * 8 is exact
__device__ __forceinline__ int z_bin(float z) {
    // fmaxf/fminf drop NaN (-> bin 0; such a keypoint fails every gate anyway)
    return (int)fminf(fmaxf(floorf(z * kZBinsPerMetre), 0.f), (float)(kZBins - 1));
}

__global__ void __launch_bounds__(kGThreads)
guided_collect_kernel(GuidedArgs A, uint32_t* __restrict__ best, int* __restrict__ count, uint2* __restrict__ cache) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    chain_begin();
    float* sxyz = reinterpret_cast<float*>(smem_raw);                    // 3 N, bin-sorted
    uint32_t* skey = reinterpret_cast<uint32_t*>(sxyz + 3 * (size_t)A.N);  // N: level << 16 | original index
    int* bin_start = reinterpret_cast<int*>(skey + A.N);                 // kZBins + 1
    int* cursor = bin_start + kZBins + 1;                                // kZBins
    const int tid = threadIdx.x;
    if (tid < kZBins) cursor[tid] = 0;
    __syncthreads();
    for (int i = tid; i < A.N; i += kGThreads) atomicAdd(&cursor[z_bin(A.cur_xyz[3 * i + 2])], 1);
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the 64 bin counts by one warp
        const int c0 = cursor[2 * tid], c1 = cursor[2 * tid + 1];
        int incl = c0 + c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= d) incl += o;
        }
        const int excl = incl - (c0 + c1);
        bin_start[2 * tid] = excl; bin_start[2 * tid + 1] = excl + c0;
        if (tid == 31) bin_start[kZBins] = incl;
        cursor[2 * tid] = excl; cursor[2 * tid + 1] = excl + c0;
    }
    __syncthreads();
    for (int i = tid; i < A.N; i += kGThreads) {
        const float x = A.cur_xyz[3 * i], y = A.cur_xyz[3 * i + 1], z = A.cur_xyz[3 * i + 2];
        const int sl = atomicAdd(&cursor[z_bin(z)], 1);   // order inside a bin is arbitrary; restored per feature below
        sxyz[3 * sl] = x; sxyz[3 * sl + 1] = y; sxyz[3 * sl + 2] = z;
        skey[sl] = ((uint32_t)A.cur_level[i] << 16) | (uint32_t)i;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int gw = blockIdx.x * kGWarps + (threadIdx.x >> 5);
    const int M = A.M_dev ? min(*A.M_dev, A.M) : A.M;
    for (int j = gw; j < M; j += gridDim.x * kGWarps) {
        const float px = A.map_xyz[3 * j], py = A.map_xyz[3 * j + 1], pz = A.map_xyz[3 * j + 2];
        const int lvl = A.map_level[j];
        const uint4* dj = A.map_desc + 2 * (size_t)j;
        uint2* slot = cache + (size_t)j * kCacheCap;
        const int s_lo = bin_start[z_bin(pz - A.z_reach)], s_hi = bin_start[z_bin(pz + A.z_reach) + 1];
        uint32_t bestp = 0xffffffffu;
        int nc = 0;
        for (int base = s_lo; base < s_hi; base += 32) {
            const int sl = base + lane;
            bool g = false;
            uint32_t key = 0;
            if (sl < s_hi) {
                key = skey[sl];
                const float dx = px - sxyz[3 * sl], dy = py - sxyz[3 * sl + 1], dz = pz - sxyz[3 * sl + 2];
                const float yy = dy * dy, zz = dz * dz;
                const float sq = dx * dx + (yy + zz);           // Eigen squaredNorm association
                const int li = (int)(key >> 16);
                g = (sq < A.sq_radius_f) && (li - 1 <= lvl) && (lvl <= li + 1);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, g);
            if (bal == 0u) continue;
            if (g) {
                const uint32_t i = key & 0xffffu;
                const uint32_t v = desc_distance(dj, A.cur_desc + 2 * (size_t)i, A.mode);
                const int pos = nc + __popc(bal & lt);
                if (pos < kCacheCap) slot[pos] = make_uint2(i, v);
                bestp = min(bestp, (v << 16) | i);
            }
            nc += __popc(bal);
        }
        bestp = warp_min_u32(bestp);
        if (lane == 0) best[j] = bestp;
        if (nc == 0) {
            if (lane == 0) count[j] = 0;
            continue;
        }
        const double bestVal = (double)(float)(bestp >> 16);
        int cnt = 0;
        if (nc <= kCacheCap) {
            __syncwarp();
            uint2 e = make_uint2(0xffffffffu, 0);
            if (lane < nc) e = slot[lane];
            const bool emit = lane < nc && __dmul_rn(A.accept_ratio, (double)(float)e.y) <= bestVal;
            // back to ascending keypoint index, the reference's loop order (indices are distinct): position among
            // the emitted candidates = number of emitted candidates with a smaller index
            int epos = 0, ecnt = 0;
            for (int k = 0; k < nc; ++k) {
                const uint32_t ok = __shfl_sync(0xffffffffu, (uint32_t)emit, k);
                const uint32_t ik = __shfl_sync(0xffffffffu, e.x, k);
                epos += (ok && ik < e.x) ? 1 : 0;
                ecnt += ok ? 1 : 0;
            }
            __syncwarp();
            if (emit) slot[epos] = e;
            cnt = ecnt;
            if (lane == 0) count[j] = cnt;
        } else {  // rare: recount by recomputation, flag for the emit pass
            for (int base = s_lo; base < s_hi; base += 32) {
                const int sl = base + lane;
                bool emit = false;
                if (sl < s_hi) {
                    const uint32_t key = skey[sl];
                    const float dx = px - sxyz[3 * sl], dy = py - sxyz[3 * sl + 1], dz = pz - sxyz[3 * sl + 2];
                    const float yy = dy * dy, zz = dz * dz;
                    const float sq = dx * dx + (yy + zz);
                    const int li = (int)(key >> 16);
                    if ((sq < A.sq_radius_f) && (li - 1 <= lvl) && (lvl <= li + 1)) {
                        const uint32_t v = desc_distance(dj, A.cur_desc + 2 * (size_t)(key & 0xffffu), A.mode);
                        emit = __dmul_rn(A.accept_ratio, (double)(float)v) <= bestVal;
                    }
                }
                cnt += __popc(__ballot_sync(0xffffffffu, emit));
            }
            if (lane == 0) count[j] = cnt | (int)0x80000000;
        }
    }
}

// Pass 2 (grid): ordered emission.  Every CTA owns a contiguous range of map features: it sums the match counts of
// its range, publishes the sum stamped with the launch's epoch, obtains its output offset from the stamped sums of all
// preceding CTAs (decoupled look-back, see mapprep.cu), then scans its range 256 features at a time; the thread that
// scanned feature j copies j's cached matches to their place in the (j, i)-ordered output, flagged features are
// recomputed by whole warps.  header[0] = total, header[1] = perfect matches (best value 0; reference
// matcher.cpp:729-731), both written by the last CTA.  The grid never exceeds the SM count (co-residency).
constexpr int kEmitThreads = 256;
__global__ void __launch_bounds__(kEmitThreads)
guided_emit_kernel(GuidedArgs A, const uint32_t* __restrict__ best, const int* __restrict__ count,
                   const uint2* __restrict__ cache, int* __restrict__ offsets, int cap, int* __restrict__ header,
                   int* __restrict__ out_q, int* __restrict__ out_t, float* __restrict__ out_d, int per_cta,
                   unsigned long long* __restrict__ cta_sums /* 2 x gridDim.x: matches | perfect; [-1] = ticket word */, unsigned int epoch) {
    constexpr int kW = kEmitThreads / 32;
    __shared__ int warp_tot[kW], warp_perf[kW];
    __shared__ int s_base, n_ovf;
    __shared__ int ovf_list[kEmitThreads];
    __shared__ unsigned int s_bid;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = take_cta_ticket(cta_sums - 1, epoch);
    __syncthreads();
    const int bid = (int)s_bid;                  // logical CTA index: arrival order (common.cuh take_cta_ticket)
    const int M = A.M_dev ? min(*A.M_dev, A.M) : A.M;
    const int lo = min(M, bid * per_cta), hi = min(M, lo + per_cta);
    const uint32_t lt = (1u << lane) - 1u;

    // range totals
    int mine = 0, my_perfect = 0;
    for (int j = lo + tid; j < hi; j += kEmitThreads) {
        mine += count[j] & 0x7fffffff;
        my_perfect += (best[j] >> 16) == 0u ? 1 : 0;
    }
    mine = (int)warp_add_u32((uint32_t)mine);
    my_perfect = (int)warp_add_u32((uint32_t)my_perfect);
    if (lane == 0) { warp_tot[warp] = mine; warp_perf[warp] = my_perfect; }
    if (tid == 0) n_ovf = 0;
    __syncthreads();
    if (warp == 0) {
        int total = 0, perfect = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) { total += warp_tot[w]; perfect += warp_perf[w]; }
        volatile unsigned long long* sums = cta_sums;
        volatile unsigned long long* perfs = cta_sums + gridDim.x;
        if (lane == 0) {
            sums[bid] = ((unsigned long long)epoch << 32) | (unsigned int)total;
            perfs[bid] = ((unsigned long long)epoch << 32) | (unsigned int)perfect;
        }
        int before = 0, perf_before = 0;
        for (int b = lane; b < bid; b += 32) {
            unsigned long long v, p;
            do { v = sums[b]; } while ((unsigned int)(v >> 32) != epoch);
            do { p = perfs[b]; } while ((unsigned int)(p >> 32) != epoch);
            before += (int)(unsigned int)(v & 0xffffffffu);
            perf_before += (int)(unsigned int)(p & 0xffffffffu);
        }
        before = (int)warp_add_u32((uint32_t)before);
        perf_before = (int)warp_add_u32((uint32_t)perf_before);
        if (lane == 0) {
            s_base = before;
            if (bid == (int)gridDim.x - 1) {
                offsets[M] = before + total; header[0] = before + total; header[1] = perf_before + perfect;
            }
        }
    }
    __syncthreads();

    int carry = s_base;
    for (int base = lo; base < hi; base += kEmitThreads) {
        const int j = base + tid;
        const int cj = (j < hi) ? count[j] : 0;
        const int c = cj & 0x7fffffff;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();   // warp_tot / ovf_list of the previous pass consumed
        if (lane == 31) warp_tot[warp] = incl;
        if (tid == 0) n_ovf = 0;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
            const int cw = warp_tot[w];
            if (w < warp) woff += cw;
            tot += cw;
        }
        const int off = carry + woff + incl - c;
        if (j < hi) {
            offsets[j] = off;
            if (cj >= 0) {  // the thread that scanned feature j also copies its (few) cached matches
                for (int e = 0; e < c; ++e) {
                    const uint2 m = cache[(size_t)j * kCacheCap + e];
                    const int pos = off + e;
                    if (pos < cap) { out_q[pos] = j; out_t[pos] = (int)m.x; out_d[pos] = (float)m.y; }
                }
            } else if (c > 0) {
                ovf_list[atomicAdd(&n_ovf, 1)] = j;   // at most kEmitThreads flagged features per pass
            }
        }
        carry += tot;
        __syncthreads();
        // features with more candidates than the cache holds: recomputed by whole warps, in any order
        // (their output positions are fixed by `offsets`)
        const int novf = n_ovf;
        for (int a = warp; a < novf; a += kW) {
            const int jj = ovf_list[a];
            const float px = A.map_xyz[3 * jj], py = A.map_xyz[3 * jj + 1], pz = A.map_xyz[3 * jj + 2];
            const int lvl = A.map_level[jj];
            const uint4* dj = A.map_desc + 2 * (size_t)jj;
            const double bestVal = (double)(float)(best[jj] >> 16);
            int run = offsets[jj];
            for (int b0 = 0; b0 < A.N; b0 += 32) {
                const int i = b0 + lane;
                bool emit = false;
                uint32_t v = 0;
                if (i < A.N && gate(A, px, py, pz, lvl, A.cur_xyz, A.cur_level, i)) {
                    v = desc_distance(dj, A.cur_desc + 2 * (size_t)i, A.mode);
                    emit = __dmul_rn(A.accept_ratio, (double)(float)v) <= bestVal;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, emit);
                if (emit) {
                    const int pos = run + __popc(bal & lt);
                    if (pos < cap) { out_q[pos] = jj; out_t[pos] = i; out_d[pos] = (float)v; }
                }
                run += __popc(bal);
            }
        }
    }
}

// Predicted ORB pyramid levels (reference src/Matcher/matcher.cpp:639-651 for current keypoints, :682-692 for map
// features) and the double->float cast of the map positions (:665), on the device.
//   level = clamp(ceil(log(pow(1.2, octave) * detDist / curDist) / log(1.2)), 0, 7)       (all double)
// pow(1.2, k) and the level for the exact case s == pow(1.2, k) come from host-libm tables, because that case --
// detDist == curDist, common for keypoints matched in the frame they were detected in -- sits exactly on a ceil()
// boundary and depends on the last bit of the host's log().  Everywhere else the argument is at least one float
// ulp (6e-8) away from the boundary, eight orders of magnitude more than any log() implementation's error.
struct LevelTables {
    double pow_tab[16];     // pow(1.2, k) from the host libm
    int lvl_tab[16];        // (int)ceil(log(pow(1.2, k)) / log(1.2)) from the host libm
    double log_sf;          // log(1.2) from the host libm
};
__device__ __forceinline__ int predict_level(const LevelTables& T, int octave, double det_dist, double cur_dist) {
    const double ps = (octave >= 0 && octave < 16) ? T.pow_tab[octave] : pow(1.2, (double)octave);
    const double sfac = __ddiv_rn(ps * det_dist, cur_dist);
    int lvl;
    if (octave >= 0 && octave < 16 && sfac == ps) lvl = T.lvl_tab[octave];
    else lvl = (int)ceil(__ddiv_rn(log(sfac), T.log_sf));
    lvl = max(0, lvl);
    return min(7, lvl);
}
__global__ void predict_levels_kernel(const double* __restrict__ map_xyz, const int* __restrict__ map_oct,
                                      const double* __restrict__ map_det, int M, const float* __restrict__ cur_xyz,
                                      const int* __restrict__ cur_oct, const double* __restrict__ cur_det, int N,
                                      LevelTables T, float* __restrict__ map_xyz_f, int* __restrict__ map_level,
                                      int* __restrict__ cur_level, const int* __restrict__ M_dev) {
    chain_begin();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        if (M_dev && i >= *M_dev) return;   // M is the capacity; the filtered count lives on the device
        const double x = map_xyz[3 * i], y = map_xyz[3 * i + 1], z = map_xyz[3 * i + 2];
        const double cur_dist = __dsqrt_rn(x * x + y * y + z * z);
        map_level[i] = predict_level(T, map_oct[i], map_det[i], cur_dist);
        map_xyz_f[3 * i] = (float)x; map_xyz_f[3 * i + 1] = (float)y; map_xyz_f[3 * i + 2] = (float)z;
    } else if (i < M + N) {
        const int k = i - M;
        const float x = cur_xyz[3 * k], y = cur_xyz[3 * k + 1], z = cur_xyz[3 * k + 2];
        const double cur_dist = (double)norm3(x, y, z);     // Eigen Vector3f::norm(), widened
        cur_level[k] = predict_level(T, cur_oct[k], cur_det[k], cur_dist);
    }
}

cudaError_t launch_predict_levels(const double* d_map_xyz, const int* d_map_oct, const double* d_map_det, int M,
                                  const float* d_cur_xyz, const int* d_cur_oct, const double* d_cur_det, int N,
                                  const double* pow_tab, const int* lvl_tab, double log_sf, float* d_map_xyz_f,
                                  int* d_map_level, int* d_cur_level, cudaStream_t st, int* launches, const int* d_M) {
    if (M + N <= 0) return cudaSuccess;
    LevelTables T;
    for (int k = 0; k < 16; ++k) { T.pow_tab[k] = pow_tab[k]; T.lvl_tab[k] = lvl_tab[k]; }
    T.log_sf = log_sf;
    const cudaError_t e = launch_chained(predict_levels_kernel, dim3((unsigned)((M + N + 255) / 256)), dim3(256), 0, st,
                                         d_map_xyz, d_map_oct, d_map_det, M, d_cur_xyz, d_cur_oct, d_cur_det, N, T,
                                         d_map_xyz_f, d_map_level, d_cur_level, d_M);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

size_t guided_cache_bytes(int M) { return sizeof(uint2) * (size_t)kCacheCap * (size_t)(M > 0 ? M : 1); }

cudaError_t launch_guided_match(const float* d_map_xyz, const uint8_t* d_map_desc, const int* d_map_level, int M,
                                const float* d_cur_xyz, const uint8_t* d_cur_desc, const int* d_cur_level, int N,
                                float sq_radius_f, double accept_ratio, int mode, int* d_count, int* d_best, void* d_cache,
                                int* d_out, int cap, unsigned long long* d_cta_sums, unsigned int epoch, int sm_count,
                                cudaStream_t st, int* launches, const int* d_M) {
    GuidedArgs A;
    A.map_xyz = d_map_xyz; A.map_desc = reinterpret_cast<const uint4*>(d_map_desc); A.map_level = d_map_level; A.M = M;
    A.M_dev = d_M;
    A.cur_xyz = d_cur_xyz; A.cur_desc = reinterpret_cast<const uint4*>(d_cur_desc); A.cur_level = d_cur_level; A.N = N;
    A.sq_radius_f = sq_radius_f; A.accept_ratio = accept_ratio; A.mode = mode;
    {   // a keypoint inside the sphere has fl(dz*dz) <= fl(dx*dx + (dy*dy + dz*dz)) < sq_radius_f (rounding is monotone),
        // hence |dz| < sqrt(sq_radius_f) (1 + 2^-23); the margins below also cover the rounding of pz -+ z_reach
        const double reach = sqrt((double)sq_radius_f) * 1.001 + 1e-4;
        A.z_reach = (float)reach;
        if ((double)A.z_reach < reach) A.z_reach = nextafterf(A.z_reach, INFINITY);
        if (!(A.z_reach == A.z_reach)) A.z_reach = INFINITY;   // NaN radius: no gate can pass; visit everything
    }
    // d_count: M counts followed by M+1 offsets
    int* d_offsets = d_count + M;
    int* out_q = d_out + 2;
    int* out_t = d_out + 2 + cap;
    float* out_d = reinterpret_cast<float*>(d_out + 2 + 2 * cap);
    const size_t smem = sizeof(float) * 4 * (size_t)N + sizeof(int) * (2 * kZBins + 1);
    int grid = (M + kGWarps - 1) / kGWarps;
    if (grid > 4 * PSLAM_SM_COUNT_HINT) grid = 4 * PSLAM_SM_COUNT_HINT;
    if (grid < 1) grid = 1;
    cudaError_t e;
    if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(guided_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    }
    if ((e = launch_chained(guided_collect_kernel, dim3((unsigned)grid), dim3(kGThreads), smem, st, A,
                            reinterpret_cast<uint32_t*>(d_best), d_count, reinterpret_cast<uint2*>(d_cache))) != cudaSuccess)
        return e;
    // emission: contiguous feature ranges (a multiple of the CTA size), at most one CTA per SM
    int egrid = (M + kEmitThreads - 1) / kEmitThreads;
    if (egrid > sm_count) egrid = sm_count;
    if (egrid < 1) egrid = 1;
    int per_cta = (M + egrid - 1) / egrid;
    per_cta = (per_cta + kEmitThreads - 1) / kEmitThreads * kEmitThreads;
    if (per_cta < kEmitThreads) per_cta = kEmitThreads;
    egrid = (M + per_cta - 1) / per_cta;
    if (egrid < 1) egrid = 1;
    if ((e = launch_chained(guided_emit_kernel, dim3((unsigned)egrid), dim3(kEmitThreads), 0, st, A,
                            reinterpret_cast<const uint32_t*>(d_best), (const int*)d_count,
                            reinterpret_cast<const uint2*>(d_cache), d_offsets, cap, d_out, out_q, out_t, out_d, per_cta,
                            d_cta_sums, epoch)) != cudaSuccess)
        return e;
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace pslam
