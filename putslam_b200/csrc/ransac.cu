// ransac.cu -- RANSAC hypothesis scoring and model selection/refit (sm_100a).
//
//   ransac_filter_kernel  : drop matches with NaN / z outside [0.1, 6] and gather the surviving pairs
//                           into SoA form (reference src/TransformEst/RANSAC.cpp:65-80)
//   K4a ransac_model_kernel: one thread per hypothesis -- counter-based sample of 3 matches, 3-point Umeyama
//   K4b ransac_score_kernel: one lane per hypothesis, matches broadcast from shared memory, integer RED per hypothesis
//                           (reference RANSAC.cpp:87-150 loop body, :180-281, :325-436)
//   K5 ransac_select_kernel: replay of saveBetterModel / iterationCount (RANSAC.cpp:438-461) over the
//                           per-hypothesis counts, inlier list of the winner, Umeyama refit over all its
//                           inliers, Euclidean recount restricted to them, ratio gate (RANSAC.cpp:152-164)
//
// All float arithmetic is single-rounding (-fmad=false); summation orders are sequential exactly as
// written in DESIGN.md, so inlier sets are reproducible bit for bit.
#include <math.h>

#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

namespace pslam {

constexpr int kHdrInts = 24;  // result header: see layout below
// result layout (ints): [0] n_inliers  [1] hyp_used  [2] n_filtered  [3] best_count  [4..19] T (col-major float)
//                       [20..21] best_ratio (double)  [22] winner hypothesis  [23] flags   [24..] inlier match idx

size_t ransac_result_ints(int m_cap) { return (size_t)kHdrInts + (size_t)(m_cap > 0 ? m_cap : 1); }

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
ransac_filter_kernel(const float* __restrict__ prev, const float* __restrict__ cur, const int* __restrict__ mq,
                     const int* __restrict__ mt, const int* __restrict__ d_m, int m_host, int m_cap,
                     float* __restrict__ pts, int* __restrict__ keep, int* __restrict__ n_filtered) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int m = d_m ? *d_m : m_host;
    if (m > m_cap) m = m_cap;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int k = base + tid;
        bool ok = false;
        float p[3] = {0, 0, 0}, c[3] = {0, 0, 0};
        if (k < m) {
            const int qi = mq[k], ti = mt[k];
#pragma unroll
            for (int a = 0; a < 3; ++a) { p[a] = prev[3 * qi + a]; c[a] = cur[3 * ti + a]; }
            const bool bad = isnan(p[0]) || isnan(p[1]) || isnan(p[2]) || isnan(c[0]) || isnan(c[1]) || isnan(c[2]) ||
                             (double)p[2] < 0.1 || p[2] > 6.f || (double)c[2] < 0.1 || c[2] > 6.f;
            ok = !bad;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int cw = warp_tot[w];
            if (w < warp) woff += cw;
            tot += cw;
        }
        if (ok) {
            const int pos = carry + woff + wpre;
            keep[pos] = k;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pts[(size_t)a * m_cap + pos] = p[a];
                pts[(size_t)(3 + a) * m_cap + pos] = c[a];
            }
        }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) *n_filtered = carry;
}

// ------------------------------------------------------------------------------------------------
struct Scorer {
    int ev;
    float thr_f;
    float sq_thr_f;   // smallest float T with sqrtf(T) >= thr_f:  sqrtf(s) < thr_f  <=>  s < T  (sqrt is monotone)
    double thr, thr_reproj;
    float fx, fy, cx, cy;
};

__device__ __forceinline__ void project(const Scorer& S, float x, float y, float z, float& u, float& v) {
    u = __fdiv_rn(x * S.fx, z) + S.cx;  // RGBD::point3Dto2D, reference src/RGBD/RGBD.cpp:92-98
    v = __fdiv_rn(y * S.fy, z) + S.cy;
}

// One inlier test; force_euclid reproduces the refit recount (always the Euclidean routine, which still
// applies the ADAPTIVE depth scaling).  Tinv is R|t of the general inverse, only read for ev 1 and 2.
__device__ __forceinline__ bool inlier_test(const Scorer& S, const float (&R)[9], const float (&t)[3],
                                            const float (&Ri)[9], const float (&ti)[3], float px, float py, float pz,
                                            float cx, float cy, float cz, bool force_euclid) {
    float ex, ey, ez;
    rigid_apply(R, t, cx, cy, cz, ex, ey, ez);
    if (force_euclid || S.ev == 0 || S.ev == 4) {
        const float dx = ex - px, dy = ey - py, dz = ez - pz;
        if (S.ev == 4) return (double)norm3(dx, dy, dz) < S.thr * (double)pz;
        // (double)sqrtf(s) < thr  <=>  sqrtf(s) < thr_f  <=>  s < sq_thr_f : same predicate, no square root
        const float yy = dy * dy, zz = dz * dz;
        return dx * dx + (yy + zz) < S.sq_thr_f;
    }
    float nx, ny, nz;
    rigid_apply(Ri, ti, px, py, pz, nx, ny, nz);
    float pnu, pnv, rnu, rnv, pou, pov, rou, rov;
    project(S, nx, ny, nz, pnu, pnv);
    project(S, cx, cy, cz, rnu, rnv);
    project(S, ex, ey, ez, pou, pov);
    project(S, px, py, pz, rou, rov);
    const float ax = pnu - rnu, ay = pnv - rnv, bx = pou - rou, by = pov - rov;
    const double e0 = __dsqrt_rn((double)ax * (double)ax + (double)ay * (double)ay);
    const double e1 = __dsqrt_rn((double)bx * (double)bx + (double)by * (double)by);
    const bool ok2d = e0 < S.thr_reproj && e1 < S.thr_reproj;
    if (S.ev == 1) return ok2d;
    const double e3 = (double)norm3(ex - px, ey - py, ez - pz);
    return e3 < S.thr && ok2d;
}

__device__ __forceinline__ void model_inverse(const Rigid3f& M, float (&Ri)[9], float (&ti)[3]) {
    float T[16], Tin[16];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) T[4 * i + j] = M.R[3 * i + j];
        T[4 * i + 3] = M.t[i];
    }
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    inverse4(T, Tin);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Ri[3 * i + j] = Tin[4 * i + j];
        ti[i] = Tin[4 * i + 3];
    }
}

__device__ __forceinline__ Rigid3f hypothesis_model(const float* __restrict__ pts, int m_cap, int mf, uint32_t seed_lo,
                                                    uint32_t seed_hi, uint32_t h) {
    int s[3];
    sample3(seed_lo, seed_hi, h, (uint32_t)mf, s);
    float src[3][3], dst[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            dst[k][a] = pts[(size_t)a * m_cap + s[k]];
            src[k][a] = pts[(size_t)(3 + a) * m_cap + s[k]];
        }
    return umeyama3(src, dst);
}

constexpr int kScoreThreads = 256;
constexpr int kModelThreads = 64;

// K4a: one THREAD per hypothesis -- counter-based sample of 3 matches and the 3-point Umeyama model
// (Jacobi SVD is ~2000 dependent instructions: running it once per thread instead of redundantly on the
// 32 lanes of a scoring warp removes most of the stage's instruction count).  counts[h] = -1 marks a
// degenerate model (NaN, reference RANSAC.cpp:238-242), 0 otherwise.
__global__ void __launch_bounds__(kModelThreads)
ransac_model_kernel(const float* __restrict__ pts, int m_cap, const int* __restrict__ n_filtered, int min_matches,
                    uint32_t seed_lo, uint32_t seed_hi, int H, int* __restrict__ counts, float* __restrict__ models) {
    chain_begin();
    const int mf = *n_filtered;
    if (mf < min_matches || mf < 3) return;
    const int h = blockIdx.x * kModelThreads + threadIdx.x;
    if (h >= H) return;
    const Rigid3f M = hypothesis_model(pts, m_cap, mf, seed_lo, seed_hi, (uint32_t)h);
    counts[h] = M.ok ? 0 : -1;
    float4* mp = reinterpret_cast<float4*>(models + 12 * (size_t)h);
    mp[0] = make_float4(M.R[0], M.R[1], M.R[2], M.R[3]);
    mp[1] = make_float4(M.R[4], M.R[5], M.R[6], M.R[7]);
    mp[2] = make_float4(M.R[8], M.t[0], M.t[1], M.t[2]);
}

// K4b: one LANE per hypothesis, the matches broadcast from shared memory.  A CTA owns 32 hypotheses (lane = hypothesis,
// model in registers) and one slice of the match list (blockIdx.y); it stages the slice from the SoA arrays into
// shared memory as {prev.xyz, cur.x | cur.yz}, and its 8 warps walk interleaved matches: every LDS is a warp-wide
// broadcast that serves 32 hypothesis x match tests, so there is no global load and no warp reduction in the loop
// (the previous warp-per-hypothesis layout spent 78 % of its cycles on the per-match global loads, ncu).  The 8 partial
// counts per hypothesis are added in shared memory, then one integer RED per hypothesis into counts[] (zeroed by the
// model kernel) -- integer sums, so the result does not depend on the order.
constexpr int kScoreTile = 512;   // matches staged per pass: 12 KB
__global__ void __launch_bounds__(kScoreThreads)
ransac_score_kernel(const float* __restrict__ pts, int m_cap, const int* __restrict__ n_filtered, int min_matches,
                    Scorer S, int H, int* __restrict__ counts, const float* __restrict__ models) {
    __shared__ float4 sA[kScoreTile];   // prev.x prev.y prev.z cur.x
    __shared__ float2 sB[kScoreTile];   // cur.y cur.z
    __shared__ int s_cnt[32];
    chain_begin();
    const int mf = *n_filtered;
    if (mf < min_matches || mf < 3) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kScoreThreads / 32;
    const int h = blockIdx.x * 32 + lane;
    const int per = (mf + (int)gridDim.y - 1) / (int)gridDim.y;
    const int lo = blockIdx.y * per;
    const int hi = min(mf, lo + per);
    if (lo >= hi) return;

    bool live = h < H;
    if (live) live = counts[h] >= 0;   // -1: degenerate model
    Rigid3f M;
    {
        const float4* mp = reinterpret_cast<const float4*>(models + 12 * (size_t)(h < H ? h : 0));
        const float4 m0 = mp[0], m1 = mp[1], m2 = mp[2];
        M.R[0] = m0.x; M.R[1] = m0.y; M.R[2] = m0.z; M.R[3] = m0.w; M.R[4] = m1.x; M.R[5] = m1.y; M.R[6] = m1.z;
        M.R[7] = m1.w; M.R[8] = m2.x; M.t[0] = m2.y; M.t[1] = m2.z; M.t[2] = m2.w; M.ok = true;
    }
    float Ri[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ti[3] = {0, 0, 0};
    if (S.ev == 1 || S.ev == 2) model_inverse(M, Ri, ti);
    if (tid < 32) s_cnt[tid] = 0;

    int c = 0;
    for (int base = lo; base < hi; base += kScoreTile) {
        const int nt = min(kScoreTile, hi - base);
        __syncthreads();   // previous tile consumed (and s_cnt initialised)
        for (int k = tid; k < nt; k += kScoreThreads) {
            const int g = base + k;
            sA[k] = make_float4(pts[g], pts[(size_t)m_cap + g], pts[2 * (size_t)m_cap + g], pts[3 * (size_t)m_cap + g]);
            sB[k] = make_float2(pts[4 * (size_t)m_cap + g], pts[5 * (size_t)m_cap + g]);
        }
        __syncthreads();
#pragma unroll 2
        for (int k = warp; k < nt; k += kWarps) {
            const float4 a = sA[k];
            const float2 b = sB[k];
            c += inlier_test(S, M.R, M.t, Ri, ti, a.x, a.y, a.z, a.w, b.x, b.y, false) ? 1 : 0;
        }
    }
    if (live && c) atomicAdd(&s_cnt[lane], c);
    __syncthreads();
    if (tid < 32 && live && s_cnt[tid]) atomicAdd(&counts[h], s_cnt[tid]);
}

// computeRANSACIteration (reference RANSAC.cpp:457-461): int(log(1-0.98) / log(1 - w^3)).  The
// reference's conversion is undefined behaviour once the quotient leaves the int range (tiny ratios);
// it is defined here as the saturating conversion (deliberate divergences).
__device__ __forceinline__ int ransac_iterations(double w) {
    const double v = log(1 - 0.98) / log(1 - pow(w, 3.0));
    if (v != v) return (int)0x80000000;
    if (v >= 2147483648.0) return 0x7fffffff;
    if (v <= -2147483649.0) return (int)0x80000000;
    return (int)v;
}

// USAC<T>::updateStandardStopping (reference include/putslam/USAC/USAC.h:944-971), sample size 3.
__device__ __forceinline__ unsigned usac_standard_stopping(unsigned inl, unsigned tot, double conf, unsigned max_hyp) {
    double n_inl = 1.0, n_pts = 1.0;
#pragma unroll
    for (unsigned i = 0; i < 3; ++i) {
        n_inl = n_inl * (double)(inl - i);   // unsigned wrap-around like the reference
        n_pts = n_pts * (double)(tot - i);
    }
    const double prob = n_inl / n_pts;
    if (prob < 2.220446049250313e-16) return max_hyp;
    if (1 - prob < 2.220446049250313e-16) return 1u;
    return (unsigned)ceil(log(1 - conf) / log(1 - prob));
}

constexpr int kSelThreads = 1024;

__global__ void __launch_bounds__(kSelThreads, 1)
ransac_select_kernel(const float* __restrict__ pts, int m_cap, const int* __restrict__ keep,
                     const int* __restrict__ n_filtered, const int* __restrict__ counts,
                     const float* __restrict__ models, int H, int adaptive, int stop_rule, double usac_conf,
                     int min_matches, double min_ratio, Scorer S, uint32_t seed_lo, uint32_t seed_hi,
                     int* __restrict__ inl_tmp /* m_cap scratch */, int stage_cap, int* __restrict__ result) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    __shared__ unsigned long long best_key;
    __shared__ int s_win, s_used, s_cnt;
    __shared__ float s_mean[6];
    __shared__ float s_sig[9];
    __shared__ float s_R[9], s_t[3];
    __shared__ int s_ok;

    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mf = *n_filtered;
    float* Tout = reinterpret_cast<float*>(result + 4);
    int* inl_out = result + kHdrInts;

    auto write_identity = [&](int used) {
        if (tid < 16) Tout[tid] = (tid % 5 == 0) ? 1.f : 0.f;
        if (tid == 0) {
            result[0] = 0; result[1] = used; result[2] = mf; result[3] = 0; result[22] = -1; result[23] = 0;
            *reinterpret_cast<double*>(result + 20) = 0.0;
        }
    };
    if (mf < min_matches || mf < 3) {  // RANSAC.cpp:77-80 (3 is the sample size; guards the sampler)
        write_identity(0);
        return;
    }

    // ---- winner: replay of the sequential loop over the scored hypotheses ----
    if (tid == 0) { best_key = 0ull; s_win = -1; s_used = H; s_cnt = 0; carry = 0; }
    __syncthreads();
    if (adaptive) {
        // the replay is sequential by definition; stage the counts in shared memory first so that thread 0 does not
        // pay a global-memory latency per step (the adaptive bound is at most 487, USAC replay uses the budget H)
        __shared__ int s_counts[1024];
        const int n_stage_c = H < 1024 ? H : 1024;
        for (int i = tid; i < n_stage_c; i += kSelThreads) s_counts[i] = counts[i];
        __syncthreads();
        if (tid == 0) {
            int bound = H, win = -1, bc = 0;
            double best = 0.0;
            int i = 0;
            for (; i < bound; ++i) {
                const int c = i < n_stage_c ? s_counts[i] : counts[i];
                if (c < 0) continue;
                const float ratio = __fdiv_rn((float)c, (float)mf);
                if ((double)ratio > best) {
                    best = (double)ratio; win = i; bc = c;
                    if (stop_rule == 1) {
                        const unsigned b = usac_standard_stopping((unsigned)c, (unsigned)mf, usac_conf, (unsigned)H);
                        bound = (int)(b < (unsigned)H ? b : (unsigned)H);
                    } else if (adaptive == 1) {
                        const int a = ransac_iterations(min_ratio), b = ransac_iterations(best);
                        bound = a < b ? a : b;
                    }
                }
            }
            s_win = win; s_used = i; s_cnt = bc;
        }
    } else {
        // fixed bound: first maximum of float(c)/float(mf) == first maximum of c (monotone, c <= mf < 2^24)
        unsigned long long loc = 0ull;
        for (int i = tid; i < H; i += kSelThreads) {
            const int c = counts[i];
            if (c > 0) {
                const unsigned long long key = ((unsigned long long)(uint32_t)c << 32) | (0xffffffffu - (uint32_t)i);
                if (key > loc) loc = key;
            }
        }
        // warp maximum first: a 64-bit shared atomicMax is a CAS loop, and 1024 threads contending on one address
        // made this step the longest of the kernel
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, loc, d);
            if (o > loc) loc = o;
        }
        if (lane == 0 && loc) atomicMax(&best_key, loc);
        __syncthreads();
        if (tid == 0 && best_key) {
            s_cnt = (int)(best_key >> 32);
            s_win = (int)(0xffffffffu - (uint32_t)(best_key & 0xffffffffu));
        }
    }
    __syncthreads();
    const int win = s_win, used = s_used, best_cnt = s_cnt;
    if (win < 0) {  // no hypothesis scored above zero: refit on the empty set fails -> identity (RANSAC.cpp:152-164)
        write_identity(used);
        return;
    }
    const float best_ratio_f = __fdiv_rn((float)best_cnt, (float)mf);
    const double best_ratio = (double)best_ratio_f;

    const float* px = pts;
    const float* py = pts + (size_t)m_cap;
    const float* pz = pts + 2 * (size_t)m_cap;
    const float* cx = pts + 3 * (size_t)m_cap;
    const float* cy = pts + 4 * (size_t)m_cap;
    const float* cz = pts + 5 * (size_t)m_cap;

    // ---- inlier list of the winner (ordered); its model was stored by the scoring kernel ----
    Rigid3f M;
#pragma unroll
    for (int i = 0; i < 9; ++i) M.R[i] = models[12 * (size_t)win + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) M.t[i] = models[12 * (size_t)win + 9 + i];
    M.ok = true;
    float Ri[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ti[3] = {0, 0, 0};
    if (S.ev == 1 || S.ev == 2) model_inverse(M, Ri, ti);
    for (int base = 0; base < mf; base += kSelThreads) {
        const int k = base + tid;
        const bool in = (k < mf) && inlier_test(S, M.R, M.t, Ri, ti, px[k], py[k], pz[k], cx[k], cy[k], cz[k], false);
        const uint32_t bal = __ballot_sync(0xffffffffu, in);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int cw = warp_tot[w];
            if (w < warp) woff += cw;
            tot += cw;
        }
        if (in) inl_tmp[carry + woff + wpre] = k;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    const int n_in = carry;
    __syncthreads();

    // ---- refit: Umeyama over all inliers; every sum is a sequential float chain, one thread per chain ----
    // The inlier coordinates are first gathered into shared memory (all threads, coalesced index reads) so
    // that the 15 chain threads stream them with pipelined LDS instead of dependent global loads.
    extern __shared__ float s_pts[];                      // 6 x n_stage (dst xyz | src xyz), SoA
    const int n_stage = n_in <= stage_cap ? n_in : 0;     // too many inliers for shared memory: read global
    for (int k = tid; k < n_stage; k += kSelThreads) {
        const int id = inl_tmp[k];
#pragma unroll
        for (int a = 0; a < 6; ++a) s_pts[a * n_stage + k] = pts[(size_t)a * m_cap + id];
    }
    __syncthreads();
    const float one_over_n = __fdiv_rn(1.f, (float)n_in);
    if (tid < 6) {
        float s = 0.f;
        if (n_stage) {
            const float* col = s_pts + tid * n_stage;
#pragma unroll 8
            for (int k = 0; k < n_in; ++k) s = s + col[k];
        } else {
            const float* col = pts + (size_t)tid * m_cap;  // 0..2 prev (dst), 3..5 cur (src)
            for (int k = 0; k < n_in; ++k) s = s + col[inl_tmp[k]];
        }
        s_mean[tid] = s * one_over_n;
    }
    __syncthreads();
    if (tid < 9) {
        const int i = tid / 3, j = tid % 3;
        const float dmean = s_mean[i], smean = s_mean[3 + j];
        float s = 0.f;
        if (n_stage) {
            const float* dcol = s_pts + i * n_stage;
            const float* scol = s_pts + (3 + j) * n_stage;
#pragma unroll 8
            for (int k = 0; k < n_in; ++k) s = s + (dcol[k] - dmean) * (scol[k] - smean);
        } else {
            const float* dcol = pts + (size_t)i * m_cap;
            const float* scol = pts + (size_t)(3 + j) * m_cap;
            for (int k = 0; k < n_in; ++k) {
                const int id = inl_tmp[k];
                s = s + (dcol[id] - dmean) * (scol[id] - smean);
            }
        }
        s_sig[tid] = one_over_n * s;
    }
    __syncthreads();
    if (tid == 0) {
        float sig[9], sm[3], dm[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) sig[i] = s_sig[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) { dm[i] = s_mean[i]; sm[i] = s_mean[3 + i]; }
        const Rigid3f F = umeyama_from_sigma(sig, sm, dm);
        s_ok = F.ok ? 1 : 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) s_R[i] = F.ok ? F.R[i] : ((i % 4 == 0) ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) s_t[i] = F.ok ? F.t[i] : 0.f;
        carry = 0;
    }
    __syncthreads();
    float R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = s_R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = s_t[i];

    // ---- Euclidean recount restricted to the winner's inliers (RANSAC.cpp:155-157), ordered ----
    const bool gate = !(best_ratio < min_ratio);  // RANSAC.cpp:161
    for (int base = 0; base < n_in; base += kSelThreads) {
        const int a = base + tid;
        int k = 0;
        bool in = false;
        if (a < n_in) {
            k = inl_tmp[a];
            in = inlier_test(S, R, t, Ri, ti, px[k], py[k], pz[k], cx[k], cy[k], cz[k], true);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, in);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int cw = warp_tot[w];
            if (w < warp) woff += cw;
            tot += cw;
        }
        if (in && gate) inl_out[carry + woff + wpre] = keep[k];
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid < 16) {  // column-major 4x4 (Eigen::Matrix4f layout)
        const int r = tid % 4, c = tid / 4;
        float v = (r == c) ? 1.f : 0.f;
        if (gate) {
            if (r < 3 && c < 3) v = R[3 * r + c];
            else if (r < 3 && c == 3) v = t[r];
        }
        Tout[tid] = v;
    }
    if (tid == 0) {
        result[0] = gate ? carry : 0;
        result[1] = used;
        result[2] = mf;
        result[3] = best_cnt;
        *reinterpret_cast<double*>(result + 20) = best_ratio;
        result[22] = win;
        result[23] = s_ok;
    }
}

cudaError_t launch_ransac(const float* d_prev, const float* d_cur, const int* d_mq, const int* d_mt, const int* d_m,
                          int m_host, const RansacDeviceParams& P, const RansacWorkspace& ws, int sm_count,
                          cudaStream_t st, int* launches) {
    Scorer S;
    S.ev = P.error_version;
    S.thr_f = P.thr_euclid_f;
    S.sq_thr_f = P.sq_thr_euclid_f;
    S.thr = P.thr_euclid;
    S.thr_reproj = P.thr_reproj;
    S.fx = P.fx; S.fy = P.fy; S.cx = P.cx; S.cy = P.cy;
    // sequential replay is needed for the reference's adaptive bound and for USAC stopping (2 = replay with a fixed
    // budget); plain fixed-H is a parallel first-max
    const int adaptive = P.num_hyp <= 0 ? 1 : (P.stop_rule == 1 ? 2 : 0);
    const int H = adaptive ? 487 : P.num_hyp;  // int(log(0.02)/log(1-0.2^3)), reference RANSAC.cpp:30
    cudaError_t e;
    if ((e = launch_chained(ransac_filter_kernel, dim3(1), dim3(1024), 0, st, d_prev, d_cur, d_mq, d_mt, d_m, m_host,
                            ws.m_cap, ws.pts, ws.keep, ws.n_filtered)) != cudaSuccess) return e;
    if ((e = launch_chained(ransac_model_kernel, dim3((unsigned)((H + kModelThreads - 1) / kModelThreads)),
                            dim3(kModelThreads), 0, st, (const float*)ws.pts, ws.m_cap, (const int*)ws.n_filtered,
                            P.min_matches, P.seed_lo, P.seed_hi, H, ws.counts, ws.models)) != cudaSuccess) return e;
    // 32 hypotheses per CTA; the match list is cut into as many slices as put up to 4 CTAs on every SM, but
    // no slice shorter than ~64 matches (the filtered count is only known on the device: bound it by the number of
    // matches handed in, or by the capacity of the list when that number is device-resident too)
    const int hyp_ctas = (H + 31) / 32;
    int slices = (4 * sm_count) / hyp_ctas;   // rounded down: 4 CTAs of 256 threads x 64 registers fit an SM -> one wave
    const int m_bound = d_m ? ws.m_cap : m_host;
    const int max_slices = m_bound / 64 > 1 ? m_bound / 64 : 1;
    if (slices > max_slices) slices = max_slices;
    if (slices > 64) slices = 64;
    if (slices < 1) slices = 1;
    if ((e = launch_chained(ransac_score_kernel, dim3((unsigned)hyp_ctas, (unsigned)slices), dim3(kScoreThreads), 0, st,
                            (const float*)ws.pts, ws.m_cap, (const int*)ws.n_filtered, P.min_matches, S, H, ws.counts,
                            (const float*)ws.models)) != cudaSuccess) return e;
    // shared-memory staging of the winner's inliers for the refit: up to 8192 inliers (192 KB)
    int stage_cap = ws.m_cap < 8192 ? ws.m_cap : 8192;
    const size_t sel_smem = sizeof(float) * 6 * (size_t)stage_cap;
    // per-device function attribute; setting it again is harmless, so no cross-thread state is kept
    if ((e = cudaFuncSetAttribute(ransac_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 8192 * 4)) != cudaSuccess)
        return e;
    if ((e = launch_chained(ransac_select_kernel, dim3(1), dim3(kSelThreads), sel_smem, st, ws.pts, ws.m_cap, ws.keep,
                            ws.n_filtered, ws.counts, ws.models, H, adaptive, P.stop_rule, P.usac_conf, P.min_matches,
                            P.min_inlier_ratio, S, P.seed_lo, P.seed_hi,
                            ws.keep + ws.m_cap /* scratch: second half of keep */, stage_cap, ws.result)) != cudaSuccess)
        return e;
    if (launches) *launches += 4;
    return cudaGetLastError();
}

}  // namespace pslam
