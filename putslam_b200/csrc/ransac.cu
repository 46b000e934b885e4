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

constexpr int kHdrInts = kRansacHdrInts;  // result header: layout in kernels.h

size_t ransac_result_ints(int m_cap) { return (size_t)kHdrInts + (size_t)(m_cap > 0 ? m_cap : 1); }

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
ransac_filter_kernel(const float* __restrict__ prev, const float* __restrict__ cur, const int* __restrict__ mq,
                     const int* __restrict__ mt, const int* __restrict__ d_m, int m_host, int m_cap,
                     float* __restrict__ pts, int* __restrict__ keep, int* __restrict__ n_filtered) {
    // Ordered compaction, 2048 matches per pass: every thread takes match k of the first 1024 and match k + 1024 of the
    // second, two ballots per warp, one barrier pair per pass (a pass is all latency -- dependent loads and barriers --
    // so the second match per thread is almost free; 1000-2000 matches, the usual frame, need one pass).
    __shared__ int warp_tot[2][32];
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int m = d_m ? *d_m : m_host;
    if (m > m_cap) m = m_cap;
    int carry = 0;   // identical in every thread
    for (int base = 0; base < m; base += 2048) {
        bool ok[2] = {false, false};
        float p[2][3], c[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = base + 1024 * h + tid;
#pragma unroll
            for (int a = 0; a < 3; ++a) { p[h][a] = 0.f; c[h][a] = 0.f; }
            if (k < m) {
                const int qi = mq[k], ti = mt[k];
#pragma unroll
                for (int a = 0; a < 3; ++a) { p[h][a] = prev[3 * qi + a]; c[h][a] = cur[3 * ti + a]; }
                const bool bad = isnan(p[h][0]) || isnan(p[h][1]) || isnan(p[h][2]) || isnan(c[h][0]) || isnan(c[h][1]) ||
                                 isnan(c[h][2]) || (double)p[h][2] < 0.1 || p[h][2] > 6.f || (double)c[h][2] < 0.1 ||
                                 c[h][2] > 6.f;
                ok[h] = !bad;
            }
        }
        const uint32_t bal0 = __ballot_sync(0xffffffffu, ok[0]), bal1 = __ballot_sync(0xffffffffu, ok[1]);
        __syncthreads();   // warp_tot of the previous pass consumed
        if (lane == 0) { warp_tot[0][warp] = __popc(bal0); warp_tot[1][warp] = __popc(bal1); }
        __syncthreads();
        int woff0 = 0, woff1 = 0, tot0 = 0, tot1 = 0;
        for (int w = 0; w < 32; ++w) {
            const int c0 = warp_tot[0][w], c1 = warp_tot[1][w];
            if (w < warp) { woff0 += c0; woff1 += c1; }
            tot0 += c0; tot1 += c1;
        }
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
                const int pos = h == 0 ? carry + woff0 + __popc(bal0 & lt) : carry + tot0 + woff1 + __popc(bal1 & lt);
                keep[pos] = base + 1024 * h + tid;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    pts[(size_t)a * m_cap + pos] = p[h][a];
                    pts[(size_t)(3 + a) * m_cap + pos] = c[h][a];
                }
            }
        }
        carry += tot0 + tot1;
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

// Sequential (order-exact) float sums over shared-memory columns.  The additions form one dependent chain -- that is
// the definition of the result -- but the loads and the per-element products do not have to sit on it: batch k+1 is
// fetched (and multiplied) while batch k is added.  The empty asm statements pin the loads of the next batch in front
// of the additions of the current one (left alone, ptxas reuses the registers and issues the loads after the chain,
// which then waits a shared-memory latency per batch: 11.7 cycles per element measured, FADD latency is 4).
constexpr int kSumBatch = 8;
__device__ __forceinline__ float seq_sum(const float* __restrict__ col, int n) {
    float s = 0.f;
    int k = 0;
    if (n >= 2 * kSumBatch) {
        float a[kSumBatch], b[kSumBatch];
#pragma unroll
        for (int u = 0; u < kSumBatch; ++u) a[u] = col[u];
        for (; k + 2 * kSumBatch <= n; k += 2 * kSumBatch) {
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) b[u] = col[k + kSumBatch + u];
            asm volatile("" ::: "memory");
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) s = s + a[u];
            const int k2 = (k + 3 * kSumBatch <= n) ? k + 2 * kSumBatch : 0;   // clamped prefetch, unused past the end
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) a[u] = col[k2 + u];
            asm volatile("" ::: "memory");
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) s = s + b[u];
        }
    }
    for (; k < n; ++k) s = s + col[k];
    return s;
}
// sum_k (d[k] - dmean) * (c[k] - cmean), same order
__device__ __forceinline__ float seq_cov_sum(const float* __restrict__ d, float dmean, const float* __restrict__ c, float cmean,
                                             int n) {
    float s = 0.f;
    int k = 0;
    if (n >= 2 * kSumBatch) {
        float a[kSumBatch], b[kSumBatch];
#pragma unroll
        for (int u = 0; u < kSumBatch; ++u) a[u] = (d[u] - dmean) * (c[u] - cmean);
        for (; k + 2 * kSumBatch <= n; k += 2 * kSumBatch) {
            float rd[kSumBatch], rc[kSumBatch];
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) { rd[u] = d[k + kSumBatch + u]; rc[u] = c[k + kSumBatch + u]; }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) { s = s + a[u]; b[u] = (rd[u] - dmean) * (rc[u] - cmean); }
            const int k2 = (k + 3 * kSumBatch <= n) ? k + 2 * kSumBatch : 0;
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) { rd[u] = d[k2 + u]; rc[u] = c[k2 + u]; }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int u = 0; u < kSumBatch; ++u) { s = s + b[u]; a[u] = (rd[u] - dmean) * (rc[u] - cmean); }
        }
    }
    for (; k < n; ++k) s = s + (d[k] - dmean) * (c[k] - cmean);
    return s;
}

constexpr int kSelThreads = 1024;

// phase clocks of the selection kernel (debug builds only: make -C putslam_b200/csrc dbg)
#ifdef PSLAM_SELECT_TIMING
__device__ long long g_sel_clk[16];
#define SEL_CLK(k) do { if (threadIdx.x == 0) g_sel_clk[k] = clock64(); } while (0)
#else
#define SEL_CLK(k) do { } while (0)
#endif

__device__ __forceinline__ void
ransac_select_body(const float* __restrict__ pts, int m_cap, const int* __restrict__ keep,
                   const int* __restrict__ n_filtered, const int* __restrict__ counts,
                   const float* __restrict__ models, int H, int adaptive, int stop_rule, double usac_conf,
                   int min_matches, double min_ratio, int iters_min, Scorer S, uint32_t seed_lo, uint32_t seed_hi,
                   int* __restrict__ inl_tmp /* m_cap scratch */, int stage_cap, int* __restrict__ result) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    __shared__ unsigned long long best_key;
    __shared__ int s_win, s_used, s_cnt, s_last;
    __shared__ float s_mean[6];
    __shared__ float s_sig[9];
    __shared__ float s_R[9], s_t[3];
    __shared__ int s_ok;

    SEL_CLK(0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mf = *n_filtered;
    float* Tout = reinterpret_cast<float*>(result + 4);
    int* inl_out = result + kHdrInts;

    auto write_identity = [&](int used) {
        if (tid < 16) Tout[tid] = (tid % 5 == 0) ? 1.f : 0.f;
        if (tid == 0) {
            result[0] = 0; result[1] = used; result[2] = mf; result[3] = 0; result[22] = -1; result[23] = 0;
            result[24] = iters_min; result[25] = -1; result[26] = H;
            *reinterpret_cast<double*>(result + 20) = 0.0;
        }
    };
    if (mf < min_matches || mf < 3) {  // RANSAC.cpp:77-80 (3 is the sample size; guards the sampler)
        write_identity(0);
        return;
    }

    // ---- winner: replay of the sequential loop over the scored hypotheses ----
    if (tid == 0) { best_key = 0ull; s_win = -1; s_used = H; s_cnt = 0; s_last = -1; carry = 0; }
    __syncthreads();
    if (adaptive) {
        // Replay of the reference's sequential loop  for (i = 0; i < iterationCount; ++i) { if (ratio_i > best) {...} }.
        // Its state only changes at "records", hypotheses whose count exceeds every earlier count (float(c)/float(mf) is
        // non-decreasing in c, so ratio_i > best implies c_i > all earlier c).  All threads find the records of a
        // 1024-hypothesis chunk with a prefix maximum and compact them in order; thread 0 then replays just those
        // (re-checking the float comparison, updating the bound exactly like saveBetterModel / updateStandardStopping).
        __shared__ int s_rec_idx[kSelThreads], s_rec_c[kSelThreads];
        __shared__ int s_wmax[32], s_wcnt[32];
        __shared__ int s_runmax, s_bound, s_done;
        // the reference's loop starts with computeRANSACIteration(0.20) = 487 (RANSAC.cpp:30); H is larger when
        // minimalInlierRatioThreshold < 0.2 lets the bound grow after the first improvement (kernels.h)
        const int bound0 = (adaptive == 1 && H > 487) ? 487 : H;
        if (tid == 0) { s_runmax = 0; s_bound = bound0; s_done = 0; }
        // thread 0's replay state.  With the reference rule the bound after an improvement is
        // min(iters_min, computeRANSACIteration(best)): two double-precision log/pow chains (~5k cycles) per record if
        // evaluated eagerly.  Only the comparison "record index < bound" needs it, so a float estimate with a +-0.1 % +-2
        // bracket decides, the exact value is computed only inside the bracket, and hyp_used itself is finished on the
        // host (kernels.h), with the libm the reference would use.
        const bool lazy = adaptive == 1 && stop_rule != 1;
        int bound_lo = bound0, bound_hi = bound0;   // bracket of the current bound (equal when it is known exactly)
        int win = -1, bc = 0, last = -1;
        double best = 0.0;
        for (int base = 0; base < H; base += kSelThreads) {
            __syncthreads();
            if (s_done || base >= s_bound) break;
            const int idx = base + tid;
            const int c = idx < H ? counts[idx] : -1;
            const int cc = c > 0 ? c : 0;          // degenerate (-1) and empty hypotheses can never be records
            int incl = cc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d && o > incl) incl = o;
            }
            int prev = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) prev = 0;
            if (lane == 31) s_wmax[warp] = incl;
            __syncthreads();
            int before = s_runmax, chunk_max = 0;
            for (int w = 0; w < 32; ++w) {
                const int m = s_wmax[w];
                if (w < warp && m > before) before = m;
                if (m > chunk_max) chunk_max = m;
            }
            if (prev > before) before = prev;
            const bool rec = cc > before;
            const uint32_t bal = __ballot_sync(0xffffffffu, rec);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            int off = 0, nrec = 0;
            for (int w = 0; w < 32; ++w) {
                const int n = s_wcnt[w];
                if (w < warp) off += n;
                nrec += n;
            }
            if (rec) {
                const int pos = off + __popc(bal & ((1u << lane) - 1u));
                s_rec_idx[pos] = idx; s_rec_c[pos] = c;
            }
            __syncthreads();
            if (tid == 0) {
                for (int r = 0; r < nrec; ++r) {
                    const int i = s_rec_idx[r];
                    if (i >= bound_hi) { s_done = 1; break; }
                    if (i >= bound_lo) {   // inside the bracket: evaluate the bound exactly
                        const int b = ransac_iterations(best);
                        bound_lo = bound_hi = iters_min < b ? iters_min : b;
                        if (bound_hi > H) bound_lo = bound_hi = H;
                        if (i >= bound_hi) { s_done = 1; break; }
                    }
                    const int ci = s_rec_c[r];
                    const float ratio = __fdiv_rn((float)ci, (float)mf);
                    if ((double)ratio > best) {
                        best = (double)ratio; win = i; bc = ci; last = i;
                        if (stop_rule == 1) {
                            const unsigned b = usac_standard_stopping((unsigned)ci, (unsigned)mf, usac_conf, (unsigned)H);
                            bound_lo = bound_hi = (int)(b < (unsigned)H ? b : (unsigned)H);
                        } else if (lazy) {
                            // float estimate of log(0.02) / log(1 - w^3)
                            const float x = ratio * ratio * ratio;
                            const float bf = __fdiv_rn(-3.912023005f, log1pf(-x));
                            if (x < 1.f && bf >= 0.f && bf < 1.0e9f) {
                                bound_lo = (int)floorf(bf * 0.999f) - 2;
                                bound_hi = (int)ceilf(bf * 1.001f) + 2;
                            } else if (x < 1.f && bf >= 1.0e9f) {
                                bound_lo = bound_hi = 0x7fffffff;     // far beyond iters_min either way
                            } else {
                                const int b = ransac_iterations(best);
                                bound_lo = bound_hi = b;
                            }
                            if (bound_lo > iters_min) bound_lo = iters_min;
                            if (bound_hi > iters_min) bound_hi = iters_min;
                        }
                        if (bound_lo > H) bound_lo = H;   // only H hypotheses were scored
                        if (bound_hi > H) bound_hi = H;
                    }
                }
                if (chunk_max > s_runmax) s_runmax = chunk_max;
                s_bound = bound_hi;
            }
        }
        if (tid == 0) {
            // the sequential loop leaves at the first i >= bound, and it is at last + 1 when the bound drops behind it
            s_win = win; s_cnt = bc; s_last = last;
            if (lazy && win >= 0) s_used = -1;   // host: max(last + 1, min(iters_min, computeRANSACIteration(best)))
            else s_used = (last + 1 > bound_hi) ? last + 1 : bound_hi;
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
    SEL_CLK(1);
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
    {   // ordered compaction, 2 x 1024 matches per pass (see ransac_filter_kernel)
        __shared__ int wt2[2][32];
        int run = 0;   // identical in every thread
        for (int base = 0; base < mf; base += 2 * kSelThreads) {
            bool in[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = base + kSelThreads * h + tid;
                in[h] = (k < mf) && inlier_test(S, M.R, M.t, Ri, ti, px[k], py[k], pz[k], cx[k], cy[k], cz[k], false);
            }
            const uint32_t bal0 = __ballot_sync(0xffffffffu, in[0]), bal1 = __ballot_sync(0xffffffffu, in[1]);
            __syncthreads();
            if (lane == 0) { wt2[0][warp] = __popc(bal0); wt2[1][warp] = __popc(bal1); }
            __syncthreads();
            int woff0 = 0, woff1 = 0, tot0 = 0, tot1 = 0;
            for (int w = 0; w < 32; ++w) {
                const int c0 = wt2[0][w], c1 = wt2[1][w];
                if (w < warp) { woff0 += c0; woff1 += c1; }
                tot0 += c0; tot1 += c1;
            }
            const uint32_t ltm = (1u << lane) - 1u;
            if (in[0]) inl_tmp[run + woff0 + __popc(bal0 & ltm)] = base + tid;
            if (in[1]) inl_tmp[run + tot0 + woff1 + __popc(bal1 & ltm)] = base + kSelThreads + tid;
            run += tot0 + tot1;
        }
        if (tid == 0) carry = run;
        __syncthreads();
    }
    const int n_in = carry;
    __syncthreads();
    SEL_CLK(2);

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
    SEL_CLK(3);
    const float one_over_n = __fdiv_rn(1.f, (float)n_in);
    if (tid < 6) {
        float s = 0.f;
        if (n_stage) {
            s = seq_sum(s_pts + tid * n_stage, n_in);
        } else {
            const float* col = pts + (size_t)tid * m_cap;  // 0..2 prev (dst), 3..5 cur (src)
            for (int k = 0; k < n_in; ++k) s = s + col[inl_tmp[k]];
        }
        s_mean[tid] = s * one_over_n;
    }
    __syncthreads();
    SEL_CLK(4);
    if (tid < 9) {
        const int i = tid / 3, j = tid % 3;
        const float dmean = s_mean[i], smean = s_mean[3 + j];
        float s = 0.f;
        if (n_stage) {
            s = seq_cov_sum(s_pts + i * n_stage, dmean, s_pts + (3 + j) * n_stage, smean, n_in);
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
    SEL_CLK(5);
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

    SEL_CLK(6);
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
    SEL_CLK(7);
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
        result[24] = iters_min;
        result[25] = s_last;
        result[26] = H;
        *reinterpret_cast<double*>(result + 20) = best_ratio;
        result[22] = win;
        result[23] = s_ok;
    }
}

// the selection, then (optionally) the copy of the call's output arena into page-locked host memory by the same CTA
__global__ void __launch_bounds__(kSelThreads, 1)
ransac_select_kernel(const float* __restrict__ pts, int m_cap, const int* __restrict__ keep,
                     const int* __restrict__ n_filtered, const int* __restrict__ counts,
                     const float* __restrict__ models, int H, int adaptive, int stop_rule, double usac_conf,
                     int min_matches, double min_ratio, int iters_min, Scorer S, uint32_t seed_lo, uint32_t seed_hi,
                     int* __restrict__ inl_tmp, int stage_cap, int* __restrict__ result, uint4* __restrict__ out_host,
                     const uint4* __restrict__ out_dev, int out_n16, int out_cap, int out_res_ints) {
    chain_begin();
    ransac_select_body(pts, m_cap, keep, n_filtered, counts, models, H, adaptive, stop_rule, usac_conf, min_matches, min_ratio,
                       iters_min, S, seed_lo, seed_hi, inl_tmp, stage_cap, result);
    if (out_host) {
        __syncthreads();          // the result written above (and the match list of the earlier kernels) is visible
        if (out_cap > 0) {        // match list + result: only what is in use crosses the bus
            const int* gd = reinterpret_cast<const int*>(out_dev);
            int* gh = reinterpret_cast<int*>(out_host);
            const int total = __ldcg(gd);
            const int n = total < out_cap ? total : out_cap;
            if (threadIdx.x < 2) gh[threadIdx.x] = __ldcg(gd + threadIdx.x);
            for (int i = threadIdx.x; i < n; i += kSelThreads) {
                gh[2 + i] = __ldcg(gd + 2 + i);
                gh[2 + out_cap + i] = __ldcg(gd + 2 + out_cap + i);
                gh[2 + 2 * out_cap + i] = __ldcg(gd + 2 + 2 * out_cap + i);
            }
            const int* rd = gd + out_res_ints;
            int* rh = gh + out_res_ints;
            const int n_res = kHdrInts + __ldcg(rd);
            for (int i = threadIdx.x; i < n_res; i += kSelThreads) rh[i] = __ldcg(rd + i);
        } else {
            for (int i = threadIdx.x; i < out_n16; i += kSelThreads) out_host[i] = __ldcg(out_dev + i);
        }
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
    const int H = ransac_hypothesis_budget(P);   // adaptive: 487 = int(log(0.02)/log(1-0.2^3)) (RANSAC.cpp:30) or more
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
    // per-device function attribute, set once per device (a relaxed flag: setting it twice is harmless)
    {
        static bool configured[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            if ((e = cudaFuncSetAttribute(ransac_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 8192 * 4)) != cudaSuccess)
                return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    if ((e = launch_chained(ransac_select_kernel, dim3(1), dim3(kSelThreads), sel_smem, st, ws.pts, ws.m_cap, ws.keep,
                            ws.n_filtered, ws.counts, ws.models, H, adaptive, P.stop_rule, P.usac_conf, P.min_matches,
                            P.min_inlier_ratio, P.iters_min_ratio, S, P.seed_lo, P.seed_hi,
                            ws.keep + ws.m_cap /* scratch: second half of keep */, stage_cap, ws.result,
                            reinterpret_cast<uint4*>(ws.out_host), reinterpret_cast<const uint4*>(ws.out_dev),
                            (int)((ws.out_bytes + 15) / 16), ws.out_cap, ws.out_res_ints)) != cudaSuccess)
        return e;
    if (launches) *launches += 4;
    return cudaGetLastError();
}

#ifdef PSLAM_SELECT_TIMING
extern "C" __attribute__((visibility("default"))) int pslam_debug_select_clocks(long long* out16) {
    return (int)cudaMemcpyFromSymbol(out16, g_sel_clk, sizeof(long long) * 16);
}
#endif

}  // namespace pslam
