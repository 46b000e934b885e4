// orb.cu -- K9: ORB descriptors for caller-provided keypoints, i.e. what MatcherOpenCV::describeFeatures obtains from
// `descriptorExtractor->compute(rgbImage, features, descriptors)` with cv::ORB::create() defaults (reference
// src/Matcher/matcherOpenCV.cpp:83-84,181-195).  The arithmetic is OpenCV's (un-vendored dependency); it is restated
// and pinned bit for bit against cv2 4.13.0 by the test suite's numpy restatement.  On the device:
//   orb_gray_kernel        COLOR_BGR2GRAY, fixed point (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   orb_resize_kernel      level l from level l-1, INTER_LINEAR_EXACT: 8.8 coefficient tables (host, double precision),
//                          horizontal 8.8, vertical 16.16, round half up -- integer, one launch per level (a cascade)
//   orb_blur_rows_kernel   all levels in one launch: 7-tap float row pass, s = k0*x0, s = fma(k_i, x_i, s)
//   orb_frame_blur_kernel  all levels in one launch: every pixel of the framed level -- interior = column pass
//                          s = k3*c, s = fma(k_{3+d}, below + above, s), round-half-even, saturate; 32-pixel frame =
//                          BORDER_REFLECT_101 of the unblurred level (ORB blurs the sub-matrix in place, the frame keeps
//                          the unblurred pixels, and rotated patches of coarse-level keypoints do reach into it)
//   orb_describe_kernel    one warp per keypoint, one descriptor byte per lane: 8 tests x 2 rotated, rounded sample points
// Images are a few hundred KB: every kernel is latency-bound, the pyramid lives in L2 between them.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace pslam {

__constant__ float c_gauss7[7];   // getGaussianKernel(7, 2, CV_32F)
static const uint32_t kGauss7Bits[7] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};

static const signed char kPatternHost[1024] = {
#include "orb_pattern.inc"
};

__device__ __forceinline__ int reflect101(int i, int n) {   // valid for -n < i < 2n - 1 (frames are 32 px, levels >= 33)
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// rgb_order 0: COLOR_BGR2GRAY (what cv::ORB does to a colour input); 1: COLOR_RGB2GRAY (detectFeatures, :122)
__global__ void orb_gray_kernel(const uint8_t* __restrict__ bgr, int W, int H, int row_bytes, uint8_t* __restrict__ gray,
                                int rgb_order) {
    chain_begin();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t* p = bgr + (size_t)y * row_bytes + 3 * (size_t)x;
    const int c0 = rgb_order ? p[2] : p[0], c2 = rgb_order ? p[0] : p[2];
    gray[(size_t)y * W + x] = (uint8_t)((c0 * 3735 + (int)p[1] * 19235 + c2 * 9798 + (1 << 14)) >> 15);
}

// xtab / ytab: per destination index {source offset, weight of the first sample (8.8)}; second weight = 256 - first
__global__ void orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int src_stride, uint8_t* __restrict__ dst,
                                  int dw, int dh, const int2* __restrict__ xtab, const int2* __restrict__ ytab) {
    chain_begin();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const int2 cx = xtab[x], cy = ytab[y];
    const int x0 = cx.x, x1 = min(cx.x + 1, sw - 1), y0 = cy.x, y1 = min(cy.x + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)y0 * src_stride;
    const uint8_t* r1 = src + (size_t)y1 * src_stride;
    const uint32_t h0 = (uint32_t)cx.y * r0[x0] + (uint32_t)(256 - cx.y) * r0[x1];
    const uint32_t h1 = (uint32_t)cx.y * r1[x0] + (uint32_t)(256 - cx.y) * r1[x1];
    const uint32_t v = (uint32_t)cy.y * h0 + (uint32_t)(256 - cy.y) * h1;
    dst[(size_t)y * dw + x] = (uint8_t)((v + 32768u) >> 16);
}

struct OrbLevels {
    int n;
    int w[kOrbMaxLevels], h[kOrbMaxLevels];
    int plain_off[kOrbMaxLevels];    // level image (w x h, tight), bytes from the pyramid base
    int ext_off[kOrbMaxLevels];      // framed level ((w + 64) x (h + 64), tight)
    int pix_start[kOrbMaxLevels + 1];   // prefix of w*h        (row-pass launch)
    int ext_start[kOrbMaxLevels + 1];   // prefix of (w+64)(h+64) (frame/blur launch)
    int edge;                           // border band of the corner filter: 31 for ORB (edgeThreshold), 0 for plain cv::FAST
};

__global__ void orb_blur_rows_kernel(OrbLevels L, const uint8_t* __restrict__ plain, float* __restrict__ rowbuf) {
    chain_begin();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= L.pix_start[L.n]) return;
    int l = 0;
    while (g >= L.pix_start[l + 1]) ++l;
    const int p = g - L.pix_start[l], w = L.w[l];
    const int y = p / w, x = p - y * w;
    const uint8_t* row = plain + L.plain_off[l] + (size_t)y * w;
    float s = __fmul_rn(c_gauss7[0], (float)row[reflect101(x - 3, w)]);
#pragma unroll
    for (int i = 1; i < 7; ++i) s = __fmaf_rn(c_gauss7[i], (float)row[reflect101(x - 3 + i, w)], s);
    rowbuf[g] = s;
}

constexpr int kOrbBorder = 32;
__global__ void orb_frame_blur_kernel(OrbLevels L, const uint8_t* __restrict__ plain, const float* __restrict__ rowbuf,
                                      uint8_t* __restrict__ ext) {
    chain_begin();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= L.ext_start[L.n]) return;
    int l = 0;
    while (g >= L.ext_start[l + 1]) ++l;
    const int p = g - L.ext_start[l], w = L.w[l], h = L.h[l], ew = w + 2 * kOrbBorder;
    const int ey = p / ew, ex = p - ey * ew;
    const int x = ex - kOrbBorder, y = ey - kOrbBorder;
    uint8_t out;
    if (x >= 0 && x < w && y >= 0 && y < h) {
        const float* col = rowbuf + L.pix_start[l] + x;
        float v = __fadd_rn(__fmul_rn(c_gauss7[3], col[(size_t)y * w]), 0.f);
#pragma unroll
        for (int d = 1; d <= 3; ++d) {
            const float t = __fadd_rn(col[(size_t)reflect101(y + d, h) * w], col[(size_t)reflect101(y - d, h) * w]);
            v = __fmaf_rn(c_gauss7[3 + d], t, v);
        }
        const int r = __float2int_rn(v);
        out = (uint8_t)min(max(r, 0), 255);
    } else {
        out = plain[L.plain_off[l] + (size_t)reflect101(y, h) * w + reflect101(x, w)];
    }
    ext[L.ext_off[l] + p] = out;
}

// rec: per kept keypoint {cx, cy (level pixel, frame not included), level, a bits, b bits}
__global__ void __launch_bounds__(256)
orb_describe_kernel(OrbLevels L, const uint8_t* __restrict__ ext, const int* __restrict__ rec, int n,
                    const char4* __restrict__ pattern, uint8_t* __restrict__ desc) {
    __shared__ char4 spat[256];
    chain_begin();
    spat[threadIdx.x] = pattern[threadIdx.x];
    __syncthreads();
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    const int cx = rec[5 * k], cy = rec[5 * k + 1], l = rec[5 * k + 2];
    const float a = __int_as_float(rec[5 * k + 3]), b = __int_as_float(rec[5 * k + 4]);
    const int ew = L.w[l] + 2 * kOrbBorder;
    const uint8_t* centre = ext + L.ext_off[l] + (size_t)(cy + kOrbBorder) * ew + (cx + kOrbBorder);
    uint32_t byte = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const char4 pt = spat[8 * lane + t];
        const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int v0 = centre[iy0 * ew + ix0], v1 = centre[iy1 * ew + ix1];
        byte |= (v0 < v1 ? 1u : 0u) << t;
    }
    desc[32 * (size_t)k + lane] = (uint8_t)byte;
}

// ---- K10: detection (cv::ORB::detect == the call inside MatcherOpenCV::detectFeatures, reference
// src/Matcher/matcherOpenCV.cpp:118-176).  Device: FAST-9/16 corner score of every pixel of every (unblurred) level,
// 3x3 non-maximum suppression + 31-px border filter + ordered (raster) compaction, Harris response and
// intensity-centroid angle of every surviving corner.  Host (ctx.cu): OpenCV's two retainBest passes per level with the
// same libstdc++ algorithms, so that even the order inside a level is OpenCV's.
__constant__ int c_umax[16];
static const int kUmaxHost[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// score = largest t for which 9 contiguous ring pixels are all > v + t or all < v - t (cv::cornerScore<16>); 0 unless
// that holds for `threshold`.  Window minima by doubling: w2, w4, w8, then w9 = min(w8[k], d[k + 8]).
__global__ void orb_fast_score_kernel(OrbLevels L, const uint8_t* __restrict__ plain, uint8_t* __restrict__ score,
                                      int threshold) {
    chain_begin();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= L.pix_start[L.n]) return;
    int l = 0;
    while (g >= L.pix_start[l + 1]) ++l;
    const int p = g - L.pix_start[l], w = L.w[l], h = L.h[l];
    const int y = p / w, x = p - y * w;
    uint8_t out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const uint8_t* c = plain + L.plain_off[l] + (size_t)y * w + x;
        const int v = c[0];
        int d[16];
        // any 9 contiguous ring pixels contain at least two of the four compass points: most pixels stop here
        d[0] = v - c[3 * w]; d[4] = v - c[3]; d[8] = v - c[-3 * w]; d[12] = v - c[-3];
        const int nb = (d[0] > threshold) + (d[4] > threshold) + (d[8] > threshold) + (d[12] > threshold);
        const int nd = (d[0] < -threshold) + (d[4] < -threshold) + (d[8] < -threshold) + (d[12] < -threshold);
        if (nb < 2 && nd < 2) { score[g] = 0; return; }
        d[0] = v - c[3 * w];          d[1] = v - c[3 * w + 1];      d[2] = v - c[2 * w + 2];      d[3] = v - c[w + 3];
        d[4] = v - c[3];              d[5] = v - c[-w + 3];         d[6] = v - c[-2 * w + 2];     d[7] = v - c[-3 * w + 1];
        d[8] = v - c[-3 * w];         d[9] = v - c[-3 * w - 1];     d[10] = v - c[-2 * w - 2];    d[11] = v - c[-w - 3];
        d[12] = v - c[-3];            d[13] = v - c[w - 3];         d[14] = v - c[2 * w - 2];     d[15] = v - c[3 * w - 1];
        int lo[16], hi[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { lo[k] = min(d[k], d[(k + 1) & 15]); hi[k] = max(d[k], d[(k + 1) & 15]); }
        int lo4[16], hi4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { lo4[k] = min(lo[k], lo[(k + 2) & 15]); hi4[k] = max(hi[k], hi[(k + 2) & 15]); }
        int amin = -(1 << 20), bmax = 1 << 20;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
            const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
            amin = max(amin, lo9);
            bmax = min(bmax, hi9);
        }
        if (amin > threshold || bmax < -threshold) out = (uint8_t)(max(amin, -bmax) - 1);
    }
    score[g] = out;
}

__device__ __forceinline__ bool orb_is_candidate(const OrbLevels& L, const uint8_t* __restrict__ score, int g, int& lvl,
                                                 int& x, int& y, int& sc) {
    const uint8_t* s = score + g;
    sc = s[0];
    if (sc == 0) return false;                 // not a corner: the common case, decided by one coalesced byte
    int l = 0;
    while (g >= L.pix_start[l + 1]) ++l;
    const int p = g - L.pix_start[l], w = L.w[l], h = L.h[l];
    y = p / w; x = p - y * w; lvl = l;
    if (x < L.edge || x >= w - L.edge || y < L.edge || y >= h - L.edge) return false;   // runByImageBorder(edgeThreshold)
    return sc > s[-1] && sc > s[1] && sc > s[-w - 1] && sc > s[-w] && sc > s[-w + 1] && sc > s[w - 1] && sc > s[w] &&
           sc > s[w + 1];
}

// cand: 6 ints per candidate {level, x, y, FAST score, Harris bits, angle bits}; header[0] = count (may exceed cap).
// Grid-wide ordered (raster) compaction, same scheme as map_prepare_kernel: contiguous pixel ranges per CTA (grid <= SM
// count), pass 1 counts, stamped per-CTA counts + look-back give the offset, pass 2 writes.  Every thread takes four
// consecutive score bytes with one 32-bit load (zero = no corner, the common case); per-thread counts are scanned with
// shuffles.  (First version: one byte per thread per trip, 30 dependent trips per pass -> 47 us.)
constexpr int kCandThreads = 1024;
__device__ __forceinline__ int orb_quad_candidates(const OrbLevels& L, const uint8_t* __restrict__ score, int g, int hi,
                                                   int (&rec)[2][4]) {
    const uint32_t word = *reinterpret_cast<const uint32_t*>(score + g);   // g is a multiple of 4, the buffer is padded
    if (word == 0u) return 0;
    int n = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int l, x, y, sc;
        if (((word >> (8 * q)) & 0xffu) != 0u && g + q < hi && orb_is_candidate(L, score, g + q, l, x, y, sc)) {
            if (n < 2) { rec[n][0] = l; rec[n][1] = x; rec[n][2] = y; rec[n][3] = sc; }   // two strict 3x3 maxima cannot be adjacent
            ++n;
        }
    }
    return n;
}

__global__ void __launch_bounds__(kCandThreads)
orb_candidates_kernel(OrbLevels L, const uint8_t* __restrict__ score, int per_cta, int* __restrict__ cand, int cap,
                      int* __restrict__ header, unsigned long long* __restrict__ cta_counts, unsigned int epoch) {
    constexpr int kW = kCandThreads / 32;
    __shared__ int warp_tot[kW];
    __shared__ int s_base;
    __shared__ unsigned int s_bid;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = take_cta_ticket(cta_counts - 1, epoch);
    __syncthreads();
    const int bid = (int)s_bid;                  // logical CTA index: arrival order (common.cuh take_cta_ticket)
    const int total = L.pix_start[L.n];
    const int lo = min(total, bid * per_cta), hi = min(total, lo + per_cta);   // per_cta is a multiple of 4096
    int rec[2][4];
    int mine = 0;
    for (int g = lo + 4 * tid; g < hi; g += 4 * kCandThreads) mine += orb_quad_candidates(L, score, g, hi, rec);
    mine = (int)warp_add_u32((uint32_t)mine);
    if (lane == 0) warp_tot[warp] = mine;
    __syncthreads();
    if (warp == 0) {
        int tot = warp_tot[lane];   // kW == 32
        tot = (int)warp_add_u32((uint32_t)tot);
        volatile unsigned long long* sums = cta_counts;
        if (lane == 0) sums[bid] = ((unsigned long long)epoch << 32) | (unsigned int)tot;
        int before = 0;
        for (int b = lane; b < bid; b += 32) {
            unsigned long long v;
            do { v = sums[b]; } while ((unsigned int)(v >> 32) != epoch);
            before += (int)(unsigned int)(v & 0xffffffffu);
        }
        before = (int)warp_add_u32((uint32_t)before);
        if (lane == 0) {
            s_base = before;
            if (bid == (int)gridDim.x - 1) header[0] = before + tot;
        }
    }
    __syncthreads();
    int carry = s_base;
    for (int base = lo; base < hi; base += 4 * kCandThreads) {
        const int g = base + 4 * tid;
        const int n = g < hi ? orb_quad_candidates(L, score, g, hi, rec) : 0;
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
            const int c = warp_tot[w];
            if (w < warp) woff += c;
            tot += c;
        }
        int pos = carry + woff + incl - n;
        for (int q = 0; q < n && q < 2; ++q, ++pos) {
            if (pos < cap) {
                int* r = cand + 6 * (size_t)pos;
                r[0] = rec[q][0]; r[1] = rec[q][1]; r[2] = rec[q][2]; r[3] = rec[q][3];
            }
        }
        carry += tot;
    }
}

// cv::fastAtan2 (degrees): 7th-order odd polynomial, float32, no fused operations (the library is built with -fmad=false)
__device__ __forceinline__ float orb_fast_atan2(float y, float x) {
    const float rad = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * rad, p3 = -0.3258083974640975f * rad, p5 = 0.1555786518463281f * rad,
                p7 = -0.04432655554792128f * rad;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, ax + eps); c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = __fdiv_rn(ax, ay + eps); c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// one warp per candidate: HarrisResponses (block 7, k 0.04) and the ICAngles moments, integer sums by warp reduction
__global__ void __launch_bounds__(256)
orb_harris_angle_kernel(OrbLevels L, const uint8_t* __restrict__ plain, int* __restrict__ cand, int cap,
                        const int* __restrict__ header) {
    chain_begin();
    const int n = min(header[0], cap);
    const int lane = threadIdx.x & 31;
    for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < n; k += gridDim.x * 8) {
        int* r = cand + 6 * (size_t)k;
        const int l = r[0], x = r[1], y = r[2], w = L.w[l];
        const uint8_t* c = plain + L.plain_off[l] + (size_t)y * w + x;
        int a = 0, b = 0, cc = 0;
        for (int q = lane; q < 49; q += 32) {
            const int i = q / 7 - 3, j = q % 7 - 3;
            const uint8_t* p = c + i * w + j;
            const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
            const int Iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
            a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
        }
        int m01 = 0, m10 = 0;
        if (lane < 31) {
            const int v = lane - 15, d = c_umax[v < 0 ? -v : v];
            const uint8_t* row = c + v * w;
            int s = 0, su = 0;
            for (int u = -d; u <= d; ++u) { const int val = row[u]; s += val; su += u * val; }
            m01 = v * s; m10 = su;
        }
        a = (int)warp_add_u32((uint32_t)a); b = (int)warp_add_u32((uint32_t)b); cc = (int)warp_add_u32((uint32_t)cc);
        m01 = (int)warp_add_u32((uint32_t)m01); m10 = (int)warp_add_u32((uint32_t)m10);
        if (lane == 0) {
            const float scale = __fdiv_rn(1.f, (float)(4 * 7) * 255.f);
            const float s4 = scale * scale * scale * scale;
            const float fa = (float)a, fb = (float)b, fc = (float)cc;
            const float resp = (fa * fb - fc * fc - 0.04f * (fa + fb) * (fa + fb)) * s4;
            r[4] = __float_as_int(resp);
            r[5] = __float_as_int(orb_fast_atan2((float)m01, (float)m10));
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
float orb_level_scale(int level) { return (float)pow((double)1.2f, (double)level); }   // ORB's getScale()

void orb_level_size(int W, int H, int level, int* w, int* h) {
    const float s = orb_level_scale(level);
    *w = (int)nearbyintf((float)W / s);   // cvRound(float): round half to even
    *h = (int)nearbyintf((float)H / s);
}

// interpolationLinear<uchar>::getCoeffs (OpenCV resize, INTER_LINEAR_EXACT): {offset, first weight in 8.8}
void orb_linear_exact_table(int src, int dst, int* tab /* 2 * dst */) {
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        const double fval = scale * ((double)d + 0.5) - 0.5;
        const double fl = floor(fval);
        int off = 0, c0 = 256;
        if (fl >= 0 && src > 1) {
            if (fl < (double)(src - 1)) {
                off = (int)fl;
                c0 = 256 - (int)nearbyint((fval - fl) * 256.0);
            } else {
                off = src - 1;
            }
        }
        tab[2 * d] = off; tab[2 * d + 1] = c0;
    }
}

size_t orb_plan(int W, int H, int nlevels, OrbPlan* P) {
    P->n = nlevels;
    size_t plain = 0, ext = 0, pix = 0, tab = 0;
    for (int l = 0; l < nlevels; ++l) {
        orb_level_size(W, H, l, &P->w[l], &P->h[l]);
        P->plain_off[l] = (int)plain; P->ext_off[l] = (int)ext; P->pix_start[l] = (int)pix; P->ext_start[l] = (int)ext;
        P->tab_off[l] = (int)tab;
        plain += (size_t)P->w[l] * P->h[l];
        ext += (size_t)(P->w[l] + 64) * (P->h[l] + 64);
        pix += (size_t)P->w[l] * P->h[l];
        tab += 2 * (size_t)(P->w[l] + P->h[l]);
    }
    P->pix_start[nlevels] = (int)pix; P->ext_start[nlevels] = (int)ext;
    P->plain_bytes = plain; P->ext_bytes = ext; P->row_floats = pix; P->tab_ints = tab;
    return plain + ext + 4 * pix + 4 * tab;
}

void orb_fill_tables(const OrbPlan& P, int* tab) {
    for (int l = 1; l < P.n; ++l) {
        orb_linear_exact_table(P.w[l - 1], P.w[l], tab + P.tab_off[l]);
        orb_linear_exact_table(P.h[l - 1], P.h[l], tab + P.tab_off[l] + 2 * P.w[l]);
    }
}

cudaError_t orb_upload_constants(void* d_pattern, cudaStream_t st) {
    float k[7];
    memcpy(k, kGauss7Bits, sizeof(k));
    cudaError_t e = cudaMemcpyToSymbolAsync(c_gauss7, k, sizeof(k), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_umax, kUmaxHost, sizeof(kUmaxHost), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    return cudaMemcpyAsync(d_pattern, kPatternHost, 1024, cudaMemcpyHostToDevice, st);
}

// the (unblurred) pyramid: level 0 in d_plain (or made from d_bgr), levels 1.. by the resize cascade
static cudaError_t launch_orb_pyramid(const uint8_t* d_bgr, int rgb_order, int W, int H, int row_bytes, const OrbPlan& P,
                                      uint8_t* d_plain, const int* d_tab, cudaStream_t st, int* nl) {
    cudaError_t e;
    if (d_bgr) {
        if ((e = launch_chained(orb_gray_kernel, dim3((unsigned)((W + 255) / 256), (unsigned)H), dim3(256), 0, st, d_bgr, W, H,
                                row_bytes, d_plain, rgb_order)) != cudaSuccess) return e;
        ++*nl;
    }
    for (int l = 1; l < P.n; ++l) {
        const int2* xt = reinterpret_cast<const int2*>(d_tab + P.tab_off[l]);
        const int2* yt = reinterpret_cast<const int2*>(d_tab + P.tab_off[l] + 2 * P.w[l]);
        if ((e = launch_chained(orb_resize_kernel, dim3((unsigned)((P.w[l] + 255) / 256), (unsigned)P.h[l]), dim3(256), 0, st,
                                (const uint8_t*)(d_plain + P.plain_off[l - 1]), P.w[l - 1], P.h[l - 1], P.w[l - 1],
                                d_plain + P.plain_off[l], P.w[l], P.h[l], xt, yt)) != cudaSuccess) return e;
        ++*nl;
    }
    return cudaSuccess;
}

static void orb_levels_from_plan(const OrbPlan& P, OrbLevels& L) {
    L.n = P.n;
    for (int l = 0; l < P.n; ++l) {
        L.w[l] = P.w[l]; L.h[l] = P.h[l]; L.plain_off[l] = P.plain_off[l]; L.ext_off[l] = P.ext_off[l];
        L.pix_start[l] = P.pix_start[l]; L.ext_start[l] = P.ext_start[l];
    }
    L.pix_start[P.n] = P.pix_start[P.n]; L.ext_start[P.n] = P.ext_start[P.n];
    L.edge = 31;
}

// d_bgr: H rows of row_bytes (3 bytes per pixel), or nullptr when level 0 (tight W x H gray) is already in d_plain
cudaError_t launch_orb_describe(const uint8_t* d_bgr, int W, int H, int row_bytes, const OrbPlan& P, uint8_t* d_plain,
                                uint8_t* d_ext, float* d_rowbuf, const int* d_tab, const void* d_pattern, const int* d_rec,
                                int n_kp, uint8_t* d_desc, cudaStream_t st, int* launches) {
    OrbLevels L;
    orb_levels_from_plan(P, L);
    int nl = 0;
    cudaError_t e = launch_orb_pyramid(d_bgr, 0, W, H, row_bytes, P, d_plain, d_tab, st, &nl);
    if (e != cudaSuccess) return e;
    if ((e = launch_chained(orb_blur_rows_kernel, dim3((unsigned)((P.row_floats + 255) / 256)), dim3(256), 0, st, L,
                            (const uint8_t*)d_plain, d_rowbuf)) != cudaSuccess) return e;
    if ((e = launch_chained(orb_frame_blur_kernel, dim3((unsigned)((P.ext_bytes + 255) / 256)), dim3(256), 0, st, L,
                            (const uint8_t*)d_plain, (const float*)d_rowbuf, d_ext)) != cudaSuccess) return e;
    nl += 2;
    if (n_kp > 0) {
        if ((e = launch_chained(orb_describe_kernel, dim3((unsigned)((n_kp + 7) / 8)), dim3(256), 0, st, L, (const uint8_t*)d_ext,
                                d_rec, n_kp, reinterpret_cast<const char4*>(d_pattern), d_desc)) != cudaSuccess) return e;
        ++nl;
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
}

cudaError_t launch_orb_detect(const uint8_t* d_bgr, int rgb_order, int W, int H, int row_bytes, const OrbPlan& P,
                              uint8_t* d_plain, uint8_t* d_score, const int* d_tab, int fast_threshold, int* d_cand, int cap,
                              int* d_header, unsigned long long* d_cta_counts, unsigned int epoch, int sm_count,
                              cudaStream_t st, int* launches) {
    OrbLevels L;
    orb_levels_from_plan(P, L);
    int nl = 0;
    cudaError_t e = launch_orb_pyramid(d_bgr, rgb_order, W, H, row_bytes, P, d_plain, d_tab, st, &nl);
    if (e != cudaSuccess) return e;
    const int total = (int)P.row_floats;
    if ((e = launch_chained(orb_fast_score_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, L,
                            (const uint8_t*)d_plain, d_score, fast_threshold)) != cudaSuccess) return e;
    int grid = sm_count > 0 ? sm_count : 1;
    int per_cta = (total + grid - 1) / grid;
    per_cta = (per_cta + 4 * kCandThreads - 1) / (4 * kCandThreads) * (4 * kCandThreads);
    grid = (total + per_cta - 1) / per_cta;
    if (grid < 1) grid = 1;
    if ((e = launch_chained(orb_candidates_kernel, dim3((unsigned)grid), dim3(kCandThreads), 0, st, L, (const uint8_t*)d_score, per_cta,
                            d_cand, cap, d_header, d_cta_counts, epoch)) != cudaSuccess) return e;
    int hgrid = (cap + 7) / 8;
    if (hgrid > 8 * (sm_count > 0 ? sm_count : 1)) hgrid = 8 * (sm_count > 0 ? sm_count : 1);
    if ((e = launch_chained(orb_harris_angle_kernel, dim3((unsigned)hgrid), dim3(256), 0, st, L, (const uint8_t*)d_plain, d_cand,
                            cap, (const int*)d_header)) != cudaSuccess) return e;
    if (launches) *launches += nl + 3;
    return cudaGetLastError();
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression = true, TYPE_9_16) on one (gray) image: score map, strict 3x3
// maxima in raster order.  Records as launch_orb_detect ({0, x, y, score, -, -}).
cudaError_t launch_fast_detect(const uint8_t* d_bgr, int rgb_order, int W, int H, int row_bytes, const OrbPlan& P,
                               uint8_t* d_plain, uint8_t* d_score, int threshold, int* d_cand, int cap, int* d_header,
                               unsigned long long* d_cta_counts, unsigned int epoch, int sm_count, cudaStream_t st,
                               int* launches) {
    OrbLevels L;
    orb_levels_from_plan(P, L);
    L.edge = 0;
    int nl = 0;
    cudaError_t e;
    if (d_bgr) {
        if ((e = launch_chained(orb_gray_kernel, dim3((unsigned)((W + 255) / 256), (unsigned)H), dim3(256), 0, st, d_bgr, W, H,
                                row_bytes, d_plain, rgb_order)) != cudaSuccess) return e;
        ++nl;
    }
    const int total = (int)P.row_floats;
    if ((e = launch_chained(orb_fast_score_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, L,
                            (const uint8_t*)d_plain, d_score, threshold)) != cudaSuccess) return e;
    int grid = sm_count > 0 ? sm_count : 1;
    int per_cta = (total + grid - 1) / grid;
    per_cta = (per_cta + 4 * kCandThreads - 1) / (4 * kCandThreads) * (4 * kCandThreads);
    grid = (total + per_cta - 1) / per_cta;
    if (grid < 1) grid = 1;
    if ((e = launch_chained(orb_candidates_kernel, dim3((unsigned)grid), dim3(kCandThreads), 0, st, L, (const uint8_t*)d_score, per_cta,
                            d_cand, cap, d_header, d_cta_counts, epoch)) != cudaSuccess) return e;
    if (launches) *launches += nl + 2;
    return cudaGetLastError();
}

}  // namespace pslam
