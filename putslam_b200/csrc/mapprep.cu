// mapprep.cu -- K8: map-side preparation of the frame-to-map matcher's input, the body of
// PUTSLAM::getAndFilterFeaturesFromMap after getCovisibleFeatures (reference src/PUTSLAM/PUTSLAM.cpp:624-674):
//   FeaturesMap::findNearestFrame (src/Map/featuresMap.cpp:528-563)   view-angle test against the descriptor's view
//   moveMapFeaturesToLocalCordinateSystem (PUTSLAM.cpp:28-51)         p_local = cameraPose^-1 p, (u,v) = inverseModel
//   DepthSensorModel::inverseModel (src/Grabber/depthSensorModel.cpp:18-25), RGBD::removeFarMapFeatures (RGBD.cpp:232-252)
// One thread per map feature, ordered compaction in a single CTA (the map has a few thousand visible features), so
// that the feature set can stay in HBM in SoA form between the map and pslam_frame_to_map.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace pslam {

struct MapPrepArgs {
    double Li[9];     // inverse of the pose's linear part (row-major)
    double ti[3];     // translation of the inverse pose
    float zc[3];      // current optical axis (third column of the rotation), float like the reference
    double fx, fy, cx, cy, img_w, img_h, max_angle, max_z;
};

// attributes that travel with a kept feature when the map is resident in HBM (pslam_frame_to_resident_map): the
// compacted copies are what the level prediction and the guided matcher read
struct MapGather {
    const uint4* desc_in; uint4* desc_out;      // 32-byte descriptors
    const int* oct_in; int* oct_out;            // ExtendedDescriptor::octave
    const double* det_in; double* det_out;      // ExtendedDescriptor::detDist
};

struct MapFeatureEval {
    bool keep;
    double p[3], u, v, ang;
};

// everything the reference computes for one map feature (see the file header), in its operation order
__device__ __forceinline__ MapFeatureEval eval_feature(const double* __restrict__ xyz, const float* __restrict__ view_axis,
                                                       int i, const MapPrepArgs& A) {
    MapFeatureEval e;
    e.p[0] = e.p[1] = e.p[2] = 0; e.u = -1; e.v = -1;
    const float z0 = view_axis[3 * i], z1 = view_axis[3 * i + 1], z2 = view_axis[3 * i + 2];
    const float dot = z0 * A.zc[0] + (z1 * A.zc[1] + z2 * A.zc[2]);
    const float a1 = z1 * z1, a2 = z2 * z2, b1 = A.zc[1] * A.zc[1], b2 = A.zc[2] * A.zc[2];
    const float na = __fsqrt_rn(z0 * z0 + (a1 + a2)), nb = __fsqrt_rn(A.zc[0] * A.zc[0] + (b1 + b2));
    e.ang = fabs(acos((double)__fdiv_rn(dot, na * nb)));
    e.keep = (e.ang < 10.0) && !(e.ang > A.max_angle);
    if (e.keep) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double s = A.Li[3 * r] * xyz[3 * i];
            s = s + A.Li[3 * r + 1] * xyz[3 * i + 1];
            s = s + A.Li[3 * r + 2] * xyz[3 * i + 2];
            e.p[r] = s + A.ti[r];
        }
        e.keep = !(e.p[2] > A.max_z);
        e.u = __ddiv_rn(A.fx * e.p[0], e.p[2]) + A.cx;
        e.v = __ddiv_rn(A.fy * e.p[1], e.p[2]) + A.cy;
        if (e.u < 0 || e.u > A.img_w || e.v < 0 || e.v > A.img_h || e.p[2] < 0.8 || e.p[2] > 6.0) { e.u = -1; e.v = -1; }
    }
    return e;
}

// Ordered compaction over the whole grid.  Every CTA owns a contiguous range of features.  Pass 1 counts the kept
// ones and publishes the count stamped with this launch's epoch; warp 0 then reads the stamped counts of all
// preceding CTAs in parallel (a decoupled look-back: one L2 round trip after the slowest predecessor) to get the
// CTA's output offset; pass 2 re-evaluates and writes in feature order.  CTAs are numbered by
// arrival (take_cta_ticket), so a CTA only waits on CTAs that are already running; the stamps make a reset between launches
// unnecessary.
constexpr int kPrepThreads = 256;
__global__ void __launch_bounds__(kPrepThreads)
map_prepare_kernel(const double* __restrict__ xyz, const float* __restrict__ view_axis, int M, int per_cta, MapPrepArgs A,
                   int* __restrict__ kept, double* __restrict__ xyz_local, double* __restrict__ uv,
                   double* __restrict__ angles, int* __restrict__ n_out, MapGather G,
                   unsigned long long* __restrict__ cta_counts /* [-1] = ticket word */, unsigned int epoch) {
    __shared__ int warp_tot[kPrepThreads / 32];
    __shared__ int s_base;
    __shared__ unsigned int s_bid;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = take_cta_ticket(cta_counts - 1, epoch);
    __syncthreads();
    const int bid = (int)s_bid;                  // logical CTA index: arrival order (see take_cta_ticket)
    const int lo = min(M, bid * per_cta), hi = min(M, lo + per_cta);

    // pass 1: count
    int mine = 0;
    for (int i = lo + tid; i < hi; i += kPrepThreads) mine += eval_feature(xyz, view_axis, i, A).keep ? 1 : 0;
    mine = (int)warp_add_u32((uint32_t)mine);
    if (lane == 0) warp_tot[warp] = mine;
    __syncthreads();
    if (warp == 0) {
        int total = 0;
        for (int w = 0; w < kPrepThreads / 32; ++w) total += warp_tot[w];
        if (lane == 0) {
            const unsigned long long stamped = ((unsigned long long)epoch << 32) | (unsigned int)total;
            *reinterpret_cast<volatile unsigned long long*>(cta_counts + bid) = stamped;
        }
        int before = 0;
        for (int b = lane; b < bid; b += 32) {
            unsigned long long v;
            do {
                v = *reinterpret_cast<volatile unsigned long long*>(cta_counts + b);
            } while ((unsigned int)(v >> 32) != epoch);
            before += (int)(unsigned int)(v & 0xffffffffu);
        }
        before = (int)warp_add_u32((uint32_t)before);
        if (lane == 0) {
            s_base = before;
            if (bid == (int)gridDim.x - 1) *n_out = before + total;
        }
    }
    __syncthreads();

    // pass 2: write, in feature order
    int carry = s_base;
    for (int base = lo; base < hi; base += kPrepThreads) {
        const int i = base + tid;
        MapFeatureEval e;
        e.keep = false;
        if (i < hi) e = eval_feature(xyz, view_axis, i, A);
        const uint32_t bal = __ballot_sync(0xffffffffu, e.keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        __syncthreads();   // warp_tot of the previous pass consumed
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kPrepThreads / 32; ++w) {
            const int c = warp_tot[w];
            if (w < warp) woff += c;
            tot += c;
        }
        if (e.keep) {
            const int pos = carry + woff + wpre;
            kept[pos] = i;
            xyz_local[3 * pos] = e.p[0]; xyz_local[3 * pos + 1] = e.p[1]; xyz_local[3 * pos + 2] = e.p[2];
            uv[2 * pos] = e.u; uv[2 * pos + 1] = e.v;
            angles[pos] = e.ang;
            if (G.desc_in) {
                G.desc_out[2 * (size_t)pos] = G.desc_in[2 * (size_t)i];
                G.desc_out[2 * (size_t)pos + 1] = G.desc_in[2 * (size_t)i + 1];
                G.oct_out[pos] = G.oct_in[i];
                G.det_out[pos] = G.det_in[i];
            }
        }
        carry += tot;
    }
}

// 3x3 inverse by cofactors, host side (Eigen's closed form for the affine inverse's linear part)
static void inverse3_host(const double* m, double* r) {
#define CF(i, j) (m[3 * (((i) + 1) % 3) + (((j) + 1) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 2) % 3)] - \
                  m[3 * (((i) + 1) % 3) + (((j) + 2) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 1) % 3)])
    const double c00 = CF(0, 0), c10 = CF(1, 0), c20 = CF(2, 0);
    const double det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
    const double invdet = 1.0 / det;
    r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
    r[3] = CF(0, 1) * invdet; r[4] = CF(1, 1) * invdet; r[5] = CF(2, 1) * invdet;
    r[6] = CF(0, 2) * invdet; r[7] = CF(1, 2) * invdet; r[8] = CF(2, 2) * invdet;
#undef CF
}

cudaError_t launch_map_prepare(const double* d_xyz, const float* d_view_axis, int M, const double* pose_colmajor,
                               double fx, double fy, double cx, double cy, double img_w, double img_h, double max_angle,
                               double max_z, int* d_kept, double* d_xyz_local, double* d_uv, double* d_angles, int* d_n,
                               unsigned long long* d_cta_counts, unsigned int epoch, int sm_count, cudaStream_t st,
                               int* launches, const uint8_t* d_desc_in, uint8_t* d_desc_out, const int* d_oct_in,
                               int* d_oct_out, const double* d_det_in, double* d_det_out) {
    MapPrepArgs A;
    double L[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) L[3 * r + c] = pose_colmajor[4 * c + r];
    inverse3_host(L, A.Li);
    for (int r = 0; r < 3; ++r) {
        double s = A.Li[3 * r] * pose_colmajor[12];
        s = s + A.Li[3 * r + 1] * pose_colmajor[13];
        s = s + A.Li[3 * r + 2] * pose_colmajor[14];
        A.ti[r] = -s;
    }
    A.zc[0] = (float)pose_colmajor[8]; A.zc[1] = (float)pose_colmajor[9]; A.zc[2] = (float)pose_colmajor[10];
    A.fx = fx; A.fy = fy; A.cx = cx; A.cy = cy; A.img_w = img_w; A.img_h = img_h; A.max_angle = max_angle; A.max_z = max_z;
    MapGather G;
    G.desc_in = reinterpret_cast<const uint4*>(d_desc_in); G.desc_out = reinterpret_cast<uint4*>(d_desc_out);
    G.oct_in = d_oct_in; G.oct_out = d_oct_out; G.det_in = d_det_in; G.det_out = d_det_out;
    // contiguous feature ranges, a multiple of the CTA size, at most one CTA per SM (co-residency, see the kernel)
    int grid = (M + kPrepThreads - 1) / kPrepThreads;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    int per_cta = (M + grid - 1) / grid;
    per_cta = (per_cta + kPrepThreads - 1) / kPrepThreads * kPrepThreads;
    grid = (M + per_cta - 1) / per_cta;
    if (grid < 1) grid = 1;
    const cudaError_t e = launch_chained(map_prepare_kernel, dim3((unsigned)grid), dim3(kPrepThreads), 0, st, d_xyz,
                                         d_view_axis, M, per_cta, A, d_kept, d_xyz_local, d_uv, d_angles, d_n, G,
                                         d_cta_counts, epoch);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
