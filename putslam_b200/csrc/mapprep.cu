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

__global__ void __launch_bounds__(1024, 1)
map_prepare_kernel(const double* __restrict__ xyz, const float* __restrict__ view_axis, int M, MapPrepArgs A,
                   int* __restrict__ kept, double* __restrict__ xyz_local, double* __restrict__ uv,
                   double* __restrict__ angles, int* __restrict__ n_out) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < M; base += 1024) {
        const int i = base + tid;
        bool keep = false;
        double p[3] = {0, 0, 0}, u = -1, v = -1, ang = 0;
        if (i < M) {
            const float z0 = view_axis[3 * i], z1 = view_axis[3 * i + 1], z2 = view_axis[3 * i + 2];
            const float dot = z0 * A.zc[0] + (z1 * A.zc[1] + z2 * A.zc[2]);
            const float a1 = z1 * z1, a2 = z2 * z2, b1 = A.zc[1] * A.zc[1], b2 = A.zc[2] * A.zc[2];
            const float na = __fsqrt_rn(z0 * z0 + (a1 + a2)), nb = __fsqrt_rn(A.zc[0] * A.zc[0] + (b1 + b2));
            ang = fabs(acos((double)__fdiv_rn(dot, na * nb)));
            keep = (ang < 10.0) && !(ang > A.max_angle);
            if (keep) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double s = A.Li[3 * r] * xyz[3 * i];
                    s = s + A.Li[3 * r + 1] * xyz[3 * i + 1];
                    s = s + A.Li[3 * r + 2] * xyz[3 * i + 2];
                    p[r] = s + A.ti[r];
                }
                keep = !(p[2] > A.max_z);
                u = __ddiv_rn(A.fx * p[0], p[2]) + A.cx;
                v = __ddiv_rn(A.fy * p[1], p[2]) + A.cy;
                if (u < 0 || u > A.img_w || v < 0 || v > A.img_h || p[2] < 0.8 || p[2] > 6.0) { u = -1; v = -1; }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = warp_tot[w];
            if (w < warp) woff += c;
            tot += c;
        }
        if (keep) {
            const int pos = carry + woff + wpre;
            kept[pos] = i;
            xyz_local[3 * pos] = p[0]; xyz_local[3 * pos + 1] = p[1]; xyz_local[3 * pos + 2] = p[2];
            uv[2 * pos] = u; uv[2 * pos + 1] = v;
            angles[pos] = ang;
        }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) *n_out = carry;
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
                               cudaStream_t st, int* launches) {
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
    map_prepare_kernel<<<1, 1024, 0, st>>>(d_xyz, d_view_axis, M, A, d_kept, d_xyz_local, d_uv, d_angles, d_n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
