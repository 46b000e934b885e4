// backproject.cu -- K1: keypoint undistortion, depth back-projection, detection distance and the
// per-point sensor covariance.
//   cv::undistortPoints + re-projection : reference src/RGBD/RGBD.cpp:254-314 (OpenCV model: doubles,
//                                         5 fixed-point iterations, result cast to float)
//   RGBD::point2Dto3D / keypoints2Dto3D : reference src/RGBD/RGBD.cpp:30-65 (roundSize :10-16)
//   detDist                             : reference src/Matcher/matcher.cpp:51-58
//   DepthSensorModel::computeCov        : reference src/Grabber/depthSensorModel.cpp:28-36
// One thread per keypoint; the only memory traffic is one 2-byte depth gather per keypoint.
#include "common.cuh"
#include "kernels.h"

namespace pslam {

__device__ __forceinline__ int round_size(double x, int size) {
    // reference returns `size` (one past the edge) for x > size-1; clamped to size-1 here, see DESIGN.md
    if (x < 0) x = 0;
    else if (x > size - 1) x = size - 1;
    return (int)round(x);
}

// d^3 with a single rounding of the exact product (what a correctly rounded pow(d, 3.0) returns).
__device__ __forceinline__ double cube_rn(double d) {
    const double p = d * d;
    const double pe = fma(d, d, -p);        // exact error of d*d
    const double h = p * d;
    const double he = fma(p, d, -h);        // exact error of p*d
    return h + (he + pe * d);
}

__global__ void backproject_kernel(const float* __restrict__ uv, int n, const uint16_t* __restrict__ depth, int W, int H,
                                   int stride, pslam_camera cam, int undistort, double depth_scale,
                                   float* __restrict__ uv_und, float* __restrict__ xyz, double* __restrict__ det_dist,
                                   double* __restrict__ cov, pslam_cov_params cp, int want_cov) {
    chain_begin();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u2 = uv[2 * i], v2 = uv[2 * i + 1];
    if (undistort) {
        const double k1 = cam.dist[0], k2 = cam.dist[1], p1 = cam.dist[2], p2 = cam.dist[3], k3 = cam.dist[4];
        const double ifx = 1. / (double)cam.fx, ify = 1. / (double)cam.fy;
        double x = ((double)u2 - (double)cam.cx) * ifx;
        double y = ((double)v2 - (double)cam.cy) * ify;
        const double x0 = x, y0 = y;
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
            const double r2 = x * x + y * y;
            const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const float xf = (float)x, yf = (float)y;
        u2 = xf * cam.fx + cam.cx;
        v2 = yf * cam.fy + cam.cy;
    }
    if (uv_und) { uv_und[2 * i] = u2; uv_und[2 * i + 1] = v2; }
    const int uR = round_size((double)u2, W), vR = round_size((double)v2, H);
    const uint16_t dz = depth[(size_t)vR * stride + uR];
    const float Z = (float)(((double)dz) / depth_scale);
    const float u = __fdiv_rn(u2 - cam.cx, cam.fx);
    const float v = __fdiv_rn(v2 - cam.cy, cam.fy);
    const float X = u * Z, Y = v * Z;
    xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
    if (det_dist) {
        const float s = X * X + Y * Y + Z * Z;     // float products and adds, left to right
        det_dist[i] = (double)__fsqrt_rn(s);       // std::sqrt(float) overload, then widened
    }
    if (want_cov) {
        // FeaturesMap::addFeatures feeds (u, v, depth) of the undistorted keypoint, u/v truncated to integers
        const unsigned ui = (unsigned)u2, vi = (unsigned)v2;
        const double d = (double)Z;
        const double J00 = d / cp.fx, J02 = ((double)ui / cp.fx) - (cp.cx / cp.fx);
        const double J11 = d / cp.fy, J12 = ((double)vi / cp.fy) - (cp.cy / cp.fy);
        const double r0 = cp.var_u, r1 = cp.var_v;
        const double r2 = cp.dist_var_coefs[0] * cube_rn(d) + cp.dist_var_coefs[1] * (d * d) + cp.dist_var_coefs[2] * d +
                          cp.dist_var_coefs[3];
        const double J[9] = {J00, 0.0, J02, 0.0, J11, J12, 0.0, 0.0, 1.0};
        const double R[3] = {r0, r1, r2};
        double JR[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                double s = J[3 * a + 0] * (b == 0 ? R[0] : 0.0);
                s = s + J[3 * a + 1] * (b == 1 ? R[1] : 0.0);
                s = s + J[3 * a + 2] * (b == 2 ? R[2] : 0.0);
                JR[3 * a + b] = s;
            }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                double s = JR[3 * a + 0] * J[3 * b + 0];
                s = s + JR[3 * a + 1] * J[3 * b + 1];
                s = s + JR[3 * a + 2] * J[3 * b + 2];
                cov[9 * (size_t)i + 3 * a + b] = s;
            }
    }
}

// 3x3 double inverse by cofactors (the closed form Eigen uses for Matrix3d::inverse()).
__device__ __forceinline__ void inverse3d(const double (&m)[9], double (&r)[9]) {
#define CF(i, j) (m[3 * (((i) + 1) % 3) + (((j) + 1) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 2) % 3)] - \
                  m[3 * (((i) + 1) % 3) + (((j) + 2) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 1) % 3)])
    const double c00 = CF(0, 0), c10 = CF(1, 0), c20 = CF(2, 0);
    const double det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
    const double invdet = __ddiv_rn(1.0, det);
    r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
    r[3] = CF(0, 1) * invdet; r[4] = CF(1, 1) * invdet; r[5] = CF(2, 1) * invdet;
    r[6] = CF(0, 2) * invdet; r[7] = CF(1, 2) * invdet; r[8] = CF(2, 2) * invdet;
#undef CF
}

// DepthSensorModel::informationMatrixFromImageCoordinates (reference src/Grabber/depthSensorModel.cpp:55-59),
// batched: uvz = n x {u, v, depth} doubles; u, v truncated to unsigned like the reference's cast.
__global__ void information_kernel(const double* __restrict__ uvz, int n, pslam_cov_params cp, double* __restrict__ cov_out,
                                   double* __restrict__ info_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned ui = (unsigned)(unsigned long long)uvz[3 * i], vi = (unsigned)(unsigned long long)uvz[3 * i + 1];
    const double d = uvz[3 * i + 2];
    const double J[9] = {d / cp.fx, 0.0, ((double)ui / cp.fx) - (cp.cx / cp.fx), 0.0, d / cp.fy,
                         ((double)vi / cp.fy) - (cp.cy / cp.fy), 0.0, 0.0, 1.0};
    const double R[3] = {cp.var_u, cp.var_v,
                         cp.dist_var_coefs[0] * cube_rn(d) + cp.dist_var_coefs[1] * (d * d) + cp.dist_var_coefs[2] * d +
                             cp.dist_var_coefs[3]};
    double JR[9], cov[9], info[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            double s = J[3 * a + 0] * (b == 0 ? R[0] : 0.0);
            s = s + J[3 * a + 1] * (b == 1 ? R[1] : 0.0);
            s = s + J[3 * a + 2] * (b == 2 ? R[2] : 0.0);
            JR[3 * a + b] = s;
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            double s = JR[3 * a + 0] * J[3 * b + 0];
            s = s + JR[3 * a + 1] * J[3 * b + 1];
            s = s + JR[3 * a + 2] * J[3 * b + 2];
            cov[3 * a + b] = s;
        }
    inverse3d(cov, info);
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        if (cov_out) cov_out[9 * (size_t)i + a] = cov[a];
        info_out[9 * (size_t)i + a] = info[a];
    }
}

__device__ __forceinline__ void px_to_3d(float fu, float fv, const uint16_t* __restrict__ depth, int W, int H, int stride,
                                         const pslam_camera& cam, double depth_scale, float (&o)[3]) {
    const int uR = round_size((double)fu, W), vR = round_size((double)fv, H);
    const float Z = (float)(((double)depth[(size_t)vR * stride + uR]) / depth_scale);
    const float u = __fdiv_rn(fu - cam.cx, cam.fx), v = __fdiv_rn(fv - cam.cy, cam.fy);
    o[0] = u * Z; o[1] = v * Z; o[2] = Z;
}
__device__ __forceinline__ void normalize3(double (&v)[3]) {
    const double yy = v[1] * v[1], zz = v[2] * v[2];
    const double n = __dsqrt_rn(v[0] * v[0] + (yy + zz));
    v[0] = __ddiv_rn(v[0], n); v[1] = __ddiv_rn(v[1], n); v[2] = __ddiv_rn(v[2], n);
}
__device__ __forceinline__ void matmul3(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = A[3 * i] * B[j];
            s = s + A[3 * i + 1] * B[3 + j];
            s = s + A[3 * i + 2] * B[6 + j];
            C[3 * i + j] = s;
        }
}

// cov and (optionally) cov.inverse() -- the information matrix FeaturesMap attaches for models 1 / 2
// (reference src/Map/featuresMap.cpp:115-120, 268-273); Eigen's 3x3 cofactor inverse.
__device__ __forceinline__ void store_cov_info(const double (&cov)[9], int i, double* __restrict__ cov_out,
                                               double* __restrict__ info_out) {
    if (cov_out) {
#pragma unroll
        for (int a = 0; a < 9; ++a) cov_out[9 * (size_t)i + a] = cov[a];
    }
    if (info_out) {
        double inf[9];
        inverse3d(cov, inf);
#pragma unroll
        for (int a = 0; a < 9; ++a) info_out[9 * (size_t)i + a] = inf[a];
    }
}

// Uncertainty model 1: RGBD::computeNormal (reference src/RGBD/RGBD.cpp:101-144) + DepthSensorModel::
// uncertinatyFromNormal (src/Grabber/depthSensorModel.cpp:62-76).  px = n x {u, v} integer pixels
// ((int)it->u, (int)it->v in RGBD::computeNormals, include/putslam/RGBD/RGBD.h:91-95).
__global__ void normal_cov_kernel(const int* __restrict__ px, int n, const uint16_t* __restrict__ depth, int W, int H,
                                  int stride, pslam_camera cam, double depth_scale, double scale_unc,
                                  double* __restrict__ normals, double* __restrict__ cov_out,
                                  double* __restrict__ info_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = px[2 * i], v = px[2 * i + 1];
    float c[3];
    px_to_3d((float)u, (float)v, depth, W, H, stride, cam, depth_scale, c);
    double first[3] = {0, 0, 0}, prev[3] = {0, 0, 0};
    double sx = 0, sy = 0, sz = 0;
    int nv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int du = (k < 3) ? -1 : ((k == 3 || k == 7) ? 0 : 1);
        const int dv = (k == 0 || k == 6 || k == 7) ? -1 : ((k == 1 || k == 5) ? 0 : 1);
        float e[3];
        px_to_3d((float)(u + du), (float)(v + dv), depth, W, H, stride, cam, depth_scale, e);
        if (e[2] > 0.f) {
            const double cur[3] = {(double)(e[0] - c[0]), (double)(e[1] - c[1]), (double)(e[2] - c[2])};
            if (nv == 0) { first[0] = cur[0]; first[1] = cur[1]; first[2] = cur[2]; }
            else {   // cross(prev, cur), accumulated in order
                sx += prev[1] * cur[2] - prev[2] * cur[1];
                sy += prev[2] * cur[0] - prev[0] * cur[2];
                sz += prev[0] * cur[1] - prev[1] * cur[0];
            }
            prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
            ++nv;
        }
    }
    if (nv > 0) {   // closing cross(last, first)
        sx += prev[1] * first[2] - prev[2] * first[1];
        sy += prev[2] * first[0] - prev[0] * first[2];
        sz += prev[0] * first[1] - prev[1] * first[0];
    }
    double nr[3] = {__ddiv_rn(sx, (double)nv), __ddiv_rn(sy, (double)nv), __ddiv_rn(sz, (double)nv)};
    normalize3(nr);
    if (normals) { normals[3 * (size_t)i] = nr[0]; normals[3 * (size_t)i + 1] = nr[1]; normals[3 * (size_t)i + 2] = nr[2]; }
    if (cov_out || info_out) {
        double nn[3] = {nr[0], nr[1], nr[2]};
        double y[3] = {nn[1] * 0.0 - nn[2] * 0.0, nn[2] * 1.0 - nn[0] * 0.0, nn[0] * 0.0 - nn[1] * 1.0};
        normalize3(nn);
        double x[3] = {y[1] * nn[2] - y[2] * nn[1], y[2] * nn[0] - y[0] * nn[2], y[0] * nn[1] - y[1] * nn[0]};
        normalize3(x);
        const double R[9] = {x[0], y[0], nn[0], x[1], y[1], nn[1], x[2], y[2], nn[2]};
        const double S[9] = {1, 0, 0, 0, 1, 0, 0, 0, scale_unc};
        double RS[9], RSS[9], Rinv[9], cov[9];
        matmul3(R, S, RS);
        matmul3(RS, S, RSS);
        inverse3d(R, Rinv);
        matmul3(RSS, Rinv, cov);
        store_cov_info(cov, i, cov_out, info_out);
    }
}

// Uncertainty model 2: RGBD::computeRGBGradient (reference src/RGBD/RGBD.cpp:147-187) + DepthSensorModel::
// uncertinatyFromRGBGradient (src/Grabber/depthSensorModel.cpp:79-95).  The colour patch is read as 16-bit words at
// byte offsets 2c of the patch rows, like the reference's at<uint16_t> (:156-159).  Direction offsets
// coord1 = (int(sqrt2 sin a), int(sqrt2 cos a)), a = atan2(gy, gx) + pi/2, i.e. (trunc(sqrt2 gx/r), trunc(-sqrt2 gy/r)):
// decided with exact integer comparisons; the diagonals |gx| == |gy| come from the host-libm table diag
// (q = (gx<0) + 2(gy<0) -> {c1x, c1y, c2x, c2y}).
struct GradDiag { int t[16]; };
__global__ void gradient_cov_kernel(const int* __restrict__ px, int n, const uint8_t* __restrict__ rgb, int rgb_row_bytes,
                                    const uint16_t* __restrict__ depth, int W, int H, int stride, pslam_camera cam,
                                    double depth_scale, double scale_unc, GradDiag diag, double* __restrict__ grads,
                                    double* __restrict__ cov_out, double* __restrict__ info_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = px[2 * i], v = px[2 * i + 1];
    double g[3];
    if (!((u - 1 > 0) && (v - 1 > 0) && (u + 1 < W) && (v + 1 < H))) {
        g[0] = 1; g[1] = 1; g[2] = 1;
    } else {
        int p[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint8_t* b = rgb + (size_t)(v - 1 + r) * rgb_row_bytes + 3 * (size_t)(u - 1) + 2 * c;
                p[r][c] = (int)b[0] | ((int)b[1] << 8);
            }
        const int gx = -3 * p[0][0] - 10 * p[0][1] - 3 * p[0][2] + 3 * p[2][0] + 10 * p[2][1] + 3 * p[2][2];
        const int gy = -3 * p[0][0] - 10 * p[1][0] - 3 * p[2][0] + 3 * p[0][2] + 10 * p[1][2] + 3 * p[2][2];
        const int ax = gx < 0 ? -gx : gx, ay = gy < 0 ? -gy : gy;
        int c1x, c1y, c2x, c2y;
        if (ax == ay && ax != 0) {
            const int q = (gx < 0 ? 1 : 0) + (gy < 0 ? 2 : 0);
            c1x = diag.t[4 * q]; c1y = diag.t[4 * q + 1]; c2x = diag.t[4 * q + 2]; c2y = diag.t[4 * q + 3];
        } else {
            // sqrt2*|gx|/r >= 1  <=>  |gx| >= |gy|; gx == gy == 0 behaves like atan2(+0,+0) = 0 -> (1, 0)
            c1x = (ax > ay || (ax == 0 && ay == 0)) ? (gx < 0 ? -1 : 1) : 0;
            c1y = (ay > ax) ? (gy < 0 ? 1 : -1) : 0;
            c2x = -c1x; c2y = -c1y;
        }
        float pc[3], pe[3], pb[3];
        px_to_3d((float)u, (float)v, depth, W, H, stride, cam, depth_scale, pc);
        px_to_3d((float)(u + c1x), (float)(v + c1y), depth, W, H, stride, cam, depth_scale, pe);
        px_to_3d((float)(u + c2x), (float)(v + c2y), depth, W, H, stride, cam, depth_scale, pb);
        if (pe[2] > 0.f && pb[2] != 0.f) {
            g[0] = (double)(pe[0] - pb[0]); g[1] = (double)(pe[1] - pb[1]); g[2] = (double)(pe[2] - pb[2]);
        } else if (pc[2] > 0.f && pb[2] != 0.f) {
            g[0] = (double)(pc[0] - pb[0]); g[1] = (double)(pc[1] - pb[1]); g[2] = (double)(pc[2] - pb[2]);
        } else if (pc[2] > 0.f && pe[2] != 0.f) {
            g[0] = (double)(pe[0] - pc[0]); g[1] = (double)(pe[1] - pc[1]); g[2] = (double)(pe[2] - pc[2]);
        } else {
            g[0] = (double)c1x; g[1] = (double)c1y; g[2] = 0.0;
        }
        normalize3(g);
    }
    if (grads) { grads[3 * (size_t)i] = g[0]; grads[3 * (size_t)i + 1] = g[1]; grads[3 * (size_t)i + 2] = g[2]; }
    if (cov_out || info_out) {
        double y[3] = {0.0 * g[2] - 1.0 * g[1], 1.0 * g[0] - 0.0 * g[2], 0.0 * g[1] - 0.0 * g[0]};
        normalize3(y);
        double z[3] = {y[1] * g[2] - y[2] * g[1], y[2] * g[0] - y[0] * g[2], y[0] * g[1] - y[1] * g[0]};
        normalize3(z);
        const double R[9] = {g[0], y[0], z[0], g[1], y[1], z[1], g[2], y[2], z[2]};
        const double S[9] = {1, 0, 0, 0, scale_unc, 0, 0, 0, 1};
        double RS[9], RSS[9], Rinv[9], cov[9];
        matmul3(R, S, RS);
        matmul3(RS, S, RSS);
        inverse3d(R, Rinv);
        matmul3(RSS, Rinv, cov);
        store_cov_info(cov, i, cov_out, info_out);
    }
}

cudaError_t launch_normal_cov(const int* d_px, int n, const uint16_t* d_depth, int W, int H, int stride,
                              const pslam_camera& cam, double depth_scale, double scale_unc, double* d_normals,
                              double* d_cov, double* d_info, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    normal_cov_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_px, n, d_depth, W, H, stride, cam, depth_scale, scale_unc,
                                                       d_normals, d_cov, d_info);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_gradient_cov(const int* d_px, int n, const uint8_t* d_rgb, int rgb_row_bytes, const uint16_t* d_depth,
                                int W, int H, int stride, const pslam_camera& cam, double depth_scale, double scale_unc,
                                const int* diag16, double* d_grads, double* d_cov, double* d_info, cudaStream_t st,
                                int* launches) {
    if (n <= 0) return cudaSuccess;
    GradDiag dg;
    for (int k = 0; k < 16; ++k) dg.t[k] = diag16[k];
    gradient_cov_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_px, n, d_rgb, rgb_row_bytes, d_depth, W, H, stride, cam,
                                                         depth_scale, scale_unc, dg, d_grads, d_cov, d_info);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_information(const double* d_uvz, int n, const pslam_cov_params& cp, double* d_cov, double* d_info,
                               cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    information_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_uvz, n, cp, d_cov, d_info);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_backproject(const float* d_uv, int n, const uint16_t* d_depth, int W, int H, int stride,
                               const pslam_camera& cam, int undistort, double depth_scale, float* d_uv_und,
                               float* d_xyz, double* d_det_dist, double* d_cov, const pslam_cov_params* cov,
                               cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    pslam_cov_params cp = {};
    if (cov) cp = *cov;
    const cudaError_t e = launch_chained(backproject_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, st, d_uv, n,
                                         d_depth, W, H, stride, cam, undistort, depth_scale, d_uv_und, d_xyz, d_det_dist,
                                         d_cov, cp, (cov && d_cov) ? 1 : 0);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
