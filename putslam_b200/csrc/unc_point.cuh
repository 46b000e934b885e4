// unc_point.cuh -- K13: covariance of an estimated rigid transform from the covariances of the matched points,
// TransformEst::computeUncertainty (Euler angles; reference include/putslam/TransformEst/transformEst.h:29-144) and
// computeUncertaintyG2O (quaternion vector part; :147-272), SURVEY 8f rank 4 (callers: demos/demoKabsch.cpp:655,731,1028).
//
// The reference spells the derivatives out as machine-generated scalar expressions.  Here they are derived:
//   J = sum_i |r_i|^2, r_i = a_i - R b_i - t;  g = dJ/dtheta = -2 sum M_i^T r_i,  M_i = [I | D_1 b_i  D_2 b_i  D_3 b_i]
//   H = dg/dtheta = 2 sum (M_i^T M_i - S_i),  S_i[3+k][3+l] = r_i . (D_kl b_i)
//   Ga_i = -2 M_i;  Gb_i = [2 R^T | -2 D_k^T r_i + 2 R^T D_k b_i]            (3 x 6 each)
//   U = H^-1 (sum Ga^T CA Ga + Gb^T CB Gb) H^-1, H and the G's scaled by 1/n as the reference scales them
// with D_k = dR/dparam_k, D_kl the second derivatives; equal to the reference's expressions to 1e-15 relative
// (tests/golden/uncertainty_ref.npz holds the reference's own expressions evaluated from its header).
// Host + device source: the kernel (uncertainty.cu) reduces unc_point() over the points of a problem; the CPU test suite
// runs the same functions sequentially (tests/unc_emul.cpp, test infrastructure).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define UNC_HD __host__ __device__ __forceinline__
#else
#define UNC_HD inline
#endif

namespace pslam {

constexpr int kUncEuler = 0, kUncQuat = 1;
constexpr int kUncAcc = 21 + 36;   // upper triangle of H (row-major order) + Q (6 x 6 row-major)

struct UncRot {
    double R[9];        // row-major
    double D[3][9];     // dR / dparam_k
    double DD[6][9];    // d2R / dparam_k dparam_l for (k,l) = (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
};

// Eigen::Quaternion(Matrix3): trace / largest-diagonal method.  m row-major, q = (w, x, y, z)
UNC_HD void unc_quaternion(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (m[7] - m[5]) * t; q[2] = (m[2] - m[6]) * t; q[3] = (m[3] - m[1]) * t;
        return;
    }
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
}

UNC_HD void unc_mul3(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// rotation about axis k by angle a, and its first and second derivative with respect to a
UNC_HD void unc_axis(double a, int k, double* R, double* d, double* dd) {
    const double c = cos(a), s = sin(a);
    const int i = (k + 1) % 3, j = (k + 2) % 3;
    for (int e = 0; e < 9; ++e) { R[e] = 0.0; d[e] = 0.0; dd[e] = 0.0; }
    R[4 * k] = 1.0;
    R[4 * i] = c; R[3 * i + j] = -s; R[3 * j + i] = s; R[4 * j] = c;
    d[4 * i] = -s; d[3 * i + j] = -c; d[3 * j + i] = c; d[4 * j] = -s;
    dd[4 * i] = -c; dd[3 * i + j] = s; dd[3 * j + i] = -s; dd[4 * j] = -c;
}
// R = Rz(yaw) Ry(pitch) Rx(roll); parameters (roll, pitch, yaw) read off the quaternion as the reference does (:32-39)
UNC_HD void unc_rot_euler(const double* q, UncRot& o) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double roll = atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
    const double pitch = asin(2.0 * (w * y - z * x));
    const double yaw = atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    double X[9], dX[9], ddX[9], Y[9], dY[9], ddY[9], Z[9], dZ[9], ddZ[9], t[9];
    unc_axis(roll, 0, X, dX, ddX); unc_axis(pitch, 1, Y, dY, ddY); unc_axis(yaw, 2, Z, dZ, ddZ);
    unc_mul3(Z, Y, t); unc_mul3(t, X, o.R); unc_mul3(t, dX, o.D[0]); unc_mul3(t, ddX, o.DD[0]);
    unc_mul3(Z, dY, t); unc_mul3(t, X, o.D[1]); unc_mul3(t, dX, o.DD[1]);
    unc_mul3(dZ, Y, t); unc_mul3(t, X, o.D[2]); unc_mul3(t, dX, o.DD[2]);
    unc_mul3(Z, ddY, t); unc_mul3(t, X, o.DD[3]);
    unc_mul3(dZ, dY, t); unc_mul3(t, X, o.DD[4]);
    unc_mul3(ddZ, Y, t); unc_mul3(t, X, o.DD[5]);
}
// quaternion rotation matrix with 1 - 2(..) on the diagonal, differentiated in qx, qy, qz with qw held constant
UNC_HD void unc_rot_quat(const double* q, UncRot& o) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double R[9] = {1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
                         2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w,
                         2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y};
    const double D0[9] = {0, 2 * y, 2 * z, 2 * y, -4 * x, -2 * w, 2 * z, 2 * w, -4 * x};
    const double D1[9] = {-4 * y, 2 * x, 2 * w, 2 * x, 0, 2 * z, -2 * w, 2 * z, -4 * y};
    const double D2[9] = {-4 * z, -2 * w, 2 * x, 2 * w, -4 * z, 2 * y, 2 * x, 2 * y, 0};
    for (int e = 0; e < 9; ++e) {
        o.R[e] = R[e]; o.D[0][e] = D0[e]; o.D[1][e] = D1[e]; o.D[2][e] = D2[e];
        for (int p = 0; p < 6; ++p) o.DD[p][e] = 0.0;
    }
    o.DD[0][4] = -4; o.DD[0][8] = -4;           // d2/dx2
    o.DD[1][1] = 2; o.DD[1][3] = 2;             // d2/dxdy
    o.DD[2][2] = 2; o.DD[2][6] = 2;             // d2/dxdz
    o.DD[3][0] = -4; o.DD[3][8] = -4;           // d2/dy2
    o.DD[4][5] = 2; o.DD[4][7] = 2;             // d2/dydz
    o.DD[5][0] = -4; o.DD[5][4] = -4;           // d2/dz2
}
UNC_HD void unc_mv(const double* M, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2];
}
UNC_HD void unc_mtv(const double* M, const double* v, double* o) {   // M^T v
    for (int i = 0; i < 3; ++i) o[i] = M[i] * v[0] + M[3 + i] * v[1] + M[6 + i] * v[2];
}

// adds point i's contribution: acc[0..20] += upper triangle of 2 (M^T M - S), acc[21..56] += Ga^T CA Ga + Gb^T CB Gb
// CA, CB: 3 x 3 row-major
UNC_HD void unc_point(const UncRot& rot, const double* t, const double* a, const double* b, const double* CA, const double* CB,
                      double* acc) {
    double Rb[3], r[3], M[18], Gb[18];                           // M, Gb: 3 x 6 row-major
    unc_mv(rot.R, b, Rb);
    for (int i = 0; i < 3; ++i) r[i] = a[i] - Rb[i] - t[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { M[6 * i + j] = i == j ? 1.0 : 0.0; Gb[6 * i + j] = 2.0 * rot.R[3 * j + i]; }
    for (int k = 0; k < 3; ++k) {
        double Db[3], Dtr[3], RtDb[3];
        unc_mv(rot.D[k], b, Db); unc_mtv(rot.D[k], r, Dtr); unc_mtv(rot.R, Db, RtDb);
        for (int i = 0; i < 3; ++i) { M[6 * i + 3 + k] = Db[i]; Gb[6 * i + 3 + k] = -2.0 * Dtr[i] + 2.0 * RtDb[i]; }
    }
    double S[6];
    for (int p = 0; p < 6; ++p) {
        double v[3];
        unc_mv(rot.DD[p], b, v);
        S[p] = r[0] * v[0] + r[1] * v[1] + r[2] * v[2];
    }
    int e = 0, p = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j, ++e) {
            double m = M[i] * M[j] + M[6 + i] * M[6 + j] + M[12 + i] * M[12 + j];
            if (i >= 3) m -= S[p++];
            acc[e] += 2.0 * m;
        }
    // Ga = -2 M:  Ga^T CA Ga = 4 M^T CA M
    double CM[18], CG[18];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) {
            CM[6 * i + j] = CA[3 * i] * M[j] + CA[3 * i + 1] * M[6 + j] + CA[3 * i + 2] * M[12 + j];
            CG[6 * i + j] = CB[3 * i] * Gb[j] + CB[3 * i + 1] * Gb[6 + j] + CB[3 * i + 2] * Gb[12 + j];
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j)
            acc[21 + 6 * i + j] += 4.0 * (M[i] * CM[j] + M[6 + i] * CM[6 + j] + M[12 + i] * CM[12 + j]) +
                                   (Gb[i] * CG[j] + Gb[6 + i] * CG[6 + j] + Gb[12 + i] * CG[12 + j]);
}

// U = H^-1 Q H^-1 with H, Q scaled as the reference scales dgdTheta and dgdX (k = 1/n); 6 x 6 inverse by Gauss-Jordan
// elimination with partial pivoting.  U row-major; returns false when H is singular (U is then left untouched).
UNC_HD bool unc_finish(const double* acc, int n, double* U) {
    const double k = 1.0 / (double)n;
    double H[36], I[36], Q[36];
    int e = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j, ++e) { H[6 * i + j] = k * acc[e]; H[6 * j + i] = k * acc[e]; }
    for (int i = 0; i < 36; ++i) { I[i] = (i % 7 == 0) ? 1.0 : 0.0; Q[i] = (k * k) * acc[21 + i]; }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(H[6 * r + c]) > fabs(H[6 * piv + c])) piv = r;
        if (H[6 * piv + c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 6; ++j) {
                double s = H[6 * c + j]; H[6 * c + j] = H[6 * piv + j]; H[6 * piv + j] = s;
                s = I[6 * c + j]; I[6 * c + j] = I[6 * piv + j]; I[6 * piv + j] = s;
            }
        const double d = 1.0 / H[6 * c + c];
        for (int j = 0; j < 6; ++j) { H[6 * c + j] *= d; I[6 * c + j] *= d; }
        for (int r = 0; r < 6; ++r) {
            if (r == c) continue;
            const double f = H[6 * r + c];
            for (int j = 0; j < 6; ++j) { H[6 * r + j] -= f * H[6 * c + j]; I[6 * r + j] -= f * I[6 * c + j]; }
        }
    }
    double P[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int m = 0; m < 6; ++m) s += I[6 * i + m] * Q[6 * m + j];
            P[6 * i + j] = s;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int m = 0; m < 6; ++m) s += P[6 * i + m] * I[6 * m + j];
            U[6 * i + j] = s;
        }
    return true;
}

}  // namespace pslam
