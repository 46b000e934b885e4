// Stand-in for the Eigen 3 headers -- TEST INFRASTRUCTURE ONLY (nothing in the product includes this file).
//
// Purpose: compile the REFERENCE's own translation units where they lie under /root/reference
// (src/TransformEst/RANSAC.cpp, src/RGBD/RGBD.cpp, src/TransformEst/kabschEst.cpp, src/Grabber/depthSensorModel.cpp,
// src/Matcher/matcher.cpp; see oracle/Makefile, target ref) into oracle/_ref/, so that the CPU oracle's restatement of
// the reference's scalar loops (RANSAC driver, filters, thresholds, guided gate, back-projection, covariance) is checked
// against the reference's compiled code rather than against a reading of it.  Eigen is not vendored by the reference
// and does not exist in this image; this header supplies exactly the Eigen API those files use.
//
// WHAT IS LEFT TO THIS SHIM (and therefore NOT pinned by the reference build): the arithmetic ORDER inside Eigen's own
// operations.  Each one below names the Eigen 3.3 code it models.  The reference needs Eigen >= 3.3 (it uses
// Eigen::Index, include/putslam/TransformEst/transformEst.h:45) and builds with -DEIGEN_DONT_VECTORIZE, -std=c++11
// (scalar IEEE-754, no FMA; CMakeLists.txt:23,147).  Evaluation is eager (every operator returns a plain Matrix);
// Eigen's lazy evaluation yields the same per-coefficient expressions for the uses in those files.
//
// Compile-time switches (used by tools/umeyama_sensitivity.py to bound what cannot be pinned here):
//   SHIM_FIXED_REDUX_TREE   1 (default): fixed-size reductions (sum, dot, norm, and -- through .sum() -- the inner
//                           product of small fixed products) split in halves like redux_novec_unroller: c0 + (c1 + c2),
//                           (c0 + c1) + (c2 + c3).  0: every reduction strictly left to right -- NO Eigen version does
//                           that to norm(); kept to show how sharp the level gate of matchXYZ is.
//   SHIM_PRODUCT_COEFF_SEQ  0 (default): product coefficients are (row . col).sum() (Eigen 3.3).  1: left to right
//                           (Eigen 3.2's product_coeff_impl) while norm()/sum() keep the halving both versions have.
//   SHIM_UMEYAMA_SCALE_LHS  0 (default): sigma = one_over_n * (dst_demean * src_demean^T) (scalar factored out of the
//                           product: GEMM path alpha, and the lazy path since 3.3.8).  1: (one_over_n * dst_demean) *
//                           src_demean^T (lazy path of 3.3.0-3.3.7 for n + 6 < 20).
//   SHIM_JACOBI_THRESHOLD32 0 (default): 3.3 sweep threshold max(considerAsZero, 2 eps * maxDiagEntry) with the input
//                           scaled by its largest coefficient.  1: Eigen 3.2 (no scaling, threshold 2 eps *
//                           max(|w_pp|, |w_qq|)).
//   SHIM_JACOBI_SWEEP_ALT   1: the sweep visits the index pairs in the opposite order (p = n-1..1, q = p-1..0) -- not an
//                           Eigen version, a perturbation that bounds the sensitivity to the SVD's rounding path.
//   SHIM_UMEYAMA_F64        1: float umeyama evaluated in double and rounded once -- the "ideal" answer.
#ifndef PSLAM_REF_SHIM_EIGEN_H
#define PSLAM_REF_SHIM_EIGEN_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

#ifndef SHIM_FIXED_REDUX_TREE
#define SHIM_FIXED_REDUX_TREE 1
#endif
#ifndef SHIM_PRODUCT_COEFF_SEQ
#define SHIM_PRODUCT_COEFF_SEQ 0
#endif
#ifndef SHIM_UMEYAMA_SCALE_LHS
#define SHIM_UMEYAMA_SCALE_LHS 0
#endif
#ifndef SHIM_JACOBI_THRESHOLD32
#define SHIM_JACOBI_THRESHOLD32 0
#endif
#ifndef SHIM_JACOBI_SWEEP_ALT
#define SHIM_JACOBI_SWEEP_ALT 0
#endif
#ifndef SHIM_UMEYAMA_F64
#define SHIM_UMEYAMA_F64 0
#endif

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <typename T, int R, int C> class Matrix;
template <typename T, int R, int C> class View;

namespace shim {

template <typename T, int N> struct Store {
    T d[N];
    Store() { for (int i = 0; i < N; ++i) d[i] = T(); }
    void resize(std::size_t) {}
    T* data() { return d; }
    const T* data() const { return d; }
};
template <typename T> struct Store<T, Dynamic> {
    std::vector<T> d;
    void resize(std::size_t n) { d.assign(n, T()); }
    T* data() { return d.data(); }
    const T* data() const { return d.data(); }
};
template <int R, int C> struct SizeOf { enum { value = (R == Dynamic || C == Dynamic) ? Dynamic : R * C }; };

// Eigen 3.3 redux_novec_unroller (Core/Redux.h): a completely unrolled fixed-size reduction splits [start, start+len)
// in halves, HalfLength = len / 2.  Dynamic sizes reduce left to right (redux_impl<DefaultTraversal, NoUnrolling>).
template <typename T, typename F> T tree(const F& c, Index start, Index len) {
    if (len == 1) return c(start);
    const Index half = len / 2;
    return tree<T>(c, start, half) + tree<T>(c, start + half, len - half);
}
template <typename T, typename F> T reduce(const F& c, Index n, bool fixed) {
    if (n == 0) return T(0);
    if (fixed && SHIM_FIXED_REDUX_TREE) return tree<T>(c, 0, n);
    T s = c(0);
    for (Index i = 1; i < n; ++i) s = s + c(i);
    return s;
}

template <typename D> struct traits;
}  // namespace shim

// ---- common read interface (CRTP) ----------------------------------------------------------------------------------
template <typename D> class Base {
public:
    typedef typename shim::traits<D>::Scalar Scalar;
    enum { Rows = shim::traits<D>::Rows, Cols = shim::traits<D>::Cols, RowsAtCompileTime = Rows, ColsAtCompileTime = Cols,
           Fixed = (Rows != Dynamic && Cols != Dynamic) };
    typedef Matrix<Scalar, Rows, Cols> Plain;
    const D& derived() const { return *static_cast<const D*>(this); }
    D& derived() { return *static_cast<D*>(this); }
    Index rows() const { return derived().rows_(); }
    Index cols() const { return derived().cols_(); }
    Index size() const { return rows() * cols(); }
    Scalar coeff(Index i, Index j) const { return derived().at_(i, j); }
    Scalar coeff(Index i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar operator()(Index i, Index j) const { return coeff(i, j); }
    Scalar operator()(Index i) const { return coeff(i); }
    Scalar operator[](Index i) const { return coeff(i); }
    Scalar x() const { return coeff(0); }
    Scalar y() const { return coeff(1); }
    Scalar z() const { return coeff(2); }
    Scalar w() const { return coeff(3); }
    Plain eval() const {
        Plain r(rows(), cols(), 0);
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.ref(i, j) = coeff(i, j);
        return r;
    }
    const D& matrix() const { return derived(); }
    D& matrix() { return derived(); }

    // reductions: column-major linear order
    Scalar sum() const {
        const Index r = rows();
        return shim::reduce<Scalar>([&](Index k) { return coeff(k % r, k / r); }, size(), Fixed);
    }
    Scalar mean() const { return sum() / Scalar(size()); }        // DenseBase::mean(): redux(sum) / size
    Scalar squaredNorm() const {                                  // cwiseAbs2().sum()
        const Index r = rows();
        return shim::reduce<Scalar>([&](Index k) { Scalar v = coeff(k % r, k / r); return v * v; }, size(), Fixed);
    }
    Scalar norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    template <typename O> Scalar dot(const Base<O>& o) const {
        return shim::reduce<Scalar>([&](Index k) { return coeff(k) * o.coeff(k); }, size(), Fixed);
    }
    bool hasNaN() const {
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) != coeff(i, j)) return true;
        return false;
    }
    Scalar maxCoeff() const {
        Scalar m = coeff(0, 0);
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) > m) m = coeff(i, j);
        return m;
    }
    Plain cwiseAbs() const {
        using std::abs;
        Plain r(rows(), cols(), 0);
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.ref(i, j) = abs(coeff(i, j));
        return r;
    }
    template <typename U> Matrix<U, Rows, Cols> cast() const {
        Matrix<U, Rows, Cols> r(rows(), cols(), 0);
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.ref(i, j) = U(coeff(i, j));
        return r;
    }
    template <typename O> Matrix<Scalar, 3, 1> cross(const Base<O>& o) const {   // Geometry/OrthoMethods.h
        Matrix<Scalar, 3, 1> r;
        r.ref(0, 0) = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
        r.ref(1, 0) = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
        r.ref(2, 0) = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
        return r;
    }
    Scalar determinant() const;
    Plain inverse() const;
};

namespace shim {
template <typename T, int R, int C> struct traits<Matrix<T, R, C> > { typedef T Scalar; enum { Rows = R, Cols = C }; };
template <typename T, int R, int C> struct traits<View<T, R, C> > { typedef T Scalar; enum { Rows = R, Cols = C }; };
}  // namespace shim

// ---- a strided window onto someone else's storage (block / row / col / transpose / head) ---------------------------
template <typename T, int R, int C> class View : public Base<View<T, R, C> > {
public:
    T* p; Index r, c, rs, cs;
    View(T* p_, Index r_, Index c_, Index rs_, Index cs_) : p(p_), r(r_), c(c_), rs(rs_), cs(cs_) {}
    Index rows_() const { return r; }
    Index cols_() const { return c; }
    T at_(Index i, Index j) const { return p[i * rs + j * cs]; }
    T& ref(Index i, Index j) { return p[i * rs + j * cs]; }
    T& operator()(Index i, Index j) { return ref(i, j); }
    T& operator()(Index i) { return c == 1 ? ref(i, 0) : ref(0, i); }
    T& operator[](Index i) { return (*this)(i); }
    T operator()(Index i, Index j) const { return at_(i, j); }
    T operator()(Index i) const { return c == 1 ? at_(i, 0) : at_(0, i); }
    T operator[](Index i) const { return (*this)(i); }
    template <typename O> View& operator=(const Base<O>& o) {
        auto v = o.eval();                       // aliasing-safe
        if ((r == 1 || c == 1) && v.rows() == c && v.cols() == r) {     // Eigen transposes a vector assigned to a vector
            for (Index k = 0; k < r * c; ++k) (*this)(k) = v.coeff(k);
            return *this;
        }
        assert(v.rows() == r && v.cols() == c);
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = v.coeff(i, j);
        return *this;
    }
    View& operator=(const View& o) { return operator=<View>(o); }
    View<T, C, R> transpose() const { return View<T, C, R>(p, c, r, cs, rs); }
    View& noalias() { return *this; }
    template <typename O> View& operator-=(const Base<O>& o) {
        auto v = o.eval();
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = ref(i, j) - v.coeff(i, j);
        return *this;
    }
    View& operator*=(T s) { for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = ref(i, j) * s; return *this; }
    template <int N> View<T, N, 1> head() const { return View<T, N, 1>(p, N, 1, c == 1 ? rs : cs, 0); }
};

// ---- comma initialiser -----------------------------------------------------------------------------------------------
template <typename M> class CommaInit {
    M& m; Index k;
public:
    CommaInit(M& m_, typename M::Scalar v) : m(m_), k(0) { put(v); }
    void put(typename M::Scalar v) { m.ref(k / m.cols(), k % m.cols()) = v; ++k; }     // row by row
    CommaInit& operator,(typename M::Scalar v) { put(v); return *this; }
};

// ---- the plain dense matrix (column-major like Eigen's default) -------------------------------------------------------
template <typename T, int R, int C> class Matrix : public Base<Matrix<T, R, C> > {
    shim::Store<T, shim::SizeOf<R, C>::value> s;
    Index r, c;
public:
    typedef T Scalar;
    typedef Base<Matrix<T, R, C> > B;
    Matrix() : r(R == Dynamic ? 0 : R), c(C == Dynamic ? 0 : C) {}
    Matrix(Index rows, Index cols, int /*tag: sized*/) : r(rows), c(cols) { s.resize((std::size_t)(rows * cols)); }
    // (a, b): sizes for a dynamic matrix, the two coefficients for a fixed 2-vector
    Matrix(double a, double b) : r(R == Dynamic ? (Index)a : R), c(C == Dynamic ? (Index)b : C) {
        if (R == Dynamic || C == Dynamic) s.resize((std::size_t)(r * c));
        else { s.data()[0] = T(a); s.data()[1] = T(b); }
    }
    Matrix(T a, T b, T cc) : r(R), c(C) { s.data()[0] = a; s.data()[1] = b; s.data()[2] = cc; }
    Matrix(T a, T b, T cc, T d) : r(R), c(C) { s.data()[0] = a; s.data()[1] = b; s.data()[2] = cc; s.data()[3] = d; }
    template <typename O> Matrix(const Base<O>& o) : r(o.rows()), c(o.cols()) {
        s.resize((std::size_t)(r * c));
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = o.coeff(i, j);
    }
    template <typename O> Matrix& operator=(const Base<O>& o) {
        Matrix t(o);
        r = t.r; c = t.c; s = t.s;
        return *this;
    }
    Index rows_() const { return r; }
    Index cols_() const { return c; }
    T at_(Index i, Index j) const { return s.data()[i + j * r]; }
    T& ref(Index i, Index j) { return s.data()[i + j * r]; }
    T* data() { return s.data(); }
    const T* data() const { return s.data(); }
    void resize(Index rows, Index cols) { r = rows; c = cols; s.resize((std::size_t)(rows * cols)); }

    using B::operator();
    using B::operator[];
    using B::x; using B::y; using B::z; using B::w;
    T& operator()(Index i, Index j) { return ref(i, j); }
    T& operator()(Index i) { return c == 1 ? ref(i, 0) : ref(0, i); }
    T& operator[](Index i) { return (*this)(i); }
    T& x() { return (*this)(0); }
    T& y() { return (*this)(1); }
    T& z() { return (*this)(2); }
    T& w() { return (*this)(3); }

    CommaInit<Matrix> operator<<(T v) { return CommaInit<Matrix>(*this, v); }
    // a 1 x 1 result (an inner product written as a matrix product) converts to its scalar
    template <typename U, typename = typename std::enable_if<R == 1 && C == 1 && std::is_arithmetic<U>::value>::type>
    operator U() const { return U(s.data()[0]); }

    // views
    template <int BR, int BC> View<T, BR, BC> block(Index i, Index j) { return View<T, BR, BC>(&ref(i, j), BR, BC, 1, r); }
    template <int BR, int BC> View<T, BR, BC> block(Index i, Index j) const { return View<T, BR, BC>(const_cast<T*>(&s.data()[i + j * r]), BR, BC, 1, r); }
    View<T, Dynamic, Dynamic> block(Index i, Index j, Index br, Index bc) { return View<T, Dynamic, Dynamic>(&ref(i, j), br, bc, 1, r); }
    View<T, Dynamic, Dynamic> block(Index i, Index j, Index br, Index bc) const { return View<T, Dynamic, Dynamic>(const_cast<T*>(&s.data()[i + j * r]), br, bc, 1, r); }
    View<T, Dynamic, Dynamic> topLeftCorner(Index br, Index bc) const { return block(0, 0, br, bc); }
    View<T, 1, C> row(Index i) const { return View<T, 1, C>(const_cast<T*>(s.data()) + i, 1, c, 1, r); }
    View<T, R, 1> col(Index j) const { return View<T, R, 1>(const_cast<T*>(s.data()) + j * r, r, 1, 1, r); }
    View<T, C, R> transpose() const { return View<T, C, R>(const_cast<T*>(s.data()), c, r, r, 1); }
    template <int N> View<T, N, 1> head() const { return View<T, N, 1>(const_cast<T*>(s.data()), N, 1, 1, r); }
    Matrix& noalias() { return *this; }

    Matrix& setZero() { for (Index k = 0; k < r * c; ++k) s.data()[k] = T(0); return *this; }
    Matrix& setIdentity() { setZero(); for (Index k = 0; k < std::min(r, c); ++k) ref(k, k) = T(1); return *this; }
    Matrix& setIdentity(Index rows, Index cols) { resize(rows, cols); return setIdentity(); }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index rows, Index cols) { Matrix m(rows, cols, 0); m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(Index rows, Index cols) { Matrix m(rows, cols, 0); m.setIdentity(); return m; }
    static Matrix Ones(Index n) { Matrix m(R == Dynamic ? n : R, C == Dynamic ? (C == 1 ? 1 : n) : C, 0); for (Index k = 0; k < m.size(); ++k) m.data()[k] = T(1); return m; }

    template <typename O> Matrix& operator+=(const Base<O>& o) {
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = ref(i, j) + o.coeff(i, j);
        return *this;
    }
    template <typename O> Matrix& operator-=(const Base<O>& o) {
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = ref(i, j) - o.coeff(i, j);
        return *this;
    }
    Matrix& operator*=(T v) { for (Index k = 0; k < r * c; ++k) s.data()[k] = s.data()[k] * v; return *this; }
    Matrix& operator/=(T v) { for (Index k = 0; k < r * c; ++k) s.data()[k] = s.data()[k] / v; return *this; }   // true division (3.3)
};

typedef Matrix<float, 2, 1> Vector2f;   typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<float, 3, 1> Vector3f;   typedef Matrix<double, 3, 1> Vector3d;   typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, 4, 1> Vector4f;   typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 2> Matrix2f;   typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<float, 3, 3> Matrix3f;   typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;   typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;   typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, 1> VectorXf;         typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> ArrayXXd;  // only ArrayXXd::Zero(r, c) assigned to a MatrixXd is used (transformEst.h:42)

// ---- arithmetic --------------------------------------------------------------------------------------------------------
template <typename A, typename Bq>
Matrix<typename A::Scalar, A::Rows, A::Cols> operator+(const Base<A>& a, const Base<Bq>& b) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) + b.coeff(i, j);
    return r;
}
template <typename A, typename Bq>
Matrix<typename A::Scalar, A::Rows, A::Cols> operator-(const Base<A>& a, const Base<Bq>& b) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) - b.coeff(i, j);
    return r;
}
template <typename A> Matrix<typename A::Scalar, A::Rows, A::Cols> operator-(const Base<A>& a) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = -a.coeff(i, j);
    return r;
}
template <typename A> Matrix<typename A::Scalar, A::Rows, A::Cols> operator*(const Base<A>& a, typename A::Scalar s) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) * s;
    return r;
}
template <typename A> Matrix<typename A::Scalar, A::Rows, A::Cols> operator*(typename A::Scalar s, const Base<A>& a) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = s * a.coeff(i, j);
    return r;
}
template <typename A> Matrix<typename A::Scalar, A::Rows, A::Cols> operator/(const Base<A>& a, typename A::Scalar s) {
    Matrix<typename A::Scalar, A::Rows, A::Cols> r(a.rows(), a.cols(), 0);
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) / s;
    return r;
}
// Matrix product.  Eigen 3.3 coefficient-based product (Core/ProductEvaluators.h, product_evaluator<LazyProduct>):
//   coeff(i,j) = (lhs.row(i).transpose().cwiseProduct(rhs.col(j))).sum()
// -> a fixed inner size reduces through the unrolled tree, a dynamic inner size left to right; the general (GEMM) path
// used for large dynamic products also accumulates each result coefficient left to right over the inner index within
// one depth block (gebp_kernel) -- blocks of depth kc (a function of the HOST's cache sizes, > 256 floats on the CPUs
// we know of) are not modelled.
template <typename A, typename Bq>
Matrix<typename A::Scalar, A::Rows, Bq::Cols> operator*(const Base<A>& a, const Base<Bq>& b) {
    typedef typename A::Scalar T;
    assert(a.cols() == b.rows());
    Matrix<T, A::Rows, Bq::Cols> r(a.rows(), b.cols(), 0);
    // the summed expression lhs.row(i).transpose().cwiseProduct(rhs.col(j)) takes its compile-time size from its LEFT operand
    // (traits<CwiseBinaryOp>: Ancestor = Lhs), so the reduction is unrolled iff the left matrix has a fixed column count
    const bool fixed_inner = (A::Cols != Dynamic);
    const Index n = a.cols();
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i)
            r.ref(i, j) = shim::reduce<T>([&](Index k) { return a.coeff(i, k) * b.coeff(k, j); }, n, fixed_inner && !SHIM_PRODUCT_COEFF_SEQ);
    return r;
}

template <typename D> std::ostream& operator<<(std::ostream& os, const Base<D>& m) {
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m.coeff(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

// ---- determinant / inverse ---------------------------------------------------------------------------------------------
namespace shim {
// partial-pivoting LU of a square matrix (row-major scratch); returns the permutation sign, 0 if singular at a pivot
template <typename T> int lu(std::vector<T>& a, Index n, std::vector<Index>& piv) {
    using std::abs;
    int sign = 1;
    piv.resize((std::size_t)n);
    for (Index k = 0; k < n; ++k) {
        Index p = k; T best = abs(a[k * n + k]);
        for (Index i = k + 1; i < n; ++i) if (abs(a[i * n + k]) > best) { best = abs(a[i * n + k]); p = i; }
        piv[k] = p;
        if (p != k) { for (Index j = 0; j < n; ++j) std::swap(a[k * n + j], a[p * n + j]); sign = -sign; }
        if (a[k * n + k] == T(0)) continue;
        for (Index i = k + 1; i < n; ++i) {
            a[i * n + k] = a[i * n + k] / a[k * n + k];
            for (Index j = k + 1; j < n; ++j) a[i * n + j] = a[i * n + j] - a[i * n + k] * a[k * n + j];
        }
    }
    return sign;
}
}  // namespace shim

template <typename D> typename Base<D>::Scalar Base<D>::determinant() const {
    typedef Scalar T;
    const Index n = rows();
    assert(n == cols());
    if (Fixed && n == 3) {   // LU/Determinant.h bruteforce_det3_helper
        auto h = [&](int a, int b, int c) { return coeff(0, a) * (coeff(1, b) * coeff(2, c) - coeff(1, c) * coeff(2, b)); };
        return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
    }
    if (Fixed && n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);
    if (n == 0) return T(1);
    // dynamic (or > 4): m.partialPivLu().determinant() = sign * prod(diag)
    std::vector<T> a((std::size_t)(n * n)); std::vector<Index> piv;
    for (Index i = 0; i < n; ++i) for (Index j = 0; j < n; ++j) a[i * n + j] = coeff(i, j);
    T d = T(shim::lu(a, n, piv));
    for (Index k = 0; k < n; ++k) d = d * a[k * n + k];
    return d;
}

template <typename D> typename Base<D>::Plain Base<D>::inverse() const {
    typedef Scalar T;
    const Index n = rows();
    assert(n == cols());
    Plain res(n, n, 0);
    if (Fixed && n == 3) {   // LU/InverseImpl.h compute_inverse<MatrixType, ResultType, 3>
        auto cf = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1);
        };
        const T c0 = cf(0, 0), c1 = cf(1, 0), c2 = cf(2, 0);
        const T cc[3] = {c0, c1, c2};
        const T det = shim::reduce<T>([&](Index k) { return cc[k] * coeff(k, 0); }, 3, true);
        const T invdet = T(1) / det;
        res.ref(0, 0) = c0 * invdet; res.ref(0, 1) = c1 * invdet; res.ref(0, 2) = c2 * invdet;
        res.ref(1, 0) = cf(0, 1) * invdet; res.ref(1, 1) = cf(1, 1) * invdet; res.ref(2, 1) = cf(1, 2) * invdet;
        res.ref(1, 2) = cf(2, 1) * invdet; res.ref(2, 0) = cf(0, 2) * invdet; res.ref(2, 2) = cf(2, 2) * invdet;
        return res;
    }
    if (Fixed && n == 4) {   // compute_inverse_size4 (generic, non-vectorised): cofactors, then /= det along column 0
        auto d3 = [&](int i1, int i2, int i3, int j1, int j2, int j3) {
            return coeff(i1, j1) * (coeff(i2, j2) * coeff(i3, j3) - coeff(i2, j3) * coeff(i3, j2));
        };
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4, j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
            return d3(i1, i2, i3, j1, j2, j3) + d3(i2, i3, i1, j1, j2, j3) + d3(i3, i1, i2, j1, j2, j3);
        };
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { T c = cof(i, j); res.ref(j, i) = ((i + j) & 1) ? -c : c; }
        const T det = shim::reduce<T>([&](Index k) { return coeff(k, 0) * res.coeff(0, k); }, 4, true);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) res.ref(i, j) = res.coeff(i, j) / det;
        return res;
    }
    if (Fixed && n == 2) {
        const T invdet = T(1) / determinant();
        res.ref(0, 0) = coeff(1, 1) * invdet; res.ref(1, 0) = -coeff(1, 0) * invdet;
        res.ref(0, 1) = -coeff(0, 1) * invdet; res.ref(1, 1) = coeff(0, 0) * invdet;
        return res;
    }
    // general: partialPivLu().inverse() = solve(Identity)
    std::vector<T> a((std::size_t)(n * n)); std::vector<Index> piv;
    for (Index i = 0; i < n; ++i) for (Index j = 0; j < n; ++j) a[i * n + j] = coeff(i, j);
    shim::lu(a, n, piv);
    for (Index col = 0; col < n; ++col) {
        std::vector<T> b((std::size_t)n, T(0));
        b[col] = T(1);
        for (Index k = 0; k < n; ++k) std::swap(b[k], b[piv[k]]);
        for (Index i = 0; i < n; ++i) for (Index k = 0; k < i; ++k) b[i] = b[i] - a[i * n + k] * b[k];
        for (Index i = n - 1; i >= 0; --i) {
            for (Index k = i + 1; k < n; ++k) b[i] = b[i] - a[i * n + k] * b[k];
            b[i] = b[i] / a[i * n + i];
        }
        for (Index i = 0; i < n; ++i) res.ref(i, col) = b[i];
    }
    return res;
}

// ---- JacobiSVD (SVD/JacobiSVD.h, Jacobi/Jacobi.h, misc/RealSvd2x2.h), square real input --------------------------------
template <typename MatrixType> class JacobiSVD {
public:
    typedef typename MatrixType::Scalar T;
    typedef Matrix<T, Dynamic, Dynamic> Dyn;
    JacobiSVD(const MatrixType& m, unsigned int = 0) { compute(m); }
    const MatrixType& matrixU() const { return U; }
    const MatrixType& matrixV() const { return V; }
    const Matrix<T, MatrixType::Rows, 1>& singularValues() const { return S; }

private:
    MatrixType U, V;
    Matrix<T, MatrixType::Rows, 1> S;
    struct Rot { T c, s; };
    // apply_rotation_in_the_plane on (x_i, y_i): x <- c x + s y ; y <- -s x + c y ; identity rotations are skipped
    template <typename FX, typename FY> static void plane(Index n, FX x, FY y, Rot j) {
        if (j.c == T(1) && j.s == T(0)) return;
        for (Index i = 0; i < n; ++i) {
            const T xi = x(i), yi = y(i);
            x(i) = j.c * xi + j.s * yi;
            y(i) = -j.s * xi + j.c * yi;
        }
    }
    static void onTheLeft(MatrixType& M, Index p, Index q, Rot j) {       // rows p, q rotated by j
        plane(M.cols(), [&](Index i) -> T& { return M(p, i); }, [&](Index i) -> T& { return M(q, i); }, j);
    }
    static void onTheRight(MatrixType& M, Index p, Index q, Rot j) {      // columns p, q rotated by j.transpose()
        Rot t = {j.c, -j.s};
        plane(M.rows(), [&](Index i) -> T& { return M(i, p); }, [&](Index i) -> T& { return M(i, q); }, t);
    }
    static Rot makeJacobi(T x, T y, T z) {                                 // JacobiRotation::makeJacobi(x, y, z)
        using std::abs; using std::sqrt;
        Rot r;
        const T deno = T(2) * abs(y);
        if (deno < (std::numeric_limits<T>::min)()) { r.c = T(1); r.s = T(0); return r; }
        const T tau = (x - z) / deno;
        const T w = sqrt(tau * tau + T(1));
        T t;
        if (tau > T(0)) t = T(1) / (tau + w); else t = T(1) / (tau - w);
        const T sign_t = t > T(0) ? T(1) : T(-1);
        const T n = T(1) / sqrt(t * t + T(1));
        r.s = -sign_t * (y / abs(y)) * abs(t) * n;
        r.c = n;
        return r;
    }
    void compute(const MatrixType& matrix) {
        using std::abs; using std::sqrt;
        const Index n = matrix.rows();
        assert(n == matrix.cols());
        const T precision = T(2) * std::numeric_limits<T>::epsilon();
        const T considerAsZero = (std::numeric_limits<T>::min)();
        MatrixType W = matrix;
        T scale = T(1);
#if !SHIM_JACOBI_THRESHOLD32
        scale = matrix.cwiseAbs().maxCoeff();
        if (scale == T(0)) scale = T(1);
        W = MatrixType(matrix / scale);
#endif
        U = MatrixType::Identity(n, n); V = MatrixType::Identity(n, n);
        T maxDiagEntry = T(0);
        for (Index i = 0; i < n; ++i) maxDiagEntry = std::max(maxDiagEntry, abs(W(i, i)));
        bool finished = false;
        while (!finished) {
            finished = true;
#if SHIM_JACOBI_SWEEP_ALT
            for (Index p = n - 1; p >= 1; --p)
                for (Index q = p - 1; q >= 0; --q) {
#else
            for (Index p = 1; p < n; ++p)
                for (Index q = 0; q < p; ++q) {
#endif
#if SHIM_JACOBI_THRESHOLD32
                    const T threshold = std::max(considerAsZero, precision * std::max(abs(W(p, p)), abs(W(q, q))));
#else
                    const T threshold = std::max(considerAsZero, precision * maxDiagEntry);
#endif
                    if (abs(W(p, q)) > threshold || abs(W(q, p)) > threshold) {
                        finished = false;
                        // real_2x2_jacobi_svd(W, p, q, &j_left, &j_right)
                        T m00 = W(p, p), m01 = W(p, q), m10 = W(q, p), m11 = W(q, q);
                        Rot rot1;
                        const T t = m00 + m11, d = m10 - m01;
                        if (abs(d) < (std::numeric_limits<T>::min)()) { rot1.s = T(0); rot1.c = T(1); }
                        else { const T u = t / d; const T tmp = sqrt(T(1) + u * u); rot1.s = T(1) / tmp; rot1.c = u / tmp; }
                        if (!(rot1.c == T(1) && rot1.s == T(0))) {   // m.applyOnTheLeft(0, 1, rot1)
                            const T x0 = m00, y0 = m10, x1 = m01, y1 = m11;
                            m00 = rot1.c * x0 + rot1.s * y0; m10 = -rot1.s * x0 + rot1.c * y0;
                            m01 = rot1.c * x1 + rot1.s * y1; m11 = -rot1.s * x1 + rot1.c * y1;
                        }
                        const Rot j_right = makeJacobi(m00, m01, m11);
                        const Rot jrt = {j_right.c, -j_right.s};
                        // *j_left = rot1 * j_right->transpose()
                        const Rot j_left = {rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
                        onTheLeft(W, p, q, j_left);
                        const Rot jlt = {j_left.c, -j_left.s};
                        onTheRight(U, p, q, jlt);
                        onTheRight(W, p, q, j_right);
                        onTheRight(V, p, q, j_right);
                        maxDiagEntry = std::max(maxDiagEntry, std::max(abs(W(p, p)), abs(W(q, q))));
                    }
                }
        }
        S = Matrix<T, MatrixType::Rows, 1>(n, 1, 0);
        for (Index i = 0; i < n; ++i) {
            const T a = W(i, i);
            S(i) = abs(a);
            if (a < T(0)) for (Index k = 0; k < n; ++k) U(k, i) = U(k, i) * T(-1);
        }
        for (Index i = 0; i < n; ++i) S(i) = S(i) * scale;
        for (Index i = 0; i < n; ++i) {
            Index pos = 0; T mx = S(i);
            for (Index k = 1; k < n - i; ++k) if (S(i + k) > mx) { mx = S(i + k); pos = k; }
            if (mx == T(0)) break;
            if (pos) {
                pos += i;
                std::swap(S(i), S(pos));
                for (Index k = 0; k < n; ++k) { std::swap(U(k, i), U(k, pos)); std::swap(V(k, i), V(k, pos)); }
            }
        }
    }
};

// ---- Eigen::umeyama (Geometry/Umeyama.h) for dynamic inputs (m x n, points in columns) -------------------------------
template <typename A, typename Bq>
Matrix<typename A::Scalar, Dynamic, Dynamic> umeyama(const Base<A>& src, const Base<Bq>& dst, bool with_scaling = true) {
    typedef typename A::Scalar T;
    typedef Matrix<T, Dynamic, Dynamic> Mx;
    const Index m = src.rows(), n = src.cols();
#if SHIM_UMEYAMA_F64
    if (sizeof(T) == sizeof(float)) {
        const Matrix<double, Dynamic, Dynamic> sd = src.template cast<double>(), dd = dst.template cast<double>();
        return Mx(umeyama(sd, dd, with_scaling).template cast<T>());
    }
#endif
    const T one_over_n = T(1) / static_cast<T>(n);
    // src.rowwise().sum() * one_over_n  (dynamic number of columns: left to right)
    Mx src_mean(m, 1, 0), dst_mean(m, 1, 0);
    for (Index i = 0; i < m; ++i) {
        src_mean(i, 0) = shim::reduce<T>([&](Index k) { return src.coeff(i, k); }, n, false) * one_over_n;
        dst_mean(i, 0) = shim::reduce<T>([&](Index k) { return dst.coeff(i, k); }, n, false) * one_over_n;
    }
    Mx src_demean(m, n, 0), dst_demean(m, n, 0);
    for (Index k = 0; k < n; ++k) for (Index i = 0; i < m; ++i) {
        src_demean(i, k) = src.coeff(i, k) - src_mean(i, 0);
        dst_demean(i, k) = dst.coeff(i, k) - dst_mean(i, 0);
    }
    // Eq. (38): sigma = one_over_n * dst_demean * src_demean.transpose()
    Mx sigma(m, m, 0);
    for (Index j = 0; j < m; ++j) for (Index i = 0; i < m; ++i) {
#if SHIM_UMEYAMA_SCALE_LHS
        sigma(i, j) = shim::reduce<T>([&](Index k) { return (one_over_n * dst_demean(i, k)) * src_demean(j, k); }, n, false);
#else
        sigma(i, j) = one_over_n * shim::reduce<T>([&](Index k) { return dst_demean(i, k) * src_demean(j, k); }, n, false);
#endif
    }
    JacobiSVD<Mx> svd(sigma, ComputeFullU | ComputeFullV);
    Mx Rt = Mx::Identity(m + 1, m + 1);
    // Eq. (39): S = 1 ; if det(U) * det(V) < 0 : S(m-1) = -1      (dynamic determinant: partialPivLu)
    Mx S = Mx::Ones(m); S.resize(m, 1); for (Index i = 0; i < m; ++i) S(i, 0) = T(1);
    if (svd.matrixU().determinant() * svd.matrixV().determinant() < T(0)) S(m - 1, 0) = T(-1);
    // Eq. (40): R = U * S.asDiagonal() * V^T      ((U * diag) first, then the product with V^T: dynamic inner size)
    Mx US(m, m, 0);
    for (Index j = 0; j < m; ++j) for (Index i = 0; i < m; ++i) US(i, j) = svd.matrixU()(i, j) * S(j, 0);
    Mx R = US * svd.matrixV().transpose();
    Rt.block(0, 0, m, m) = R;
    if (with_scaling) {
        T src_var = T(0);
        for (Index k = 0; k < n; ++k) for (Index i = 0; i < m; ++i) src_var = src_var + src_demean(i, k) * src_demean(i, k);
        src_var = src_var * one_over_n;
        T tr = T(0);
        for (Index i = 0; i < m; ++i) tr = tr + svd.singularValues()(i) * S(i, 0);
        const T c = T(1) / src_var * tr;
        Mx Rs = R * src_mean;
        for (Index i = 0; i < m; ++i) Rt(i, m) = dst_mean(i, 0) - c * Rs(i, 0);
        for (Index j = 0; j < m; ++j) for (Index i = 0; i < m; ++i) Rt(i, j) = Rt(i, j) * c;
    } else {
        // Rt.col(m).head(m) = dst_mean;  Rt.col(m).head(m).noalias() -= Rt.topLeftCorner(m, m) * src_mean;
        Mx Rs = R * src_mean;
        for (Index i = 0; i < m; ++i) Rt(i, m) = dst_mean(i, 0) - Rs(i, 0);
    }
    return Rt;
}

// ---- Geometry: Translation, Quaternion, Transform -------------------------------------------------------------------
template <typename T, int Dim, int Mode> class Transform;

template <typename T, int Dim> class Translation {
    Matrix<T, Dim, 1> v;
public:
    Translation() {}
    Translation(T x, T y, T z) : v(x, y, z) {}
    template <typename O> explicit Translation(const Base<O>& o) : v(o) {}
    Matrix<T, Dim, 1>& vector() { return v; }
    const Matrix<T, Dim, 1>& vector() const { return v; }
    Matrix<T, Dim, 1>& translation() { return v; }
    const Matrix<T, Dim, 1>& translation() const { return v; }
    T& x() { return v(0); }  T& y() { return v(1); }  T& z() { return v(2); }
    T x() const { return v.coeff(0); }  T y() const { return v.coeff(1); }  T z() const { return v.coeff(2); }
};

template <typename T> class Quaternion {
    T w_, x_, y_, z_;
public:
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
    // Geometry/Quaternion.h quaternionbase_assign_impl<Other, 3, 3> (Shoemake)
    template <typename O> explicit Quaternion(const Base<O>& mat) {
        using std::sqrt;
        T t = mat.coeff(0, 0) + mat.coeff(1, 1) + mat.coeff(2, 2);
        if (t > T(0)) {
            t = sqrt(t + T(1.0));
            w_ = T(0.5) * t;
            t = T(0.5) / t;
            x_ = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
            y_ = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
            z_ = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            Index i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const Index j = (i + 1) % 3, k = (j + 1) % 3;
            t = sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + T(1.0));
            T q[3];
            q[i] = T(0.5) * t;
            t = T(0.5) / t;
            w_ = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            q[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            q[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
            x_ = q[0]; y_ = q[1]; z_ = q[2];
        }
    }
    T w() const { return w_; }  T x() const { return x_; }  T y() const { return y_; }  T z() const { return z_; }
    T& w() { return w_; }  T& x() { return x_; }  T& y() { return y_; }  T& z() { return z_; }
    Matrix<T, 3, 3> toRotationMatrix() const {   // QuaternionBase::toRotationMatrix
        Matrix<T, 3, 3> res;
        const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_;
        const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
        return res;
    }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;

template <typename T, int Dim, int Mode> class Transform {
    Matrix<T, Dim + 1, Dim + 1> m;
public:
    typedef Matrix<T, Dim + 1, Dim + 1> MatrixType;
    typedef Matrix<T, Dim, Dim> LinearMatrixType;
    Transform() { m.setIdentity(); }     // (Eigen leaves it uninitialised; every use in the reference assigns first)
    template <int OtherMode> Transform(const Transform<T, Dim, OtherMode>& o) : m(o.matrix()) {}
    template <typename O> explicit Transform(const Base<O>& o) : m(o) {}
    MatrixType& matrix() { return m; }
    const MatrixType& matrix() const { return m; }
    T& operator()(Index i, Index j) { return m(i, j); }
    T operator()(Index i, Index j) const { return m.coeff(i, j); }
    void setIdentity() { m.setIdentity(); }
    static Transform Identity() { return Transform(); }
    View<T, Dim, Dim> linear() { return m.template block<Dim, Dim>(0, 0); }
    View<T, Dim, Dim> linear() const { return m.template block<Dim, Dim>(0, 0); }
    View<T, Dim, 1> translation() { return m.template block<Dim, 1>(0, Dim); }
    View<T, Dim, 1> translation() const { return m.template block<Dim, 1>(0, Dim); }
    // Transform::rotation() for Mode != Isometry: computeRotationScaling through a JacobiSVD of linear()
    // (Geometry/Transform.h): x = det(U V^T); m = U with its last column scaled by x; rotation = m V^T.
    LinearMatrixType rotation() const {
        if (Mode == Isometry) return LinearMatrixType(linear());
        const LinearMatrixType L(linear());
        JacobiSVD<LinearMatrixType> svd(L, ComputeFullU | ComputeFullV);
        const LinearMatrixType UVt = svd.matrixU() * svd.matrixV().transpose();
        const T x = UVt.determinant();
        LinearMatrixType mm = svd.matrixU();
        for (Index i = 0; i < Dim; ++i) mm(i, Dim - 1) = mm(i, Dim - 1) * x;
        return mm * svd.matrixV().transpose();
    }
    template <typename U> Transform<U, Dim, Mode> cast() const { return Transform<U, Dim, Mode>(m.template cast<U>()); }
    Transform inverse() const { return Transform(m.inverse()); }
    Transform operator*(const Transform& o) const { return Transform(m * o.m); }
    template <typename O> Matrix<T, Dim, 1> operator*(const Base<O>& v) const {
        return LinearMatrixType(linear()) * v + Matrix<T, Dim, 1>(translation());
    }
};
// Quaternion * Translation -> Isometry transform [R | R t]
template <typename T> Transform<T, 3, Isometry> operator*(const Quaternion<T>& q, const Translation<T, 3>& t) {
    Transform<T, 3, Isometry> r;
    const Matrix<T, 3, 3> R = q.toRotationMatrix();
    r.matrix().template block<3, 3>(0, 0) = R;
    r.matrix().template block<3, 1>(0, 3) = R * t.vector();
    return r;
}
typedef Transform<float, 3, Affine> Affine3f;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 3, Isometry> Isometry3d;

template <typename T> class aligned_allocator : public std::allocator<T> {
public:
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};

}  // namespace Eigen

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif
