// oracle/shim/eigen/shim_eigen.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section 4).
//
// A minimal, EAGER stand-in for the subset of Eigen 3.3 that the reference's FIRST-PARTY sources use
// (cpp/rkhs_registration/src/cvo.cpp, src/adaptive_cvo.cpp, src/LieGroup.cpp and the headers they include), so that
// those files can be compiled UNMODIFIED, where they lie under /root/reference, in an image that has no Eigen
// (oracle/Makefile, target `refsrc`).  It is not Eigen and makes no attempt at its performance: every operator
// evaluates at once into a concrete Matrix (no expression templates), blocks are (pointer, offset) views.
// What it keeps of Eigen's SEMANTICS, because the reference's arithmetic depends on it:
//   * column-major dense storage, fixed-size objects hold their coefficients inline (std::vector<Vector3f> is an
//     array of 3-float records: thirdparty/nanoflann reads the query point through &v(0));
//   * a scalar operand of another arithmetic type is converted to the matrix's scalar type BEFORE the operation
//     (Eigen 3.3 promote_scalar_arg: `2.0 * MatrixXf` multiplies in float);
//   * matrix products, dot products and squared norms accumulate in the scalar type, left to right;
//   * SparseMatrix::setFromTriplets leaves every row sorted by column;
//   * MatrixXf::eigenvalues() is a real Hessenberg-QR in the matrix's scalar type (real roots come back with an
//     imaginary part that is exactly zero, which cvo::compute_step_size tests for, src/cvo.cpp:300);
//   * MatrixBase::log()/exp() of small matrices (inverse scaling and squaring; evaluated in double, narrowed).
// Anything the reference does not use is absent.
#ifndef CVO_ORACLE_SHIM_EIGEN_HPP
#define CVO_ORACLE_SHIM_EIGEN_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3

namespace Eigen {

constexpr int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };

template <class T, int R, int C, int O = 0, int MR = R, int MC = C> class Matrix;
template <class M, int BR, int BC> class Block;

namespace shim {
template <class X> struct is_mat : std::false_type {};
template <class T, int R, int C, int O, int MR, int MC> struct is_mat<Matrix<T, R, C, O, MR, MC>> : std::true_type {};
template <class M, int BR, int BC> struct is_mat<Block<M, BR, BC>> : std::true_type {};
template <class X> using dec = std::remove_cv_t<std::remove_reference_t<X>>;
template <class X> constexpr bool is_mat_v = is_mat<dec<X>>::value;
template <class X> constexpr bool is_scalar_v = std::is_arithmetic<dec<X>>::value;
template <class T> struct is_complex : std::false_type {};
template <class T> struct is_complex<std::complex<T>> : std::true_type {};
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }

template <class T, int N> struct Store {  // fixed size: coefficients inline
    T d[N];
    T* p() { return d; }
    const T* p() const { return d; }
    void resize(std::size_t) {}
};
template <class T> struct Store<T, Dynamic> {
    std::vector<T> d;
    T* p() { return d.data(); }
    const T* p() const { return d.data(); }
    void resize(std::size_t n) { d.resize(n); }
};
template <int R, int C> struct Dims {  // both fixed: no runtime members (Vector3f is exactly 3 floats)
    constexpr int rows() const { return R; }
    constexpr int cols() const { return C; }
    void set(int, int) {}
};
template <int C> struct Dims<Dynamic, C> {
    int r = 0;
    int rows() const { return r; }
    constexpr int cols() const { return C; }
    void set(int rr, int) { r = rr; }
};
template <int R> struct Dims<R, Dynamic> {
    int c = 0;
    constexpr int rows() const { return R; }
    int cols() const { return c; }
    void set(int, int cc) { c = cc; }
};
template <> struct Dims<Dynamic, Dynamic> {
    int r = 0, c = 0;
    int rows() const { return r; }
    int cols() const { return c; }
    void set(int rr, int cc) { r = rr; c = cc; }
};
}  // namespace shim

template <class M> class CommaInit;

// Operations shared by Matrix and Block (CRTP): everything is evaluated eagerly into a Matrix.
template <class D, class T, int R, int C> class Base {
  public:
    typedef T Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
    D& derived() { return static_cast<D&>(*this); }
    const D& derived() const { return static_cast<const D&>(*this); }
    int rows() const { return derived().rows_(); }
    int cols() const { return derived().cols_(); }
    int size() const { return rows() * cols(); }
    T& operator()(int i, int j) { return derived().at(i, j); }
    const T& operator()(int i, int j) const { return derived().at(i, j); }
    T& operator()(int i) { return vec_at(i); }
    const T& operator()(int i) const { return const_cast<Base*>(this)->vec_at(i); }
    T& operator[](int i) { return vec_at(i); }
    const T& operator[](int i) const { return const_cast<Base*>(this)->vec_at(i); }
    T& vec_at(int i) { return cols() == 1 ? derived().at(i, 0) : derived().at(0, i); }

    Matrix<T, R, C> eval() const {
        Matrix<T, R, C> m(rows(), cols(), 0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) m(i, j) = (*this)(i, j);
        return m;
    }
    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> m(cols(), rows(), 0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) m(j, i) = (*this)(i, j);
        return m;
    }
    template <class U> Matrix<U, R, C> cast() const {
        Matrix<U, R, C> m(rows(), cols(), 0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) m(i, j) = (U)(*this)(i, j);
        return m;
    }
    T value() const {
        assert(rows() == 1 && cols() == 1);
        return (*this)(0, 0);
    }
    T squaredNorm() const {  // accumulated in the scalar type, storage order
        T s = T(0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += (*this)(i, j) * (*this)(i, j);
        return s;
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    T trace() const {
        T s = T(0);
        for (int i = 0; i < rows(); ++i) s += (*this)(i, i);
        return s;
    }
    T sum() const {
        T s = T(0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += (*this)(i, j);
        return s;
    }
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> T dot(const O& o) const {
        assert(size() == o.size());
        T s = T(0);
        for (int i = 0; i < size(); ++i) s += (*this)(i) * o(i);
        return s;
    }
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> Matrix<T, R, C> cross(const O& o) const {
        assert(size() == 3 && o.size() == 3);
        Matrix<T, R, C> m(rows(), cols(), 0);
        const T a0 = (*this)(0), a1 = (*this)(1), a2 = (*this)(2), b0 = o(0), b1 = o(1), b2 = o(2);
        m(0) = a1 * b2 - a2 * b1;
        m(1) = a2 * b0 - a0 * b2;
        m(2) = a0 * b1 - a1 * b0;
        return m;
    }
    // views
    template <int BR, int BC> Block<D, BR, BC> block(int i, int j) { return Block<D, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC> Block<const D, BR, BC> block(int i, int j) const { return Block<const D, BR, BC>(derived(), i, j, BR, BC); }
    Block<D, Dynamic, Dynamic> block(int i, int j, int r, int c) { return Block<D, Dynamic, Dynamic>(derived(), i, j, r, c); }
    Block<const D, Dynamic, Dynamic> block(int i, int j, int r, int c) const { return Block<const D, Dynamic, Dynamic>(derived(), i, j, r, c); }
    Block<D, 1, C> row(int i) { return Block<D, 1, C>(derived(), i, 0, 1, cols()); }
    Block<const D, 1, C> row(int i) const { return Block<const D, 1, C>(derived(), i, 0, 1, cols()); }
    Block<D, R, 1> col(int j) { return Block<D, R, 1>(derived(), 0, j, rows(), 1); }
    Block<const D, R, 1> col(int j) const { return Block<const D, R, 1>(derived(), 0, j, rows(), 1); }
    Block<D, Dynamic, Dynamic> bottomLeftCorner(int r, int c) { return block(rows() - r, 0, r, c); }
    Block<D, Dynamic, Dynamic> topLeftCorner(int r, int c) { return block(0, 0, r, c); }
    // vector segments (column or row vectors)
    Block<D, Dynamic, Dynamic> segment(int i, int n) { return cols() == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
    Block<const D, Dynamic, Dynamic> segment(int i, int n) const { return cols() == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
    template <int N> Block<D, (C == 1 ? N : 1), (C == 1 ? 1 : N)> segment(int i) {
        return C == 1 ? Block<D, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(derived(), i, 0, N, 1)
                      : Block<D, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(derived(), 0, i, 1, N);
    }
    template <int N> Block<const D, (C == 1 ? N : 1), (C == 1 ? 1 : N)> segment(int i) const {
        return C == 1 ? Block<const D, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(derived(), i, 0, N, 1)
                      : Block<const D, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(derived(), 0, i, 1, N);
    }
    Block<D, Dynamic, Dynamic> head(int n) { return segment(0, n); }
    Block<const D, Dynamic, Dynamic> head(int n) const { return segment(0, n); }
    Block<D, Dynamic, Dynamic> tail(int n) { return segment(size() - n, n); }
    Block<const D, Dynamic, Dynamic> tail(int n) const { return segment(size() - n, n); }

    Matrix<typename std::conditional<shim::is_complex<T>::value, decltype(std::real(T())), T>::type, R, C> real() const {
        Matrix<decltype(std::real(T())), R, C> m(rows(), cols(), 0);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) m(i, j) = std::real((*this)(i, j));
        return m;
    }
    // in-place arithmetic
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> D& operator+=(const O& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) += (T)o(i, j);
        return derived();
    }
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> D& operator-=(const O& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) -= (T)o(i, j);
        return derived();
    }
    template <class S, std::enable_if_t<shim::is_scalar_v<S>, int> = 0> D& operator*=(S s) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) *= (T)s;
        return derived();
    }
    void setZero() {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = T(0);
    }
    void setIdentity() {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = i == j ? T(1) : T(0);
    }
    // comma initialiser
    template <class S, std::enable_if_t<shim::is_scalar_v<S> || shim::is_complex<S>::value, int> = 0> CommaInit<D> operator<<(const S& s);
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> CommaInit<D> operator<<(const O& o);

    // dense decompositions the reference uses
    Matrix<std::complex<T>, Dynamic, 1> eigenvalues() const;  // real general matrix: Hessenberg + shifted QR
    Matrix<T, R, C> log() const;
    Matrix<T, R, C> exp() const;
    Matrix<T, R, C> inverse() const;
};

template <class T, int R, int C, int O, int MR, int MC>
class Matrix : public Base<Matrix<T, R, C, O, MR, MC>, T, R, C> {
    static constexpr int N = (R != Dynamic && C != Dynamic) ? R * C : Dynamic;
    shim::Store<T, N> s_;
    shim::Dims<R, C> dim_;

  public:
    typedef Base<Matrix, T, R, C> B;
    Matrix() {
        if (N != Dynamic) dim_.set(R, C);
        else dim_.set(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C);
        s_.resize((std::size_t)rows_() * cols_());
    }
    // (rows, cols, tag): sized, uninitialised -- the shim's own constructor
    Matrix(int r, int c, int /*tag*/) { resize(r, c); }
    // Eigen: Matrix(n) sizes a dynamic vector; Matrix(r, c) sizes a dynamic matrix; for a FIXED 2-vector (x, y) are
    // coefficients (thirdparty/NumType.h: Vec2(a, b))
    template <class S, std::enable_if_t<std::is_integral<S>::value, int> = 0> explicit Matrix(S n) {
        if (R == Dynamic && C == 1) resize((int)n, 1);
        else if (C == Dynamic && R == 1) resize(1, (int)n);
        else {
            resize(R == Dynamic ? (int)n : R, C == Dynamic ? (int)n : C);
        }
    }
    template <class S1, class S2, std::enable_if_t<shim::is_scalar_v<S1> && shim::is_scalar_v<S2>, int> = 0> Matrix(S1 a, S2 b) {
        if (N == 2) {
            resize(R, C);
            s_.p()[0] = (T)a;
            s_.p()[1] = (T)b;
        } else {
            resize((int)a, (int)b);
        }
    }
    template <class S, std::enable_if_t<shim::is_scalar_v<S>, int> = 0> Matrix(S a, S b, S c) {
        resize(R == Dynamic ? 3 : R, C == Dynamic ? 1 : C);
        s_.p()[0] = (T)a; s_.p()[1] = (T)b; s_.p()[2] = (T)c;
    }
    Matrix(const Matrix&) = default;
    Matrix& operator=(const Matrix&) = default;
    template <class Oth, std::enable_if_t<shim::is_mat_v<Oth> && !std::is_same<shim::dec<Oth>, Matrix>::value, int> = 0>
    Matrix(const Oth& o) {
        assign(o);
    }
    template <class Oth, std::enable_if_t<shim::is_mat_v<Oth> && !std::is_same<shim::dec<Oth>, Matrix>::value, int> = 0>
    Matrix& operator=(const Oth& o) {
        // the source may alias this object (M = M.transpose() is not used by the reference; blocks of *this are copied first)
        Matrix<T, Dynamic, Dynamic> tmp(o.rows(), o.cols(), 0);
        for (int j = 0; j < o.cols(); ++j)
            for (int i = 0; i < o.rows(); ++i) tmp(i, j) = (T)o(i, j);
        assign(tmp);
        return *this;
    }
    template <class Oth> void assign(const Oth& o) {
        int r = o.rows(), c = o.cols();
        if (R != Dynamic && C != Dynamic && (r != R || c != C)) {  // vector <- transposed-shape vector of the same length
            assert(r * c == R * C && (r == 1 || c == 1));
            resize(R, C);
            for (int i = 0; i < r * c; ++i) s_.p()[i] = (T)o(i);
            return;
        }
        if ((R != Dynamic && r != R) || (C != Dynamic && c != C)) {  // e.g. VectorXf <- a 1 x n row: Eigen transposes implicitly for vectors
            assert(r == 1 || c == 1);
            std::swap(r, c);
            resize(r, c);
            for (int i = 0; i < r * c; ++i) s_.p()[i] = (T)o(i);
            return;
        }
        resize(r, c);
        for (int j = 0; j < c; ++j)
            for (int i = 0; i < r; ++i) at(i, j) = (T)o(i, j);
    }
    void resize(int r, int c) {
        assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
        dim_.set(r, c);
        s_.resize((std::size_t)r * c);
    }
    void resize(int n) {
        if (C == 1) resize(n, 1);
        else resize(1, n);
    }
    int rows_() const { return dim_.rows(); }
    int cols_() const { return dim_.cols(); }
    T& at(int i, int j) {
        assert(i >= 0 && i < rows_() && j >= 0 && j < cols_());
        return s_.p()[(std::size_t)j * rows_() + i];
    }
    const T& at(int i, int j) const {
        assert(i >= 0 && i < rows_() && j >= 0 && j < cols_());
        return s_.p()[(std::size_t)j * rows_() + i];
    }
    T* data() { return s_.p(); }
    const T* data() const { return s_.p(); }

    static Matrix Zero() { return Zero(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C); }
    static Matrix Zero(int n) {
        Matrix m((long)n);
        m.setZero();
        return m;
    }
    static Matrix Zero(int r, int c) {
        Matrix m(r, c, 0);
        m.setZero();
        return m;
    }
    static Matrix Identity() { return Identity(R, C); }
    static Matrix Identity(int r, int c) {
        Matrix m(r, c, 0);
        m.setIdentity();
        return m;
    }
    static Matrix Ones(int r, int c) {
        Matrix m(r, c, 0);
        for (int j = 0; j < c; ++j)
            for (int i = 0; i < r; ++i) m(i, j) = T(1);
        return m;
    }
};

// A rectangular view into a Matrix (or into another view).  BR / BC: compile-time size or Dynamic.
template <class M, int BR, int BC>
class Block : public Base<Block<M, BR, BC>, typename M::Scalar, BR, BC> {
    M* m_;
    int i0_, j0_, r_, c_;

  public:
    typedef typename M::Scalar T;
    Block(M& m, int i0, int j0, int r, int c) : m_(&m), i0_(i0), j0_(j0), r_(r), c_(c) {
        assert(i0 >= 0 && j0 >= 0 && i0 + r <= m.rows() && j0 + c <= m.cols());
    }
    Block(const Block&) = default;
    int rows_() const { return r_; }
    int cols_() const { return c_; }
    T& at(int i, int j) { return const_cast<T&>(static_cast<const M*>(m_)->operator()(i0_ + i, j0_ + j)); }
    const T& at(int i, int j) const { return static_cast<const M*>(m_)->operator()(i0_ + i, j0_ + j); }
    template <class Oth, std::enable_if_t<shim::is_mat_v<Oth>, int> = 0> Block& operator=(const Oth& o) {
        static_assert(!std::is_const<M>::value, "assignment to a block of a const matrix");
        Matrix<T, Dynamic, Dynamic> tmp(o.rows(), o.cols(), 0);  // the source may alias the target
        for (int j = 0; j < o.cols(); ++j)
            for (int i = 0; i < o.rows(); ++i) tmp(i, j) = (T)o(i, j);
        if (tmp.rows() == r_ && tmp.cols() == c_) {
            for (int j = 0; j < c_; ++j)
                for (int i = 0; i < r_; ++i) at(i, j) = tmp(i, j);
        } else {  // vector shapes: row <- column of the same length
            assert(tmp.size() == r_ * c_ && (r_ == 1 || c_ == 1));
            for (int i = 0; i < r_ * c_; ++i) this->vec_at(i) = tmp(i);
        }
        return *this;
    }
    Block& operator=(const Block& o) { return this->template operator=<Block>(o); }
};

template <class M> class CommaInit {
    M m_;  // a Matrix& target is held by reference through a Block-like pointer; blocks are cheap copies
    int k_ = 0;

  public:
    explicit CommaInit(M& m) : m_(m) {}
    template <class S> void put(const S& v) {
        const int r = m_.rows(), c = m_.cols();
        assert(k_ < r * c);
        if (c == 1 || r == 1) m_.vec_at(k_) = (typename M::Scalar)v;
        else m_.at(k_ / c, k_ % c) = (typename M::Scalar)v;  // matrices fill row by row
        ++k_;
    }
    template <class S, std::enable_if_t<shim::is_scalar_v<S> || shim::is_complex<S>::value, int> = 0> CommaInit& operator,(const S& s) {
        put(s);
        return *this;
    }
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> CommaInit& operator,(const O& o) {
        for (int i = 0; i < o.size(); ++i) put(o(i));  // the reference only stacks vectors
        return *this;
    }
};
// Matrix targets are initialised through a whole-matrix view so that CommaInit can hold its target by value
template <class D, class T, int R, int C>
template <class S, std::enable_if_t<shim::is_scalar_v<S> || shim::is_complex<S>::value, int>>
CommaInit<D> Base<D, T, R, C>::operator<<(const S& s) {
    static_assert(sizeof(D) == 0 || true, "");
    CommaInit<D> ci(derived());
    ci.put(s);
    return ci;
}
template <class D, class T, int R, int C>
template <class O, std::enable_if_t<shim::is_mat_v<O>, int>>
CommaInit<D> Base<D, T, R, C>::operator<<(const O& o) {
    CommaInit<D> ci(derived());
    for (int i = 0; i < o.size(); ++i) ci.put(o(i));
    return ci;
}
// CommaInit<Matrix> must write through to the ORIGINAL matrix: specialise the holder for Matrix as a pointer
template <class T, int R, int C, int O, int MR, int MC> class CommaInit<Matrix<T, R, C, O, MR, MC>> {
    typedef Matrix<T, R, C, O, MR, MC> M;
    M* m_;
    int k_ = 0;

  public:
    explicit CommaInit(M& m) : m_(&m) {}
    template <class S> void put(const S& v) {
        const int r = m_->rows(), c = m_->cols();
        assert(k_ < r * c);
        if (c == 1 || r == 1) m_->vec_at(k_) = (T)v;
        else m_->at(k_ / c, k_ % c) = (T)v;
        ++k_;
    }
    template <class S, std::enable_if_t<shim::is_scalar_v<S> || shim::is_complex<S>::value, int> = 0> CommaInit& operator,(const S& s) {
        put(s);
        return *this;
    }
    template <class Oth, std::enable_if_t<shim::is_mat_v<Oth>, int> = 0> CommaInit& operator,(const Oth& o) {
        for (int i = 0; i < o.size(); ++i) put(o(i));
        return *this;
    }
};

// ---------------------------------------------------------------------------------------------------------
// free operators (eager)
// ---------------------------------------------------------------------------------------------------------
#define SHIM_RES(A, B) Matrix<typename shim::dec<A>::Scalar, shim::pick(shim::dec<A>::RowsAtCompileTime, shim::dec<B>::RowsAtCompileTime), \
                              shim::pick(shim::dec<A>::ColsAtCompileTime, shim::dec<B>::ColsAtCompileTime)>

template <class A, class B, std::enable_if_t<shim::is_mat_v<A> && shim::is_mat_v<B>, int> = 0> SHIM_RES(A, B) operator+(const A& a, const B& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    SHIM_RES(A, B) m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) + b(i, j);
    return m;
}
template <class A, class B, std::enable_if_t<shim::is_mat_v<A> && shim::is_mat_v<B>, int> = 0> SHIM_RES(A, B) operator-(const A& a, const B& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    SHIM_RES(A, B) m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) - b(i, j);
    return m;
}
template <class A, std::enable_if_t<shim::is_mat_v<A>, int> = 0>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator-(const A& a) {
    Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = -a(i, j);
    return m;
}
// scalar operands are converted to the matrix's scalar type first (Eigen 3.3 promote_scalar_arg)
template <class S, class A, std::enable_if_t<shim::is_scalar_v<S> && shim::is_mat_v<A>, int> = 0>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator*(S s, const A& a) {
    typedef typename A::Scalar T;
    const T t = (T)s;
    Matrix<T, A::RowsAtCompileTime, A::ColsAtCompileTime> m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = t * a(i, j);
    return m;
}
template <class S, class A, std::enable_if_t<shim::is_scalar_v<S> && shim::is_mat_v<A>, int> = 0>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator*(const A& a, S s) {
    typedef typename A::Scalar T;
    const T t = (T)s;
    Matrix<T, A::RowsAtCompileTime, A::ColsAtCompileTime> m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) * t;
    return m;
}
template <class S, class A, std::enable_if_t<shim::is_scalar_v<S> && shim::is_mat_v<A>, int> = 0>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator/(const A& a, S s) {
    typedef typename A::Scalar T;
    const T t = (T)s;
    Matrix<T, A::RowsAtCompileTime, A::ColsAtCompileTime> m(a.rows(), a.cols(), 0);
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) / t;
    return m;
}
template <class A, class B, std::enable_if_t<shim::is_mat_v<A> && shim::is_mat_v<B>, int> = 0>
Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> operator*(const A& a, const B& b) {
    typedef typename A::Scalar T;
    assert(a.cols() == b.rows());
    Matrix<T, A::RowsAtCompileTime, B::ColsAtCompileTime> m(a.rows(), b.cols(), 0);
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            T s = T(0);
            for (int k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j);  // scalar type, k ascending
            m(i, j) = s;
        }
    return m;
}
template <class A, std::enable_if_t<shim::is_mat_v<A>, int> = 0> std::ostream& operator<<(std::ostream& os, const A& a) {
    for (int i = 0; i < a.rows(); ++i) {
        for (int j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a(i, j);
        if (i + 1 < a.rows()) os << "\n";
    }
    return os;
}

typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<std::complex<float>, Dynamic, 1> VectorXcf;
typedef Matrix<std::complex<double>, Dynamic, 1> VectorXcd;

// ---------------------------------------------------------------------------------------------------------
// eigenvalues of a real general matrix: reduction to Hessenberg form + shifted QR (EISPACK elmhes / hqr as in
// Press et al.), in the matrix's own scalar type.  Real eigenvalues come out with imaginary part exactly 0.
// ---------------------------------------------------------------------------------------------------------
template <class D, class T, int R, int C> Matrix<std::complex<T>, Dynamic, 1> Base<D, T, R, C>::eigenvalues() const {
    const int n = rows();
    assert(n == cols());
    std::vector<std::vector<T>> a(n, std::vector<T>(n));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = (*this)(i, j);
    // Hessenberg reduction by stabilised elementary similarity transformations
    for (int m = 1; m < n - 1; ++m) {
        T x = T(0);
        int i = m;
        for (int j = m; j < n; ++j)
            if (std::fabs(a[j][m - 1]) > std::fabs(x)) {
                x = a[j][m - 1];
                i = j;
            }
        if (i != m) {
            for (int j = m - 1; j < n; ++j) std::swap(a[i][j], a[m][j]);
            for (int j = 0; j < n; ++j) std::swap(a[j][i], a[j][m]);
        }
        if (x != T(0)) {
            for (i = m + 1; i < n; ++i) {
                T y = a[i][m - 1];
                if (y != T(0)) {
                    y /= x;
                    a[i][m - 1] = y;
                    for (int j = m; j < n; ++j) a[i][j] -= y * a[m][j];
                    for (int j = 0; j < n; ++j) a[j][m] += y * a[j][i];
                }
            }
        }
    }
    for (int i = 2; i < n; ++i)
        for (int j = 0; j < i - 1; ++j) a[i][j] = T(0);
    Matrix<std::complex<T>, Dynamic, 1> out((long)n);
    std::vector<T> wr(n), wi(n);
    int nn = n - 1, its, l;
    T p = 0, q = 0, r = 0, s = 0, t = T(0), u, v, w, x, y, z, anorm = T(0);
    for (int i = 0; i < n; ++i)
        for (int j = std::max(i - 1, 0); j < n; ++j) anorm += std::fabs(a[i][j]);
    bool failed = !(anorm == anorm) || std::isinf(anorm);
    while (nn >= 0 && !failed) {
        its = 0;
        do {
            for (l = nn; l >= 1; --l) {
                s = std::fabs(a[l - 1][l - 1]) + std::fabs(a[l][l]);
                if (s == T(0)) s = anorm;
                if (std::fabs(a[l][l - 1]) + s == s) {
                    a[l][l - 1] = T(0);
                    break;
                }
            }
            x = a[nn][nn];
            if (l == nn) {  // one real root
                wr[nn] = x + t;
                wi[nn--] = T(0);
            } else {
                y = a[nn - 1][nn - 1];
                w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {  // two roots
                    p = T(0.5) * (y - x);
                    q = p * p + w;
                    z = std::sqrt(std::fabs(q));
                    x += t;
                    if (q >= T(0)) {  // a real pair
                        z = p + (p >= T(0) ? std::fabs(z) : -std::fabs(z));
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != T(0)) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = T(0);
                    } else {  // a complex pair
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = -(wi[nn] = z);
                    }
                    nn -= 2;
                } else {  // no roots yet: one QR step
                    if (its == 60) {
                        failed = true;
                        break;
                    }
                    if (its == 10 || its == 20) {  // exceptional shift
                        t += x;
                        for (int i = 0; i <= nn; ++i) a[i][i] -= x;
                        s = std::fabs(a[nn][nn - 1]) + std::fabs(a[nn - 1][nn - 2]);
                        y = x = T(0.75) * s;
                        w = T(-0.4375) * s * s;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; --m) {
                        z = a[m][m];
                        r = x - z;
                        s = y - z;
                        p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - s;
                        r = a[m + 2][m + 1];
                        s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        u = std::fabs(a[m][m - 1]) * (std::fabs(q) + std::fabs(r));
                        v = std::fabs(p) * (std::fabs(a[m - 1][m - 1]) + std::fabs(z) + std::fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) {
                        a[i][i - 2] = T(0);
                        if (i != m + 2) a[i][i - 3] = T(0);
                    }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = a[k][k - 1];
                            q = a[k + 1][k - 1];
                            r = T(0);
                            if (k != nn - 1) r = a[k + 2][k - 1];
                            if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != T(0)) {
                                p /= x; q /= x; r /= x;
                            }
                        }
                        const T sq = std::sqrt(p * p + q * q + r * r);
                        if ((s = (p >= T(0) ? sq : -sq)) != T(0)) {
                            if (k == m) {
                                if (l != m) a[k][k - 1] = -a[k][k - 1];
                            } else {
                                a[k][k - 1] = -s * x;
                            }
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) {
                                    p += r * a[k + 2][j];
                                    a[k + 2][j] -= p * z;
                                }
                                a[k + 1][j] -= p * y;
                                a[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) {
                                    p += z * a[i][k + 2];
                                    a[i][k + 2] -= p * r;
                                }
                                a[i][k + 1] -= p * q;
                                a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1 && !failed);
    }
    for (int i = 0; i < n; ++i) {
        if (failed) out(i) = std::complex<T>(std::nan(""), std::nan(""));  // Eigen reports NoConvergence; NaN roots are never selected
        else out(i) = std::complex<T>(wr[i], wi[i]);
    }
    return out;
}

namespace shim {
typedef std::vector<double> dmat;  // n x n row-major
inline dmat dmul(const dmat& a, const dmat& b, int n) {
    dmat c((std::size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            const double aik = a[i * n + k];
            for (int j = 0; j < n; ++j) c[i * n + j] += aik * b[k * n + j];
        }
    return c;
}
inline dmat dinv(dmat a, int n) {  // Gauss-Jordan with partial pivoting
    dmat inv((std::size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(a[r * n + c]) > std::fabs(a[piv * n + c])) piv = r;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                std::swap(a[piv * n + j], a[c * n + j]);
                std::swap(inv[piv * n + j], inv[c * n + j]);
            }
        const double d = a[c * n + c];
        for (int j = 0; j < n; ++j) {
            a[c * n + j] /= d;
            inv[c * n + j] /= d;
        }
        for (int r = 0; r < n; ++r)
            if (r != c) {
                const double f = a[r * n + c];
                if (f != 0.0)
                    for (int j = 0; j < n; ++j) {
                        a[r * n + j] -= f * a[c * n + j];
                        inv[r * n + j] -= f * inv[c * n + j];
                    }
            }
    }
    return inv;
}
inline dmat deye(int n) {
    dmat e((std::size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) e[i * n + i] = 1.0;
    return e;
}
inline double dnorm1(const dmat& a, int n) {
    double best = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += std::fabs(a[i * n + j]);
        best = std::max(best, s);
    }
    return best;
}
inline dmat dsqrt(const dmat& a, int n) {  // Denman-Beavers iteration (matrices near the identity)
    dmat y = a, z = deye(n);
    for (int it = 0; it < 60; ++it) {
        const dmat yi = dinv(y, n), zi = dinv(z, n);
        dmat yn((std::size_t)n * n), zn((std::size_t)n * n);
        double delta = 0.0;
        for (int i = 0; i < n * n; ++i) {
            yn[i] = 0.5 * (y[i] + zi[i]);
            zn[i] = 0.5 * (z[i] + yi[i]);
            delta = std::max(delta, std::fabs(yn[i] - y[i]));
        }
        y.swap(yn);
        z.swap(zn);
        if (delta < 1e-16 * std::max(1.0, dnorm1(y, n))) break;
    }
    return y;
}
inline dmat dlog(dmat a, int n) {  // inverse scaling and squaring + the series of log(I + X)
    int k = 0;
    dmat x((std::size_t)n * n);
    while (true) {
        for (int i = 0; i < n * n; ++i) x[i] = a[i] - (i / n == i % n ? 1.0 : 0.0);
        if (dnorm1(x, n) < 0.05 || k > 40) break;
        a = dsqrt(a, n);
        ++k;
    }
    dmat term = x, sum = x;
    for (int p = 2; p < 40; ++p) {
        term = dmul(term, x, n);
        const double c = ((p & 1) ? 1.0 : -1.0) / p;
        double big = 0.0;
        for (int i = 0; i < n * n; ++i) {
            sum[i] += c * term[i];
            big = std::max(big, std::fabs(term[i]));
        }
        if (big < 1e-20) break;
    }
    const double scale = std::ldexp(1.0, k);
    for (double& v : sum) v *= scale;
    return sum;
}
inline dmat dexp(dmat a, int n) {  // scaling and squaring + Taylor
    int k = 0;
    double nrm = dnorm1(a, n);
    while (nrm > 0.25 && k < 60) {
        nrm *= 0.5;
        ++k;
    }
    const double scale = std::ldexp(1.0, -k);
    for (double& v : a) v *= scale;
    dmat term = deye(n), sum = deye(n);
    for (int p = 1; p < 30; ++p) {
        term = dmul(term, a, n);
        for (double& v : term) v /= p;
        for (int i = 0; i < n * n; ++i) sum[i] += term[i];
    }
    for (int i = 0; i < k; ++i) sum = dmul(sum, sum, n);
    return sum;
}
template <class Mt> dmat to_d(const Mt& m) {
    const int n = m.rows();
    dmat a((std::size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i * n + j] = (double)m(i, j);
    return a;
}
}  // namespace shim

template <class D, class T, int R, int C> Matrix<T, R, C> Base<D, T, R, C>::log() const {
    const int n = rows();
    assert(n == cols());
    const shim::dmat l = shim::dlog(shim::to_d(*this), n);
    Matrix<T, R, C> m(n, n, 0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m(i, j) = (T)l[i * n + j];
    return m;
}
template <class D, class T, int R, int C> Matrix<T, R, C> Base<D, T, R, C>::exp() const {
    const int n = rows();
    assert(n == cols());
    const shim::dmat l = shim::dexp(shim::to_d(*this), n);
    Matrix<T, R, C> m(n, n, 0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m(i, j) = (T)l[i * n + j];
    return m;
}
template <class D, class T, int R, int C> Matrix<T, R, C> Base<D, T, R, C>::inverse() const {
    const int n = rows();
    assert(n == cols());
    const shim::dmat l = shim::dinv(shim::to_d(*this), n);
    Matrix<T, R, C> m(n, n, 0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m(i, j) = (T)l[i * n + j];
    return m;
}

// ---------------------------------------------------------------------------------------------------------
// Transform<float, 3, Affine>
// ---------------------------------------------------------------------------------------------------------
enum { Affine = 2 };
template <class T, int Dim, int Mode> class Transform {
    Matrix<T, Dim + 1, Dim + 1> m_;

  public:
    typedef Matrix<T, Dim + 1, Dim + 1> MatrixType;
    Transform() { m_.setIdentity(); }  // (Eigen leaves it uninitialised; the reference always assigns Identity())
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> Transform(const O& o) : m_(o) {}
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> Transform& operator=(const O& o) {
        m_ = o;
        return *this;
    }
    static Transform Identity() { return Transform(); }
    MatrixType& matrix() { return m_; }
    const MatrixType& matrix() const { return m_; }
    Block<MatrixType, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
    Block<const MatrixType, Dim, Dim> linear() const { return m_.template block<Dim, Dim>(0, 0); }
    Block<MatrixType, Dim, Dim> rotation() { return linear(); }
    Block<MatrixType, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
    Block<const MatrixType, Dim, 1> translation() const { return m_.template block<Dim, 1>(0, Dim); }
    T& operator()(int i, int j) { return m_(i, j); }
    const T& operator()(int i, int j) const { return m_(i, j); }
    Transform operator*(const Transform& o) const { return Transform(m_ * o.m_); }
    // Transform * (Dim+1 x Dim+1 matrix): the product of the two homogeneous matrices
    template <class O, std::enable_if_t<shim::is_mat_v<O>, int> = 0> Transform operator*(const O& o) const { return Transform(m_ * o); }
};
typedef Transform<float, 3, Affine> Affine3f;
typedef Transform<double, 3, Affine> Affine3d;

// ---------------------------------------------------------------------------------------------------------
// SparseMatrix<float, RowMajor> (CSR) + Triplet
// ---------------------------------------------------------------------------------------------------------
template <class T, class I = int> class Triplet {
    I r_, c_;
    T v_;

  public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(const I& r, const I& c, const T& v = T(0)) : r_(r), c_(c), v_(v) {}
    const I& row() const { return r_; }
    const I& col() const { return c_; }
    const T& value() const { return v_; }
};

template <class T, int Opt = 0> class SparseMatrix {
    static_assert(Opt == RowMajor, "the reference only uses row-major sparse matrices");
    int rows_ = 0, cols_ = 0;
    std::vector<int> ptr_ = std::vector<int>(1, 0), idx_;
    std::vector<T> val_;

  public:
    struct InnerVectorRef {
        int n;
        int nonZeros() const { return n; }
    };
    class InnerIterator {
        const SparseMatrix* m_;
        int k_, end_, row_;

      public:
        InnerIterator(const SparseMatrix& m, int outer) : m_(&m), k_(m.ptr_[outer]), end_(m.ptr_[outer + 1]), row_(outer) {}
        InnerIterator& operator++() {
            ++k_;
            return *this;
        }
        operator bool() const { return k_ < end_; }
        int col() const { return m_->idx_[k_]; }
        int row() const { return row_; }
        int index() const { return m_->idx_[k_]; }
        const T& value() const { return m_->val_[k_]; }
    };
    SparseMatrix() {}
    SparseMatrix(int r, int c) { resize(r, c); }
    void resize(int r, int c) {
        rows_ = r;
        cols_ = c;
        setZero();
    }
    void setZero() {
        ptr_.assign((std::size_t)rows_ + 1, 0);
        idx_.clear();
        val_.clear();
    }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int outerSize() const { return rows_; }
    long nonZeros() const { return (long)idx_.size(); }
    InnerVectorRef innerVector(int i) const { return InnerVectorRef{ptr_[i + 1] - ptr_[i]}; }
    void makeCompressed() {}
    // Eigen builds the matrix through a transposed copy: every row ends up sorted by column, duplicates summed
    template <class It> void setFromTriplets(It begin, It end) {
        struct E {
            int r, c;
            T v;
        };
        std::vector<E> e;
        for (It it = begin; it != end; ++it) {
            assert(it->row() >= 0 && it->row() < rows_ && it->col() >= 0 && it->col() < cols_);
            e.push_back(E{(int)it->row(), (int)it->col(), (T)it->value()});
        }
        std::stable_sort(e.begin(), e.end(), [](const E& a, const E& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
        ptr_.assign((std::size_t)rows_ + 1, 0);
        idx_.clear();
        val_.clear();
        for (std::size_t k = 0; k < e.size(); ++k) {
            if (k > 0 && e[k].r == e[k - 1].r && e[k].c == e[k - 1].c) {
                val_.back() += e[k].v;
                continue;
            }
            idx_.push_back(e[k].c);
            val_.push_back(e[k].v);
            ptr_[(std::size_t)e[k].r + 1] += 1;
        }
        for (int r = 0; r < rows_; ++r) ptr_[(std::size_t)r + 1] += ptr_[r];
    }
};

template <class T> class aligned_allocator : public std::allocator<T> {};
template <class T> class Quaternion {
  public:
    T x_, y_, z_, w_;
    T x() const { return x_; }
    T y() const { return y_; }
    T z() const { return z_; }
    T w() const { return w_; }
};
typedef Quaternion<float> Quaternionf;

}  // namespace Eigen
#endif
