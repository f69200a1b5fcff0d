// tiny_linalg.hpp -- the minimal dense-matrix vocabulary the host API mirror needs.
//
// The reference's C++ API passes Eigen objects by value (tinympc/TinyMPC/src/tinympc/types.hpp:15-17:
// tinyMatrix = Matrix<double,Dynamic,Dynamic>, tinyVector = Matrix<double,Dynamic,1>, VectorXi).
// Eigen is not a dependency of this repository.  Two ways to build the mirror:
//   * default: the small column-major classes below (same member names for the subset the API and
//     its callers use: rows(), cols(), size(), data(), operator()(i,j), operator()(i), setZero(),
//     Zero/Constant/Identity/Ones, col(j), transpose-free arithmetic helpers);
//   * -DTINYMPC_B200_USE_EIGEN with Eigen on the include path: the exact reference typedefs, so
//     reference user code (e.g. T/examples/quadrotor_hovering.cpp) compiles unchanged.
#pragma once

#ifdef TINYMPC_B200_USE_EIGEN
#include <Eigen/Core>
#include <Eigen/LU>
using namespace Eigen;
typedef double tinytype;
typedef Matrix<tinytype, Dynamic, Dynamic> tinyMatrix;
typedef Matrix<tinytype, Dynamic, 1> tinyVector;
#else

#include <cstddef>
#include <initializer_list>
#include <vector>

typedef double tinytype;

template <typename S>
class TinyDense {
public:
    TinyDense() : r_(0), c_(0) {}
    TinyDense(int rows, int cols) : r_(rows), c_(cols), d_((size_t)rows * cols, S(0)) {}
    explicit TinyDense(int rows) : r_(rows), c_(1), d_((size_t)rows, S(0)) {}
    TinyDense(int rows, int cols, const S* colmajor) : r_(rows), c_(cols), d_(colmajor, colmajor + (size_t)rows * cols) {}

    static TinyDense Zero(int rows, int cols = 1) { return TinyDense(rows, cols); }
    static TinyDense Constant(int rows, int cols, S v) { TinyDense m(rows, cols); for (auto& e : m.d_) e = v; return m; }
    static TinyDense Ones(int rows, int cols = 1) { return Constant(rows, cols, S(1)); }
    static TinyDense Identity(int rows, int cols) { TinyDense m(rows, cols); for (int i = 0; i < rows && i < cols; ++i) m(i, i) = S(1); return m; }
    // diagonal matrix from a vector (what the reference callers write as v.asDiagonal())
    static TinyDense Diagonal(const TinyDense& v) { TinyDense m((int)v.size(), (int)v.size()); for (int i = 0; i < (int)v.size(); ++i) m(i, i) = v.d_[i]; return m; }
    // row-major initialiser (what the reference callers write as Map<Matrix<.., RowMajor>>(data))
    static TinyDense FromRowMajor(int rows, int cols, const S* p) { TinyDense m(rows, cols); for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) m(i, j) = p[(size_t)i * cols + j]; return m; }

    int rows() const { return r_; }
    int cols() const { return c_; }
    size_t size() const { return d_.size(); }
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }
    S& operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
    const S& operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
    S& operator()(int i) { return d_[i]; }
    const S& operator()(int i) const { return d_[i]; }
    void setZero() { for (auto& e : d_) e = S(0); }
    void setConstant(S v) { for (auto& e : d_) e = v; }
    void resize(int rows, int cols) { r_ = rows; c_ = cols; d_.assign((size_t)rows * cols, S(0)); }

    TinyDense col(int j) const { return TinyDense(r_, 1, d_.data() + (size_t)j * r_); }
    void set_col(int j, const TinyDense& v) { for (int i = 0; i < r_; ++i) (*this)(i, j) = v.d_[i]; }
    TinyDense diagonal() const { int n = r_ < c_ ? r_ : c_; TinyDense v(n, 1); for (int i = 0; i < n; ++i) v(i) = (*this)(i, i); return v; }
    TinyDense transpose() const { TinyDense t(c_, r_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j); return t; }
    TinyDense replicate(int rf, int cf) const {
        TinyDense m(r_ * rf, c_ * cf);
        for (int a = 0; a < rf; ++a) for (int b = 0; b < cf; ++b) for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) m(a * r_ + i, b * c_ + j) = (*this)(i, j);
        return m;
    }

private:
    int r_, c_;
    std::vector<S> d_;
};

typedef TinyDense<tinytype> tinyMatrix;
typedef TinyDense<tinytype> tinyVector;   // a column: cols() == 1
typedef TinyDense<int> VectorXi;

inline tinyMatrix operator*(const tinyMatrix& a, const tinyMatrix& b) {
    tinyMatrix c(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            tinytype s = 0;
            for (int k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j);
            c(i, j) = s;
        }
    return c;
}
inline tinyMatrix operator+(const tinyMatrix& a, const tinyMatrix& b) { tinyMatrix c = a; for (size_t i = 0; i < c.size(); ++i) c.data()[i] += b.data()[i]; return c; }
inline tinyMatrix operator-(const tinyMatrix& a, const tinyMatrix& b) { tinyMatrix c = a; for (size_t i = 0; i < c.size(); ++i) c.data()[i] -= b.data()[i]; return c; }
inline tinyMatrix operator*(tinytype s, const tinyMatrix& a) { tinyMatrix c = a; for (size_t i = 0; i < c.size(); ++i) c.data()[i] *= s; return c; }

#endif  // TINYMPC_B200_USE_EIGEN
