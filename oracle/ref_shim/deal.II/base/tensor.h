// Minimal stand-in for <deal.II/base/tensor.h>: rank-1 tensors (possibly nested) with the operations the reference's
// src/five_moment/euler.h and src/tensor_utils.h use, with deal.II's semantics (value-initialised to zero, scalar
// product accumulated in index order).  Rank-2 is declared only far enough for tensor_utils.h to parse.
#pragma once
#include <cmath>
#include <initializer_list>
#include <type_traits>
#include "table_indices.h"
#include "vectorization.h"
namespace dealii {
// what may multiply / divide a tensor: built-in arithmetic types and (one-lane) VectorizedArray
template <typename S> struct is_tensor_scalar : std::is_arithmetic<S> {};
template <typename N, std::size_t w> struct is_tensor_scalar<VectorizedArray<N, w>> : std::true_type {};
template <int rank, int dim, typename Number = double>
class Tensor;

template <int dim, typename Number>
class Tensor<1, dim, Number> {
   public:
    Tensor() { for (int i = 0; i < dim; i++) v[i] = Number(); }
    Tensor(std::initializer_list<Number> l) { int i = 0; for (const Number& x : l) v[i++] = x; for (; i < dim; i++) v[i] = Number(); }
    Number& operator[](unsigned i) { return v[i]; }
    const Number& operator[](unsigned i) const { return v[i]; }
    Tensor& operator+=(const Tensor& o) { for (int i = 0; i < dim; i++) v[i] += o.v[i]; return *this; }
    Tensor& operator-=(const Tensor& o) { for (int i = 0; i < dim; i++) v[i] -= o.v[i]; return *this; }
    template <typename S, typename = typename std::enable_if<is_tensor_scalar<S>::value>::type>
    Tensor& operator*=(const S s) { for (int i = 0; i < dim; i++) v[i] *= s; return *this; }
    template <typename S, typename = typename std::enable_if<is_tensor_scalar<S>::value>::type>
    Tensor& operator/=(const S s) { for (int i = 0; i < dim; i++) v[i] /= s; return *this; }
    Number norm_square() const { Number s = v[0] * v[0]; for (int i = 1; i < dim; i++) s += v[i] * v[i]; return s; }
    Number norm() const { using std::sqrt; return sqrt(norm_square()); }
   private:
    Number v[dim > 0 ? dim : 1];
};

template <int dim, typename Number>
class Tensor<2, dim, Number> {
   public:
    Number& operator[](const TableIndices<2>& i) { return v[i[0]][i[1]]; }
    const Number& operator[](const TableIndices<2>& i) const { return v[i[0]][i[1]]; }
   private:
    Number v[dim > 0 ? dim : 1][dim > 0 ? dim : 1] = {};
};

template <int dim, typename Number>
inline Tensor<1, dim, Number> operator+(Tensor<1, dim, Number> a, const Tensor<1, dim, Number>& b) { a += b; return a; }
template <int dim, typename Number>
inline Tensor<1, dim, Number> operator-(Tensor<1, dim, Number> a, const Tensor<1, dim, Number>& b) { a -= b; return a; }
template <int dim, typename Number>
inline Tensor<1, dim, Number> operator-(Tensor<1, dim, Number> a) { for (int i = 0; i < dim; i++) a[i] = -a[i]; return a; }
template <int dim, typename Number, typename S, typename = typename std::enable_if<is_tensor_scalar<S>::value>::type>
inline Tensor<1, dim, Number> operator*(const S s, Tensor<1, dim, Number> a) { for (int i = 0; i < dim; i++) a[i] = s * a[i]; return a; }
template <int dim, typename Number, typename S, typename = typename std::enable_if<is_tensor_scalar<S>::value>::type>
inline Tensor<1, dim, Number> operator*(Tensor<1, dim, Number> a, const S s) { for (int i = 0; i < dim; i++) a[i] = a[i] * s; return a; }
template <int dim, typename Number, typename S, typename = typename std::enable_if<is_tensor_scalar<S>::value>::type>
inline Tensor<1, dim, Number> operator/(Tensor<1, dim, Number> a, const S s) { for (int i = 0; i < dim; i++) a[i] = a[i] / s; return a; }
// scalar product of two rank-1 tensors of arithmetic type
template <int dim, typename Number, typename = typename std::enable_if<is_tensor_scalar<Number>::value>::type>
inline Number operator*(const Tensor<1, dim, Number>& a, const Tensor<1, dim, Number>& b) {
    Number s = a[0] * b[0];
    for (int i = 1; i < dim; i++) s += a[i] * b[i];
    return s;
}
}  // namespace dealii
