// Minimal stand-in for <deal.II/base/vectorization.h>: VectorizedArray<Number> with ONE lane, plus the std:: overloads
// deal.II provides for it (the reference's euler.h calls std::log / std::abs / std::max / std::sqrt on its Number type).
// One lane is what the reference's own per-lane fallbacks see (fluid_flux_es_dgsem_operator.h:282-293); lanes never
// interact in the volume / subcell drivers, so a 1-lane array evaluates exactly the arithmetic of any lane.
#pragma once
#include <cmath>
#include <cstddef>
namespace dealii {
template <typename Number, std::size_t width = 1>
class VectorizedArray {
   public:
    VectorizedArray() : v(Number()) {}
    VectorizedArray(const Number x) : v(x) {}
    Number& operator[](std::size_t) { return v; }
    const Number& operator[](std::size_t) const { return v; }
    static constexpr std::size_t size() { return 1; }
    VectorizedArray& operator+=(const VectorizedArray& o) { v += o.v; return *this; }
    VectorizedArray& operator-=(const VectorizedArray& o) { v -= o.v; return *this; }
    VectorizedArray& operator*=(const VectorizedArray& o) { v *= o.v; return *this; }
    VectorizedArray& operator/=(const VectorizedArray& o) { v /= o.v; return *this; }
    Number v;
};
template <typename N> inline VectorizedArray<N> operator+(const VectorizedArray<N>& a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a.v + b.v); }
template <typename N> inline VectorizedArray<N> operator-(const VectorizedArray<N>& a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a.v - b.v); }
template <typename N> inline VectorizedArray<N> operator*(const VectorizedArray<N>& a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a.v * b.v); }
template <typename N> inline VectorizedArray<N> operator/(const VectorizedArray<N>& a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a.v / b.v); }
template <typename N> inline VectorizedArray<N> operator-(const VectorizedArray<N>& a) { return VectorizedArray<N>(-a.v); }
template <typename N> inline VectorizedArray<N> operator+(const double a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a + b.v); }
template <typename N> inline VectorizedArray<N> operator+(const VectorizedArray<N>& a, const double b) { return VectorizedArray<N>(a.v + b); }
template <typename N> inline VectorizedArray<N> operator-(const double a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a - b.v); }
template <typename N> inline VectorizedArray<N> operator-(const VectorizedArray<N>& a, const double b) { return VectorizedArray<N>(a.v - b); }
template <typename N> inline VectorizedArray<N> operator*(const double a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a * b.v); }
template <typename N> inline VectorizedArray<N> operator*(const VectorizedArray<N>& a, const double b) { return VectorizedArray<N>(a.v * b); }
template <typename N> inline VectorizedArray<N> operator/(const double a, const VectorizedArray<N>& b) { return VectorizedArray<N>(a / b.v); }
template <typename N> inline VectorizedArray<N> operator/(const VectorizedArray<N>& a, const double b) { return VectorizedArray<N>(a.v / b); }
}  // namespace dealii
namespace std {
template <typename N> inline dealii::VectorizedArray<N> sqrt(const dealii::VectorizedArray<N>& a) { return dealii::VectorizedArray<N>(std::sqrt(a.v)); }
template <typename N> inline dealii::VectorizedArray<N> abs(const dealii::VectorizedArray<N>& a) { return dealii::VectorizedArray<N>(std::abs(a.v)); }
template <typename N> inline dealii::VectorizedArray<N> log(const dealii::VectorizedArray<N>& a) { return dealii::VectorizedArray<N>(std::log(a.v)); }
template <typename N> inline dealii::VectorizedArray<N> max(const dealii::VectorizedArray<N>& a, const dealii::VectorizedArray<N>& b) {
    return dealii::VectorizedArray<N>(std::max(a.v, b.v));
}
}  // namespace std
