#pragma once
#include <deal.II/base/tensor.h>   // (deal.II's lac/vector.h brings Tensor along; solution_vec.h relies on that)
#include <cstddef>
#include <vector>
namespace dealii {
template <typename Number>
class Vector {
   public:
    Vector() {}
    void reinit(std::size_t n) { a.assign(n, Number()); }
    void reinit(const Vector& o) { a.assign(o.a.size(), Number()); }
    std::size_t size() const { return a.size(); }
    Number& operator[](std::size_t i) { return a[i]; }
    const Number& operator[](std::size_t i) const { return a[i]; }
    void sadd(const Number s, const Number x, const Vector& V) { for (std::size_t i = 0; i < a.size(); i++) a[i] = s * a[i] + x * V.a[i]; }
    Number* begin() { return a.data(); }
    const Number* begin() const { return a.data(); }
    std::vector<Number> a;
};
}  // namespace dealii
