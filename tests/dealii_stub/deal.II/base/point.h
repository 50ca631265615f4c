#pragma once
#include <deal.II/base/tensor.h>
namespace dealii {
template <int dim, typename Number = double>
class Point : public Tensor<1, dim, Number> {
   public:
    Point() {}
    Point(const Tensor<1, dim, Number>& t) : Tensor<1, dim, Number>(t) {}
};
}  // namespace dealii
