#pragma once
#include <deal.II/lac/vector.h>
namespace dealii { namespace LinearAlgebra { namespace distributed {
template <typename Number>
class Vector : public dealii::Vector<Number> {
   public:
    void reinit(std::size_t n) { dealii::Vector<Number>::reinit(n); }
    void reinit(const Vector& o) { dealii::Vector<Number>::reinit(o); }
};
}}}  // namespace dealii::LinearAlgebra::distributed
