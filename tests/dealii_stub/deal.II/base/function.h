#pragma once
#include <deal.II/base/exceptions.h>
#include <deal.II/base/point.h>
#include <map>
#include <memory>
#include <set>
#include <string>
namespace dealii {
template <int dim>
class Function {
   public:
    explicit Function(unsigned n_components = 1) : n_components(n_components), time(0.0) {}
    virtual ~Function() {}
    virtual double value(const Point<dim>& p, const unsigned int component = 0) const = 0;
    virtual void set_time(const double t) { time = t; }
    double get_time() const { return time; }
    const unsigned int n_components;
   protected:
    double time;
};
}  // namespace dealii
