#pragma once
#include <utility>
namespace dealii {
// FE_DGQ(p)^nc.  deal.II numbers a cell's DoFs component-blocked; the stub deliberately uses the OTHER legal layout
// (node-major), so that an adapter which assumed a layout instead of asking component_to_system_index would fail its test.
template <int dim>
class FESystem {
   public:
    FESystem(unsigned degree, unsigned n_components) : degree(degree), nc(n_components) {}
    unsigned component_to_system_index(unsigned component, unsigned index) const { return index * nc + component; }
    unsigned n_components() const { return nc; }
    unsigned degree;
   private:
    unsigned nc;
};
}  // namespace dealii
