// Minimal stand-in for <deal.II/base/table_indices.h>.
#pragma once
#include <cstddef>
namespace dealii {
template <int N>
class TableIndices {
   public:
    TableIndices() { for (int i = 0; i < N; i++) idx[i] = 0; }
    TableIndices(std::size_t a, std::size_t b) { static_assert(N == 2, "two-index form"); idx[0] = a; idx[1] = b; }
    std::size_t operator[](unsigned i) const { return idx[i]; }
    std::size_t& operator[](unsigned i) { return idx[i]; }
   private:
    std::size_t idx[N];
};
}  // namespace dealii
