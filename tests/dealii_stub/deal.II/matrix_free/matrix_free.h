#pragma once
namespace dealii {
template <int dim, typename Number = double> class MatrixFree {};
}
