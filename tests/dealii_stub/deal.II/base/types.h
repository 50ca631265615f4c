// tests/dealii_stub: just enough of deal.II's interface to COMPILE include/gpu_es_dgsem_operator.h (the adapter a WarpII
// maintainer adds) together with the reference's own src/rk.h, src/five_moment/solution_vec.{h,cc}, bc_helper.h and
// dof_utils.h, in an image that has no deal.II.  Test infrastructure; nothing here is shipped.  Tensor, TableIndices and the
// Assert macros come from oracle/ref_shim (same purpose).
#pragma once
#include <cstdint>
namespace dealii { namespace types {
typedef unsigned int boundary_id;
typedef std::uint64_t global_dof_index;
}}  // namespace dealii::types
