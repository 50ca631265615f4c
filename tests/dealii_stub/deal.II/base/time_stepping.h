#pragma once
#include <vector>
namespace dealii { namespace TimeStepping {
enum runge_kutta_method { LOW_STORAGE_RK_STAGE3_ORDER3, LOW_STORAGE_RK_STAGE5_ORDER4, LOW_STORAGE_RK_STAGE7_ORDER4, LOW_STORAGE_RK_STAGE9_ORDER5 };
template <typename VectorType>
class LowStorageRungeKutta {
   public:
    explicit LowStorageRungeKutta(runge_kutta_method) {}
    void get_coefficients(std::vector<double>&, std::vector<double>&, std::vector<double>&) const {}
};
}}  // namespace dealii::TimeStepping
#define AssertDimension(a, b) do { } while (0)
