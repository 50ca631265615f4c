// C entry points for the reference templates compiled by `make -C oracle ref` (see README.md in this directory).
// This file contains no physics: it only instantiates and forwards to the reference's own functions.
#include <functional>
#include <vector>

#include "src/dof_utils.h"
#include "src/five_moment/euler.h"
#include "src/timestepper.h"

using namespace dealii;
using namespace warpii;
using namespace warpii::five_moment;

namespace {
Tensor<1, 5, double> state(const double* q) { Tensor<1, 5, double> s; for (int i = 0; i < 5; i++) s[i] = q[i]; return s; }
template <int dim> Tensor<1, dim, double> vec(const double* n) { Tensor<1, dim, double> v; for (int i = 0; i < dim; i++) v[i] = n[i]; return v; }

template <int dim> void flux_t(const double* q, double g, double* F) {
    const auto f = euler_flux<dim>(state(q), g);
    for (int c = 0; c < 5; c++) for (int d = 0; d < dim; d++) F[c * dim + d] = f[c][d];
}
template <int dim> void ec_t(const double* a, const double* b, double g, double* F) {
    const auto f = euler_CH_EC_flux<dim>(state(a), state(b), g);
    for (int c = 0; c < 5; c++) for (int d = 0; d < dim; d++) F[c * dim + d] = f[c][d];
}
template <int dim> void es_t(const double* a, const double* b, const double* n, double g, double* out) {
    const auto f = euler_CH_entropy_dissipating_flux<dim>(state(a), state(b), vec<dim>(n), g);
    for (int c = 0; c < 5; c++) out[c] = f[c];
}
template <int dim> void lf_t(const double* a, const double* b, const double* n, double g, double* out) {
    const auto f = euler_numerical_flux<dim>(state(a), state(b), vec<dim>(n), g);
    for (int c = 0; c < 5; c++) out[c] = f[c];
}
}  // namespace

extern "C" {
double ref_ln_avg(double a, double b) { return ln_avg(a, b); }
double ref_pressure(const double* q, double g) { return euler_pressure<1>(state(q), g); }
void ref_euler_flux(int dim, const double* q, double g, double* F) { dim == 1 ? flux_t<1>(q, g, F) : dim == 2 ? flux_t<2>(q, g, F) : flux_t<3>(q, g, F); }
void ref_ec_flux(int dim, const double* a, const double* b, double g, double* F) { dim == 1 ? ec_t<1>(a, b, g, F) : dim == 2 ? ec_t<2>(a, b, g, F) : ec_t<3>(a, b, g, F); }
void ref_es_flux(int dim, const double* a, const double* b, const double* n, double g, double* out) { dim == 1 ? es_t<1>(a, b, n, g, out) : dim == 2 ? es_t<2>(a, b, n, g, out) : es_t<3>(a, b, n, g, out); }
void ref_lf_flux(int dim, const double* a, const double* b, const double* n, double g, double* out) { dim == 1 ? lf_t<1>(a, b, n, g, out) : dim == 2 ? lf_t<2>(a, b, n, g, out) : lf_t<3>(a, b, n, g, out); }
void ref_entropy_variables(const double* q, double g, double* w) { const auto v = euler_entropy_variables<3>(state(q), g); for (int i = 0; i < 5; i++) w[i] = v[i]; }
double ref_mathematical_entropy(const double* q, double g) { return euler_mathematical_entropy<3>(state(q), g); }

unsigned ref_pencil_stride(unsigned Np, unsigned d) { return pencil_stride(Np, d); }
unsigned ref_pencil_base(int dim, unsigned q, unsigned Np, unsigned d) { return dim == 1 ? pencil_base<1>(q, Np, d) : pencil_base<2>(q, Np, d); }
unsigned ref_quadrature_point_neighbor(int dim, unsigned q, unsigned k, unsigned Np, unsigned d) { return dim == 1 ? quadrature_point_neighbor<1>(q, k, Np, d) : quadrature_point_neighbor<2>(q, k, Np, d); }
unsigned ref_quad_point_1d_index(int dim, unsigned q, unsigned Np, unsigned d) { return dim == 1 ? quad_point_1d_index<1>(q, Np, d) : quad_point_1d_index<2>(q, Np, d); }
int ref_pencil_starts(int dim, unsigned Np, unsigned d, unsigned* out) {
    const std::vector<unsigned> v = dim == 1 ? pencil_starts<1>(Np, d) : pencil_starts<2>(Np, d);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

typedef int (*step_fn)(double, double, void*);
typedef double (*dt_fn)(void*);
typedef void (*cb_fn)(double, int, void*);
void ref_advance(step_fn step, double t_end, dt_fn rdt, int n, const double* intervals, const int* zeroth, const int* final_, cb_fn cb, void* user) {
    std::vector<TimestepCallback> cbs;
    for (int i = 0; i < n; i++) cbs.emplace_back(intervals[i], [=](double t) { cb(t, i, user); }, zeroth[i] != 0, final_[i] != 0);
    advance([&](double t, double dt) { return step(t, dt, user) != 0; }, t_end, [&]() { return rdt(user); }, cbs);
}
}
