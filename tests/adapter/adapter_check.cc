// tests/adapter/adapter_check.cc -- compile check and driver of include/gpu_es_dgsem_operator.h (TEST INFRASTRUCTURE).
//
// Compiles, in one translation unit:
//   * the adapter header a WarpII maintainer adds (include/gpu_es_dgsem_operator.h), verbatim;
//   * the reference's own, unmodified sources where they lie under /root/reference: src/rk.h (SSPRK2Integrator),
//     src/five_moment/solution_vec.{h,cc} (FiveMSolutionVec), src/five_moment/bc_helper.h (EulerBCMap), src/dof_utils.h;
//   * a minimal deal.II stand-in (tests/dealii_stub/ + oracle/ref_shim/ for Tensor / Assert) and 20-line stand-ins for
//     NodalDGDiscretization<dim> and Species<dim> (their real headers pull in the whole application),
// and INSTANTIATES the reference's SSPRK2Integrator<double, FiveMSolutionVec, GpuESDGSEMOperator<dim>> exactly as
// src/five_moment/dg_solver.h:71-72 would after the splice.  adapter_run() then drives a few time steps the way
// FiveMomentDGSolver::solve does (dg_solver.cc:23-38): recommend_dt -> evolve_one_time_step.  On a machine without a GPU
// the constructor throws the library's "no CUDA device" error (CPU tier); the GPU tier compares the result with the
// product's own solver.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <deal.II/base/function.h>
#include <deal.II/base/types.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/fe/fe_system.h>
#include <deal.II/matrix_free/matrix_free.h>

#include "bc_helper.h"   // the reference's EulerBCMap<dim>

namespace warpii {
// what the adapter asks of NodalDGDiscretization<dim> (src/dgsem/nodal_dg_discretization.h:22-88)
template <int dim>
class NodalDGDiscretization {
   public:
    NodalDGDiscretization(const int* nx, const double* left, const double* right, const int* periodic, unsigned n_components,
                          unsigned fe_degree)
        : fe_degree(fe_degree), n_components(n_components), fe(fe_degree, n_components), dof_handler(nx, left, right, periodic) {}
    unsigned int get_n_components() { return n_components; }
    unsigned int get_fe_degree() { return fe_degree; }
    dealii::DoFHandler<dim>& get_dof_handler() { return dof_handler; }
    dealii::FESystem<dim>& get_fe() { return fe; }
    dealii::MatrixFree<dim>& get_matrix_free() { return mf; }

   private:
    unsigned int fe_degree, n_components;
    dealii::FESystem<dim> fe;
    dealii::DoFHandler<dim> dof_handler;

   public:
    dealii::MatrixFree<dim> mf;
};
namespace five_moment {
// what the adapter asks of Species<dim> (src/five_moment/species.h:22-62)
template <int dim>
class Species {
   public:
    Species(std::string name, double charge, double mass, EulerBCMap<dim> bc_map) : name(name), charge(charge), mass(mass), bc_map(bc_map) {}
    std::string name;
    double charge, mass;
    EulerBCMap<dim> bc_map;
};
}  // namespace five_moment
}  // namespace warpii

#define WARPII_GPU_ADAPTER_STANDALONE
#include "gpu_es_dgsem_operator.h"   // the adapter, verbatim
#include "rk.h"                       // the reference's integrators, unmodified

using namespace dealii;
using namespace warpii;
using namespace warpii::five_moment;

namespace {
std::string g_error;

// a constant inflow state, as a Function<dim> the way Species::create_from_parameters hands them to EulerBCMap
template <int dim>
class ConstantState : public Function<dim> {
   public:
    explicit ConstantState(const double* q) : Function<dim>(5) { for (int c = 0; c < 5; c++) v[c] = q[c]; }
    double value(const Point<dim>&, const unsigned int component) const override { return v[component]; }
    double v[5];
};

template <int dim>
int run(int fe_degree, const int* nx, const double* left, const double* right, const int* periodic, int n_species, int fields,
        double gas_gamma, const int* bc_kind /*[n_species][2 dim]*/, const double* inflow /*[n_species][2 dim][5]*/, double* state,
        int n_steps, double* t_out, double* bif_out) {
    const unsigned nc = 5 * n_species + (fields ? 8 : 0);
    auto disc = std::make_shared<NodalDGDiscretization<dim>>(nx, left, right, periodic, nc, (unsigned)fe_degree);
    std::vector<std::shared_ptr<Species<dim>>> species;
    bool bounded = false;
    for (int d = 0; d < dim; d++) bounded = bounded || !periodic[d];
    for (int s = 0; s < n_species; s++) {
        EulerBCMap<dim> bc;
        for (int b = 0; bounded && b < 2 * dim; b++) {
            const int k = bc_kind[s * 2 * dim + b];
            if (k == WARPII_BC_WALL) bc.set_wall_boundary((types::boundary_id)b);
            else if (k == WARPII_BC_INFLOW)
                bc.set_inflow_boundary((types::boundary_id)b, std::make_unique<ConstantState<dim>>(inflow + (s * 2 * dim + b) * 5));
            else if (k == WARPII_BC_SUBSONIC_OUTFLOW)   // (the prescribed total energy is component 4 of the function)
                bc.set_subsonic_outflow_boundary((types::boundary_id)b, std::make_unique<ConstantState<dim>>(inflow + (s * 2 * dim + b) * 5));
            else bc.set_supersonic_outflow_boundary((types::boundary_id)b);
        }
        species.push_back(std::make_shared<Species<dim>>("species" + std::to_string(s), 1.0, 1.0, bc));
    }
    // dg_solver.h:71-72 after the splice
    GpuESDGSEMOperator<dim> fluid_flux_operator(disc, gas_gamma, species);
    SSPRK2Integrator<double, FiveMSolutionVec, GpuESDGSEMOperator<dim>> ssp_integrator;

    std::size_t n_dofs = nc;
    for (int d = 0; d < dim; d++) n_dofs *= (std::size_t)nx[d] * (fe_degree + 1);
    FiveMSolutionVec solution;
    solution.mesh_sol.reinit(n_dofs);
    solution.boundary_integrated_fluxes.reinit(bounded ? 2 * dim : 0, dim);
    for (std::size_t i = 0; i < n_dofs; i++) solution.mesh_sol[i] = state[i];
    ssp_integrator.reinit(solution, 3);                  // dg_solver.cc:11
    fluid_flux_operator.to_device(solution);              // after project_initial_condition (dg_solver.cc:15-20)
    double t = 0.0;
    for (int i = 0; i < n_steps; i++) {                   // the loop of advance() (timestepper.cc:34-42), no callbacks
        const double dt = fluid_flux_operator.recommend_dt(disc->get_matrix_free(), solution);
        ssp_integrator.evolve_one_time_step(fluid_flux_operator, solution, dt, t);
        t += dt;
    }
    fluid_flux_operator.to_host(solution);                // the writeout callback
    for (std::size_t i = 0; i < n_dofs; i++) state[i] = solution.mesh_sol[i];
    for (std::size_t i = 0; bif_out && i < solution.boundary_integrated_fluxes.data.size(); i++)
        bif_out[i] = solution.boundary_integrated_fluxes.data[i];
    *t_out = t;
    return 0;
}
}  // namespace

extern "C" {
const char* adapter_last_error(void) { return g_error.c_str(); }
// state: the host vector in the STUB DoFHandler's numbering: [cell (lexicographic)][node][component] (node-major, see
// tests/dealii_stub/deal.II/fe/fe_system.h); updated in place.  Returns 0, or 1 with adapter_last_error() set.
int adapter_run(int dim, int fe_degree, const int* nx, const double* left, const double* right, const int* periodic, int n_species,
                int fields, double gas_gamma, const int* bc_kind, const double* inflow, double* state, int n_steps, double* t_out,
                double* bif_out) {
    try {
        if (dim == 1)
            return run<1>(fe_degree, nx, left, right, periodic, n_species, fields, gas_gamma, bc_kind, inflow, state, n_steps, t_out, bif_out);
        if (dim == 2)
            return run<2>(fe_degree, nx, left, right, periodic, n_species, fields, gas_gamma, bc_kind, inflow, state, n_steps, t_out, bif_out);
        if (dim == 3)
            return run<3>(fe_degree, nx, left, right, periodic, n_species, fields, gas_gamma, bc_kind, inflow, state, n_steps, t_out, bif_out);
        g_error = "dim must be 1, 2 or 3";
    } catch (const std::exception& e) {
        g_error = e.what();
    }
    return 1;
}
}
