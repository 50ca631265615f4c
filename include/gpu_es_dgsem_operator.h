// GpuESDGSEMOperator<dim>: the adapter a WarpII maintainer adds as src/five_moment/gpu_es_dgsem_operator.h.
//
// It has the interface of FluidFluxESDGSEMOperator<dim> (src/five_moment/fluid_flux_es_dgsem_operator.h:49-71: the
// constructor, perform_forward_euler_step, recommend_dt), so the whole splice is the operator TYPE in
// src/five_moment/dg_solver.h:71-72 (the integrator takes its operator as a duck-typed template parameter, src/rk.h:79-80),
// plus to_device() after project_initial_condition() and to_host() at the top of output_results().  Everything it does is a
// call into the C ABI of warpii_gpu.h; the state lives in HBM between calls.
//
// This header is COMPILED by tests/adapter/ (tests/test_adapter_cpu.py) against a minimal deal.II stand-in
// (tests/dealii_stub/) together with the reference's own, unmodified src/rk.h, src/five_moment/solution_vec.{h,cc},
// src/five_moment/bc_helper.h and src/dof_utils.h, and the reference's SSPRK2Integrator is instantiated over it and run
// (GPU tier: tests/test_gpu_adapter.py).  With WARPII_GPU_ADAPTER_STANDALONE defined it does not include
// nodal_dg_discretization.h / species.h itself (they pull in the whole application: ParameterHandler, Triangulation,
// MatrixFree); the including file then provides NodalDGDiscretization<dim> and Species<dim> with the members used here.
//
// Scope: HyperRectangle grids (every cell the same box).  Other triangulations additionally hand their metric terms to
// warpii_gpu_set_geometry once, right after warpii_gpu_create.
#pragma once
#include <deal.II/base/quadrature_lib.h>
#include <deal.II/base/utilities.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/fe/fe_system.h>
#include <deal.II/matrix_free/matrix_free.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <vector>

#include "warpii_gpu.h"   // this repository, include/

#include "../dof_utils.h"   // ZeroOutPolicy
#include "solution_vec.h"
#ifndef WARPII_GPU_ADAPTER_STANDALONE
#include "../dgsem/nodal_dg_discretization.h"
#include "species.h"
#endif

namespace warpii {
namespace five_moment {

template <int dim>
class GpuESDGSEMOperator {
   public:
    GpuESDGSEMOperator(std::shared_ptr<NodalDGDiscretization<dim>> discretization, double gas_gamma,
                       std::vector<std::shared_ptr<Species<dim>>> species, int device = 0)
        : discretization(discretization), species(species) {
        using namespace dealii;
        const unsigned fe_degree = discretization->get_fe_degree();
        Np = fe_degree + 1;
        NN = Utilities::pow(Np, dim);
        nc = discretization->get_n_components();
        const auto& dh = discretization->get_dof_handler();
        const auto& fe = discretization->get_fe();

        // ---- cells, numbered patch by patch (warpii_gpu_elems_per_block): any order is correct, this one is fast ----------
        std::vector<typename DoFHandler<dim>::active_cell_iterator> cells;
        for (const auto& cell : dh.active_cell_iterators()) cells.push_back(cell);
        AssertThrow(!cells.empty(), ExcMessage("GpuESDGSEMOperator: empty triangulation"));
        double h[3] = {1.0, 1.0, 1.0}, lo[3] = {0.0, 0.0, 0.0};
        for (int d = 0; d < dim; ++d) h[d] = cells[0]->extent_in_direction(d);
        for (int d = 0; d < dim; ++d) lo[d] = cells[0]->center()[d];
        for (const auto& cell : cells) {
            for (int d = 0; d < dim; ++d) {
                AssertThrow(std::fabs(cell->extent_in_direction(d) - h[d]) <= 1e-12 * h[d],
                            ExcMessage("GpuESDGSEMOperator: not a uniform box; pass the metric terms to warpii_gpu_set_geometry"));
                lo[d] = std::min(lo[d], cell->center()[d]);
            }
        }
        int shape[3] = {1, 1, 1};
        for (int g = warpii_gpu_elems_per_block(dim, (int)fe_degree), d = 0; g > 1; g /= 2, d = (d + 1) % dim) shape[d] *= 2;
        auto patch_key = [&](const typename DoFHandler<dim>::active_cell_iterator& cell) {
            long long idx[3] = {0, 0, 0}, patch = 0, within = 0;
            for (int d = 0; d < dim; ++d) idx[d] = std::llround((cell->center()[d] - lo[d]) / h[d]);
            for (int d = dim - 1; d >= 0; --d) {
                patch = patch * (1LL << 20) + idx[d] / shape[d];
                within = within * shape[d] + idx[d] % shape[d];
            }
            return std::make_pair(patch, within);
        };
        std::stable_sort(cells.begin(), cells.end(), [&](const auto& a, const auto& b) { return patch_key(a) < patch_key(b); });

        // ---- flat tables: face pairs, boundary faces, DoF translation ---------------------------------------------------------
        std::map<CellId, int> id_of;
        for (int e = 0; e < (int)cells.size(); ++e) id_of[cells[e]->id()] = e;
        const unsigned nf = 2 * dim;
        std::vector<int32_t> nbr(cells.size() * nf);
        dof_index.resize(cells.size() * (size_t)nc * NN);
        std::vector<types::global_dof_index> dofs((size_t)nc * NN);
        for (int e = 0; e < (int)cells.size(); ++e) {
            const auto& cell = cells[e];
            for (unsigned f = 0; f < nf; ++f) {
                if (cell->at_boundary(f) && !cell->has_periodic_neighbor(f)) {
                    nbr[e * nf + f] = -1 - (int32_t)bf_elem.size();
                    bf_elem.push_back(e);
                    bf_side.push_back((int32_t)f);
                    bf_id.push_back((int32_t)cell->face(f)->boundary_id());
                    bf_corner.push_back(cell->vertex(0));
                } else {
                    nbr[e * nf + f] = id_of[cell->neighbor_or_periodic_neighbor(f)->id()];
                }
            }
            // FE_DGQ^nc: ask deal.II where (component, node) sits instead of assuming a layout; device layout [elem][comp][node]
            cell->get_dof_indices(dofs);
            for (unsigned c = 0; c < nc; ++c)
                for (unsigned j = 0; j < NN; ++j) dof_index[((size_t)e * nc + c) * NN + j] = (int64_t)dofs[fe.component_to_system_index(c, j)];
        }
        n_boundaries = 0;
        for (int32_t b : bf_id) n_boundaries = std::max(n_boundaries, (unsigned)b + 1);
        std::vector<int32_t> bc_kind;   // [species][boundary id]: bc_helper.h
        for (const auto& sp : species) {
            for (unsigned b = 0; b < n_boundaries; ++b) {
                int32_t kind = WARPII_BC_OUTFLOW;   // supersonic outflow: ghost = inside state (species.cc:48-49)
                if (sp->bc_map.is_wall(b)) kind = WARPII_BC_WALL;
                else if (sp->bc_map.is_inflow(b)) kind = WARPII_BC_INFLOW;
                else if (sp->bc_map.is_subsonic_outflow(b)) kind = WARPII_BC_SUBSONIC_OUTFLOW;   // (:385-390)
                else AssertThrow(sp->bc_map.is_supersonic_outflow(b) || bf_id.empty(),
                                 ExcMessage("Unknown boundary id, did you set a boundary condition for this part of the domain boundary?"));
                bc_kind.push_back(kind);
            }
        }

        warpii_gpu_mesh m{};
        m.dim = dim;
        m.fe_degree = (int)fe_degree;
        m.n_species = (int)species.size();
        m.fields_enabled = nc > 5 * species.size() ? 1 : 0;
        m.gas_gamma = gas_gamma;
        m.n_elems = (int64_t)cells.size();
        m.n_ghost_faces = 0;
        m.n_boundary_faces = (int64_t)bf_elem.size();
        m.n_boundaries = (int)n_boundaries;
        for (int d = 0; d < 3; ++d) m.h[d] = h[d];
        m.face_neighbor = nbr.data();
        m.boundary_face_elem = bf_elem.data();
        m.boundary_face_side = bf_side.data();
        m.boundary_face_id = bf_id.data();
        m.bc_kind = bc_kind.empty() ? nullptr : bc_kind.data();
        m.n_vectors = 2;
        ok(warpii_gpu_create(&m, device, &ctx));

        // ---- quadrature points of the boundary faces (Gauss(fe_degree+2)^(dim-1), lower tangential dimension fastest): where
        // the reference evaluates its inflow functions (fluid_flux_es_dgsem_operator.h:355-358, 381-384)
        const QGauss<1> gauss(fe_degree + 2);
        nq = Utilities::pow(fe_degree + 2, dim - 1);
        bpts.resize(bf_elem.size() * nq);
        for (size_t f = 0; f < bf_elem.size(); ++f) {
            const int d = bf_side[f] / 2, side = bf_side[f] % 2;
            for (unsigned q = 0; q < nq; ++q) {
                Point<dim> x = bf_corner[f];
                unsigned r = q;
                for (int a = 0; a < dim; ++a) {
                    if (a == d) { x[a] += side ? h[a] : 0.0; continue; }
                    x[a] += gauss.point(r % (fe_degree + 2))[0] * h[a];
                    r /= (fe_degree + 2);
                }
                bpts[f * nq + q] = x;
            }
        }
        inflow_table.resize(bf_elem.size() * nq * 5);
        for (unsigned s = 0; s < species.size(); ++s) tabulate_inflow(s, 0.0);
    }
    ~GpuESDGSEMOperator() { warpii_gpu_destroy(ctx); }
    GpuESDGSEMOperator(const GpuESDGSEMOperator&) = delete;
    GpuESDGSEMOperator& operator=(const GpuESDGSEMOperator&) = delete;

    // FluidFluxESDGSEMOperator<dim>::perform_forward_euler_step (:62-68): dst = beta dst + alpha (u + dt M^-1 R(u)).  The
    // register vectors of the caller are not needed (the fused kernel never materialises M du/dt or du/dt).
    void perform_forward_euler_step(FiveMSolutionVec& dst, const FiveMSolutionVec& u, std::vector<FiveMSolutionVec>& /*sol_registers*/,
                                    const double dt, const double t, const double alpha = 1.0, const double beta = 0.0,
                                    const ZeroOutPolicy /*zero_out_register*/ = DO_NOT_ZERO_DST_VECTOR) {
        // set_time(t) on the inflow functions + their values at the boundary quadrature points (:139-144, :381-384)
        for (unsigned s = 0; s < species.size(); ++s)
            if (!species[s]->bc_map.inflow_boundaries().empty() || !species[s]->bc_map.subsonic_outflow_boundaries().empty()) tabulate_inflow(s, t);
        ok(warpii_gpu_forward_euler_step_ex(ctx, slot(dst), slot(u), dt, t, alpha, beta, beta != 0.0 ? WARPII_FUSE_CFL : 0));
    }

    // FluidFluxESDGSEMOperator<dim>::recommend_dt (:70-71, :442-448)
    double recommend_dt(const dealii::MatrixFree<dim, double>& /*mf*/, const FiveMSolutionVec& sol) {
        double dt = 0.0;
        ok(warpii_gpu_recommend_dt(ctx, slot(sol), &dt));
        return dt;
    }

    // The only host <-> device traffic of a run: at t = 0 after project_initial_condition() (dg_solver.cc:15-20) and in the
    // writeout callback, at the top of FiveMomentApp::output_results (five_moment.h:245)
    void to_device(const FiveMSolutionVec& v) {
        ok(warpii_gpu_upload_state(ctx, slot(v), v.mesh_sol.begin(), dof_index.data()));
        if (n_boundaries > 0 && v.boundary_integrated_fluxes.data.size() == 5 * n_boundaries)
            ok(warpii_gpu_set_boundary_fluxes(ctx, slot(v), v.boundary_integrated_fluxes.data.begin()));
    }
    void to_host(FiveMSolutionVec& v) {
        ok(warpii_gpu_download_state(ctx, slot(v), v.mesh_sol.begin(), dof_index.data()));
        if (n_boundaries > 0 && v.boundary_integrated_fluxes.data.size() == 5 * n_boundaries)
            ok(warpii_gpu_boundary_fluxes(ctx, slot(v), v.boundary_integrated_fluxes.data.begin()));
    }
    warpii_gpu_ctx* context() { return ctx; }

   private:
    // FiveMSolutionVec objects are mapped to device vector ids by address: the solver's `solution` and the integrator's f_1
    int slot(const FiveMSolutionVec& v) {
        auto it = ids.find(&v);
        if (it == ids.end()) {
            AssertThrow(ids.size() < 2, dealii::ExcMessage("GpuESDGSEMOperator: more than two solution vectors in flight"));
            it = ids.emplace(&v, (int)ids.size()).first;
        }
        return it->second;
    }
    void tabulate_inflow(unsigned s, double t) {
        const auto& bc = species[s]->bc_map;
        if ((bc.inflow_boundaries().empty() && bc.subsonic_outflow_boundaries().empty()) || bf_elem.empty()) return;
        for (const auto& entry : bc.inflow_boundaries()) entry.second->set_time(t);
        std::fill(inflow_table.begin(), inflow_table.end(), 0.0);
        for (size_t f = 0; f < bf_elem.size(); ++f) {
            const auto id = (dealii::types::boundary_id)bf_id[f];
            if (bc.is_inflow(id)) {
                const auto fn = bc.get_inflow(id);
                for (unsigned q = 0; q < nq; ++q)
                    for (unsigned c = 0; c < 5; ++c) inflow_table[(f * nq + q) * 5 + c] = fn->value(bpts[f * nq + q], c);
            } else if (bc.is_subsonic_outflow(id)) {   // only the energy component is read (:387-389)
                const auto fn = bc.get_subsonic_outflow_energy(id);
                for (unsigned q = 0; q < nq; ++q) inflow_table[(f * nq + q) * 5 + 4] = fn->value(bpts[f * nq + q], 4);
            }
        }
        ok(warpii_gpu_set_inflow_table(ctx, (int)s, inflow_table.data()));
    }
    static void ok(int status) { AssertThrow(status == 0, dealii::ExcMessage(warpii_gpu_last_error())); }

    std::shared_ptr<NodalDGDiscretization<dim>> discretization;
    std::vector<std::shared_ptr<Species<dim>>> species;
    unsigned Np = 2, NN = 2, nc = 5, n_boundaries = 0, nq = 1;
    std::vector<int64_t> dof_index;
    std::vector<int32_t> bf_elem, bf_side, bf_id;
    std::vector<dealii::Point<dim>> bf_corner, bpts;
    std::vector<double> inflow_table;
    std::map<const FiveMSolutionVec*, int> ids;
    warpii_gpu_ctx* ctx = nullptr;
};

}  // namespace five_moment
}  // namespace warpii
