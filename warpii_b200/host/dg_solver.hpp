// FiveMomentGpuSolver: host-side mirror of five_moment::FiveMomentDGSolver<dim> (dg_solver.h:34-74, dg_solver.cc:7-38)
// for box grids: owns the mesh tables, the GPU context, the solution vector, the SSPRK2 integrator and the
// operator; reinit / project_initial_condition / solve keep the reference's meaning.
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <vector>

#include "box_mesh.hpp"
#include "general_mesh.hpp"
#include "gpu_operator.hpp"
#include "mapped_mesh.hpp"
#include "reference_element.hpp"

namespace warpii_b200 {

struct SpeciesBC {
    std::vector<int32_t> kind;                     // per boundary id: WARPII_BC_*
    std::vector<std::array<double, 5>> inflow;     // conserved inflow state per boundary id
    // optional: a function of (x, t) per boundary id instead of the constant state (species.cc:51-57); an empty
    // function means "use the constant".  time_dependent[b] = false tabulates it once.
    std::vector<InflowFunction> inflow_function;
    std::vector<bool> time_dependent;
};

class FiveMomentGpuSolver {
   public:
    using Integrator = SSPRK2Integrator<double, GpuSolutionVec, GpuFluidFluxESDGSEMOperator>;

    FiveMomentGpuSolver(const BoxDescription& box, int fe_degree, int n_species, bool fields_enabled, double gas_gamma,
                        double t_end, int n_boundaries, std::vector<SpeciesBC> bcs, int rank, int n_ranks, int device)
        : t_end_(t_end), fe_degree_(fe_degree), n_species_(n_species), fields_enabled_(fields_enabled), gas_gamma_(gas_gamma),
          n_boundaries_(n_boundaries), bcs_(std::move(bcs)), device_(device), rank_(rank), n_ranks_(n_ranks),
          tables_(box, rank, n_ranks, warpii_gpu_elems_per_block(box.dim, fe_degree)), element_(fe_degree) {
        nc_ = 5 * n_species + (fields_enabled ? 8 : 0);
        nn_ = 1;
        for (int d = 0; d < box.dim; d++) nn_ *= fe_degree + 1;
        // a box with non-periodic sides has boundary ids up to 2*dim-1 whether or not the input declared them;
        // the reference throws "Unknown boundary id" from the face loop in that case, the ABI does so at create.
    }

    // The same solver on a mesh that is not a Cartesian box (GridType = Extension, or a mapped box): tables and
    // Gauss-Lobatto support points from GeneralMesh, metric terms from mapped_mesh.hpp, warpii_gpu_set_geometry.
    // One rank (the reference does not shard either: serial Triangulation, src/grid.h:40).
    FiveMomentGpuSolver(GeneralMesh mesh, int n_species, bool fields_enabled, double gas_gamma, double t_end, int n_boundaries,
                        std::vector<SpeciesBC> bcs, int device)
        : t_end_(t_end), fe_degree_(mesh.fe_degree), n_species_(n_species), fields_enabled_(fields_enabled), gas_gamma_(gas_gamma),
          n_boundaries_(n_boundaries), bcs_(std::move(bcs)), device_(device), rank_(mesh.rank), n_ranks_(mesh.n_ranks),
          tables_(unit_box(mesh.dim), 0, 1, 1), element_(mesh.fe_degree), general_(std::make_unique<GeneralMesh>(std::move(mesh))) {
        nc_ = 5 * n_species + (fields_enabled ? 8 : 0);
        nn_ = 1;
        for (int d = 0; d < general_->dim; d++) nn_ *= fe_degree_ + 1;
    }

    // dg_solver.cc:7-12
    void reinit() {
        std::vector<int32_t> bc_kind((size_t)n_species_ * (n_boundaries_ > 0 ? n_boundaries_ : 0), WARPII_BC_WALL);
        for (int s = 0; s < n_species_ && s < (int)bcs_.size(); s++)
            for (int b = 0; b < n_boundaries_ && b < (int)bcs_[s].kind.size(); b++) bc_kind[(size_t)s * n_boundaries_ + b] = bcs_[s].kind[b];
        warpii_gpu_mesh mesh;
        tables_.fill(mesh, fe_degree_, n_species_, fields_enabled_, gas_gamma_, n_boundaries_, bc_kind, 6);   // ids (solution, f_1, low-storage RK registers); HBM is taken on first use
        if (general_) {
            mesh.dim = general_->dim;
            mesh.n_elems = general_->n_elems;
            mesh.n_ghost_faces = general_->n_ghost_faces;
            mesh.n_boundary_faces = (int64_t)general_->bf_elem.size();
            mesh.face_neighbor = general_->face_neighbor.data();
            mesh.boundary_face_elem = general_->bf_elem.data();
            mesh.boundary_face_side = general_->bf_side.data();
            mesh.boundary_face_id = general_->bf_id.data();
        }
        ctx_ = std::make_shared<GpuContext>(mesh, device_);
        if (general_) {
            metrics_ = build_mapped_metrics(general_->dim, fe_degree_, general_->n_elems, general_->xyz.data(),
                                            general_->face_neighbor.data(),
                                            general_->neighbor_face.empty() ? nullptr : general_->neighbor_face.data(),
                                            (int64_t)general_->bf_elem.size(), general_->bf_elem.data(), general_->bf_side.data());
            warpii_gpu_geometry g{};
            g.inverse_jacobian = metrics_.inverse_jacobian.data();
            g.face_normal = metrics_.face_normal.data();
            g.face_jacobian = metrics_.face_jacobian.data();
            g.neighbor_face = general_->neighbor_face.empty() ? nullptr : general_->neighbor_face.data();
            g.boundary_normal = metrics_.boundary_normal.data();
            g.boundary_jacobian = metrics_.boundary_jacobian.data();
            check(warpii_gpu_set_geometry(ctx_->get(), &g));
        }
        solution_ = std::make_unique<GpuSolutionVec>(ctx_);
        op_ = std::make_unique<GpuFluidFluxESDGSEMOperator>(ctx_);
        integrator_ = std::make_unique<Integrator>();
        integrator_->reinit(*solution_, 3);
        if (general_) op_->set_boundary_points(metrics_.boundary_points, general_->bf_id, general_->dim, n_species_);
        else op_->set_boundary_points(boundary_quadrature_points(), tables_.boundary_face_id(), tables_.box().dim, n_species_);
        for (int s = 0; s < n_species_ && s < (int)bcs_.size(); s++)
            for (int b = 0; b < n_boundaries_ && b < (int)bcs_[s].kind.size(); b++) {
                if (bcs_[s].kind[b] != WARPII_BC_INFLOW && bcs_[s].kind[b] != WARPII_BC_SUBSONIC_OUTFLOW) continue;
                if (b < (int)bcs_[s].inflow_function.size() && bcs_[s].inflow_function[b])
                    op_->set_inflow_function(s, b, bcs_[s].inflow_function[b],
                                             b < (int)bcs_[s].time_dependent.size() ? (bool)bcs_[s].time_dependent[b] : true);
                else if (b < (int)bcs_[s].inflow.size())
                    op_->set_inflow(s, b, bcs_[s].inflow[b].data());
            }
    }

    // xyz[face][point][dim] of the boundary quadrature points of this rank: Gauss(fe_degree+2) abscissae in the
    // tangential dimensions (lower dimension fastest), the face's coordinate in the normal one.  Faces in the order of
    // the mesh tables.  These are the phi.quadrature_point(q) of fluid_flux_es_dgsem_operator.h:381-384.
    std::vector<double> boundary_quadrature_points() const {
        const BoxDescription& box = tables_.box();
        const int dim = box.dim, Ng = element_.Ng;
        int nq = 1;
        for (int d = 1; d < dim; d++) nq *= Ng;
        const auto& bf_elem = tables_.boundary_face_elem();
        const auto& bf_side = tables_.boundary_face_side();
        std::vector<double> xyz(bf_elem.size() * nq * dim);
        for (size_t f = 0; f < bf_elem.size(); f++) {
            int idx[3];
            tables_.elem_multi_index(tables_.local_to_global()[bf_elem[f]], idx);
            const int d = bf_side[f] / 2, side = bf_side[f] % 2;
            for (int q = 0; q < nq; q++) {
                int t = q;
                for (int a = 0; a < dim; a++) {
                    double x;
                    if (a == d) x = box.left[a] + (idx[a] + side) * tables_.h(a);
                    else { x = box.left[a] + (idx[a] + element_.xg[t % Ng]) * tables_.h(a); t /= Ng; }
                    xyz[(f * nq + q) * dim + a] = x;
                }
            }
        }
        return xyz;
    }

    // Node coordinates of the owned elements in device order, xyz[elem][node][dim].
    std::vector<double> node_coords() const {
        if (general_) return general_->xyz;
        const BoxDescription& box = tables_.box();
        const int Np = fe_degree_ + 1;
        std::vector<double> xyz((size_t)tables_.n_local() * nn_ * box.dim);
        for (int64_t l = 0; l < tables_.n_local(); l++) {
            int idx[3];
            tables_.elem_multi_index(tables_.local_to_global()[l], idx);
            for (int j = 0; j < nn_; j++) {
                int t = j;
                for (int d = 0; d < box.dim; d++) {
                    xyz[((size_t)l * nn_ + j) * box.dim + d] = box.left[d] + (idx[d] + element_.x[t % Np]) * tables_.h(d);
                    t /= Np;
                }
            }
        }
        return xyz;
    }

    // Nodal interpolation of the initial condition (project_fluid_quantities, dg_solution_helper.cc:24-48), with the
    // Primitive -> conserved conversion of SpeciesFunc::value (species_func.cc:9-30).  f(x, out5) fills
    // [rho, ux, uy, uz, p] (primitive = true) or the conserved state.
    void project_initial_condition(int species, const std::function<void(const double* x, double* out5)>& f, bool primitive = true) {
        if (host_.empty()) host_.assign((size_t)ctx_->n_dofs(), 0.0);
        const std::vector<double> xyz = node_coords();
        const int dim = n_dims();
        for (int64_t l = 0; l < n_local_elems(); l++)
            for (int j = 0; j < nn_; j++) {
                double v[5], q[5];
                f(&xyz[((size_t)l * nn_ + j) * dim], v);
                if (primitive) {
                    const double rho = v[0];
                    q[0] = rho;
                    double ke = 0.0;
                    for (int d = 0; d < 3; d++) { q[d + 1] = rho * v[d + 1]; ke += 0.5 * rho * v[d + 1] * v[d + 1]; }
                    q[4] = ke + v[4] / (gas_gamma_ - 1);
                } else {
                    for (int k = 0; k < 5; k++) q[k] = v[k];
                }
                for (int k = 0; k < 5; k++) host_[((size_t)l * nc_ + 5 * species + k) * nn_ + j] = q[k];
            }
        solution_->upload(host_.data());
    }

    // FiveMomentDGSolutionHelper::compute_global_error (dg_solution_helper.cc:50-69): L2 norm of (solution - f) in one of
    // the 5 components of a species, by VectorTools::integrate_difference with QGauss<dim>(fe_degree).  A diagnostic that
    // runs once after a solve (test/input_test.cc:57-62): the state is downloaded and integrated on the host, element by
    // element in a fixed order.  f(x, out5) returns the 5 exact values at a point.  Box grids (Cartesian JxW).
    double compute_global_error(const std::function<void(const double* x, double* out5)>& f, unsigned int component, int species = 0) {
        if (general_) throw std::runtime_error("compute_global_error: box grids only");
        if (component >= 5) throw std::invalid_argument("compute_global_error: component must be one of the 5 fluid components");
        const BoxDescription& box = tables_.box();
        const int dim = box.dim, Np = fe_degree_ + 1, nq = fe_degree_;
        std::vector<double> xq, wq, I;
        element_.gauss_rule(nq, xq, wq, I);
        int NQ = 1;
        for (int d = 0; d < dim; d++) NQ *= nq;
        std::vector<double> host((size_t)ctx_->n_dofs());
        solution_->download(host.data());
        double jdet = 1.0;
        for (int d = 0; d < dim; d++) jdet *= tables_.h(d);
        double err2 = 0.0;
        for (int64_t l = 0; l < tables_.n_local(); l++) {
            int idx[3];
            tables_.elem_multi_index(tables_.local_to_global()[l], idx);
            const double* ue = &host[((size_t)l * nc_ + 5 * species + component) * nn_];
            double cell = 0.0;
            for (int q = 0; q < NQ; q++) {
                int qi[3] = {0, 0, 0}, t = q;
                for (int d = 0; d < dim; d++) { qi[d] = t % nq; t /= nq; }
                double uh = 0.0;
                for (int j = 0; j < nn_; j++) {
                    int tt = j;
                    double phi = 1.0;
                    for (int d = 0; d < dim; d++) { phi *= I[(size_t)qi[d] * Np + tt % Np]; tt /= Np; }
                    uh += phi * ue[j];
                }
                double x[3] = {0, 0, 0}, w = jdet, exact[5];
                for (int d = 0; d < dim; d++) { x[d] = box.left[d] + (idx[d] + xq[qi[d]]) * tables_.h(d); w *= wq[qi[d]]; }
                f(x, exact);
                const double diff = uh - exact[component];
                cell += diff * diff * w;
            }
            err2 += cell;
        }
        return std::sqrt(err2);   // (a sharded run sums err2 over ranks before the root: the caller's reduction)
    }

    // dg_solver.cc:23-38.  Without time-dependent inflow the inner loop of advance() runs resident on the device side
    // of the ABI (warpii_gpu_advance_to: same dt sequence, no host round trip per step); otherwise every stage goes
    // through the operator so that the inflow tables can follow the stage time.
    void solve(TimestepCallback writeout_callback) {
        steps_ = 0;
        std::vector<TimestepCallback> cbs = {writeout_callback};
        if (device_loop_ && !op_->has_time_dependent_inflow()) {
            op_->refresh_inflow(0.0);
            advance_segments(
                [&](double t, double stop) {
                    int64_t n = 0;
                    check(warpii_gpu_advance_to(ctx_->get(), solution_->id(), f1_id(), &t, stop, fixed_dt_, 0, &n));
                    steps_ += n;
                    return t;
                },
                t_end_, cbs);
            return;
        }
        auto step = [&](double t, double dt) -> bool {
            integrator_->evolve_one_time_step(*op_, *solution_, dt, t);
            steps_++;
            return true;
        };
        auto recommend_dt = [&]() -> double { return fixed_dt_ > 0 ? fixed_dt_ : op_->recommend_dt(*solution_); };
        std::vector<TimestepCallback> callbacks = {writeout_callback};
        advance(step, t_end_, recommend_dt, callbacks);
    }

    // One process per GPU: hand the halo lists of this rank's slab and the communicator id to the context
    void attach_comm(const char id[WARPII_GPU_NCCL_ID_BYTES]) {
        warpii_gpu_halo halo;
        if (general_) {
            if (general_->n_ranks <= 1) throw std::runtime_error("attach_comm: this mesh is not sharded (extension grids run on one GPU)");
            general_->fill(halo);
        } else {
            tables_.fill(halo);
        }
        check(warpii_gpu_attach_comm(ctx_->get(), id, rank_, n_ranks_, &halo));
    }

    GpuSolutionVec& get_solution() { return *solution_; }
    GpuFluidFluxESDGSEMOperator& get_fluid_flux_operator() { return *op_; }
    const BoxMeshTables& tables() const { return tables_; }
    std::shared_ptr<GpuContext> context() const { return ctx_; }
    int64_t steps_taken() const { return steps_; }
    void set_t_end(double t) { t_end_ = t; }
    void set_fixed_dt(double dt) { fixed_dt_ = dt; }
    void set_device_loop(bool on) { device_loop_ = on; }
    int f1_id() const { return integrator_->stage_vector().id(); }
    bool general_geometry() const { return (bool)general_; }
    int n_dims() const { return general_ ? general_->dim : tables_.box().dim; }
    int64_t n_local_elems() const { return general_ ? general_->n_elems : tables_.n_local(); }
    const GeneralMesh* general_mesh() const { return general_.get(); }
    const MappedMeshMetrics& metrics() const { return metrics_; }
    int n_components() const { return nc_; }
    int n_species() const { return n_species_; }
    int nodes_per_elem() const { return nn_; }

   private:
    double t_end_;
    int fe_degree_, n_species_;
    bool fields_enabled_;
    double gas_gamma_;
    int n_boundaries_;
    std::vector<SpeciesBC> bcs_;
    int device_, rank_ = 0, n_ranks_ = 1;
    BoxMeshTables tables_;
    ReferenceElement element_;
    int nc_ = 5, nn_ = 1;
    std::shared_ptr<GpuContext> ctx_;
    std::unique_ptr<GpuSolutionVec> solution_;
    std::unique_ptr<GpuFluidFluxESDGSEMOperator> op_;
    std::unique_ptr<Integrator> integrator_;
    std::vector<double> host_;
    std::unique_ptr<GeneralMesh> general_;
    MappedMeshMetrics metrics_;
    static BoxDescription unit_box(int dim) {
        BoxDescription b;
        b.dim = dim;
        return b;
    }
    int64_t steps_ = 0;
    double fixed_dt_ = 0.0;
    bool device_loop_ = true;
};

}  // namespace warpii_b200
