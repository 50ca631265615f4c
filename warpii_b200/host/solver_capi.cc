// C wrappers of the C++ host layer (include/warpii_host.h).
#include "warpii_host.h"

#include <cstring>
#include <exception>
#include <string>

#include "dg_solver.hpp"
#include "five_moment_app.hpp"
#include "mapped_mesh.hpp"

using namespace warpii_b200;

struct warpii_box_solver {
    std::shared_ptr<FiveMomentGpuSolver> solver;
    int rank = 0, n_ranks = 1;
};

struct warpii_app {
    std::unique_ptr<FiveMomentGpuApp> app;
    warpii_box_solver solver_view;   // the app's solver behind the warpii_box_solver_* accessors
};

namespace {
thread_local std::string g_host_error;

// Host-layer failures are reported through the same accessor as the GPU ABI: stash the text where
// warpii_gpu_last_error() can see it by routing through a failing ABI call is not possible, so the host layer
// keeps its own message and warpii_host_last_error() exposes it.
int host_fail(const std::exception& e) {
    g_host_error = e.what();
    return 1;
}
}  // namespace

extern "C" {

const char* warpii_host_last_error(void) { return g_host_error.c_str(); }

#define GUARD(...)                                    \
    try {                                             \
        __VA_ARGS__;                                  \
        return 0;                                     \
    } catch (const std::exception& e) {               \
        return host_fail(e);                          \
    }

int warpii_box_solver_create(int dim, int fe_degree, int n_species, int fields_enabled, double gas_gamma, const int32_t* nx,
                             const double* left, const double* right, const int32_t* periodic, int n_boundaries,
                             const int32_t* bc_kinds, int rank, int n_ranks, int device, warpii_box_solver** out) {
    GUARD({
        if (!out || !nx || !left || !right) throw std::invalid_argument("warpii_box_solver_create: null argument");
        if (dim < 1 || dim > 3) throw std::invalid_argument("n_dims must be 1, 2, or 3");
        BoxDescription box;
        box.dim = dim;
        for (int d = 0; d < dim; d++) {
            box.nx[d] = nx[d];
            box.left[d] = left[d];
            box.right[d] = right[d];
            box.periodic[d] = periodic ? periodic[d] != 0 : true;
        }
        std::vector<SpeciesBC> bcs(n_species);
        for (int s = 0; s < n_species; s++) {
            bcs[s].kind.assign(n_boundaries, WARPII_BC_WALL);   // species.cc:17: default "Wall"
            bcs[s].inflow.assign(n_boundaries, std::array<double, 5>{{0, 0, 0, 0, 0}});
            if (bc_kinds)
                for (int b = 0; b < n_boundaries; b++) bcs[s].kind[b] = bc_kinds[s * n_boundaries + b];
        }
        auto* s = new warpii_box_solver();
        s->rank = rank;
        s->n_ranks = n_ranks;
        try {
            s->solver = std::make_shared<FiveMomentGpuSolver>(box, fe_degree, n_species, fields_enabled != 0, gas_gamma, 0.0,
                                                              n_boundaries, bcs, rank, n_ranks, device);
            s->solver->reinit();
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    })
}

int warpii_mapped_box_solver_create(int dim, int fe_degree, int n_species, int fields_enabled, double gas_gamma, const int32_t* nx,
                                    const double* left, const double* right, const int32_t* periodic, int n_boundaries,
                                    const int32_t* bc_kinds, warpii_mapping_fn mapping, void* user, int rank, int n_ranks,
                                    int device, warpii_box_solver** out) {
    GUARD({
        if (!out || !nx || !left || !right) throw std::invalid_argument("warpii_mapped_box_solver_create: null argument");
        if (dim < 1 || dim > 3) throw std::invalid_argument("n_dims must be 1, 2, or 3");
        BoxDescription box;
        box.dim = dim;
        for (int d = 0; d < dim; d++) {
            box.nx[d] = nx[d];
            box.left[d] = left[d];
            box.right[d] = right[d];
            box.periodic[d] = periodic ? periodic[d] != 0 : true;
        }
        std::vector<SpeciesBC> bcs(n_species);
        for (int sp = 0; sp < n_species; sp++) {
            bcs[sp].kind.assign(n_boundaries, WARPII_BC_WALL);
            bcs[sp].inflow.assign(n_boundaries, std::array<double, 5>{{0, 0, 0, 0, 0}});
            for (int b = 0; b < n_boundaries && bc_kinds; b++) bcs[sp].kind[b] = bc_kinds[sp * n_boundaries + b];
        }
        std::function<void(const double*, double*)> map_fn;
        if (mapping) map_fn = [=](const double* x, double* y) { mapping(x, y, user); };
        GeneralMesh mesh = GeneralMesh::mapped_box(box, fe_degree, warpii_gpu_elems_per_block(dim, fe_degree), map_fn, rank, n_ranks);
        auto* s = new warpii_box_solver();
        try {
            s->solver = std::make_shared<FiveMomentGpuSolver>(std::move(mesh), n_species, fields_enabled != 0, gas_gamma, 0.0,
                                                             n_boundaries, bcs, device);
            s->solver->reinit();
        } catch (...) {
            delete s;
            throw;
        }
        s->rank = rank;
        s->n_ranks = n_ranks;
        *out = s;
    })
}

int warpii_box_solver_destroy(warpii_box_solver* s) {
    delete s;
    return 0;
}

warpii_gpu_ctx* warpii_box_solver_ctx(warpii_box_solver* s) { return s ? s->solver->context()->get() : nullptr; }
int64_t warpii_box_solver_n_local_elems(const warpii_box_solver* s) { return s->solver->n_local_elems(); }
int64_t warpii_box_solver_n_interface_elems(const warpii_box_solver* s) {
    return s->solver->general_geometry() ? s->solver->general_mesh()->n_interface : s->solver->tables().n_interface();
}
int64_t warpii_box_solver_n_ghost_faces(const warpii_box_solver* s) {
    return s->solver->general_geometry() ? s->solver->general_mesh()->n_ghost_faces : s->solver->tables().n_ghost_faces();
}
int warpii_box_solver_n_components(const warpii_box_solver* s) { return s->solver->n_components(); }
int warpii_box_solver_nodes_per_elem(const warpii_box_solver* s) { return s->solver->nodes_per_elem(); }

int warpii_box_solver_local_to_global(const warpii_box_solver* s, int64_t* out) {
    GUARD({
        if (s->solver->general_geometry()) {   // an extension's cells keep their order; a mapped box has its slab's numbering
            const auto& l2g = s->solver->general_mesh()->local_to_global;
            for (int64_t i = 0; i < s->solver->n_local_elems(); i++) out[i] = l2g.empty() ? i : l2g[i];
            return 0;
        }
        const auto& v = s->solver->tables().local_to_global();
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    })
}

int warpii_box_solver_node_coords(const warpii_box_solver* s, double* xyz) {
    GUARD({
        const std::vector<double> v = s->solver->node_coords();
        std::memcpy(xyz, v.data(), v.size() * sizeof(double));
    })
}

int warpii_box_solver_set_state(warpii_box_solver* s, const double* host) { GUARD({ s->solver->get_solution().upload(host); }) }
int warpii_box_solver_get_state(warpii_box_solver* s, double* host) { GUARD({ s->solver->get_solution().download(host); }) }
int warpii_box_solver_set_inflow(warpii_box_solver* s, int species, int boundary_id, const double q[5]) {
    GUARD({ s->solver->get_fluid_flux_operator().set_inflow(species, boundary_id, q); })
}

int warpii_box_solver_set_inflow_function(warpii_box_solver* s, int species, int boundary_id, warpii_inflow_fn fn, void* user,
                                          int time_dependent) {
    GUARD({
        if (!fn) throw std::invalid_argument("set_inflow_function: null function");
        s->solver->get_fluid_flux_operator().set_inflow_function(
            species, boundary_id, [=](const double* x, double t, double* q5) { fn(x, t, q5, user); }, time_dependent != 0);
    })
}
int warpii_box_solver_set_sources(warpii_box_solver* s, int enabled, double epsilon0, double chi, const double* charge_over_mass) {
    GUARD({
        const int n = s->solver->n_species();
        std::vector<double> qm(n, 0.0);
        if (charge_over_mass) qm.assign(charge_over_mass, charge_over_mass + n);
        s->solver->get_fluid_flux_operator().set_sources(enabled != 0, epsilon0, chi, qm);
    })
}
int warpii_box_solver_set_maxwell(warpii_box_solver* s, int enabled, double light_speed, double chi, double gamma) {
    GUARD({ s->solver->get_fluid_flux_operator().set_maxwell(enabled != 0, light_speed, chi, gamma); })
}
int64_t warpii_box_solver_n_boundary_faces(const warpii_box_solver* s) {
    return s->solver->general_geometry() ? (int64_t)s->solver->general_mesh()->bf_elem.size()
                                         : (int64_t)s->solver->tables().boundary_face_elem().size();
}
int warpii_box_solver_boundary_points(const warpii_box_solver* s, double* xyz, int32_t* face_boundary_id) {
    GUARD({
        const bool general = s->solver->general_geometry();
        const std::vector<double> v = general ? s->solver->metrics().boundary_points : s->solver->boundary_quadrature_points();
        if (xyz && !v.empty()) std::memcpy(xyz, v.data(), v.size() * sizeof(double));
        const auto& ids = general ? s->solver->general_mesh()->bf_id : s->solver->tables().boundary_face_id();
        if (face_boundary_id && !ids.empty()) std::memcpy(face_boundary_id, ids.data(), ids.size() * sizeof(int32_t));
    })
}

int warpii_box_solver_attach_comm(warpii_box_solver* s, const char id[WARPII_GPU_NCCL_ID_BYTES]) {
    GUARD({ s->solver->attach_comm(id); })
}

int warpii_box_solver_solve(warpii_box_solver* s, double t_end, double fixed_dt, double callback_interval, warpii_callback_fn cb,
                            void* user, int64_t* steps_out) {
    GUARD({
        s->solver->set_t_end(t_end);
        s->solver->set_fixed_dt(fixed_dt);
        // FiveMomentApp::run (five_moment.h:233-243): the writeout callback skips t = 0
        const double interval = (cb && callback_interval > 0) ? callback_interval : t_end;
        TimestepCallback callback(interval, [&](double t) { if (cb) cb(t, user); }, false, true);
        s->solver->solve(callback);
        if (steps_out) *steps_out = s->solver->steps_taken();
    })
}

int warpii_box_solver_step(warpii_box_solver* s, double dt, double t) {
    GUARD({
        // one evolve_one_time_step through the same integrator solve() uses
        s->solver->set_t_end(0.0);
        check(warpii_gpu_ssprk2_step(s->solver->context()->get(), s->solver->get_solution().id(), s->solver->f1_id(), dt, t));
    })
}

int warpii_box_solver_global_error(warpii_box_solver* s, warpii_inflow_fn exact, void* user, int species, int component,
                                   double* error_out) {
    GUARD({
        if (!exact || !error_out) throw std::invalid_argument("global_error: null argument");
        if (species < 0 || species >= s->solver->n_species()) throw std::invalid_argument("global_error: species out of range");
        if (component < 0) throw std::invalid_argument("global_error: component out of range");
        *error_out = s->solver->compute_global_error([=](const double* x, double* q5) { exact(x, 0.0, q5, user); },
                                                     (unsigned)component, species);
    })
}

int warpii_box_solver_lsrk_step(warpii_box_solver* s, int scheme, double dt, double t, double* coefficients, int* n_stages_out) {
    GUARD({
        if (scheme < 0 || scheme > 3) throw std::invalid_argument("lsrk_step: scheme must be 0..3");
        const LowStorageRungeKuttaIntegrator integrator((LowStorageRungeKuttaScheme)scheme);
        if (n_stages_out) *n_stages_out = (int)integrator.n_stages();
        if (coefficients) {
            double* o = coefficients;
            for (double v : integrator.get_bi()) *o++ = v;
            for (double v : integrator.get_ai()) *o++ = v;
            for (double v : integrator.get_ci()) *o++ = v;
        }
        if (dt > 0.0) {
            GpuSolutionVec& sol = s->solver->get_solution();
            GpuSolutionVec vec_ri(sol.context());
            GpuSolutionVec vec_ki(sol.context());
            integrator.perform_time_step(s->solver->get_fluid_flux_operator(), t, dt, sol, vec_ri, vec_ki);
        }
    })
}

int warpii_box_solver_recommend_dt(warpii_box_solver* s, double* dt_out) {
    GUARD({ *dt_out = s->solver->get_fluid_flux_operator().recommend_dt(s->solver->get_solution()); })
}

// ---- the FiveMoment application from an input file --------------------------------------------------------------
int warpii_app_create(const char* input_text, int rank, int n_ranks, int device, warpii_app** out) {
    GUARD({
        if (!input_text || !out) throw std::invalid_argument("warpii_app_create: null argument");
        auto* a = new warpii_app();
        try {
            a->app = FiveMomentGpuApp::create_from_input(input_text, rank, n_ranks, device);
        } catch (...) {
            delete a;
            throw;
        }
        a->solver_view.solver = a->app->solver_ptr();
        a->solver_view.rank = rank;
        a->solver_view.n_ranks = n_ranks;
        *out = a;
    })
}
namespace {
// an extension that hands over arrays (the C ABI cannot carry a C++ object)
class TableExtension : public GridExtension {
   public:
    Triangulation2D tria;
    void populate_triangulation(Triangulation2D& out, const ParameterFile&) override { out = tria; }
};
}  // namespace

static void fill_triangulation(Triangulation2D& tria, int64_t n_vertices, const double* vertices, int64_t n_cells,
                               const int32_t* cells, const int32_t* face_boundary_ids) {
    for (int64_t v = 0; v < n_vertices; v++) tria.vertices.push_back({{vertices[2 * v], vertices[2 * v + 1]}});
    for (int64_t c = 0; c < n_cells; c++) {
        tria.cells.push_back({{cells[4 * c], cells[4 * c + 1], cells[4 * c + 2], cells[4 * c + 3]}});
        for (int f = 0; f < 4 && face_boundary_ids; f++)
            if (face_boundary_ids[4 * c + f] >= 0) tria.boundary_ids[{(int)c, f}] = face_boundary_ids[4 * c + f];
    }
}

int warpii_host_triangulation_tables(int64_t n_vertices, const double* vertices, int64_t n_cells, const int32_t* cells,
                                     const int32_t* face_boundary_ids, int fe_degree, int32_t* face_neighbor,
                                     int32_t* neighbor_face, double* xyz, int32_t* bf_elem, int32_t* bf_side, int32_t* bf_id,
                                     int64_t* n_boundary_faces_out) {
    GUARD({
        if (!vertices || !cells || !face_neighbor || !neighbor_face || !xyz) throw std::invalid_argument("triangulation_tables: null argument");
        Triangulation2D tria;
        fill_triangulation(tria, n_vertices, vertices, n_cells, cells, face_boundary_ids);
        const GeneralMesh m = GeneralMesh::from_triangulation(tria, fe_degree);
        std::memcpy(face_neighbor, m.face_neighbor.data(), m.face_neighbor.size() * sizeof(int32_t));
        std::memcpy(neighbor_face, m.neighbor_face.data(), m.neighbor_face.size() * sizeof(int32_t));
        std::memcpy(xyz, m.xyz.data(), m.xyz.size() * sizeof(double));
        if (bf_elem) std::memcpy(bf_elem, m.bf_elem.data(), m.bf_elem.size() * sizeof(int32_t));
        if (bf_side) std::memcpy(bf_side, m.bf_side.data(), m.bf_side.size() * sizeof(int32_t));
        if (bf_id) std::memcpy(bf_id, m.bf_id.data(), m.bf_id.size() * sizeof(int32_t));
        if (n_boundary_faces_out) *n_boundary_faces_out = (int64_t)m.bf_elem.size();
    })
}

int warpii_app_create_with_triangulation(const char* input_text, int64_t n_vertices, const double* vertices, int64_t n_cells,
                                         const int32_t* cells, const int32_t* face_boundary_ids, int device, warpii_app** out) {
    GUARD({
        if (!input_text || !out || !vertices || !cells) throw std::invalid_argument("warpii_app_create_with_triangulation: null argument");
        auto ext = std::make_shared<TableExtension>();
        fill_triangulation(ext->tria, n_vertices, vertices, n_cells, cells, face_boundary_ids);
        auto* a = new warpii_app();
        try {
            a->app = FiveMomentGpuApp::create_from_input(input_text, 0, 1, device, ext);
        } catch (...) {
            delete a;
            throw;
        }
        a->solver_view.solver = a->app->solver_ptr();
        a->solver_view.rank = 0;
        a->solver_view.n_ranks = 1;
        *out = a;
    })
}
int warpii_app_destroy(warpii_app* a) {
    delete a;
    return 0;
}
warpii_box_solver* warpii_app_solver(warpii_app* a) { return a ? &a->solver_view : nullptr; }
int warpii_app_describe(const warpii_app* a, int32_t ints[16], double dbls[16]) {
    GUARD({
        const FiveMomentGpuApp& app = *a->app;
        for (int i = 0; i < 16; i++) { ints[i] = 0; dbls[i] = 0.0; }
        ints[0] = app.n_dims();
        ints[1] = app.n_species();
        ints[2] = app.n_boundaries();
        ints[3] = app.fe_degree();
        ints[4] = app.fields_enabled();
        ints[5] = app.write_output();
        ints[6] = app.n_writeout_frames();
        for (int d = 0; d < 3; d++) { ints[7 + d] = app.box().nx[d]; ints[10 + d] = app.box().periodic[d]; }
        dbls[0] = app.gas_gamma();
        dbls[1] = app.t_end();
        for (int d = 0; d < 3; d++) { dbls[2 + d] = app.box().left[d]; dbls[5 + d] = app.box().right[d]; }
    })
}
int warpii_app_species(const warpii_app* a, int species, char name[16], double* charge, double* mass, int32_t* bc_kinds) {
    GUARD({
        const SpeciesDescription& sp = a->app->species().at(species);
        std::snprintf(name, 16, "%s", sp.name.c_str());
        *charge = sp.charge;
        *mass = sp.mass;
        for (size_t b = 0; b < sp.bc_kind.size(); b++) bc_kinds[b] = sp.bc_kind[b];
    })
}
int warpii_app_eval_function(const warpii_app* a, int species, int boundary_id, int64_t n, const double* xyz, double t, double* q5_out,
                             int32_t* time_dependent_out) {
    GUARD({
        const SpeciesDescription& sp = a->app->species().at(species);
        const SpeciesFunc* f = boundary_id < 0 ? sp.initial_condition.get() : sp.inflow.at(boundary_id).get();
        if (!f) throw std::invalid_argument("boundary " + std::to_string(boundary_id) + " has no inflow function");
        const int dim = a->app->n_dims();
        for (int64_t i = 0; i < n; i++) f->conserved(xyz + i * dim, t, q5_out + i * 5);
        if (time_dependent_out) *time_dependent_out = f->time_dependent();
    })
}
int warpii_app_set_output_dir(warpii_app* a, const char* dir) { GUARD({ a->app->set_output_dir(dir ? dir : ""); }) }
int warpii_app_format_workdir(const warpii_app* a, const char* input_name, char* out, int out_len) {
    GUARD({
        // the name of an input FILE goes through remove_file_extension first, like format_workdir's caller (warpii.cc:209-210)
        const std::string name = input_name ? input_name : "";
        const std::string stem = name == "STDIN" ? name : FiveMomentGpuApp::remove_file_extension(name);
        std::snprintf(out, out_len, "%s", a->app->format_workdir(stem).c_str());
    })
}
int warpii_app_set_device_loop(warpii_app* a, int on) { GUARD({ a->app->get_solver().set_device_loop(on != 0); }) }
int warpii_app_setup(warpii_app* a) { GUARD({ a->app->setup(); }) }
int warpii_app_run(warpii_app* a, warpii_frame_fn cb, void* user, int64_t* steps_out) {
    GUARD({
        if (cb) a->app->set_frame_callback([=](unsigned frame, double t) { cb(frame, t, user); });
        a->app->run();
        if (steps_out) *steps_out = a->app->get_solver().steps_taken();
    })
}

int warpii_host_write_vtu(const char* path, int dim, int fe_degree, int64_t n_elems, int nc, int n_species, const char* species_names,
                          int fields_enabled, double gas_gamma, int owner_rank, const double* state, const double* xyz) {
    GUARD({
        std::vector<VtuSpecies> names;
        for (const std::string& n : ParameterFile::split(species_names ? species_names : "", ',')) names.push_back({ParameterFile::trim(n)});
        if ((int)names.size() != n_species) throw std::invalid_argument("write_vtu: need one name per species");
        VtuWriter::write(path, dim, fe_degree, n_elems, nc, names, fields_enabled != 0, gas_gamma, owner_rank, state, xyz);
    })
}

int warpii_host_advance(warpii_step_fn step, double t_end, warpii_dt_fn recommend_dt, int n_callbacks, const double* intervals,
                        const int32_t* perform_zeroth, const int32_t* perform_final, warpii_cb_index_fn cb, void* user) {
    GUARD({
        std::vector<TimestepCallback> cbs;
        for (int i = 0; i < n_callbacks; i++)
            cbs.emplace_back(intervals[i], [=](double t) { cb(t, i, user); }, perform_zeroth[i] != 0, perform_final[i] != 0);
        advance([&](double t, double dt) { return step(t, dt, user) != 0; }, t_end, [&]() { return recommend_dt(user); }, cbs);
    })
}

int warpii_host_mapped_metrics(int dim, int fe_degree, int64_t n_elems, const double* xyz, const int32_t* face_neighbor,
                               const int32_t* neighbor_face, int64_t n_boundary_faces, const int32_t* bf_elem,
                               const int32_t* bf_side, double* inverse_jacobian, double* face_normal, double* face_jacobian,
                               double* boundary_normal, double* boundary_jacobian, double* boundary_points) {
    GUARD({
        if (!xyz || !face_neighbor || !inverse_jacobian || !face_normal || !face_jacobian)
            throw std::invalid_argument("warpii_host_mapped_metrics: null argument");
        if (fe_degree < 1 || fe_degree > 6) throw std::invalid_argument("fe_degree must be in [1,6]");
        const MappedMeshMetrics M = build_mapped_metrics(dim, fe_degree, n_elems, xyz, face_neighbor, neighbor_face,
                                                         n_boundary_faces, bf_elem, bf_side);
        std::memcpy(inverse_jacobian, M.inverse_jacobian.data(), M.inverse_jacobian.size() * sizeof(double));
        std::memcpy(face_normal, M.face_normal.data(), M.face_normal.size() * sizeof(double));
        std::memcpy(face_jacobian, M.face_jacobian.data(), M.face_jacobian.size() * sizeof(double));
        if (boundary_normal) std::memcpy(boundary_normal, M.boundary_normal.data(), M.boundary_normal.size() * sizeof(double));
        if (boundary_jacobian) std::memcpy(boundary_jacobian, M.boundary_jacobian.data(), M.boundary_jacobian.size() * sizeof(double));
        if (boundary_points) std::memcpy(boundary_points, M.boundary_points.data(), M.boundary_points.size() * sizeof(double));
    });
}

int warpii_host_box_tables(int dim, const int32_t* nx, const int32_t* periodic, int rank, int n_ranks, int elems_per_block, int64_t counts[6],
                           int64_t* local_to_global, int32_t* face_neighbor, int32_t* bf_elem, int32_t* bf_side, int32_t* bf_id,
                           int32_t* peer_rank, int64_t* send_offset, int64_t* recv_offset, int32_t* send_elem, int32_t* send_side,
                           int64_t* ghost_global_elem, int32_t* ghost_side) {
    GUARD({
        BoxDescription box;
        box.dim = dim;
        for (int d = 0; d < dim; d++) {
            box.nx[d] = nx[d];
            box.left[d] = 0.0;
            box.right[d] = 1.0;
            box.periodic[d] = periodic ? periodic[d] != 0 : true;
        }
        BoxMeshTables t(box, rank, n_ranks, elems_per_block);
        counts[0] = t.n_local();
        counts[1] = t.n_interface();
        counts[2] = t.n_ghost_faces();
        counts[3] = (int64_t)t.boundary_face_elem().size();
        counts[4] = (int64_t)t.peer_rank().size();
        counts[5] = (int64_t)t.send_elem().size();
        auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
        cp(local_to_global, t.local_to_global());
        cp(face_neighbor, t.face_neighbor());
        cp(bf_elem, t.boundary_face_elem());
        cp(bf_side, t.boundary_face_side());
        cp(bf_id, t.boundary_face_id());
        cp(peer_rank, t.peer_rank());
        cp(send_offset, t.send_offset());
        cp(recv_offset, t.recv_offset());
        cp(send_elem, t.send_elem());
        cp(send_side, t.send_side());
        cp(ghost_global_elem, t.ghost_global_elem());
        cp(ghost_side, t.ghost_side());
    })
}

}  // extern "C"
