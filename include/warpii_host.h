/* ============================================================================
 * warpii_host.h -- C entry points of the C++ host layer (box grids)
 *
 * The host layer (warpii_b200/host/) mirrors the reference's C++ classes above the
 * operator: FiveMomentDGSolver (src/five_moment/dg_solver.{h,cc}), SSPRK2Integrator
 * (src/rk.h:79-117), advance() (src/timestepper.cc:6-56) and the HyperRectangle grid
 * (src/grid_descriptions.cc:51-74).  These wrappers exist so that non-C++ callers
 * (the Python tests, bench.py) can drive that layer; a C++ caller includes the
 * headers in warpii_b200/host/ directly.  Status/err conventions as warpii_gpu.h.
 * ==========================================================================*/
#ifndef WARPII_HOST_H
#define WARPII_HOST_H

#include <stdint.h>

#include "warpii_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct warpii_box_solver warpii_box_solver;

/* message of the last failing warpii_box_solver_* / warpii_host_* call on this thread (C++ exceptions of the host
 * layer, including the ones that carry a warpii_gpu_last_error() text, end up here) */
const char* warpii_host_last_error(void);

/* Box grid description + discretisation parameters (the input-file keys of five_moment.h:99-147 and
 * grid_descriptions.cc:27-49 that reach the hot path). bc_kinds[n_species][n_boundaries] may be NULL when
 * n_boundaries == 0.  rank/n_ranks shard the elements in slabs along the last dimension. */
int warpii_box_solver_create(int dim, int fe_degree, int n_species, int fields_enabled, double gas_gamma,
                             const int32_t* nx, const double* left, const double* right,
                             const int32_t* periodic, int n_boundaries, const int32_t* bc_kinds, int rank,
                             int n_ranks, int device, warpii_box_solver** out);
/* The same solver on a MAPPED box: box connectivity (and slab sharding), element support points pushed through
 * x' = mapping(x) (NULL = identity), i.e. curved elements on the general-geometry kernels (warpii_gpu_set_geometry).
 * Every warpii_box_solver_* call works on the result; node_coords returns the mapped points. */
typedef void (*warpii_mapping_fn)(const double* x, double* x_out, void* user);
int warpii_mapped_box_solver_create(int dim, int fe_degree, int n_species, int fields_enabled, double gas_gamma,
                                    const int32_t* nx, const double* left, const double* right, const int32_t* periodic,
                                    int n_boundaries, const int32_t* bc_kinds, warpii_mapping_fn mapping, void* user, int rank,
                                    int n_ranks, int device, warpii_box_solver** out);
int warpii_box_solver_destroy(warpii_box_solver* s);
/* the operator context underneath, for direct use of the warpii_gpu_* ABI (vector 0 = solution, 1 = f_1) */
warpii_gpu_ctx* warpii_box_solver_ctx(warpii_box_solver* s);

int64_t warpii_box_solver_n_local_elems(const warpii_box_solver* s);
int64_t warpii_box_solver_n_interface_elems(const warpii_box_solver* s);
int64_t warpii_box_solver_n_ghost_faces(const warpii_box_solver* s);
int warpii_box_solver_n_components(const warpii_box_solver* s);
int warpii_box_solver_nodes_per_elem(const warpii_box_solver* s);
/* global (lexicographic, x fastest) index of every owned element, in device order */
int warpii_box_solver_local_to_global(const warpii_box_solver* s, int64_t* out);
/* xyz[elem][node][dim] of the owned elements, in device order */
int warpii_box_solver_node_coords(const warpii_box_solver* s, double* xyz);

/* state <-> host, in device order [local elem][comp][node] */
int warpii_box_solver_set_state(warpii_box_solver* s, const double* host);
int warpii_box_solver_get_state(warpii_box_solver* s, double* host);
int warpii_box_solver_set_inflow(warpii_box_solver* s, int species, int boundary_id, const double q[5]);

/* Space/time-dependent inflow: q5 = fn(x[dim], t), the Function<dim> of EulerBCMap::set_inflow_boundary
 * (bc_helper.h:52-64).  The host layer tabulates it at this rank's boundary quadrature points with the stage time before
 * every stage (time_dependent != 0; the reference's set_time(t), fluid_flux_es_dgsem_operator.h:139-144) or once, and
 * uploads the table through warpii_gpu_set_inflow_table. */
typedef void (*warpii_inflow_fn)(const double* x, double t, double* q5, void* user);
int warpii_box_solver_set_inflow_function(warpii_box_solver* s, int species, int boundary_id, warpii_inflow_fn fn,
                                          void* user, int time_dependent);
/* two-fluid source terms on/off (warpii_gpu_set_sources); charge_over_mass[n_species] */
int warpii_box_solver_set_sources(warpii_box_solver* s, int enabled, double epsilon0, double chi,
                                  const double* charge_over_mass);
/* perfectly hyperbolic Maxwell fluxes for the field components on/off (warpii_gpu_set_maxwell) */
int warpii_box_solver_set_maxwell(warpii_box_solver* s, int enabled, double light_speed, double chi, double gamma);
/* this rank's boundary faces: xyz[face][point][dim] of the quadrature points (Gauss(fe_degree+2)^(dim-1) per face) and
 * the boundary id of every face, in the order of the inflow table; either output may be NULL */
int64_t warpii_box_solver_n_boundary_faces(const warpii_box_solver* s);
int warpii_box_solver_boundary_points(const warpii_box_solver* s, double* xyz, int32_t* face_boundary_id);

/* attach the NCCL communicator (id from warpii_gpu_nccl_unique_id on rank 0, broadcast by the caller) */
int warpii_box_solver_attach_comm(warpii_box_solver* s, const char id[WARPII_GPU_NCCL_ID_BYTES]);

/* FiveMomentDGSolver::solve: advance() with SSPRK2 steps and recommend_dt every step, one callback of the given
 * interval (NULL = none).  fixed_dt > 0 replaces recommend_dt (for parity runs).  Returns steps in *steps_out. */
typedef void (*warpii_callback_fn)(double t, void* user);
int warpii_box_solver_solve(warpii_box_solver* s, double t_end, double fixed_dt, double callback_interval,
                            warpii_callback_fn cb, void* user, int64_t* steps_out);
/* SSPRK2Integrator::evolve_one_time_step / operator.recommend_dt on the solver's own solution vector */
int warpii_box_solver_step(warpii_box_solver* s, double dt, double t);
int warpii_box_solver_recommend_dt(warpii_box_solver* s, double* dt_out);
/* FiveMomentDGSolutionHelper::compute_global_error (dg_solution_helper.cc:50-69): L2 norm of (solution - exact) in one fluid
 * component of a species, QGauss(fe_degree) quadrature; exact(x, t = 0, q5, user) returns the five exact values at x. */
int warpii_box_solver_global_error(warpii_box_solver* s, warpii_inflow_fn exact, void* user, int species, int component,
                                   double* error_out);
/* LowStorageRungeKuttaIntegrator(scheme).perform_time_step on the solver's solution vector (rk.h:10-77);
 * scheme: 0 = stage_3_order_3, 1 = stage_5_order_4, 2 = stage_7_order_4, 3 = stage_9_order_5 (rk.h:10-15).
 * coefficients (may be NULL) receives b[n], a[n-1], c[n] back to back; n_stages_out the stage count. */
int warpii_box_solver_lsrk_step(warpii_box_solver* s, int scheme, double dt, double t, double* coefficients, int* n_stages_out);

/* ---- the FiveMoment application driven by a WarpII input file ------------------------------------------------------
 * Replaces Warpii::setup/run for Application = FiveMoment (warpii.cc:126-196, five_moment.cc:22-52, five_moment.h:99-243)
 * on HyperRectangle grids: the same entries, defaults and patterns, parsed by an independent reader
 * (warpii_b200/host/parameter_file.hpp) and expression evaluator (expression.hpp).  create needs no GPU; setup and run do.
 * The frame callback fires where the reference writes solution_<n>.vtu (frame 0 in setup). */
typedef struct warpii_app warpii_app;
typedef void (*warpii_frame_fn)(unsigned frame, double t, void* user);
int warpii_app_create(const char* input_text, int rank, int n_ranks, int device, warpii_app** out);
/* GridType = Extension (src/extensions/extension.h:24-45; the reference passes a GridExtension object to
 * Warpii::create_from_cli): the triangulation an extension would populate, as arrays.  vertices[n_vertices][2];
 * cells[n_cells][4] in deal.II vertex order (v00, v10, v01, v11), counter-clockwise; face_boundary_ids[n_cells][4] (may be
 * NULL): boundary id of each local face 0..3 = x-low, x-high, y-low, y-high, < 0 = leave the default (0), ignored on interior
 * faces.  Elements are straight-sided (MappingQ without a manifold); one GPU.  C++ callers derive from
 * warpii_b200::GridExtension instead (warpii_b200/host/general_mesh.hpp). */
int warpii_app_create_with_triangulation(const char* input_text, int64_t n_vertices, const double* vertices, int64_t n_cells,
                                         const int32_t* cells, const int32_t* face_boundary_ids, int device, warpii_app** out);
int warpii_app_destroy(warpii_app* app);
/* the app's solver, valid until warpii_app_destroy (state access, node coordinates, communicator attachment ...) */
warpii_box_solver* warpii_app_solver(warpii_app* app);
/* parsed parameters: ints = {n_dims, n_species, n_boundaries, fe_degree, fields_enabled, write_output,
 * n_writeout_frames, nx[3], periodic[3]}, dbls = {gas_gamma, t_end, left[3], right[3]} */
int warpii_app_describe(const warpii_app* app, int32_t ints[16], double dbls[16]);
int warpii_app_species(const warpii_app* app, int species, char name[16], double* charge, double* mass,
                       int32_t* bc_kinds /* [n_boundaries] */);
/* evaluate a parsed SpeciesFunc on the host: boundary_id < 0 = the initial condition, else that boundary's inflow
 * function.  xyz[n][n_dims] -> q5_out[n][5] conserved (species_func.cc:9-30). */
int warpii_app_eval_function(const warpii_app* app, int species, int boundary_id, int64_t n, const double* xyz,
                             double t, double* q5_out, int32_t* time_dependent_out);
/* with write_output, frames go to <dir>/solution_<frame>.vtu (the fields of five_moment.h:245-315 and of the
 * post-processor, postprocessor.h:33-62) + a frames.txt index; "" = no files */
int warpii_app_set_output_dir(warpii_app* app, const char* dir);
/* WorkDir format of the input (%A__%I by default) expanded as format_workdir does (warpii.cc:205-219) */
int warpii_app_format_workdir(const warpii_app* app, const char* input_name, char* out, int out_len);
/* 0: drive every stage from the host like the reference's solve(); 1 (default): device-resident loop when no inflow
 * function depends on t */
int warpii_app_set_device_loop(warpii_app* app, int on);
int warpii_app_setup(warpii_app* app);
int warpii_app_run(warpii_app* app, warpii_frame_fn cb, void* user, int64_t* steps_out);

/* the frame writer alone (warpii_b200/host/vtu_writer.hpp; stands in for FiveMomentApp::output_results + DataOut,
 * five_moment.h:245-315): state[elem][comp][node], xyz[elem][node][dim], species_names comma-separated */
int warpii_host_write_vtu(const char* path, int dim, int fe_degree, int64_t n_elems, int nc, int n_species,
                          const char* species_names, int fields_enabled, double gas_gamma, int owner_rank,
                          const double* state, const double* xyz);

/* the time loop alone (timestepper.cc:6-56) with C callbacks, for the reference's TimestepperTest cases */
typedef int (*warpii_step_fn)(double t, double dt, void* user);
typedef double (*warpii_dt_fn)(void* user);
typedef void (*warpii_cb_index_fn)(double t, int index, void* user);
int warpii_host_advance(warpii_step_fn step, double t_end, warpii_dt_fn recommend_dt, int n_callbacks,
                        const double* intervals, const int32_t* perform_zeroth, const int32_t* perform_final,
                        warpii_cb_index_fn cb, void* user);

/* metric terms of a mesh of curved quadrilaterals / hexahedra from the elements' Gauss-Lobatto support points
 * xyz[n_elems][Np^dim][dim] (warpii_b200/host/mapped_mesh.hpp; stands in for MappingQ(fe_degree) + MatrixFree's mapping
 * info): fills the tables of warpii_gpu_geometry, plus the physical coordinates of the boundary Gauss points
 * (boundary_points, may be NULL).  No GPU needed. */
int warpii_host_mapped_metrics(int dim, int fe_degree, int64_t n_elems, const double* xyz, const int32_t* face_neighbor,
                               const int32_t* neighbor_face, int64_t n_boundary_faces, const int32_t* bf_elem,
                               const int32_t* bf_side, double* inverse_jacobian, double* face_normal, double* face_jacobian,
                               double* boundary_normal, double* boundary_jacobian, double* boundary_points);

/* connectivity builder alone (no GPU; warpii_b200/host/general_mesh.hpp::GeneralMesh::from_triangulation): the flat tables a
 * triangulation of quadrilaterals turns into.  Outputs: face_neighbor[n_cells][4], neighbor_face[n_cells][4],
 * xyz[n_cells][(fe_degree+1)^2][2]; boundary faces in bf_elem/bf_side/bf_id (each sized 4*n_cells by the caller), their
 * number in *n_boundary_faces_out. */
int warpii_host_triangulation_tables(int64_t n_vertices, const double* vertices, int64_t n_cells, const int32_t* cells,
                                     const int32_t* face_boundary_ids, int fe_degree, int32_t* face_neighbor,
                                     int32_t* neighbor_face, double* xyz, int32_t* bf_elem, int32_t* bf_side, int32_t* bf_id,
                                     int64_t* n_boundary_faces_out);

/* mesh-table builder alone (no GPU): fills caller-provided arrays for tests of the partitioning logic.
 * Pass NULL output pointers to query sizes through the counts array:
 * counts = {n_local, n_interface, n_ghost_faces, n_boundary_faces, n_peers, n_send}. */
int warpii_host_box_tables(int dim, const int32_t* nx, const int32_t* periodic, int rank, int n_ranks,
                           int elems_per_block /* patch size of the element numbering, 1 = lexicographic */,
                           int64_t counts[6], int64_t* local_to_global, int32_t* face_neighbor,
                           int32_t* bf_elem, int32_t* bf_side, int32_t* bf_id, int32_t* peer_rank,
                           int64_t* send_offset, int64_t* recv_offset, int32_t* send_elem, int32_t* send_side,
                           int64_t* ghost_global_elem, int32_t* ghost_side);

#ifdef __cplusplus
}
#endif
#endif /* WARPII_HOST_H */
