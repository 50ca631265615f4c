// ============================================================================
// oracle/dgsem_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain-C++ restatement of the algorithm of WarpII's entropy-stable DGSEM
// operator + SSPRK2 time loop (the hot path named by BASELINE.json), written
// from the reference's formulas, NOT copied from its sources.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  The product (warpii_b200/) never links or calls it.
//
// Parity status (see DESIGN.md "Oracle"):
//   * point physics (ln_avg, EC / ES / LF fluxes), index maps, time loop:
//     PINNED by the reference's own golden vectors (test/euler_test.cc,
//     test/dof_utils_test.cc, test/timestepper_test.cc) and by oracle/_ref,
//     which compiles the reference's euler.h / dof_utils.cc / timestepper.cc
//     themselves against a minimal dealii::Tensor shim.
//   * GLL quadrature / projection / D / Q / JxW / face lifting / inverse mass:
//     pinned by the reference's integration tests restated in tests/.
//   * shock indicator alpha, recommend_dt, every 3D result, everything on general
//     (curved / unstructured) geometry, low-storage RK stages: PARITY UNPINNED
//     (deal.II is not available; no reference fixture exists).
//
// All arrays are C-contiguous doubles.  State layout (as deal.II FE_DGQ^nc
// with MatrixFree numbering):  u[elem][comp][node], nodes
// lexicographic with x fastest, elements lexicographic with x fastest.
// ============================================================================
#pragma once
#include <cstdint>

#ifdef __cplusplus
extern "C" {
#endif

// ---- point physics (reference: src/five_moment/euler.h) --------------------
double orc_ln_avg(double a, double b);                                   // euler.h:118-125
// logarithm used inside ln_avg: 1 = oracle/det_log.h (default; bit-identical to the device's), 0 = libm like the reference
void orc_set_log_impl(int deterministic);
int orc_get_log_impl(void);
double orc_det_log(double x);
double orc_pressure(const double q[5], double gamma);                    // euler.h:32-44
void orc_euler_flux(int dim, const double q[5], double gamma, double* F /*[5][dim]*/);   // :46-63
void orc_lf_flux(int dim, const double qin[5], const double qout[5], const double* n,
                 double gamma, double out[5]);                           // :68-89
void orc_ec_flux(int dim, const double qj[5], const double ql[5], double gamma,
                 double* F /*[5][dim]*/);                                // :186-228
void orc_es_flux(int dim, const double qj[5], const double ql[5], const double* n,
                 double gamma, double out[5]);                           // :232-284
void orc_entropy_variables(const double q[5], double gamma, double w[5]);   // :150-169
double orc_mathematical_entropy(const double q[5], double gamma);        // :143-148
void orc_entropy_flux(int dim, const double q[5], double gamma, double* out /*[dim]*/);  // :171-181
void orc_primitive_to_conserved(const double prim[5], double gamma, double cons[5]);  // species_func.cc:15-28

// ---- index maps (reference: src/dof_utils.cc) -------------------------------
unsigned orc_pencil_stride(unsigned Np, unsigned d);                     // :7-18
unsigned orc_pencil_base(int dim, unsigned q, unsigned Np, unsigned d);  // :20-45
unsigned orc_quadrature_point_neighbor(int dim, unsigned q, unsigned k, unsigned Np, unsigned d);  // :47-58
unsigned orc_quad_point_1d_index(int dim, unsigned q, unsigned Np, unsigned d);   // :60-75
int orc_pencil_starts(int dim, unsigned Np, unsigned d, unsigned* out);  // :77-96 (returns count)

// ---- time loop (reference: src/timestepper.cc:6-56) -------------------------
typedef int (*orc_step_fn)(double t, double dt, void* user);
typedef double (*orc_dt_fn)(void* user);
typedef void (*orc_cb_fn)(double t, int cb_index, void* user);
void orc_advance(orc_step_fn step, double t_end, orc_dt_fn recommend_dt, int n_callbacks,
                 const double* intervals, const int* perform_zeroth, const int* perform_final,
                 orc_cb_fn cb, void* user);

// ---- reference element tables ------------------------------------------------
// x,w: GLL(Np) on [0,1]; D[j*Np+l] = l_l'(x_j) on [0,1]; Q = diag(w) D.
void orc_gll(int Np, double* x, double* w);
void orc_gauss(int n, double* x, double* w);
void orc_diff_matrix(int Np, double* D);
void orc_legendre_analysis_1d(int Np, double* V);  // V[k*Np+q] = (k+1/2) w_q sqrt2 P_k(2x_q-1)

// ---- discretised operator on a Cartesian box ---------------------------------
// bc_kinds[species][2*dim]: 0 = Wall, 1 = Outflow (supersonic), 2 = Inflow; ignored on periodic dims.
enum { ORC_BC_WALL = 0, ORC_BC_OUTFLOW = 1, ORC_BC_INFLOW = 2, ORC_BC_SUBSONIC_OUTFLOW = 3 };   // 3: fluid_flux_es_dgsem_operator.h:385-390
void* orc_create(int dim, int fe_degree, int n_species, int fields_enabled, double gamma,
                 const int* nx, const double* left, const double* right, const int* periodic,
                 const int* bc_kinds);
// The same operator on an arbitrary conforming mesh of (curved) quadrilaterals / hexahedra: the
// caller supplies what the reference reads from deal.II's MappingQ(p) + MatrixFree per point.
//   face_neighbor[n_elems][2*dim]   neighbour element, or -1 - (boundary face number)
//   neighbor_face[n_elems][2*dim]   local face of the neighbour + 8 * (tangential order reversed, 2D); NULL = opposite face
//   bf_id[n_bfaces]                 boundary id per boundary face; bc_kinds[n_species][n_boundaries]
//   inverse_jacobian[n_elems][NN][dim][dim]   J^{-T} at the GLL nodes (FEEvaluation::inverse_jacobian)
//   face_normal[n_elems][2*dim][nF][dim], face_jacobian[n_elems][2*dim][nF]   unit outward normal and surface Jacobian
//       at the face GLL nodes (FEFaceEvaluation::normal_vector, JxW / weight), identical on both sides up to sign
//   boundary_normal[n_bfaces][nG][dim], boundary_jacobian[n_bfaces][nG]   the same at the Gauss(p+2) points of boundary faces
// PARITY UNPINNED: the reference's tests never leave Cartesian meshes, and deal.II is not available here.
void* orc_create_general(int dim, int fe_degree, int n_species, int fields_enabled, double gamma, int64_t n_elems,
                         int n_boundaries, const int64_t* face_neighbor, const int32_t* neighbor_face, int64_t n_bfaces,
                         const int32_t* bf_id, const int* bc_kinds, const double* inverse_jacobian,
                         const double* face_normal, const double* face_jacobian, const double* boundary_normal,
                         const double* boundary_jacobian);
void orc_destroy(void* h);
void orc_set_threads(void* h, int n);   // OpenMP threads for the cell/face loops (default 1, like the reference)
int64_t orc_n_elems(void* h);
int64_t orc_n_dofs(void* h);      // n_elems * nc * Np^dim
int orc_n_components(void* h);
int orc_nodes_per_elem(void* h);
int orc_n_boundaries(void* h);    // 2*dim
// Node coordinates xyz[elem][node][dim].
void orc_node_coords(void* h, double* xyz);
// Constant conserved inflow state for (species, boundary id).
void orc_set_inflow(void* h, int species, int boundary_id, const double q[5]);
// Space/time-dependent conserved inflow state q5 = fn(x[dim], t): the Function<dim> the reference evaluates at every
// boundary quadrature point after set_time(t) with the stage time (fluid_flux_es_dgsem_operator.h:139-144, 381-384).
// NULL restores the constant state.
// Two-fluid source terms (Lorentz force, E.J work, -J/eps0 on E, chi rho_c/eps0 on phi): north_star kernel 4.  The
// reference has no such terms (field components are carried through unchanged), so this is new physics defined in
// dgsem_oracle.cc::add_sources; off unless enabled here.  charge_over_mass[n_species].
void orc_set_sources(void* h, int enabled, double epsilon0, double chi, const double* charge_over_mass);
// Perfectly hyperbolic Maxwell fluxes for the 8 field components (curl, divergence cleaning; new physics, see
// dgsem_oracle.cc::add_maxwell): light speed c, cleaning speeds chi (electric) and gamma (magnetic) in units of c.  With
// it the transport speed of recommend_dt also covers c max(1, chi, gamma) and (sources on) the plasma / cyclotron frequency.
// Returns non-zero if the context has no field components, is not a Cartesian box or light_speed <= 0.
int orc_set_maxwell(void* h, int enabled, double light_speed, double chi, double gamma);
typedef void (*orc_inflow_fn)(const double* x, double t, double* q5, void* user);
void orc_set_inflow_function(void* h, int species, int boundary_id, orc_inflow_fn fn, void* user);
// dudt = M^-1 R(u) (fluid comps only; field comps 0); bif_rate[5*n_boundaries] per species summed as the reference does.
void orc_rhs(void* h, const double* u, double t, double* dudt, double* bif_rate);
// Integrated cell residual (volume + subcell FV, value*JxW as integrate_scatter leaves it) of ONE cell/species with a
// prescribed blending factor; ue[5][NN], R[5][NN] (zeroed here).  split_form_volume_flux.h:61-99, subcell_finite_volume_flux.h:68-159.
void orc_cell_residual(void* h, const double* ue, double alpha, double* R);
// Shock indicator of one cell given v[NN] = p*rho at the nodes (persson_peraire_shock_indicator.h:44-123).
double orc_shock_indicator(void* h, const double* v);
// Per-element, per-species blending factor alpha[elem][species].
void orc_alpha(void* h, const double* u, double* alpha);
// dst = beta*dst + a*(u + dt*M^-1 R(u)); bif likewise (fluid_flux_es_dgsem_operator.h:127-214).
void orc_forward_euler_step(void* h, double* dst, const double* u, double dt, double t, double a,
                            double beta, double* bif_dst, const double* bif_u);
// one low-storage RK stage (rk.h:53-71, tutorial-67.cc:880-899): k = M^-1 R(r_in); r_out = sol + factor_ai k; sol += factor_solution k
void orc_lsrk_stage(void* h, double* sol, double* r_out, const double* r_in, double factor_solution, double factor_ai, double t);
double orc_max_transport_speed(void* h, const double* u);    // :450-514
double orc_recommend_dt(void* h, const double* u);           // :442-448
// One SSPRK2 step (rk.h:97-106). f1/bif_f1 are scratch of the same size as u/bif.
void orc_ssprk2_step(void* h, double* u, double* f1, double dt, double t, double* bif, double* bif_f1);
// advance() to t_end with recommend_dt every step; returns number of steps.
int64_t orc_solve(void* h, double* u, double t_end, double* bif, int64_t max_steps, double fixed_dt);
void orc_global_integral(void* h, const double* u, int species, double out[5]);   // dg_solution_helper.cc:71-98

#ifdef __cplusplus
}
#endif
