/* ============================================================================
 * warpii_gpu.h -- C ABI of the B200-native ES-DGSEM operator (libwarpii_b200.so)
 *
 * This is the drop-in boundary for WarpII's hot path: everything below
 * `FluidFluxESDGSEMOperator<dim>::perform_forward_euler_step` and
 * `::recommend_dt` (reference src/five_moment/fluid_flux_es_dgsem_operator.h:62-71)
 * runs on the GPU behind these entry points.  Plain pointers and sizes only; no
 * C++ or torch types cross the boundary.  INTEGRATION.md shows the adapter a
 * WarpII maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available from warpii_gpu_last_error() (thread-local).  No exceptions
 *     cross the ABI (the reference throws dealii::ExcMessage; the adapter
 *     re-throws from the status code).
 *   - one context per rank/GPU, all calls from one host thread (the reference
 *     is single-threaded per rank: nodal_dg_discretization.cc:25-26).
 *   - state vectors live in HBM between calls and are named by small integer
 *     ids; the host only sees them through upload/download.
 *   - device state layout: u[elem][comp][node], comp = 5*species + {rho, mx,
 *     my, mz, E}, then 8 field comps if fields_enabled; nodes lexicographic
 *     (x fastest) at the Gauss-Lobatto points.  This is deal.II's FE_DGQ^nc
 *     cell-local ordering; a DoF index table passed to
 *     upload/download translates any other host numbering.
 * ==========================================================================*/
#ifndef WARPII_GPU_H
#define WARPII_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct warpii_gpu_ctx warpii_gpu_ctx;

/* Boundary-condition kinds, as Species::create_from_parameters maps them
 * (reference src/five_moment/species.cc:44-58): "Wall", "Outflow" (supersonic), "Inflow"; and the fourth kind the reference
 * operator knows but no input file can select, EulerBCMap::set_subsonic_outflow_boundary (bc_helper.h:12-35,
 * fluid_flux_es_dgsem_operator.h:385-390): ghost state = inside state with the total energy replaced by component 4 of the
 * state given through warpii_gpu_set_inflow / warpii_gpu_set_inflow_table for that boundary. */
enum { WARPII_BC_WALL = 0, WARPII_BC_OUTFLOW = 1, WARPII_BC_INFLOW = 2, WARPII_BC_SUBSONIC_OUTFLOW = 3 };

/* Flags for warpii_gpu_forward_euler_step_ex. */
enum {
    WARPII_FUSE_CFL = 1 /* also reduce the max transport speed of dst inside the stage kernel, so that the
                           next warpii_gpu_recommend_dt(dst) needs no extra sweep over the state */
};

/* Flat mesh tables, built once on the host from the triangulation
 * (replaces what dealii::MatrixFree::reinit derives in nodal_dg_discretization.cc:6-29).
 * All pointers are host pointers, copied during create. */
typedef struct warpii_gpu_mesh {
    int32_t dim;              /* 1, 2 or 3 */
    int32_t fe_degree;        /* 1..6 (five_moment.h:116) */
    int32_t n_species;        /* >= 1 */
    int32_t fields_enabled;   /* 0/1: 8 extra components [Ex,Ey,Ez,Bx,By,Bz,phi,psi] (five_moment.h:123-138) */
    double gas_gamma;
    int64_t n_elems;          /* elements owned by this rank */
    int64_t n_ghost_faces;    /* face traces received from other ranks (0 on one GPU) */
    int64_t n_boundary_faces; /* non-periodic domain-boundary faces of owned elements */
    int32_t n_boundaries;     /* number of boundary ids (n_boundaries input key) */
    double h[3];              /* Cartesian element size per dimension (HyperRectangle grid, grid_descriptions.cc:51-74) */
    /* face-pair table, [n_elems][2*dim], face f = 2*d + side (deal.II face numbering of a hypercube):
     *   0 <= v < n_elems            : owned neighbour element v (its face f^1 matches, same tangential order)
     *   n_elems <= v                : ghost trace slot v - n_elems
     *   v < 0                       : boundary face number -1 - v (index into the boundary tables below) */
    const int32_t* face_neighbor;
    const int32_t* boundary_face_elem;  /* [n_boundary_faces] owning element   */
    const int32_t* boundary_face_side;  /* [n_boundary_faces] local face index */
    const int32_t* boundary_face_id;    /* [n_boundary_faces] boundary id (< n_boundaries) */
    const int32_t* bc_kind;             /* [n_species][n_boundaries], WARPII_BC_* (bc_helper.h) */
    int32_t n_vectors;                  /* number of state vectors to allocate (>= 2: solution and f_1, rk.h:111-117) */
} warpii_gpu_mesh;

/* Halo description for an element-sharded run.  Faces are exchanged as nodal traces
 * of all fluid components: [face][5*n_species][Np^(dim-1)] doubles.  send lists are grouped by peer. */
typedef struct warpii_gpu_halo {
    int32_t n_peers;
    const int32_t* peer_rank;        /* [n_peers] */
    const int64_t* send_offset;      /* [n_peers+1] into send_elem/send_side */
    const int32_t* send_elem;        /* [n_send] owned element whose trace is sent */
    const int32_t* send_side;        /* [n_send] its local face index */
    const int64_t* recv_offset;      /* [n_peers+1] ghost slots [recv_offset[p], recv_offset[p+1]) come from peer p */
    int64_t n_interface_elems;       /* elements [0, n_interface_elems) touch a ghost face; the rest are interior
                                        and are advanced while the exchange is in flight */
} warpii_gpu_halo;

const char* warpii_gpu_last_error(void);
int warpii_gpu_abi_version(void);
/* Number of consecutive elements one thread block of the stage kernel works on.  Purely a performance hint for
 * the host's element ordering: a face neighbour inside the same group of this many consecutive elements is read
 * from the block's shared memory instead of L2/HBM, so compact patches (e.g. 4x2 in 2D p=3) should be numbered
 * consecutively.  Any ordering is correct. */
int warpii_gpu_elems_per_block(int dim, int fe_degree);
/* 1 if Cartesian contexts of this (dim, fe_degree) run the pencil-per-thread stage kernel (dgsem_pencil_kernel.cu), 0 if the
 * node-per-thread one (dgsem_stage_kernel.cu); diagnostic, used by bench.py to name the kernel in its roofline line. */
int warpii_gpu_stage_kernel_is_pencil(int dim, int fe_degree);

/* -- lifetime -------------------------------------------------------------- */
int warpii_gpu_create(const warpii_gpu_mesh* mesh, int device, warpii_gpu_ctx** out);
int warpii_gpu_destroy(warpii_gpu_ctx* ctx);
int64_t warpii_gpu_n_dofs(const warpii_gpu_ctx* ctx);            /* n_elems * n_components * Np^dim */
int warpii_gpu_synchronize(warpii_gpu_ctx* ctx);

/* -- general geometry ----------------------------------------------------------
 * Curved or non-rectangular elements and arbitrary conforming face pairing: what the reference obtains from
 * MappingQ(fe_degree) + MatrixFree (nodal_dg_discretization.h:80, nodal_dg_discretization.cc:15-28) and reads per
 * quadrature point.  The adapter copies these out of MatrixFree once; warpii_b200/host/mapped_mesh.hpp builds them
 * from the elements' Gauss-Lobatto support points when deal.II is not there.  Call once, after create and before
 * the first step; from then on every entry point runs the general-geometry kernels (warpii_gpu_mesh.h is ignored).
 * Host pointers, copied during the call. */
typedef struct warpii_gpu_geometry {
    const double* inverse_jacobian;   /* [n_elems][Np^dim][dim][dim]: J^{-T} at the GLL nodes, FEEvaluation::inverse_jacobian(q)
                                         (jacobian_utils.h:14,36; fluid_flux_es_dgsem_operator.h:476) */
    const double* face_normal;        /* [n_elems][2*dim][Np^(dim-1)][dim]: unit OUTWARD normal at the face GLL nodes,
                                         FEFaceEvaluation::normal_vector(q) of quadrature 1 (:317); the two sides of an
                                         interior face must hold exactly opposite vectors */
    const double* face_jacobian;      /* [n_elems][2*dim][Np^(dim-1)]: surface Jacobian = face JxW / tensor GLL weight */
    const int32_t* neighbor_face;     /* [n_elems][2*dim] or NULL: local face of the neighbour that matches this face, + 8 if
                                         its face nodes run in the opposite tangential order (2D).  NULL = the opposite
                                         face, same order (structured meshes) */
    const double* boundary_normal;    /* [n_boundary_faces][(fe_degree+2)^(dim-1)][dim]: unit outward normal at the
                                         Gauss(p+2) points of the boundary faces (quadrature 0, :355-361) */
    const double* boundary_jacobian;  /* [n_boundary_faces][(fe_degree+2)^(dim-1)]: surface Jacobian there */
} warpii_gpu_geometry;
int warpii_gpu_set_geometry(warpii_gpu_ctx* ctx, const warpii_gpu_geometry* geometry);

/* -- state transfer (FiveMSolutionVec::mesh_sol <-> HBM; solution_vec.h:45-51) ----------------------
 * dof_index == NULL: host array is already in device layout.  Otherwise host[dof_index[i]] <-> device[i]
 * (build it once from cell->get_dof_indices()). */
int warpii_gpu_upload_state(warpii_gpu_ctx* ctx, int vec, const double* host, const int64_t* dof_index);
int warpii_gpu_download_state(warpii_gpu_ctx* ctx, int vec, double* host, const int64_t* dof_index);
int warpii_gpu_zero_state(warpii_gpu_ctx* ctx, int vec);
int warpii_gpu_copy_state(warpii_gpu_ctx* ctx, int dst, int src);
/* raw device pointer of a state vector (for zero-copy producers/consumers such as a CUDA-aware writer) */
int warpii_gpu_device_ptr(warpii_gpu_ctx* ctx, int vec, void** out);

/* -- boundary data ----------------------------------------------------------------------------------
 * Conserved inflow state of (species, boundary id): what EulerBCMap::get_inflow evaluates
 * (fluid_flux_es_dgsem_operator.h:377-380).  Call again before a stage if the function depends on t
 * (the reference calls set_time(t) at :140-144). */
int warpii_gpu_set_inflow(warpii_gpu_ctx* ctx, int species, int boundary_id, const double q[5]);
/* Space-dependent inflow: the Function<dim> of EulerBCMap::get_inflow tabulated where the reference evaluates it,
 * phi.quadrature_point(q) of every boundary face (fluid_flux_es_dgsem_operator.h:381-384):
 * table[n_faces][points_per_face][5] conserved values for one species; faces in the order of
 * warpii_gpu_mesh.boundary_face_*, points_per_face = (fe_degree+2)^(dim-1) Gauss points, tensor-ordered with the
 * lower tangential dimension fastest.  Rows of faces whose condition is not Inflow are ignored.  Upload again before a
 * stage when the function depends on t.  After the first table, warpii_gpu_set_inflow keeps working (it overwrites
 * the rows of its boundary). */
int warpii_gpu_n_boundary_points(const warpii_gpu_ctx* ctx, int64_t* n_faces_out, int* points_per_face_out);
int warpii_gpu_set_inflow_table(warpii_gpu_ctx* ctx, int species, const double* table);

/* -- two-fluid source terms (north_star kernel 4) ------------------------------------------------------
 * NOT part of the reference operator, which carries the 8 field components [Ex,Ey,Ez,Bx,By,Bz,phi,psi]
 * (five_moment.h:131-137) through unchanged; off by default so that the drop-in stays one.  When enabled (requires
 * fields_enabled), every stage adds, per node: (q/m)_s (rho_s E + m_s x B) to the momentum and (q/m)_s m_s.E to the
 * energy of species s, -J/epsilon0 to E and chi rho_c/epsilon0 to phi, with J = sum_s (q/m)_s m_s and
 * rho_c = sum_s (q/m)_s rho_s (Species::charge / Species::mass, species.h:43-44).  The curl and divergence-cleaning
 * fluxes of Maxwell's equations are not evolved.  recommend_dt does not know about the plasma frequency. */
int warpii_gpu_set_sources(warpii_gpu_ctx* ctx, int enabled, double epsilon0, double chi, const double* charge_over_mass);
/* Perfectly hyperbolic Maxwell (PHM) fluxes for the 8 field components (north_star kernel 4, BASELINE config 5).  NOT in
 * the reference, which only allocates [Ex,Ey,Ez,Bx,By,Bz,phi,psi] (five_moment.h:123-138); the system is the one
 * SURVEY.md 8(c) names:  dE/dt - c^2 curl B + chi c^2 grad phi = -J/eps0,  dB/dt + curl E + gamma grad psi = 0,
 * dphi/dt + chi div E = chi rho_c/eps0,  dpsi/dt + gamma c^2 div B = 0  (sources: warpii_gpu_set_sources), discretised as a
 * linear DGSEM flux on the fluid's nodes with a Rusanov numerical flux.
 * With it recommend_dt also covers the wave speed c max(1, chi, gamma) and, with the sources on, the plasma and cyclotron
 * frequencies (omega dt <= 0.1); halo exchanges carry the field traces too.  Non-periodic boundaries are zero-gradient for
 * the fields.  Cartesian boxes only (not after warpii_gpu_set_geometry).  Off by default. */
int warpii_gpu_set_maxwell(warpii_gpu_ctx* ctx, int enabled, double light_speed, double chi, double gamma);

/* -- the operator ----------------------------------------------------------------------------------
 * dst = beta*dst + alpha*(u + dt * M^-1 R(u)), and the same for the boundary-integrated fluxes
 * (replaces perform_forward_euler_step, fluid_flux_es_dgsem_operator.h:127-214). dst != u. */
int warpii_gpu_forward_euler_step(warpii_gpu_ctx* ctx, int dst, int u, double dt, double t, double alpha,
                                  double beta);
int warpii_gpu_forward_euler_step_ex(warpii_gpu_ctx* ctx, int dst, int u, double dt, double t, double alpha,
                                     double beta, int flags);
/* dt = 0.5 / (vmax * (p+1)^2), vmax reduced over nodes, elements, species and ranks
 * (replaces recommend_dt + compute_cell_transport_speed, :442-514; MPI::max -> ncclAllReduce(max)). */
int warpii_gpu_recommend_dt(warpii_gpu_ctx* ctx, int vec, double* dt_out);
/* One stage of a low-storage Runge-Kutta scheme: what LowStorageRungeKuttaIntegrator::perform_time_step asks of
 * pde_operator.perform_stage(t, b_i*dt, a_i*dt, current_ri, vec_ki, solution, next_ri) (rk.h:53-71, deal.II step-67):
 *   k = M^-1 R(r_in);   sol_out = sol_in + factor_solution * k;   r_out = sol_in + factor_ai * k
 * (tutorial-67.cc:880-899: both from the OLD solution; r_out is left untouched when factor_ai == 0, the last stage)
 * fused into one launch (k is never stored).  r_in is read at neighbouring nodes while the stage writes, so sol_out and
 * r_out must differ from r_in; sol_out may be sol_in (in place) when sol_in != r_in.  Three vectors suffice for a step:
 * first stage (S, S) -> (K, R), later stages alternate r between R and S with the solution in place in K
 * (warpii_b200/host/gpu_operator.hpp::LowStorageRungeKuttaIntegrator does the bookkeeping).  Needs n_vectors >= 3. */
int warpii_gpu_lsrk_stage(warpii_gpu_ctx* ctx, int sol_out, int r_out, int sol_in, int r_in, double factor_solution,
                          double factor_ai, double t);
int warpii_gpu_max_transport_speed(warpii_gpu_ctx* ctx, int vec, double* vmax_out);
/* One SSPRK2 step (replaces SSPRK2Integrator::evolve_one_time_step, rk.h:97-106): two stages, the second with
 * the fused CFL reduction. */
int warpii_gpu_ssprk2_step(warpii_gpu_ctx* ctx, int solution, int f1, double dt, double t);
/* The same step for a state that lives in HOST memory (pinned, device layout): host_in -> HBM -> two stages -> host_out,
 * with the transfers and the stages overlapped slab by slab on three streams (upload, compute, download; PCIe is full
 * duplex and the stage kernels take element ranges), instead of upload, step, download one after the other.  dt comes from
 * the caller: warpii_gpu_recommend_dt for the first step, then *next_dt_out of the previous call (the fused CFL reduction of
 * the second stage, i.e. recommend_dt of the state just returned).  host_out may be host_in.  n_slabs = 0 picks 32 (measured: 16 and 32 tie on a 168 MB state, 32 wins on a 2.4 GB one).  Bit-identical
 * to upload + warpii_gpu_ssprk2_step + download.  Contexts with a communicator stream the interior slabs the same way and
 * keep the interface elements and the halo exchanges on the communication stream; contexts with boundary faces run the
 * plain sequence. */
int warpii_gpu_host_ssprk2_step(warpii_gpu_ctx* ctx, int solution, int f1, const double* host_in, double* host_out, double dt,
                                double t, double* next_dt_out, int n_slabs);
/* Time loop resident on the device side of the ABI: repeats {dt = min(recommend_dt, t_stop - t); ssprk2} until
 * t >= t_stop - 1e-12 (the inner loop of advance(), timestepper.cc:34-42).  fixed_dt > 0 overrides recommend_dt.
 * max_steps > 0 bounds the number of steps.  *t_inout is advanced; *steps_out receives the step count. */
int warpii_gpu_advance_to(warpii_gpu_ctx* ctx, int solution, int f1, double* t_inout, double t_stop,
                          double fixed_dt, int64_t max_steps, int64_t* steps_out);

/* -- diagnostics ------------------------------------------------------------------------------------ */
/* boundary_integrated_fluxes of a vector, 5*n_boundaries doubles (solution_vec.h:10-43) */
int warpii_gpu_boundary_fluxes(warpii_gpu_ctx* ctx, int vec, double* out);
int warpii_gpu_set_boundary_fluxes(warpii_gpu_ctx* ctx, int vec, const double* in);
/* sum_q u_q JxW_q per component of one species (compute_global_integral, dg_solution_helper.cc:71-98);
 * summed over ranks when a communicator is attached. */
int warpii_gpu_global_integral(warpii_gpu_ctx* ctx, int vec, int species, double out[5]);
/* blending factor per owned element and species, alpha[elem][species] (persson_peraire_shock_indicator.h:44-123) */
int warpii_gpu_shock_indicator(warpii_gpu_ctx* ctx, int vec, double* alpha_out);
/* bare M^-1 R(u) into vector dst (dst != u), for parity checks of one RHS evaluation */
int warpii_gpu_rhs(warpii_gpu_ctx* ctx, int dst, int u, double t);

/* -- multi-GPU (one process per GPU; NCCL over NVLink) ----------------------------------------------- */
#define WARPII_GPU_NCCL_ID_BYTES 128
int warpii_gpu_nccl_unique_id(char id[WARPII_GPU_NCCL_ID_BYTES]);
int warpii_gpu_attach_comm(warpii_gpu_ctx* ctx, const char id[WARPII_GPU_NCCL_ID_BYTES], int rank, int n_ranks,
                           const warpii_gpu_halo* halo);

/* -- measurement hooks -------------------------------------------------------------------------------- */
/* kernels launched by this context so far */
int64_t warpii_gpu_launch_count(const warpii_gpu_ctx* ctx);
/* device time in ms of the stage kernels launched since the last reset (CUDA events on the launching stream),
 * and how many there were; used by bench.py for the roofline line. */
int warpii_gpu_stage_timing(warpii_gpu_ctx* ctx, int enable, double* ms_total, int64_t* n_launches);
/* While stage timing is enabled, warpii_gpu_advance_to also measures the SM clock ON the device after every batch of
 * steps (one thread compares clock64 with the global timer for 40 us).  Returns the MHz readings collected since the
 * last call.  (NVML/nvidia-smi queries inside a timed region stall NCCL runs for milliseconds.) */
int warpii_gpu_sm_clock_probes(warpii_gpu_ctx* ctx, double* mhz_out, int max_out, int* n_out);
/* the CUDA stream (cudaStream_t) all work of this context is issued on */
int warpii_gpu_stream(warpii_gpu_ctx* ctx, void** stream_out);
/* Measured FP64 throughput of this GPU's CUDA cores, in fused multiply-adds per second (FMA-chain microbenchmark, all SMs
 * full): the denominator of the FP64 figure bench.py reports beside the HBM roofline. */
int warpii_gpu_measure_fp64_peak(int device, double* fma_per_second_out);
/* Tuning diagnostic: the same FMA rate with ONE block of threads_per_sm threads per SM and `ilp` (1,2,3,4,6,8) independent
 * chains per thread: what a kernel with that many resident warps and that much instruction-level parallelism can reach. */
int warpii_gpu_fp64_rate_probe(int device, int threads_per_sm, int ilp, double* fma_per_second_out);

/* -- point physics on the device, for known-answer tests ------------------------------------------------
 * For n state pairs (qa[i], qb[i]) of 5 conserved values: the direction-d entropy-conserving flux
 * (euler_CH_EC_flux, euler.h:186-228) and the entropy-dissipating flux across a face with normal +e_d
 * (euler_CH_entropy_dissipating_flux, :232-284), evaluated by the same device functions the stage kernel uses.
 * prim_out (nullable) receives the 12 per-node quantities of qa[i]: rho,u0,u1,u2,beta,log rho,log beta,|u|^2,p,E+p,
 * |u|+c, 1/beta.  Needs no context. */
int warpii_gpu_point_fluxes(int device, int n, const double* qa, const double* qb, int d, double gamma,
                            double* ec_out, double* es_out, double* prim_out);

/* Self-check of the kernels' branch-free correctly rounded division (csrc/det_log.cuh, div_rn_fast) against the IEEE
 * division on n pseudo-random operand pairs: mode 0 general operands, 1 the f/(2+f) of the logarithm, 2 quotients
 * close to 1, 3 short numerators (exact quotients and ties).  *mismatches_out counts results that differ in any bit;
 * first_bad_out (nullable) = {a, b, got, want} of one of them.  The bit-for-bit parity of beta and of the logarithm
 * with the CPU path rests on this being 0.  Needs no context. */
int warpii_gpu_check_division(int device, int64_t n, uint64_t seed, int mode, int64_t* mismatches_out,
                              double first_bad_out[4]);

#ifdef __cplusplus
}
#endif
#endif /* WARPII_GPU_H */
