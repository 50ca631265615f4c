// C ABI of libwarpii_b200.so (see include/warpii_gpu.h): context, HBM-resident state vectors, the operator
// entry points, NCCL halo exchange / reductions.  All device work of a context is issued on one stream;
// the halo exchange runs on a second stream and is ordered with events.
#include "../../include/warpii_gpu.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; the library is resolved with dlopen so single-GPU use needs no NCCL

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/reference_element.hpp"
#include "dgsem_kernels.cuh"

using namespace wgpu;

namespace {

thread_local std::string g_last_error;

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
}

#define CUDA_OK(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL through dlopen ----------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return 0;
    // If torch (or anything else) already mapped an NCCL, reuse it; otherwise take the system one.
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                        \
    if (!g_nccl.field) return fail("libnccl lacks symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return 0;
}

#define NCCL_OK(expr)                                                                                           \
    do {                                                                                                        \
        ncclResult_t _r = (expr);                                                                               \
        if (_r != ncclSuccess) return fail("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
    } while (0)

int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

// Which stage kernel a Cartesian context runs: the pencil-per-thread kernel where it exists (2-D / 3-D, Np = 3..5), the
// node-per-thread kernel otherwise.  WARPII_GPU_STAGE=node forces the latter (A/B measurements, profiles/README.md); the
// variable is read once per process because the element numbering (patch size) depends on the choice.
bool use_pencil(int dim, int Np) {
    static const bool forced_node = [] {
        const char* env = std::getenv("WARPII_GPU_STAGE");
        return env && std::strcmp(env, "node") == 0;
    }();
    return !forced_node && pencil_available(dim, Np);
}
int patch_elems(int dim, int Np) { return use_pencil(dim, Np) ? pencil_patch_elems(dim, Np) : elems_per_block(dim, Np); }

}  // namespace

void warpii_gpu_free_slab_plan(void* plan);

struct warpii_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_recv = nullptr;
    int dim = 1, p = 1, Np = 2, NN = 2, NF = 1, nsp = 1, nc = 5, n_boundaries = 0, n_vectors = 2;
    double gamma = 5.0 / 3.0;
    int64_t n_elems = 0, n_ghost = 0, n_bfaces = 0, n_dofs = 0;
    double h[3] = {1, 1, 1}, inv_h[3] = {1, 1, 1}, inv_hw[3] = {1, 1, 1}, Jdet = 1, max_eig = 1;
    ElemTables T;
    BoundaryParams B;
    std::vector<int32_t> h_bc_kind, h_bf_id;
    // CUDA graph of one full batch of the device-resident time loop (single GPU, timing off): the launch sequence and all
    // kernel arguments are the same for every batch because dt, t and the stop flag live in device memory
    cudaGraphExec_t batch_graph = nullptr;
    int batch_graph_solution = -1, batch_graph_f1 = -1;
    int64_t batch_graph_launches = 0;
    bool use_graphs = true;
    bool pencil = false;                    // Cartesian stage launches go to pencil_stage_kernel (dgsem_pencil_kernel.cu)
    bool src_on = false;                    // two-fluid source terms (warpii_gpu_set_sources)
    double inv_eps0 = 1.0, chi = 0.0;
    bool maxwell_on = false;                // PHM fluxes for the field components (warpii_gpu_set_maxwell)
    bool mx_fused = false;                  // ... evolved inside the pencil stage kernel (else by maxwell_kernel right behind it)
    double light_speed = 1.0, mx_chi = 0.0, mx_gamma = 0.0;
    int ncf = 5;                            // components per halo face trace: 5*nsp, or nc with the field system evolved
    double* d_qm = nullptr;                 // [nsp] charge / mass
    double* d_scratch = nullptr;            // state-sized throw-away destination of warpii_gpu_shock_indicator (lazy)
    std::vector<double> h_inflow;           // mirror of d_inflow
    std::vector<double> h_inflow_table;     // mirror of d_inflow_table (empty until warpii_gpu_set_inflow_table)
    double* d_inflow_table = nullptr;
    int32_t *d_nbr = nullptr, *d_bf_elem = nullptr, *d_bf_side = nullptr, *d_bf_id = nullptr, *d_bc_kind = nullptr;
    double *d_inflow = nullptr, *d_w = nullptr, *d_bres = nullptr, *d_bflux = nullptr, *d_ghost = nullptr,
           *d_sendbuf = nullptr, *d_partial = nullptr, *d_out5 = nullptr, *d_alpha = nullptr;
    std::vector<double*> vec;
    std::vector<double*> bif;
    unsigned long long* d_vmax = nullptr;   // one slot per vector
    double ind_T = 0, ind_sT = 0;
    std::vector<char> vmax_valid;
    double* h_pin = nullptr;                // pinned staging, n_dofs doubles (lazy)
    double* h_small = nullptr;              // pinned, 64 doubles
    DevClock* d_clock = nullptr;            // device-resident time loop state
    DevClock* h_clock = nullptr;            // pinned mirror
    double* d_probe = nullptr;              // SM clock probes (MHz), ring of 64
    int n_probes = 0;
    // general geometry (warpii_gpu_set_geometry)
    bool general = false;
    GeneralParams GP{};
    double *d_gnode = nullptr, *d_gsub = nullptr, *d_gface = nullptr, *d_jdet = nullptr, *d_bgeo = nullptr, *d_bmass = nullptr;
    int2* d_nbr2 = nullptr;
    std::vector<int32_t> h_bf_elem, h_bf_side;
    // streamed host step (warpii_gpu_host_ssprk2_step)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> slab_events;
    void* slab_plan = nullptr;
    int slab_plan_slabs = -1;
    std::vector<std::vector<int>> shard_deps;   // neighbour slabs of the sharded streamed step (host_step_sharded)
    // the whole streamed step as a CUDA graph, replayed while the caller keeps passing the same buffers and vectors
    cudaGraphExec_t host_step_graph = nullptr;
    const double* hsg_in = nullptr;
    double* hsg_out = nullptr;
    int hsg_solution = -1, hsg_f1 = -1, hsg_slabs = -1;
    double* d_host_dt = nullptr;            // dt of the streamed step (a kernel argument by pointer, so the graph is reusable)
    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    std::vector<int32_t> peer_rank;
    std::vector<int64_t> send_offset, recv_offset;
    int32_t *d_send_elem = nullptr, *d_send_side = nullptr;
    int64_t n_send = 0, n_interface = 0;
    // measurement
    int64_t launches = 0;
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

namespace {

template <typename T>
int upload(T** dptr, const T* host, size_t n) {
    *dptr = nullptr;
    if (n == 0) return 0;
    CUDA_OK(cudaMalloc((void**)dptr, n * sizeof(T)));
    if (host) CUDA_OK(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
    else CUDA_OK(cudaMemset(*dptr, 0, n * sizeof(T)));
    return 0;
}

int check_vec(warpii_gpu_ctx* c, int v, const char* what) {
    if (!c) return fail("%s: null context", what);
    if (v < 0 || v >= (int)c->vec.size()) return fail("%s: vector id %d out of range [0,%d)", what, v, (int)c->vec.size());
    if (!c->vec[v] && c->n_dofs > 0) {
        // vectors beyond the first two (solution, f_1) take HBM only once somebody names them (low-storage RK registers)
        CUDA_OK(cudaSetDevice(c->device));
        if (upload<double>(&c->vec[v], nullptr, (size_t)c->n_dofs)) return 1;
        if (upload<double>(&c->bif[v], nullptr, (size_t)5 * (c->n_boundaries > 0 ? c->n_boundaries : 1))) return 1;
    }
    return 0;
}

int get_events(warpii_gpu_ctx* c, cudaEvent_t* a, cudaEvent_t* b) {
    while (c->ev_pool.size() < c->ev_used + 2) {
        cudaEvent_t e;
        CUDA_OK(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    *a = c->ev_pool[c->ev_used++];
    *b = c->ev_pool[c->ev_used++];
    return 0;
}

// the Cartesian kernels or their general-geometry counterparts
void do_launch_stage(warpii_gpu_ctx* c, const StageParams& P, cudaStream_t s) {
    if (c->general) launch_stage_general(c->dim, c->Np, P, c->GP, s);
    else if (c->pencil) launch_pencil_stage(c->dim, c->Np, P, s);
    else launch_stage(c->dim, c->Np, P, s);
    if (c->maxwell_on && !c->mx_fused && P.elem_end > P.elem_begin) {
        // the field components of the same range, right behind the fluids (the pencil kernel evolves them itself)
        launch_maxwell(c->dim, c->Np, P, c->light_speed, c->mx_chi, c->mx_gamma, s);
        c->launches++;
    }
}
void do_launch_boundary(warpii_gpu_ctx* c, const BoundaryParams& B, cudaStream_t s) {
    if (c->general) launch_boundary_general(c->dim, c->Np, B, c->GP, s);
    else launch_boundary(c->dim, c->Np, B, s);
}
void do_launch_cfl(warpii_gpu_ctx* c, int vec) {
    if (c->general) launch_cfl_general(c->dim, c->Np, c->vec[vec], c->n_elems, c->nc, c->nsp, c->gamma, c->GP, c->d_vmax + vec, c->stream);
    else launch_cfl(c->dim, c->Np, c->vec[vec], c->n_elems, c->nc, c->nsp, c->gamma, c->inv_h, c->max_eig, c->d_vmax + vec, c->stream);
    if (c->maxwell_on) {
        launch_maxwell_cfl(c->dim, c->Np, c->vec[vec], c->n_elems, c->nc, c->nsp, c->d_qm, c->light_speed, c->mx_chi, c->mx_gamma,
                           c->inv_eps0, c->max_eig, c->src_on, c->d_vmax + vec, c->stream);
        c->launches++;
    }
}

StageParams stage_params(warpii_gpu_ctx* c, int dst, int u, double dt, double a, double beta, int mode, bool fuse_cfl) {
    StageParams P;
    P.lookahead = 0;   // (the pencil launcher sets it)
    P.u = c->vec[u];
    P.dst = c->vec[dst];
    P.nbr = c->d_nbr;
    P.ind_T = c->ind_T;
    P.ind_sT = c->ind_sT;
    P.ghost = c->d_ghost;
    P.bres = c->d_bres;
    P.alpha_out = nullptr;
    P.sol_in = nullptr;
    P.dst2 = nullptr;
    P.vmax = fuse_cfl ? c->d_vmax + dst : nullptr;
    P.elem_begin = 0;
    P.elem_end = c->n_elems;
    P.n_elems = c->n_elems;
    P.nc = c->nc;
    P.nsp = c->nsp;
    P.ncf = c->ncf;
    P.fields_skip = c->maxwell_on ? 1 : 0;
    P.mx_on = (c->maxwell_on && c->mx_fused) ? 1 : 0;
    {
        double big = 1.0;
        if (c->mx_chi > big) big = c->mx_chi;
        if (c->mx_gamma > big) big = c->mx_gamma;
        P.mx_c2 = c->light_speed * c->light_speed;
        P.mx_chi = c->mx_chi;
        P.mx_gam = c->mx_gamma;
        P.mx_lam = c->light_speed * big;
        P.mx_floor = c->max_eig * P.mx_lam;
        P.mx_omega_factor = 5.0 / (double)(c->Np * c->Np);
    }
    P.mode = mode;
    P.dt_dev = nullptr;
    P.skip_dev = nullptr;
    P.gamma = c->gamma;
    P.hig = 0.5 / (c->gamma - 1.0);
    P.dt = dt;
    P.a = a;
    P.beta = beta;
    for (int d = 0; d < 3; d++) { P.inv_h[d] = c->inv_h[d]; P.inv_hw[d] = c->inv_hw[d]; }
    P.max_eig = c->max_eig;
    P.src_on = c->src_on ? 1 : 0;
    P.inv_eps0 = c->inv_eps0;
    P.chi = c->chi;
    P.qm = c->d_qm;
    P.T = c->T;
    return P;
}

// halo exchange of the face traces of vector u: pack on the main stream, send/recv on the comm stream
int start_exchange(warpii_gpu_ctx* c, int u) {
    if (!c->comm || c->peer_rank.empty()) return 0;
    launch_pack(c->dim, c->Np, c->vec[u], c->d_send_elem, c->d_send_side, c->n_send, c->nc, c->ncf, c->d_sendbuf, c->stream);
    c->launches++;
    CUDA_OK(cudaEventRecord(c->ev_pack, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
    const size_t per_face = (size_t)c->ncf * c->NF;
    NCCL_OK(g_nccl.GroupStart());
    for (size_t p = 0; p < c->peer_rank.size(); p++) {
        const int64_t ns = c->send_offset[p + 1] - c->send_offset[p];
        const int64_t nr = c->recv_offset[p + 1] - c->recv_offset[p];
        if (ns > 0)
            NCCL_OK(g_nccl.Send(c->d_sendbuf + c->send_offset[p] * per_face, ns * per_face, ncclDouble, c->peer_rank[p], c->comm, c->comm_stream));
        if (nr > 0)
            NCCL_OK(g_nccl.Recv(c->d_ghost + c->recv_offset[p] * per_face, nr * per_face, ncclDouble, c->peer_rank[p], c->comm, c->comm_stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    CUDA_OK(cudaEventRecord(c->ev_recv, c->comm_stream));
    return 0;
}

// the same exchange issued entirely on the communication stream (pack kernel included): the caller has ordered that stream
// behind whatever produced the interface elements of u
int exchange_on_comm_stream(warpii_gpu_ctx* c, int u) {
    if (!c->comm || c->peer_rank.empty()) return 0;
    launch_pack(c->dim, c->Np, c->vec[u], c->d_send_elem, c->d_send_side, c->n_send, c->nc, c->ncf, c->d_sendbuf, c->comm_stream);
    c->launches++;
    const size_t per_face = (size_t)c->ncf * c->NF;
    NCCL_OK(g_nccl.GroupStart());
    for (size_t p = 0; p < c->peer_rank.size(); p++) {
        const int64_t ns = c->send_offset[p + 1] - c->send_offset[p];
        const int64_t nr = c->recv_offset[p + 1] - c->recv_offset[p];
        if (ns > 0)
            NCCL_OK(g_nccl.Send(c->d_sendbuf + c->send_offset[p] * per_face, ns * per_face, ncclDouble, c->peer_rank[p], c->comm, c->comm_stream));
        if (nr > 0)
            NCCL_OK(g_nccl.Recv(c->d_ghost + c->recv_offset[p] * per_face, nr * per_face, ncclDouble, c->peer_rank[p], c->comm, c->comm_stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    return 0;
}

int run_stage(warpii_gpu_ctx* c, int dst, int u, double dt, double a, double beta, int mode, bool fuse_cfl,
              bool device_clock = false, int sol_in = -1, int dst2 = -1) {
    const double* dt_dev = device_clock ? &c->d_clock->dt : nullptr;
    const int* skip_dev = device_clock ? &c->d_clock->done : nullptr;
    // boundary faces first: their contributions are consumed by the stage kernel
    if (c->n_bfaces > 0) {
        BoundaryParams B = c->B;
        B.u = c->vec[u];
        do_launch_boundary(c, B, c->stream);
        c->launches++;
    }
    if (c->n_boundaries > 0) {
        // (a low-storage stage leaves the flux RATE in bif[dst]: the reference's perform_stage callers keep no such totals)
        launch_bif_update(c->d_bflux, c->d_bf_id, c->n_bfaces, c->nsp, c->n_boundaries, c->bif[dst], c->bif[u], dt, dt_dev,
                          skip_dev, a, beta, mode == 2 ? 1 : mode, c->stream);
        c->launches++;
    }
    // (with the device clock, clock_kernel clears the slot after it has read the previous maximum)
    if (fuse_cfl && !device_clock) CUDA_OK(cudaMemsetAsync(c->d_vmax + dst, 0, sizeof(unsigned long long), c->stream));
    StageParams P = stage_params(c, dst, u, dt, a, beta, mode, fuse_cfl);
    P.dt_dev = dt_dev;
    P.skip_dev = skip_dev;
    if (mode == 2) {
        P.sol_in = c->vec[sol_in];
        P.dst2 = c->vec[dst2];
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        if (get_events(c, &e0, &e1)) return 1;
        CUDA_OK(cudaEventRecord(e0, c->stream));
    }
    const bool halo = c->comm && !c->peer_rank.empty();
    if (halo) {
        if (start_exchange(c, u)) return 1;
        // interior elements on the compute stream while the traces are in flight; the (few) interface elements follow the
        // receive on the high-priority communication stream, so they slot in between the interior kernel's blocks
        // instead of waiting for its tail.  The two launches write disjoint elements of dst.
        StageParams Pi = P;
        Pi.elem_begin = 0;
        Pi.elem_end = c->n_interface;
        do_launch_stage(c, Pi, c->comm_stream);
        if (Pi.elem_end > Pi.elem_begin) c->launches++;
        CUDA_OK(cudaEventRecord(c->ev_recv, c->comm_stream));
        P.elem_begin = c->n_interface;
        P.elem_end = c->n_elems;
        do_launch_stage(c, P, c->stream);
        if (P.elem_end > P.elem_begin) c->launches++;
        CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_recv, 0));
    } else {
        do_launch_stage(c, P, c->stream);
        c->launches++;
    }
    if (c->timing) CUDA_OK(cudaEventRecord(e1, c->stream));
    CUDA_OK(cudaGetLastError());
    c->vmax_valid[dst] = fuse_cfl ? 1 : 0;
    return 0;
}

int max_speed(warpii_gpu_ctx* c, int vec, double* out) {
    if (!c->vmax_valid[vec]) {
        CUDA_OK(cudaMemsetAsync(c->d_vmax + vec, 0, sizeof(unsigned long long), c->stream));
        do_launch_cfl(c, vec);
        c->launches++;
        c->vmax_valid[vec] = 1;
    }
    if (c->comm && c->n_ranks > 1) {
        // The slot holds the bits of a non-negative double (or of a NaN: an unphysical state must surface as a NaN time step
        // on EVERY rank).  Reduced as an unsigned integer: the order of the bit patterns is the order of the doubles, and a
        // NaN (0x7ff8...) beats every finite maximum, which ncclMax on doubles does not promise.  (Utilities::MPI::max, :511)
        NCCL_OK(g_nccl.AllReduce(c->d_vmax + vec, c->d_vmax + vec, 1, ncclUint64, ncclMax, c->comm, c->stream));
    }
    CUDA_OK(cudaMemcpyAsync(c->h_small, c->d_vmax + vec, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *out = c->h_small[0];
    return 0;
}

}  // namespace

extern "C" {

const char* warpii_gpu_last_error(void) { return g_last_error.c_str(); }
int warpii_gpu_abi_version(void) { return 1; }
int warpii_gpu_elems_per_block(int dim, int fe_degree) {
    if (dim < 1 || dim > 3 || fe_degree < 1 || fe_degree > 6) return 1;
    return patch_elems(dim, fe_degree + 1);
}

int warpii_gpu_stage_kernel_is_pencil(int dim, int fe_degree) {
    if (dim < 1 || dim > 3 || fe_degree < 1 || fe_degree > 6) return 0;
    return use_pencil(dim, fe_degree + 1) ? 1 : 0;
}

int warpii_gpu_create(const warpii_gpu_mesh* m, int device, warpii_gpu_ctx** out) {
    if (!m || !out) return fail("warpii_gpu_create: null argument");
    *out = nullptr;
    if (m->dim < 1 || m->dim > 3) return fail("n_dims must be 1, 2, or 3");   // five_moment.cc:48-50
    if (m->fe_degree < 1 || m->fe_degree > 6) return fail("fe_degree must be in [1,6]");
    if (m->n_species < 1) return fail("n_species must be >= 1");
    if (m->n_elems < 0 || m->n_elems > 2000000000LL) return fail("n_elems out of range");
    if (m->n_elems > 0 && !m->face_neighbor) return fail("face_neighbor table missing");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail("no CUDA device available: libwarpii_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail("device %d out of range (have %d)", device, n_dev);
    CUDA_OK(cudaSetDevice(device));

    warpii_gpu_ctx* c = new warpii_gpu_ctx();
    c->device = device;
    c->dim = m->dim;
    c->p = m->fe_degree;
    c->Np = m->fe_degree + 1;
    c->NN = ipow(c->Np, c->dim);
    c->NF = ipow(c->Np, c->dim - 1);
    c->nsp = m->n_species;
    c->nc = 5 * m->n_species + (m->fields_enabled ? 8 : 0);
    c->gamma = m->gas_gamma;
    c->n_elems = m->n_elems;
    c->n_ghost = m->n_ghost_faces;
    c->n_bfaces = m->n_boundary_faces;
    c->n_boundaries = m->n_boundaries;
    c->n_vectors = m->n_vectors < 2 ? 2 : m->n_vectors;
    c->n_dofs = c->n_elems * c->nc * c->NN;
    c->ncf = 5 * c->nsp;

    warpii_b200::ReferenceElement re(c->p);
    std::memset(&c->T, 0, sizeof c->T);
    for (int i = 0; i < c->Np * c->Np; i++) { c->T.D[i] = re.D[i]; c->T.V[i] = re.V[i]; }
    for (int i = 0; i < c->Np; i++) c->T.w[i] = re.w[i];

    double det = 1.0;
    for (int d = 0; d < c->dim; d++) {
        if (!(m->h[d] > 0)) { delete c; return fail("element size h[%d] must be positive", d); }
        c->h[d] = m->h[d];
        c->inv_h[d] = 1.0 / m->h[d];
        c->inv_hw[d] = 1.0 / (m->h[d] * re.w[0]);
        det *= c->inv_h[d];
    }
    c->Jdet = 1.0 / det;
    {   // largest singular value of J^-T by the reference's 5-step power iteration (:487-502)
        double ev[3] = {1, 1, 1};
        for (int it = 0; it < 5; it++) {
            double nrm = 0;
            for (int d = 0; d < c->dim; d++) { ev[d] = c->inv_h[d] * (c->inv_h[d] * ev[d]); nrm = std::fmax(nrm, std::fabs(ev[d])); }
            for (int d = 0; d < c->dim; d++) ev[d] /= nrm;
        }
        double num = 0, den = 0;
        for (int d = 0; d < c->dim; d++) { const double jv = c->inv_h[d] * ev[d]; num += jv * jv; den += ev[d] * ev[d]; }
        c->max_eig = std::sqrt(num / den);
    }

    // validate the tables before they reach the device
    const int nf = 2 * c->dim;
    for (int64_t i = 0; i < c->n_elems * nf; i++) {
        const int64_t v = m->face_neighbor[i];
        if (v >= c->n_elems + c->n_ghost || v < -c->n_bfaces) {
            delete c;
            return fail("face_neighbor[%lld] = %lld is outside the element/ghost/boundary ranges", (long long)i, (long long)v);
        }
    }
    c->h_bc_kind.assign((size_t)c->nsp * (c->n_boundaries > 0 ? c->n_boundaries : 1), WARPII_BC_WALL);
    if (c->n_bfaces > 0) {
        if (!m->boundary_face_elem || !m->boundary_face_side || !m->boundary_face_id || !m->bc_kind) {
            delete c;
            return fail("boundary tables missing");
        }
        for (int64_t b = 0; b < c->n_bfaces; b++) {
            if (m->boundary_face_id[b] < 0 || m->boundary_face_id[b] >= c->n_boundaries) {
                delete c;
                // same condition the reference reports at fluid_flux_es_dgsem_operator.h:406-411
                return fail("Unknown boundary id, did you set a boundary condition for this part of the domain boundary?");
            }
        }
    }
    if (m->bc_kind)
        for (size_t i = 0; i < (size_t)c->nsp * c->n_boundaries; i++) c->h_bc_kind[i] = m->bc_kind[i];

    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_recv, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return fail("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

    int rc = 0;
    rc |= upload(&c->d_nbr, m->face_neighbor, (size_t)c->n_elems * nf);
    rc |= upload(&c->d_bf_elem, m->boundary_face_elem, (size_t)c->n_bfaces);
    rc |= upload(&c->d_bf_side, m->boundary_face_side, (size_t)c->n_bfaces);
    rc |= upload(&c->d_bf_id, m->boundary_face_id, (size_t)c->n_bfaces);
    rc |= upload(&c->d_bc_kind, c->h_bc_kind.data(), c->h_bc_kind.size());
    rc |= upload<double>(&c->d_inflow, nullptr, (size_t)c->nsp * (c->n_boundaries > 0 ? c->n_boundaries : 1) * 5);
    c->h_inflow.assign((size_t)c->nsp * (c->n_boundaries > 0 ? c->n_boundaries : 1) * 5, 0.0);
    if (c->n_bfaces > 0) {
        c->h_bf_id.assign(m->boundary_face_id, m->boundary_face_id + c->n_bfaces);
        c->h_bf_elem.assign(m->boundary_face_elem, m->boundary_face_elem + c->n_bfaces);
        c->h_bf_side.assign(m->boundary_face_side, m->boundary_face_side + c->n_bfaces);
    }
    rc |= upload(&c->d_w, re.w.data(), (size_t)c->Np);
    rc |= upload<double>(&c->d_bres, nullptr, (size_t)c->n_bfaces * c->nsp * 5 * c->NF);
    rc |= upload<double>(&c->d_bflux, nullptr, (size_t)c->n_bfaces * c->nsp * 5);
    rc |= upload<double>(&c->d_ghost, nullptr, (size_t)c->n_ghost * c->nc * c->NF);   // (room for the field traces too)
    rc |= upload<double>(&c->d_partial, nullptr, (size_t)5 * integral_blocks(c->n_elems));
    rc |= upload<double>(&c->d_out5, nullptr, 8);
    rc |= upload<unsigned long long>(&c->d_vmax, nullptr, (size_t)c->n_vectors);
    c->vec.assign(c->n_vectors, nullptr);
    c->bif.assign(c->n_vectors, nullptr);
    c->vmax_valid.assign(c->n_vectors, 0);
    for (int v = 0; v < c->n_vectors && v < 2 && !rc; v++) {   // the others are allocated on first use (check_vec)
        rc |= upload<double>(&c->vec[v], nullptr, (size_t)c->n_dofs);
        rc |= upload<double>(&c->bif[v], nullptr, (size_t)5 * (c->n_boundaries > 0 ? c->n_boundaries : 1));
    }
    if (!rc && cudaMallocHost((void**)&c->h_small, 64 * sizeof(double)) != cudaSuccess) rc = fail("cudaMallocHost failed");
    if (!rc && cudaMallocHost((void**)&c->h_clock, sizeof(DevClock)) != cudaSuccess) rc = fail("cudaMallocHost failed");
    if (!rc && cudaMalloc((void**)&c->d_clock, sizeof(DevClock)) != cudaSuccess) rc = fail("cudaMalloc failed");
    if (!rc && cudaMalloc((void**)&c->d_probe, 64 * sizeof(double)) != cudaSuccess) rc = fail("cudaMalloc failed");
    if (rc) {
        std::string keep = g_last_error;
        warpii_gpu_destroy(c);
        g_last_error = keep;
        return 1;
    }

    if (const char* env = std::getenv("WARPII_GPU_NO_GRAPH")) c->use_graphs = !(env[0] == '1');
    // persson_peraire_shock_indicator.h:110-112
    c->ind_T = 0.5 * std::pow(10.0, -1.8 * std::pow((double)c->Np, 0.25));
    c->ind_sT = 9.21024 / c->ind_T;
    c->pencil = use_pencil(c->dim, c->Np);
    if (prepare_kernels(c->dim, c->Np) || (c->pencil && prepare_pencil_kernels(c->dim, c->Np))) {
        std::string keep = g_last_error.empty() ? std::string("kernel preparation failed") : g_last_error;
        warpii_gpu_destroy(c);
        g_last_error = keep;
        return 1;
    }

    BoundaryParams& B = c->B;
    std::memset(&B, 0, sizeof B);
    B.bres = c->d_bres;
    B.bflux = c->d_bflux;
    B.bf_elem = c->d_bf_elem;
    B.bf_side = c->d_bf_side;
    B.bf_id = c->d_bf_id;
    B.bc_kind = c->d_bc_kind;
    B.inflow = c->d_inflow;
    B.n_bfaces = c->n_bfaces;
    B.nc = c->nc;
    B.nsp = c->nsp;
    B.n_boundaries = c->n_boundaries;
    B.gamma = c->gamma;
    for (int d = 0; d < 3; d++) { B.inv_h[d] = c->inv_h[d]; B.h[d] = c->h[d]; }
    for (int i = 0; i < c->Np; i++) B.w[i] = re.w[i];
    B.Ng = re.Ng;
    for (int i = 0; i < re.Ng; i++) B.wg[i] = re.wg[i];
    for (int i = 0; i < re.Ng * c->Np; i++) B.Ig[i] = re.Ig[i];

    *out = c;
    return 0;
}

int warpii_gpu_destroy(warpii_gpu_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (double* v : c->vec) cudaFree(v);
    for (double* v : c->bif) cudaFree(v);
    cudaFree(c->d_nbr); cudaFree(c->d_bf_elem); cudaFree(c->d_bf_side); cudaFree(c->d_bf_id); cudaFree(c->d_bc_kind);
    if (c->batch_graph) cudaGraphExecDestroy(c->batch_graph);
    cudaFree(c->d_inflow); cudaFree(c->d_inflow_table); cudaFree(c->d_qm); cudaFree(c->d_scratch); cudaFree(c->d_w); cudaFree(c->d_bres); cudaFree(c->d_bflux); cudaFree(c->d_ghost);
    cudaFree(c->d_sendbuf); cudaFree(c->d_partial); cudaFree(c->d_out5); cudaFree(c->d_alpha); cudaFree(c->d_vmax);
    cudaFree(c->d_send_elem); cudaFree(c->d_send_side);
    cudaFree(c->d_gnode); cudaFree(c->d_gsub); cudaFree(c->d_gface); cudaFree(c->d_jdet); cudaFree(c->d_bgeo); cudaFree(c->d_bmass);
    cudaFree(c->d_nbr2);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->h_clock) cudaFreeHost(c->h_clock);
    cudaFree(c->d_clock);
    cudaFree(c->d_probe);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->slab_events) cudaEventDestroy(e);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    warpii_gpu_free_slab_plan(c->slab_plan);
    if (c->host_step_graph) cudaGraphExecDestroy(c->host_step_graph);
    cudaFree(c->d_host_dt);
    if (c->ev_pack) cudaEventDestroy(c->ev_pack);
    if (c->ev_recv) cudaEventDestroy(c->ev_recv);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    delete c;
    return 0;
}

int64_t warpii_gpu_n_dofs(const warpii_gpu_ctx* c) { return c ? c->n_dofs : 0; }

int warpii_gpu_synchronize(warpii_gpu_ctx* c) {
    if (!c) return fail("null context");
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaStreamSynchronize(c->comm_stream));
    return 0;
}

static int ensure_pinned(warpii_gpu_ctx* c) {
    if (c->h_pin || c->n_dofs == 0) return 0;
    CUDA_OK(cudaMallocHost((void**)&c->h_pin, (size_t)c->n_dofs * sizeof(double)));
    return 0;
}

int warpii_gpu_upload_state(warpii_gpu_ctx* c, int vec, const double* host, const int64_t* dof_index) {
    if (check_vec(c, vec, "upload_state")) return 1;
    if (!host) return fail("upload_state: null host pointer");
    CUDA_OK(cudaSetDevice(c->device));
    const double* src = host;
    if (dof_index) {
        if (ensure_pinned(c)) return 1;
        for (int64_t i = 0; i < c->n_dofs; i++) c->h_pin[i] = host[dof_index[i]];
        src = c->h_pin;
    }
    CUDA_OK(cudaMemcpyAsync(c->vec[vec], src, (size_t)c->n_dofs * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->vmax_valid[vec] = 0;
    return 0;
}

int warpii_gpu_download_state(warpii_gpu_ctx* c, int vec, double* host, const int64_t* dof_index) {
    if (check_vec(c, vec, "download_state")) return 1;
    if (!host) return fail("download_state: null host pointer");
    CUDA_OK(cudaSetDevice(c->device));
    if (dof_index) {
        if (ensure_pinned(c)) return 1;
        CUDA_OK(cudaMemcpyAsync(c->h_pin, c->vec[vec], (size_t)c->n_dofs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        for (int64_t i = 0; i < c->n_dofs; i++) host[dof_index[i]] = c->h_pin[i];
    } else {
        CUDA_OK(cudaMemcpyAsync(host, c->vec[vec], (size_t)c->n_dofs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int warpii_gpu_zero_state(warpii_gpu_ctx* c, int vec) {
    if (check_vec(c, vec, "zero_state")) return 1;
    CUDA_OK(cudaMemsetAsync(c->vec[vec], 0, (size_t)c->n_dofs * sizeof(double), c->stream));
    CUDA_OK(cudaMemsetAsync(c->bif[vec], 0, (size_t)5 * (c->n_boundaries > 0 ? c->n_boundaries : 1) * sizeof(double), c->stream));
    c->vmax_valid[vec] = 0;
    return 0;
}

int warpii_gpu_copy_state(warpii_gpu_ctx* c, int dst, int src) {
    if (check_vec(c, dst, "copy_state") || check_vec(c, src, "copy_state")) return 1;
    CUDA_OK(cudaMemcpyAsync(c->vec[dst], c->vec[src], (size_t)c->n_dofs * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->bif[dst], c->bif[src], (size_t)5 * (c->n_boundaries > 0 ? c->n_boundaries : 1) * sizeof(double),
                            cudaMemcpyDeviceToDevice, c->stream));
    c->vmax_valid[dst] = 0;
    return 0;
}

int warpii_gpu_device_ptr(warpii_gpu_ctx* c, int vec, void** out) {
    if (check_vec(c, vec, "device_ptr")) return 1;
    *out = c->vec[vec];
    return 0;
}

namespace {
// anything that changes kernel arguments (pointers, flags) makes the captured batch stale
void drop_batch_graph(warpii_gpu_ctx* c) {
    if (c->host_step_graph) cudaGraphExecDestroy(c->host_step_graph);
    c->host_step_graph = nullptr;
    if (c->batch_graph) cudaGraphExecDestroy(c->batch_graph);
    c->batch_graph = nullptr;
    c->batch_graph_solution = c->batch_graph_f1 = -1;
}

// (re)upload one species' slice of the inflow table
int push_inflow_table(warpii_gpu_ctx* c, int species) {
    const size_t per_species = (size_t)c->n_bfaces * ipow(c->Np + 1, c->dim - 1) * 5;
    if (per_species == 0) return 0;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(c->d_inflow_table + per_species * species, c->h_inflow_table.data() + per_species * species,
                       per_species * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}
}  // namespace

int warpii_gpu_set_inflow(warpii_gpu_ctx* c, int species, int boundary_id, const double q[5]) {
    if (!c) return fail("null context");
    if (species < 0 || species >= c->nsp) return fail("set_inflow: species %d out of range", species);
    if (boundary_id < 0 || boundary_id >= c->n_boundaries) return fail("set_inflow: boundary id %d out of range", boundary_id);
    for (int k = 0; k < 5; k++) c->h_small[8 + k] = c->h_inflow[((size_t)species * c->n_boundaries + boundary_id) * 5 + k] = q[k];
    CUDA_OK(cudaMemcpyAsync(c->d_inflow + ((size_t)species * c->n_boundaries + boundary_id) * 5, c->h_small + 8,
                            5 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    drop_batch_graph(c);
    if (!c->h_inflow_table.empty()) {   // a table is active: the constant state replaces this boundary's rows
        const int NG = ipow(c->Np + 1, c->dim - 1);
        for (int64_t bf = 0; bf < c->n_bfaces; bf++) {
            if (c->h_bf_id[bf] != boundary_id) continue;
            for (int g = 0; g < NG; g++)
                for (int k = 0; k < 5; k++) c->h_inflow_table[(((size_t)species * c->n_bfaces + bf) * NG + g) * 5 + k] = q[k];
        }
        return push_inflow_table(c, species);
    }
    return 0;
}

int warpii_gpu_set_sources(warpii_gpu_ctx* c, int enabled, double epsilon0, double chi, const double* charge_over_mass) {
    if (!c) return fail("null context");
    drop_batch_graph(c);
    if (!enabled) {
        c->src_on = false;
        return 0;
    }
    if (c->nc < 5 * c->nsp + 8) return fail("set_sources: the source terms need the 8 field components (fields_enabled)");
    if (!(epsilon0 > 0.0)) return fail("set_sources: epsilon0 must be positive");
    if (!charge_over_mass) return fail("set_sources: null charge_over_mass");
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->d_qm) CUDA_OK(cudaMalloc((void**)&c->d_qm, (size_t)c->nsp * sizeof(double)));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(c->d_qm, charge_over_mass, (size_t)c->nsp * sizeof(double), cudaMemcpyHostToDevice));
    c->inv_eps0 = 1.0 / epsilon0;
    c->chi = chi;
    c->src_on = true;
    return 0;
}

int warpii_gpu_set_maxwell(warpii_gpu_ctx* c, int enabled, double light_speed, double chi, double gamma) {
    if (!c) return fail("null context");
    drop_batch_graph(c);
    std::fill(c->vmax_valid.begin(), c->vmax_valid.end(), 0);
    if (!enabled) {
        c->maxwell_on = false;
        c->ncf = 5 * c->nsp;
        return 0;
    }
    if (c->nc < 5 * c->nsp + 8) return fail("set_maxwell: the field system needs the 8 field components (fields_enabled)");
    if (c->general) return fail("set_maxwell: the field system is implemented on Cartesian boxes only");
    if (!(light_speed > 0.0) || chi < 0.0 || gamma < 0.0) return fail("set_maxwell: light_speed must be positive, chi and gamma non-negative");
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->d_qm) {   // the transport-speed kernels read charge / mass only with the sources on, but want a valid pointer
        CUDA_OK(cudaMalloc((void**)&c->d_qm, (size_t)c->nsp * sizeof(double)));
        CUDA_OK(cudaMemset(c->d_qm, 0, (size_t)c->nsp * sizeof(double)));
    }
    c->light_speed = light_speed;
    c->mx_chi = chi;
    c->mx_gamma = gamma;
    c->maxwell_on = true;
    c->ncf = c->nc;
    {
        // Fused into the pencil stage kernel in 2-D; in 3-D the stand-alone kernel right behind the fluid stage is faster
        // (north-star shape, per stage: fused 4.78 ms, separate 3.81 ms; captures in profiles/README.md), although it
        // re-reads the species' densities and momenta.  WARPII_GPU_MAXWELL=fused|separate overrides.
        const char* env = std::getenv("WARPII_GPU_MAXWELL");
        bool fused = c->pencil && c->dim == 2;
        if (env && std::strcmp(env, "fused") == 0) fused = c->pencil;
        if (env && std::strcmp(env, "separate") == 0) fused = false;
        c->mx_fused = fused;
    }
    return 0;
}

int warpii_gpu_set_geometry(warpii_gpu_ctx* c, const warpii_gpu_geometry* g) {
    if (!c) return fail("null context");
    if (c->maxwell_on) return fail("set_geometry: the field system (set_maxwell) is implemented on Cartesian boxes only");
    if (!g || !g->inverse_jacobian || !g->face_normal || !g->face_jacobian) return fail("set_geometry: null table");
    if (c->n_bfaces > 0 && (!g->boundary_normal || !g->boundary_jacobian)) return fail("set_geometry: boundary tables missing");
    if (c->general) return fail("set_geometry: geometry already set");
    CUDA_OK(cudaSetDevice(c->device));
    const int dim = c->dim, Np = c->Np, NN = c->NN, NF = c->NF, K = dim * dim, nf = 2 * dim;
    const int NG = ipow(Np + 1, dim - 1);
    const int64_t ne = c->n_elems;
    warpii_b200::ReferenceElement re(c->p);
    std::vector<double> gnode((size_t)ne * (K + 2) * NN), gsub((size_t)ne * K * NN), jdet((size_t)ne * NN);
    auto stride = [&](int d) { return d == 0 ? 1 : (d == 1 ? Np : Np * Np); };
    for (int64_t e = 0; e < ne; e++) {
        double* ge = &gnode[(size_t)e * (K + 2) * NN];
        for (int q = 0; q < NN; q++) {
            const double* Kq = g->inverse_jacobian + ((size_t)e * NN + q) * K;
            double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) M[r][cc] = Kq[r * dim + cc];
            double det;   // tensor_utils.h:70-82; 3x3 by cofactors
            if (dim == 1) det = M[0][0];
            else if (dim == 2) det = M[0][0] * M[1][1] - M[0][1] * M[1][0];
            else det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                       M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
            if (!(det > 0.0) || !std::isfinite(det))
                return fail("set_geometry: element %lld node %d has a non-positive Jacobian", (long long)e, q);
            const double J = 1.0 / det;   // jacobian_utils.h:15-16
            jdet[(size_t)e * NN + q] = J;
            for (int d = 0; d < dim; d++)
                for (int r = 0; r < dim; r++) ge[(size_t)(d * dim + r) * NN + q] = J * M[r][d];   // jacobian_utils.h:36-39
            ge[(size_t)K * NN + q] = det;
            // fluid_flux_es_dgsem_operator.h:487-502: 5 power iterations on K^T K from (1,...,1)
            double ev[3] = {1, 1, 1};
            for (int it = 0; it < 5; it++) {
                double Kv[3] = {0, 0, 0}, w[3] = {0, 0, 0}, nrm = 0;
                for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) Kv[r] += M[r][cc] * ev[cc];
                for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) w[r] += M[cc][r] * Kv[cc];
                for (int r = 0; r < dim; r++) nrm = std::fmax(nrm, std::fabs(w[r]));
                for (int r = 0; r < dim; r++) ev[r] = w[r] / nrm;
            }
            double Kv[3] = {0, 0, 0}, num = 0, den = 0;
            for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) Kv[r] += M[r][cc] * ev[cc];
            for (int r = 0; r < dim; r++) { num += Kv[r] * Kv[r]; den += ev[r] * ev[r]; }
            ge[(size_t)(K + 1) * NN + q] = std::sqrt(num / den);
        }
        // subcell-face normals along every pencil (subcell_finite_volume_flux.h:101-106, 140-145), Q = diag(w) D
        double* se = &gsub[(size_t)e * K * NN];
        for (int d = 0; d < dim; d++) {
            const int st = stride(d);
            for (int ps = 0; ps < NN; ps++) {
                if ((ps / st) % Np != 0) continue;
                double n[3] = {0, 0, 0};
                for (int r = 0; r < dim; r++) n[r] = ge[(size_t)(d * dim + r) * NN + ps];
                for (int i = 0; i < Np; i++) {
                    for (int m = 0; m < Np; m++) {
                        const double Qim = re.w[i] * re.D[i * Np + m];
                        for (int r = 0; r < dim; r++) n[r] += Qim * ge[(size_t)(d * dim + r) * NN + ps + st * m];
                    }
                    for (int r = 0; r < dim; r++) se[(size_t)(d * dim + r) * NN + ps + st * i] = n[r];
                }
            }
        }
    }
    // faces: unit normal + (face Jacobian) / (Jdet * w_0) = face JxW / cell JxW at the node
    auto face_node = [&](int d, int side, int t) {
        int idx[3] = {0, 0, 0};
        for (int a = 0; a < dim; a++) { if (a == d) continue; idx[a] = t % Np; t /= Np; }
        idx[d] = side ? Np - 1 : 0;
        return idx[0] + Np * (idx[1] + Np * idx[2]);
    };
    std::vector<double> gface((size_t)ne * nf * (dim + 1) * NF);
    for (int64_t e = 0; e < ne; e++)
        for (int f = 0; f < nf; f++)
            for (int t = 0; t < NF; t++) {
                const int q = face_node(f / 2, f % 2, t);
                double* o = &gface[((size_t)(e * nf + f) * (dim + 1)) * NF + t];
                for (int r = 0; r < dim; r++) o[(size_t)r * NF] = g->face_normal[(((size_t)e * nf + f) * NF + t) * dim + r];
                o[(size_t)dim * NF] = g->face_jacobian[((size_t)e * nf + f) * NF + t] / (jdet[(size_t)e * NN + q] * re.w[0]);
            }
    std::vector<int32_t> h_nbr((size_t)ne * nf);
    CUDA_OK(cudaMemcpy(h_nbr.data(), c->d_nbr, h_nbr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<int2> nbr2((size_t)ne * nf);
    for (size_t i = 0; i < nbr2.size(); i++) {
        const int32_t code = g->neighbor_face ? g->neighbor_face[i] : (int32_t)((i % nf) ^ 1);
        if (code < 0 || code >= 16 || (code & 7) >= nf || ((code >> 3) && dim != 2))
            return fail("set_geometry: neighbor_face[%lld] = %d is not a (face, orientation) code of this dimension", (long long)i, code);
        nbr2[i] = make_int2(h_nbr[i], code);
    }
    std::vector<double> bgeo((size_t)c->n_bfaces * NG * (dim + 1)), bmass((size_t)c->n_bfaces * NF);
    for (int64_t b = 0; b < c->n_bfaces; b++) {
        for (int gq = 0; gq < NG; gq++) {
            for (int r = 0; r < dim; r++) bgeo[((size_t)b * NG + gq) * (dim + 1) + r] = g->boundary_normal[((size_t)b * NG + gq) * dim + r];
            bgeo[((size_t)b * NG + gq) * (dim + 1) + dim] = g->boundary_jacobian[(size_t)b * NG + gq];
        }
        const int e = c->h_bf_elem[b], f = c->h_bf_side[b];
        for (int t = 0; t < NF; t++) {
            const int q = face_node(f / 2, f % 2, t);
            double wF = 1.0;
            int tt = t;
            for (int a = 0; a < dim - 1; a++) { wF *= re.w[tt % Np]; tt /= Np; }
            bmass[(size_t)b * NF + t] = 1.0 / (jdet[(size_t)e * NN + q] * re.w[0] * wF);
        }
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (upload(&c->d_gnode, gnode.data(), gnode.size()) || upload(&c->d_gsub, gsub.data(), gsub.size()) ||
        upload(&c->d_gface, gface.data(), gface.size()) || upload(&c->d_jdet, jdet.data(), jdet.size()) ||
        upload(&c->d_nbr2, nbr2.data(), nbr2.size()) || upload(&c->d_bgeo, bgeo.data(), bgeo.size()) ||
        upload(&c->d_bmass, bmass.data(), bmass.size()))
        return 1;
    if (prepare_general_kernels(c->dim, c->Np))
        return fail("set_geometry: kernel preparation failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->GP.gnode = c->d_gnode;
    c->GP.gsub = c->d_gsub;
    c->GP.gface = c->d_gface;
    c->GP.nbr2 = c->d_nbr2;
    c->GP.jdet = c->d_jdet;
    c->GP.bgeo = c->d_bgeo;
    c->GP.bmass = c->d_bmass;
    c->general = true;
    std::fill(c->vmax_valid.begin(), c->vmax_valid.end(), 0);
    drop_batch_graph(c);
    return 0;
}

int warpii_gpu_n_boundary_points(const warpii_gpu_ctx* c, int64_t* n_faces_out, int* points_per_face_out) {
    if (!c) return fail("null context");
    if (n_faces_out) *n_faces_out = c->n_bfaces;
    if (points_per_face_out) *points_per_face_out = ipow(c->Np + 1, c->dim - 1);
    return 0;
}

int warpii_gpu_set_inflow_table(warpii_gpu_ctx* c, int species, const double* table) {
    if (!c) return fail("null context");
    if (species < 0 || species >= c->nsp) return fail("set_inflow_table: species %d out of range", species);
    if (!table) return fail("set_inflow_table: null table");
    const int NG = ipow(c->Np + 1, c->dim - 1);
    const size_t per_species = (size_t)c->n_bfaces * NG * 5;
    if (per_species == 0) return 0;   // this rank owns no boundary faces
    if (c->h_inflow_table.empty()) {
        // first use: every species starts from its constant states, then the kernel reads the table only
        c->h_inflow_table.resize(per_species * c->nsp);
        for (int sp = 0; sp < c->nsp; sp++)
            for (int64_t bf = 0; bf < c->n_bfaces; bf++)
                for (int g = 0; g < NG; g++)
                    for (int k = 0; k < 5; k++)
                        c->h_inflow_table[(((size_t)sp * c->n_bfaces + bf) * NG + g) * 5 + k] =
                            c->h_inflow[((size_t)sp * c->n_boundaries + c->h_bf_id[bf]) * 5 + k];
        CUDA_OK(cudaSetDevice(c->device));
        CUDA_OK(cudaMalloc((void**)&c->d_inflow_table, per_species * c->nsp * sizeof(double)));
        for (int sp = 0; sp < c->nsp; sp++)
            if (sp != species && push_inflow_table(c, sp)) return 1;
        c->B.inflow_table = c->d_inflow_table;
        drop_batch_graph(c);
    }
    std::memcpy(c->h_inflow_table.data() + per_species * species, table, per_species * sizeof(double));
    return push_inflow_table(c, species);
}

int warpii_gpu_forward_euler_step_ex(warpii_gpu_ctx* c, int dst, int u, double dt, double /*t*/, double alpha,
                                     double beta, int flags) {
    if (check_vec(c, dst, "forward_euler_step") || check_vec(c, u, "forward_euler_step")) return 1;
    if (dst == u) return fail("forward_euler_step: dst and u must be different vectors");
    CUDA_OK(cudaSetDevice(c->device));
    return run_stage(c, dst, u, dt, alpha, beta, 0, (flags & WARPII_FUSE_CFL) != 0);
}

int warpii_gpu_forward_euler_step(warpii_gpu_ctx* c, int dst, int u, double dt, double t, double alpha, double beta) {
    return warpii_gpu_forward_euler_step_ex(c, dst, u, dt, t, alpha, beta, 0);
}

int warpii_gpu_rhs(warpii_gpu_ctx* c, int dst, int u, double /*t*/) {
    if (check_vec(c, dst, "rhs") || check_vec(c, u, "rhs")) return 1;
    if (dst == u) return fail("rhs: dst and u must be different vectors");
    CUDA_OK(cudaSetDevice(c->device));
    return run_stage(c, dst, u, 0.0, 1.0, 0.0, 1, false);
}

int warpii_gpu_lsrk_stage(warpii_gpu_ctx* c, int sol_out, int r_out, int sol_in, int r_in, double factor_solution,
                          double factor_ai, double /*t*/) {
    if (check_vec(c, sol_out, "lsrk_stage") || check_vec(c, r_out, "lsrk_stage") || check_vec(c, sol_in, "lsrk_stage") ||
        check_vec(c, r_in, "lsrk_stage"))
        return 1;
    // the stage reads r_in at neighbouring nodes while it writes: nothing may be written into r_in
    if (sol_out == r_in || r_out == r_in) return fail("lsrk_stage: sol_out and r_out must differ from r_in (use a third vector)");
    if (sol_out == r_out) return fail("lsrk_stage: sol_out and r_out must be different vectors");
    if (r_out == sol_in) return fail("lsrk_stage: r_out must differ from sol_in");
    CUDA_OK(cudaSetDevice(c->device));
    if (run_stage(c, sol_out, r_in, 0.0, factor_solution, factor_ai, 2, false, false, sol_in, r_out)) return 1;
    c->vmax_valid[r_out] = 0;
    return 0;
}

int warpii_gpu_max_transport_speed(warpii_gpu_ctx* c, int vec, double* vmax_out) {
    if (check_vec(c, vec, "max_transport_speed")) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    return max_speed(c, vec, vmax_out);
}

int warpii_gpu_recommend_dt(warpii_gpu_ctx* c, int vec, double* dt_out) {
    if (check_vec(c, vec, "recommend_dt")) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    double vmax = 0;
    if (max_speed(c, vec, &vmax)) return 1;
    *dt_out = 0.5 / (vmax * (c->p + 1) * (c->p + 1));   // fluid_flux_es_dgsem_operator.h:446-447
    return 0;
}

int warpii_gpu_ssprk2_step(warpii_gpu_ctx* c, int solution, int f1, double dt, double t) {
    // rk.h:102-105
    if (warpii_gpu_forward_euler_step_ex(c, f1, solution, dt, t, 1.0, 0.0, 0)) return 1;
    return warpii_gpu_forward_euler_step_ex(c, solution, f1, dt, t + dt, 0.5, 0.5, WARPII_FUSE_CFL);
}

}  // extern "C" (the streamed host step needs helpers with C++ linkage)

// ---- SSPRK2 step for a state that lives in HOST memory: copies and compute overlapped slab by slab ---------------------------
// The adapter of INTEGRATION.md keeps FiveMSolutionVec on the device; a caller that cannot (post-processing on the host every
// step, state larger than HBM) pays two PCIe transfers of the state per step, ten times the compute.  They are independent
// directions of a full-duplex link and the stage kernels take element ranges, so the step is cut into slabs of consecutive
// elements: slab k is uploaded on one stream; first-stage launches follow as soon as the slabs holding their face neighbours
// have arrived, second-stage launches as soon as the first stage of their neighbours' slabs has been launched (one in-order
// compute stream, so "launched" implies "ordered before"), and each finished slab returns on a third stream while later slabs
// are still arriving.  Same kernels, same operands: the result is bit-identical to upload + warpii_gpu_ssprk2_step + download.
namespace {
struct SlabPlan {
    int n = 0;
    std::vector<int64_t> begin;              // [n+1] element ranges, multiples of the patch size
    std::vector<std::vector<int>> deps;      // slabs holding face neighbours of slab k (k included)
};

int build_slab_plan(warpii_gpu_ctx* c, int n_slabs, SlabPlan& plan) {
    const int nf = 2 * c->dim;
    const int G = (c->pencil && !c->general) ? pencil_patch_elems(c->dim, c->Np) : elems_per_block(c->dim, c->Np);
    const int64_t n_patches = (c->n_elems + G - 1) / G;
    int S = n_slabs > 0 ? n_slabs : 32;
    if (S > n_patches) S = (int)(n_patches > 0 ? n_patches : 1);
    plan.n = S;
    plan.begin.assign(S + 1, 0);
    for (int k = 0; k <= S; k++) plan.begin[k] = std::min<int64_t>(c->n_elems, ((n_patches * k) / S) * G);
    plan.begin[S] = c->n_elems;
    std::vector<int32_t> nbr((size_t)c->n_elems * nf);
    CUDA_OK(cudaMemcpy(nbr.data(), c->d_nbr, nbr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    auto slab_of = [&](int64_t e) { return (int)(std::upper_bound(plan.begin.begin(), plan.begin.end(), e) - plan.begin.begin()) - 1; };
    plan.deps.assign(S, {});
    for (int k = 0; k < S; k++) {
        std::vector<char> seen(S, 0);
        seen[k] = 1;
        for (int64_t e = plan.begin[k]; e < plan.begin[k + 1]; e++)
            for (int f = 0; f < nf; f++) {
                const int32_t v = nbr[(size_t)e * nf + f];
                if (v >= 0 && v < c->n_elems) seen[slab_of(v)] = 1;
            }
        for (int j = 0; j < S; j++)
            if (seen[j]) plan.deps[k].push_back(j);
    }
    return 0;
}
}  // namespace

void warpii_gpu_free_slab_plan(void* plan) { delete (SlabPlan*)plan; }

namespace {
// The streamed step on a SHARDED context (periodic mesh, NCCL halo).  Slab 0 = the interface elements (they are numbered
// first), slabs 1.. = the interior cut into pieces of whole patches.  Slab 0 is uploaded first; its two exchanges and its two
// stages run on the communication stream (as in run_stage), everything else on the main stream, ordered by events:
//   upload(0) -> pack + send/recv u -> stage 1 of slab 0 (needs the uploads of its neighbours' slabs) -> pack + send/recv f1
//   -> stage 2 of slab 0 (needs stage 1 of its neighbours' slabs) -> download(0)
//   interior slab j: stage 1 once its neighbours' slabs are uploaded, stage 2 once their stage 1 is launched (main stream
//   is in order; for slab 0 an event), download on the third stream.
// Same kernels, same operands: bit-identical to upload + warpii_gpu_ssprk2_step + download.
int host_step_sharded(warpii_gpu_ctx* c, int solution, int f1, const double* host_in, double* host_out, double dt, int n_slabs) {
    const size_t per_elem = (size_t)c->nc * c->NN;
    const int G = c->pencil && !c->general ? pencil_patch_elems(c->dim, c->Np) : elems_per_block(c->dim, c->Np);
    const int64_t n_if = c->n_interface, n_in = c->n_elems - n_if;
    int S = (n_slabs > 0 ? n_slabs : 32);
    const int64_t in_patches = (n_in + G - 1) / G;
    if (S - 1 > in_patches) S = (int)in_patches + 1;
    if (S < 2) S = 2;
    std::vector<int64_t> begin(S + 1, 0);
    begin[0] = 0;
    begin[1] = n_if;
    for (int k = 1; k < S; k++) begin[k + 1] = std::min<int64_t>(c->n_elems, n_if + ((in_patches * k) / (S - 1)) * G);
    begin[S] = c->n_elems;
    // neighbour slabs from the face table (ghost faces are slab 0's business)
    if ((int)c->shard_deps.size() != S) {
        const int nf = 2 * c->dim;
        std::vector<int32_t> nbr((size_t)c->n_elems * nf);
        CUDA_OK(cudaMemcpy(nbr.data(), c->d_nbr, nbr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        auto slab_of = [&](int64_t e) { return (int)(std::upper_bound(begin.begin(), begin.end(), e) - begin.begin()) - 1; };
        c->shard_deps.assign(S, {});
        for (int k = 0; k < S; k++) {
            std::vector<char> seen(S, 0);
            seen[k] = 1;
            for (int64_t e = begin[k]; e < begin[k + 1]; e++)
                for (int f = 0; f < nf; f++) {
                    const int32_t v = nbr[(size_t)e * nf + f];
                    if (v >= 0 && v < c->n_elems) seen[slab_of(v)] = 1;
                    else if (v >= c->n_elems && k != 0) return fail("host_ssprk2_step: a ghost face outside the interface elements");
                }
            for (int j = 0; j < S; j++)
                if (seen[j]) c->shard_deps[k].push_back(j);
        }
    }
    const std::vector<std::vector<int>>& deps = c->shard_deps;
    if (!c->h2d_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    }
    while ((int)c->slab_events.size() < 3 * S + 4) {
        cudaEvent_t e;
        CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->slab_events.push_back(e);
    }
    cudaEvent_t* up_done = c->slab_events.data();
    cudaEvent_t* s1_done = c->slab_events.data() + S;
    cudaEvent_t* s2_done = c->slab_events.data() + 2 * S;
    cudaEvent_t start = c->slab_events[3 * S], if_s1 = c->slab_events[3 * S + 1], main_s1 = c->slab_events[3 * S + 2],
                joined = c->slab_events[3 * S + 3];
    if (!c->d_host_dt) CUDA_OK(cudaMalloc((void**)&c->d_host_dt, sizeof(double)));
    c->h_small[24] = dt;
    CUDA_OK(cudaMemcpyAsync(c->d_host_dt, c->h_small + 24, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_vmax + solution, 0, sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaEventRecord(start, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->h2d_stream, start, 0));
    CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, start, 0));
    CUDA_OK(cudaStreamWaitEvent(c->comm_stream, start, 0));
    StageParams P1 = stage_params(c, f1, solution, dt, 1.0, 0.0, 0, false);        // rk.h:102-103
    StageParams P2 = stage_params(c, solution, f1, dt, 0.5, 0.5, 0, true);         // rk.h:104-105, fused CFL
    P1.dt_dev = c->d_host_dt;
    P2.dt_dev = c->d_host_dt;
    auto upload = [&](int k) -> int {
        const size_t off = (size_t)begin[k] * per_elem, cnt = (size_t)(begin[k + 1] - begin[k]) * per_elem;
        if (cnt) CUDA_OK(cudaMemcpyAsync(c->vec[solution] + off, host_in + off, cnt * sizeof(double), cudaMemcpyHostToDevice, c->h2d_stream));
        CUDA_OK(cudaEventRecord(up_done[k], c->h2d_stream));
        return 0;
    };
    auto download = [&](int k, cudaEvent_t after) -> int {
        CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, after, 0));
        const size_t off = (size_t)begin[k] * per_elem, cnt = (size_t)(begin[k + 1] - begin[k]) * per_elem;
        if (cnt) CUDA_OK(cudaMemcpyAsync(host_out + off, c->vec[solution] + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->d2h_stream));
        return 0;
    };
    // ---- uploads: the interface elements first (their exchange starts as soon as they have arrived), then the interior
    // slabs next to them; stages and downloads are issued as soon as what they read has been issued ---------------------------
    std::vector<char> uploaded(S, 0), s1(S, 0), s2(S, 0);
    auto ready = [&](int k, const std::vector<char>& have) {
        for (int j : deps[k]) if (!have[j]) return false;
        return true;
    };
    std::vector<int> order;
    order.push_back(0);
    if (S > 2) order.push_back(S - 1);
    for (int k = 1; k < (S > 2 ? S - 1 : S); k++) order.push_back(k);
    for (int k : order) {
        if (upload(k)) return 1;
        uploaded[k] = 1;
        if (k == 0) {
            CUDA_OK(cudaStreamWaitEvent(c->comm_stream, up_done[0], 0));
            if (exchange_on_comm_stream(c, solution)) return 1;
        }
        bool waited = false;
        for (int j = 1; j < S; j++) {      // stage 1 of interior slabs, main stream
            if (s1[j] || !ready(j, uploaded)) continue;
            if (!waited) { CUDA_OK(cudaStreamWaitEvent(c->stream, up_done[k], 0)); waited = true; }
            P1.elem_begin = begin[j];
            P1.elem_end = begin[j + 1];
            do_launch_stage(c, P1, c->stream);
            c->launches++;
            s1[j] = 1;
        }
        if (!s1[0] && ready(0, uploaded)) {
            // stage 1 of the interface elements on the communication stream (behind the receive), then the exchange of f1
            CUDA_OK(cudaStreamWaitEvent(c->comm_stream, up_done[k], 0));
            P1.elem_begin = begin[0];
            P1.elem_end = begin[1];
            do_launch_stage(c, P1, c->comm_stream);
            c->launches++;
            s1[0] = 1;
            CUDA_OK(cudaEventRecord(if_s1, c->comm_stream));
            if (exchange_on_comm_stream(c, f1)) return 1;
        }
        for (int j = 1; j < S; j++) {      // stage 2 of interior slabs (those next to the interface wait for its stage 1)
            if (s2[j] || !ready(j, s1)) continue;
            bool needs_if = false;
            for (int d : deps[j]) needs_if = needs_if || d == 0;
            if (needs_if) CUDA_OK(cudaStreamWaitEvent(c->stream, if_s1, 0));
            P2.elem_begin = begin[j];
            P2.elem_end = begin[j + 1];
            do_launch_stage(c, P2, c->stream);
            c->launches++;
            s2[j] = 1;
            CUDA_OK(cudaEventRecord(s2_done[j], c->stream));
            if (download(j, s2_done[j])) return 1;
        }
        if (!s2[0] && s1[0] && ready(0, s1)) {
            // stage 2 of the interface elements: behind the second receive and the stage 1 of the interior slabs next to them
            CUDA_OK(cudaEventRecord(main_s1, c->stream));   // (their stage 1 was launched before this point of the main stream)
            CUDA_OK(cudaStreamWaitEvent(c->comm_stream, main_s1, 0));
            P2.elem_begin = begin[0];
            P2.elem_end = begin[1];
            do_launch_stage(c, P2, c->comm_stream);
            c->launches++;
            s2[0] = 1;
            CUDA_OK(cudaEventRecord(s2_done[0], c->comm_stream));
            if (download(0, s2_done[0])) return 1;
        }
    }
    for (int j = 0; j < S; j++)
        if (!s1[j] || !s2[j]) return fail("host_ssprk2_step: internal scheduling error (a slab never became ready)");
    (void)s1_done;
    // join everything back into the main stream
    CUDA_OK(cudaEventRecord(joined, c->d2h_stream));
    CUDA_OK(cudaStreamWaitEvent(c->stream, joined, 0));
    CUDA_OK(cudaStreamWaitEvent(c->stream, s2_done[0], 0));
    CUDA_OK(cudaGetLastError());
    c->vmax_valid[solution] = 1;
    c->vmax_valid[f1] = 0;
    return 0;
}
}  // namespace

extern "C" int warpii_gpu_host_ssprk2_step(warpii_gpu_ctx* c, int solution, int f1, const double* host_in, double* host_out,
                                           double dt, double t, double* next_dt_out, int n_slabs) {
    if (check_vec(c, solution, "host_ssprk2_step") || check_vec(c, f1, "host_ssprk2_step")) return 1;
    if (solution == f1) return fail("host_ssprk2_step: solution and f1 must be different vectors");
    if (!host_in || !host_out) return fail("host_ssprk2_step: null host pointer");
    if (!(dt > 0.0)) return fail("host_ssprk2_step: dt must be positive (take it from warpii_gpu_recommend_dt or from the previous call)");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t per_elem = (size_t)c->nc * c->NN;
    const bool streamed = c->n_bfaces == 0 && !c->comm && c->n_elems > 0;
    const bool streamed_sharded = c->n_bfaces == 0 && c->comm && !c->peer_rank.empty() && c->n_interface > 0 &&
                                  c->n_interface < c->n_elems;
    if (streamed_sharded) {
        if (host_step_sharded(c, solution, f1, host_in, host_out, dt, n_slabs)) return 1;
    } else if (!streamed) {
        // boundary faces / sharded runs: plain sequence (the boundary kernel and the halo exchange work on the whole vector)
        CUDA_OK(cudaMemcpyAsync(c->vec[solution], host_in, (size_t)c->n_dofs * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->vmax_valid[solution] = 0;
        if (warpii_gpu_ssprk2_step(c, solution, f1, dt, t)) return 1;
        CUDA_OK(cudaMemcpyAsync(host_out, c->vec[solution], (size_t)c->n_dofs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (!c->h2d_stream) {
            CUDA_OK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
            CUDA_OK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        }
        if (c->slab_plan_slabs != n_slabs || !c->slab_plan) {
            auto* plan = new SlabPlan();
            if (build_slab_plan(c, n_slabs, *plan)) { delete plan; return 1; }
            delete (SlabPlan*)c->slab_plan;
            c->slab_plan = plan;
            c->slab_plan_slabs = n_slabs;
        }
        const SlabPlan& plan = *(const SlabPlan*)c->slab_plan;
        const int S = plan.n;
        if (!c->d_host_dt) CUDA_OK(cudaMalloc((void**)&c->d_host_dt, sizeof(double)));
        c->h_small[24] = dt;
        CUDA_OK(cudaMemcpyAsync(c->d_host_dt, c->h_small + 24, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const bool replay = c->use_graphs && c->host_step_graph && c->hsg_in == host_in && c->hsg_out == host_out &&
                            c->hsg_solution == solution && c->hsg_f1 == f1 && c->hsg_slabs == n_slabs;
        bool capturing = false;
        if (replay) {
            CUDA_OK(cudaGraphLaunch(c->host_step_graph, c->stream));
            c->launches += 2 * S;
        } else {
        if (c->host_step_graph) { cudaGraphExecDestroy(c->host_step_graph); c->host_step_graph = nullptr; }
        // ~10 API calls per slab cost as much host time as the slab's transfer: capture the step once, replay it afterwards
        // (pageable host memory cannot be captured: such a call runs eagerly)
        cudaPointerAttributes attr_in{}, attr_out{};
        const bool pinned = cudaPointerGetAttributes(&attr_in, host_in) == cudaSuccess && attr_in.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&attr_out, host_out) == cudaSuccess && attr_out.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (c->use_graphs && pinned) {
            CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            capturing = true;
        }
        while ((int)c->slab_events.size() < 2 * S + 1) {
            cudaEvent_t e;
            CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->slab_events.push_back(e);
        }
        cudaEvent_t* up_done = c->slab_events.data();
        cudaEvent_t* s2_done = c->slab_events.data() + S;
        cudaEvent_t start = c->slab_events[2 * S];
        // everything queued earlier on the main stream (previous steps, uploads) comes first
        CUDA_OK(cudaMemsetAsync(c->d_vmax + solution, 0, sizeof(unsigned long long), c->stream));
        CUDA_OK(cudaEventRecord(start, c->stream));
        CUDA_OK(cudaStreamWaitEvent(c->h2d_stream, start, 0));
        CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, start, 0));
        StageParams P1 = stage_params(c, f1, solution, dt, 1.0, 0.0, 0, false);        // rk.h:102-103
        StageParams P2 = stage_params(c, solution, f1, dt, 0.5, 0.5, 0, true);         // rk.h:104-105, fused CFL
        P1.dt_dev = c->d_host_dt;
        P2.dt_dev = c->d_host_dt;
        std::vector<char> uploaded(S, 0), s1(S, 0), s2(S, 0);
        auto ready = [&](int k, const std::vector<char>& have) {
            for (int j : plan.deps[k]) if (!have[j]) return false;
            return true;
        };
        // Upload order: the last slab first, then 0, 1, 2, ...  On a periodic mesh slab 0 needs slab S-1; sent in plain order,
        // the first and the last slabs could not start before the whole state had arrived.
        for (int step = 0; step < S; step++) {
            const int k = (S >= 3) ? (step == 0 ? S - 1 : step - 1) : step;
            const size_t off = (size_t)plan.begin[k] * per_elem, cnt = (size_t)(plan.begin[k + 1] - plan.begin[k]) * per_elem;
            CUDA_OK(cudaMemcpyAsync(c->vec[solution] + off, host_in + off, cnt * sizeof(double), cudaMemcpyHostToDevice, c->h2d_stream));
            CUDA_OK(cudaEventRecord(up_done[k], c->h2d_stream));
            uploaded[k] = 1;
            bool waited = false;
            for (int j = 0; j < S; j++) {
                if (s1[j] || !ready(j, uploaded)) continue;
                if (!waited) { CUDA_OK(cudaStreamWaitEvent(c->stream, up_done[k], 0)); waited = true; }
                P1.elem_begin = plan.begin[j];
                P1.elem_end = plan.begin[j + 1];
                do_launch_stage(c, P1, c->stream);
                c->launches++;
                s1[j] = 1;
            }
            for (int j = 0; j < S; j++) {
                if (s2[j] || !ready(j, s1)) continue;
                P2.elem_begin = plan.begin[j];
                P2.elem_end = plan.begin[j + 1];
                do_launch_stage(c, P2, c->stream);
                c->launches++;
                s2[j] = 1;
                CUDA_OK(cudaEventRecord(s2_done[j], c->stream));
                CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, s2_done[j], 0));
                const size_t o2 = (size_t)plan.begin[j] * per_elem, n2 = (size_t)(plan.begin[j + 1] - plan.begin[j]) * per_elem;
                CUDA_OK(cudaMemcpyAsync(host_out + o2, c->vec[solution] + o2, n2 * sizeof(double), cudaMemcpyDeviceToHost, c->d2h_stream));
            }
        }
        bool scheduled = true;
        for (int j = 0; j < S; j++) scheduled = scheduled && s1[j] && s2[j];
        if (capturing) {
            // join the transfer streams back into the capturing stream
            cudaEvent_t join_up = up_done[0], join_down = s2_done[0];   // (both already consumed by their waiters)
            CUDA_OK(cudaEventRecord(join_up, c->h2d_stream));
            CUDA_OK(cudaEventRecord(join_down, c->d2h_stream));
            CUDA_OK(cudaStreamWaitEvent(c->stream, join_up, 0));
            CUDA_OK(cudaStreamWaitEvent(c->stream, join_down, 0));
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
            if (ce != cudaSuccess || !scheduled) {
                if (graph) cudaGraphDestroy(graph);
                return fail("host_ssprk2_step: %s", scheduled ? cudaGetErrorString(ce) : "internal scheduling error");
            }
            const cudaError_t ci = cudaGraphInstantiate(&c->host_step_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (ci != cudaSuccess) {
                c->host_step_graph = nullptr;
                return fail("host_ssprk2_step: graph instantiation failed: %s", cudaGetErrorString(ci));
            }
            c->hsg_in = host_in; c->hsg_out = host_out; c->hsg_solution = solution; c->hsg_f1 = f1; c->hsg_slabs = n_slabs;
            CUDA_OK(cudaGraphLaunch(c->host_step_graph, c->stream));   // the captured work has not run yet
        } else if (!scheduled) {
            return fail("host_ssprk2_step: internal scheduling error (a slab never became ready)");
        }
        }   // !replay
        CUDA_OK(cudaGetLastError());
        c->vmax_valid[solution] = 1;
        c->vmax_valid[f1] = 0;
        if (!replay && !capturing) CUDA_OK(cudaStreamSynchronize(c->d2h_stream));   // (a graph ends on the main stream)
    }
    double vmax = 0.0;
    if (max_speed(c, solution, &vmax)) return 1;   // synchronises the main stream; the slot was filled by the second stage
    if (next_dt_out) *next_dt_out = 0.5 / (vmax * (c->p + 1) * (c->p + 1));
    return 0;
}

extern "C" {

int warpii_gpu_advance_to(warpii_gpu_ctx* c, int solution, int f1, double* t_inout, double t_stop, double fixed_dt,
                          int64_t max_steps, int64_t* steps_out) {
    if (check_vec(c, solution, "advance_to") || check_vec(c, f1, "advance_to")) return 1;
    if (solution == f1) return fail("advance_to: solution and f1 must be different vectors");
    if (!t_inout) return fail("advance_to: null time pointer");
    CUDA_OK(cudaSetDevice(c->device));
    // The loop of timestepper.cc:34-42 with t, dt and the stop test resident on the device: steps are enqueued in
    // batches and the host looks at the clock once per batch, so there is no host round trip (and no launch gap) per step.
    if (!c->vmax_valid[solution]) {
        CUDA_OK(cudaMemsetAsync(c->d_vmax + solution, 0, sizeof(unsigned long long), c->stream));
        do_launch_cfl(c, solution);
        c->launches++;
        c->vmax_valid[solution] = 1;
    }
    if (c->comm && c->n_ranks > 1)   // max over ranks; idempotent if the slot already holds the global maximum
        NCCL_OK(g_nccl.AllReduce(c->d_vmax + solution, c->d_vmax + solution, 1, ncclUint64, ncclMax, c->comm, c->stream));
    DevClock* hc = c->h_clock;
    *hc = DevClock{*t_inout, 0.0, t_stop, fixed_dt > 0.0 ? fixed_dt : 0.0, 0, max_steps > 0 ? max_steps : 0, 0, 0, 0, c->p + 1};
    CUDA_OK(cudaMemcpyAsync(c->d_clock, hc, sizeof(DevClock), cudaMemcpyHostToDevice, c->stream));
    const int batch = 8;
    int64_t known_steps = 0;
    for (;;) {
        int n = batch;
        if (max_steps > 0 && max_steps - known_steps < n) n = (int)(max_steps - known_steps);
        // A full batch on one GPU with timing off is replayed from a CUDA graph (captured on first use): same kernels, same
        // arguments, one launch call instead of ~30, which is what small meshes are bound by.
        const bool graphable = c->use_graphs && !c->comm && !c->timing && n == batch;
        if (graphable && c->batch_graph && c->batch_graph_solution == solution && c->batch_graph_f1 == f1) {
            CUDA_OK(cudaGraphLaunch(c->batch_graph, c->stream));
            c->launches += c->batch_graph_launches;
        } else {
            const bool capture = graphable;
            const int64_t launches_before = c->launches;
            if (capture) {
                drop_batch_graph(c);
                CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            }
            int rc = 0;
            for (int i = 0; i < n && !rc; i++) {
                launch_clock(c->d_clock, c->d_vmax + solution, 0, c->stream);
                rc = run_stage(c, f1, solution, 0.0, 1.0, 0.0, 0, false, true);                 // rk.h:102-103
                if (!rc) rc = run_stage(c, solution, f1, 0.0, 0.5, 0.5, 0, true, true);         // rk.h:104-105
                if (!rc && c->comm && c->n_ranks > 1)
                    NCCL_OK(g_nccl.AllReduce(c->d_vmax + solution, c->d_vmax + solution, 1, ncclUint64, ncclMax, c->comm, c->stream));
            }
            if (!rc) launch_clock(c->d_clock, c->d_vmax + solution, 1, c->stream);   // book the last step of the batch, test for the end
            c->launches += n + 1;
            if (capture) {
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
                if (rc || ce != cudaSuccess) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc ? 1 : fail("advance_to: graph capture failed: %s", cudaGetErrorString(ce));
                }
                const cudaError_t ci = cudaGraphInstantiate(&c->batch_graph, graph, 0);
                cudaGraphDestroy(graph);
                if (ci != cudaSuccess) {
                    c->batch_graph = nullptr;
                    return fail("advance_to: graph instantiation failed: %s", cudaGetErrorString(ci));
                }
                c->batch_graph_solution = solution;
                c->batch_graph_f1 = f1;
                c->batch_graph_launches = c->launches - launches_before;
                CUDA_OK(cudaGraphLaunch(c->batch_graph, c->stream));   // the captured work has not run yet
            } else if (rc) {
                return 1;
            }
            if (c->timing && c->n_probes < 64) launch_sm_clock_probe(c->d_probe + c->n_probes++, c->stream);
        }
        CUDA_OK(cudaMemcpyAsync(hc, c->d_clock, sizeof(DevClock), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        known_steps = hc->steps;
        if (hc->done || hc->error || (max_steps > 0 && known_steps >= max_steps)) break;
    }
    c->vmax_valid[solution] = 1;
    c->vmax_valid[f1] = 0;
    *t_inout = hc->t;
    if (steps_out) *steps_out = hc->steps;
    if (hc->error)
        return fail("advance_to: recommended dt = %g at t = %g (state is no longer physical)", hc->dt, hc->t);
    return 0;
}

int warpii_gpu_boundary_fluxes(warpii_gpu_ctx* c, int vec, double* out) {
    if (check_vec(c, vec, "boundary_fluxes")) return 1;
    if (c->n_boundaries <= 0) return 0;
    CUDA_OK(cudaMemcpyAsync(out, c->bif[vec], (size_t)5 * c->n_boundaries * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int warpii_gpu_set_boundary_fluxes(warpii_gpu_ctx* c, int vec, const double* in) {
    if (check_vec(c, vec, "set_boundary_fluxes")) return 1;
    if (c->n_boundaries <= 0) return 0;
    CUDA_OK(cudaMemcpyAsync(c->bif[vec], in, (size_t)5 * c->n_boundaries * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int warpii_gpu_global_integral(warpii_gpu_ctx* c, int vec, int species, double out[5]) {
    if (check_vec(c, vec, "global_integral")) return 1;
    if (species < 0 || species >= c->nsp) return fail("global_integral: species %d out of range", species);
    CUDA_OK(cudaSetDevice(c->device));
    if (c->general) launch_integral_general(c->dim, c->Np, c->vec[vec], c->n_elems, c->nc, species, c->GP, c->d_w, c->d_partial, c->d_out5, c->stream);
    else launch_integral(c->dim, c->Np, c->vec[vec], c->n_elems, c->nc, species, c->Jdet, c->d_w, c->d_partial, c->d_out5, c->stream);
    c->launches += 2;
    if (c->comm && c->n_ranks > 1)   // replaces Utilities::MPI::sum, dg_solution_helper.cc:96
        NCCL_OK(g_nccl.AllReduce(c->d_out5, c->d_out5, 5, ncclDouble, ncclSum, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->h_small + 16, c->d_out5, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; k++) out[k] = c->h_small[16 + k];
    return 0;
}

int warpii_gpu_shock_indicator(warpii_gpu_ctx* c, int vec, double* alpha_out) {
    if (check_vec(c, vec, "shock_indicator")) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    // run the stage kernel in rhs mode into a scratch vector-sized buffer? cheaper: use the last vector slot's
    // storage is not safe, so allocate the small alpha table and a throw-away destination lazily.
    if (!c->d_alpha) CUDA_OK(cudaMalloc((void**)&c->d_alpha, (size_t)(c->n_elems > 0 ? c->n_elems : 1) * c->nsp * sizeof(double)));
    // throw-away destination of the launch, kept for the next call and freed with the context
    if (!c->d_scratch) CUDA_OK(cudaMalloc((void**)&c->d_scratch, (size_t)(c->n_dofs > 0 ? c->n_dofs : 1) * sizeof(double)));
    double* const scratch = c->d_scratch;
    StageParams P = stage_params(c, vec, vec, 0.0, 1.0, 0.0, 1, false);
    P.dst = scratch;
    P.alpha_out = c->d_alpha;
    if (c->n_bfaces > 0) {
        BoundaryParams B = c->B;
        B.u = c->vec[vec];
        do_launch_boundary(c, B, c->stream);
    }
    if (c->comm && !c->peer_rank.empty()) {
        if (start_exchange(c, vec)) return 1;
        CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_recv, 0));
    }
    if (c->n_interface > 0) {
        P.elem_begin = 0; P.elem_end = c->n_interface;
        do_launch_stage(c, P, c->stream);
        P.elem_begin = c->n_interface; P.elem_end = c->n_elems;
    }
    do_launch_stage(c, P, c->stream);
    c->launches++;
    CUDA_OK(cudaMemcpyAsync(alpha_out, c->d_alpha, (size_t)c->n_elems * c->nsp * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int warpii_gpu_nccl_unique_id(char id[WARPII_GPU_NCCL_ID_BYTES]) {
    if (load_nccl()) return 1;
    static_assert(sizeof(ncclUniqueId) == WARPII_GPU_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId uid;
    NCCL_OK(g_nccl.GetUniqueId(&uid));
    std::memcpy(id, &uid, sizeof uid);
    return 0;
}

int warpii_gpu_attach_comm(warpii_gpu_ctx* c, const char id[WARPII_GPU_NCCL_ID_BYTES], int rank, int n_ranks,
                           const warpii_gpu_halo* halo) {
    if (!c) return fail("null context");
    if (load_nccl()) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    NCCL_OK(g_nccl.CommInitRank(&c->comm, n_ranks, uid, rank));
    c->rank = rank;
    c->n_ranks = n_ranks;
    if (halo && halo->n_peers > 0) {
        c->peer_rank.assign(halo->peer_rank, halo->peer_rank + halo->n_peers);
        c->send_offset.assign(halo->send_offset, halo->send_offset + halo->n_peers + 1);
        c->recv_offset.assign(halo->recv_offset, halo->recv_offset + halo->n_peers + 1);
        c->n_send = c->send_offset.back();
        if (c->recv_offset.back() != c->n_ghost)
            return fail("attach_comm: recv_offset covers %lld ghost faces, mesh declared %lld", (long long)c->recv_offset.back(), (long long)c->n_ghost);
        for (int64_t i = 0; i < c->n_send; i++)
            if (halo->send_elem[i] < 0 || halo->send_elem[i] >= c->n_elems || halo->send_side[i] < 0 || halo->send_side[i] >= 2 * c->dim)
                return fail("attach_comm: send list entry %lld out of range", (long long)i);
        if (halo->n_interface_elems < 0 || halo->n_interface_elems > c->n_elems) return fail("attach_comm: n_interface_elems out of range");
        c->n_interface = halo->n_interface_elems;
        if (upload(&c->d_send_elem, halo->send_elem, (size_t)c->n_send)) return 1;
        if (upload(&c->d_send_side, halo->send_side, (size_t)c->n_send)) return 1;
        if (upload<double>(&c->d_sendbuf, nullptr, (size_t)c->n_send * c->nc * c->NF)) return 1;
    }
    return 0;
}

int64_t warpii_gpu_launch_count(const warpii_gpu_ctx* c) { return c ? c->launches : 0; }

int warpii_gpu_stage_timing(warpii_gpu_ctx* c, int enable, double* ms_total, int64_t* n_launches) {
    if (!c) return fail("null context");
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double total = 0;
    int64_t n = 0;
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
        total += ms;
        n++;
    }
    if (ms_total) *ms_total = total;
    if (n_launches) *n_launches = n;
    c->ev_used = 0;
    c->timing = enable != 0;
    return 0;
}

int warpii_gpu_stream(warpii_gpu_ctx* c, void** stream_out) {
    if (!c || !stream_out) return fail("null argument");
    *stream_out = (void*)c->stream;
    return 0;
}

}  // extern "C"

namespace wgpu {
void launch_point_flux(int n, const double* qa, const double* qb, int d, double gamma, double* ec, double* es,
                       double* prim, cudaStream_t s);
}

extern "C" int warpii_gpu_point_fluxes(int device, int n, const double* qa, const double* qb, int d, double gamma,
                                       double* ec_out, double* es_out, double* prim_out) {
    if (n <= 0) return 0;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device available");
    CUDA_OK(cudaSetDevice(device));
    double *dqa = nullptr, *dqb = nullptr, *dec = nullptr, *des = nullptr, *dpr = nullptr;
    if (upload(&dqa, qa, (size_t)5 * n) || upload(&dqb, qb, (size_t)5 * n) || upload<double>(&dec, nullptr, (size_t)5 * n) ||
        upload<double>(&des, nullptr, (size_t)5 * n) || upload<double>(&dpr, nullptr, (size_t)12 * n))
        return 1;
    wgpu::launch_point_flux(n, dqa, dqb, d, gamma, dec, des, dpr, nullptr);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(ec_out, dec, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(es_out, des, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (prim_out) CUDA_OK(cudaMemcpy(prim_out, dpr, (size_t)12 * n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dqa); cudaFree(dqb); cudaFree(dec); cudaFree(des); cudaFree(dpr);
    return 0;
}

namespace wgpu {
double launch_fp64_fma_chain(int blocks, double* sink, cudaStream_t s);
}

// Measured FP64 throughput of the CUDA cores: fused multiply-adds per second over a launch that fills every SM (best of 5)
extern "C" int warpii_gpu_measure_fp64_peak(int device, double* fma_per_second_out) {
    if (!fma_per_second_out) return fail("null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device available");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    if (upload<double>(&sink, nullptr, 1)) return 1;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8 * 4;   // 8 resident blocks of 256 threads per SM, four waves
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_OK(cudaEventRecord(e0, nullptr));
        const double fmas = wgpu::launch_fp64_fma_chain(blocks, sink, nullptr);
        CUDA_OK(cudaEventRecord(e1, nullptr));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms > 0) best = std::max(best, fmas / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *fma_per_second_out = best;
    return 0;
}

namespace wgpu {
double launch_fp64_ilp(int blocks, int threads, int ilp, double* sink, cudaStream_t s);
}

// Tuning diagnostic: FMA rate (per second, thread level) with ONE block of `threads_per_sm` threads per SM and `ilp`
// independent chains per thread, i.e. the FP64 throughput a kernel with that many resident warps and that much
// instruction-level parallelism can reach (compare with warpii_gpu_measure_fp64_peak).
extern "C" int warpii_gpu_fp64_rate_probe(int device, int threads_per_sm, int ilp, double* fma_per_second_out) {
    if (!fma_per_second_out || threads_per_sm < 32 || threads_per_sm > 1024) return fail("fp64_rate_probe: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device available");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    if (upload<double>(&sink, nullptr, 1)) return 1;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_OK(cudaEventRecord(e0, nullptr));
        const double fmas = wgpu::launch_fp64_ilp(prop.multiProcessorCount, threads_per_sm, ilp, sink, nullptr);
        CUDA_OK(cudaEventRecord(e1, nullptr));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms > 0) best = std::max(best, fmas / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *fma_per_second_out = best;
    return 0;
}

namespace wgpu {
void launch_division_check(long long n, unsigned long long seed, int mode, unsigned long long* mismatches, double* first_bad,
                           cudaStream_t s);
}

extern "C" int warpii_gpu_check_division(int device, int64_t n, uint64_t seed, int mode, int64_t* mismatches_out,
                                         double first_bad_out[4]) {
    if (!mismatches_out) return fail("null argument");
    if (mode < 0 || mode > 3) return fail("check_division: mode %d out of range", mode);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device available");
    CUDA_OK(cudaSetDevice(device));
    unsigned long long* d_count = nullptr;
    double* d_bad = nullptr;
    if (upload<unsigned long long>(&d_count, nullptr, 1) || upload<double>(&d_bad, nullptr, 4)) return 1;
    wgpu::launch_division_check((long long)n, (unsigned long long)seed, mode, d_count, d_bad, nullptr);
    CUDA_OK(cudaDeviceSynchronize());
    unsigned long long count = 0;
    double bad[4] = {0, 0, 0, 0};
    CUDA_OK(cudaMemcpy(&count, d_count, sizeof count, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost));
    cudaFree(d_count);
    cudaFree(d_bad);
    *mismatches_out = (int64_t)count;
    if (first_bad_out) for (int i = 0; i < 4; i++) first_bad_out[i] = bad[i];
    return 0;
}

extern "C" int warpii_gpu_sm_clock_probes(warpii_gpu_ctx* c, double* mhz_out, int max_out, int* n_out) {
    if (!c || !n_out) return fail("null argument");
    CUDA_OK(cudaStreamSynchronize(c->stream));
    const int n = c->n_probes < max_out ? c->n_probes : max_out;
    if (n > 0) CUDA_OK(cudaMemcpy(mhz_out, c->d_probe, n * sizeof(double), cudaMemcpyDeviceToHost));
    *n_out = n;
    c->n_probes = 0;
    return 0;
}
