// Host-side mirror of the reference's operator / integrator / time-loop interface, over the C ABI.
//
//   GpuSolutionVec                      <-> five_moment::FiveMSolutionVec            (solution_vec.h:45-51)
//   GpuFluidFluxESDGSEMOperator         <-> FluidFluxESDGSEMOperator<dim>             (fluid_flux_es_dgsem_operator.h:46-125)
//   SSPRK2Integrator<Number,Vec,Op>     <-> SSPRK2Integrator                          (rk.h:79-117)
//   TimestepCallback, advance()         <-> timestepper.h:9-28, timestepper.cc:6-56
//
// Same names, argument meaning and error behaviour (errors surface as C++ exceptions carrying the ABI's
// message, as AssertThrow does in the reference), so a caller written against the reference compiles against
// these with the type names swapped.  The state never leaves HBM between calls.
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "warpii_gpu.h"

namespace warpii_b200 {

enum ZeroOutPolicy { DO_ZERO_DST_VECTOR, DO_NOT_ZERO_DST_VECTOR };   // dof_utils.h:44-47

inline void check(int status) {
    if (status != 0) throw std::runtime_error(warpii_gpu_last_error());
}

// Owns the context; shared by all vectors / operators built on it.
class GpuContext {
   public:
    GpuContext(const warpii_gpu_mesh& mesh, int device) {
        n_vectors_ = mesh.n_vectors < 2 ? 2 : mesh.n_vectors;
        check(warpii_gpu_create(&mesh, device, &ctx_));
        in_use_.assign(n_vectors_, false);
        n_boundaries_ = mesh.n_boundaries;
        fe_degree_ = mesh.fe_degree;
    }
    ~GpuContext() { warpii_gpu_destroy(ctx_); }
    GpuContext(const GpuContext&) = delete;
    GpuContext& operator=(const GpuContext&) = delete;
    warpii_gpu_ctx* get() const { return ctx_; }
    int64_t n_dofs() const { return warpii_gpu_n_dofs(ctx_); }
    int n_boundaries() const { return n_boundaries_; }
    int fe_degree() const { return fe_degree_; }
    int acquire_vector() {
        for (int i = 0; i < n_vectors_; i++)
            if (!in_use_[i]) { in_use_[i] = true; return i; }
        throw std::runtime_error("GpuContext: all device vectors are in use; raise warpii_gpu_mesh::n_vectors");
    }
    void release_vector(int id) { if (id >= 0 && id < n_vectors_) in_use_[id] = false; }

   private:
    warpii_gpu_ctx* ctx_ = nullptr;
    int n_vectors_ = 2, n_boundaries_ = 0, fe_degree_ = 1;
    std::vector<bool> in_use_;
};

// A state vector resident in HBM (mesh_sol) with its boundary-integrated fluxes.
class GpuSolutionVec {
   public:
    GpuSolutionVec() = default;
    explicit GpuSolutionVec(std::shared_ptr<GpuContext> ctx) { bind(ctx); }
    ~GpuSolutionVec() { if (ctx_) ctx_->release_vector(id_); }
    GpuSolutionVec(const GpuSolutionVec&) = delete;
    GpuSolutionVec& operator=(const GpuSolutionVec&) = delete;

    // FiveMSolutionVec::reinit(other): same size and partitioning, zero contents (solution_vec.cc:5-8)
    void reinit(const GpuSolutionVec& other) {
        if (!ctx_) bind(other.ctx_);
        check(warpii_gpu_zero_state(ctx_->get(), id_));
    }
    void upload(const double* host, const int64_t* dof_index = nullptr) { check(warpii_gpu_upload_state(ctx_->get(), id_, host, dof_index)); }
    void download(double* host, const int64_t* dof_index = nullptr) const { check(warpii_gpu_download_state(ctx_->get(), id_, host, dof_index)); }
    std::vector<double> boundary_integrated_fluxes() const {
        std::vector<double> out((size_t)5 * ctx_->n_boundaries(), 0.0);
        if (!out.empty()) check(warpii_gpu_boundary_fluxes(ctx_->get(), id_, out.data()));
        return out;
    }
    int id() const { return id_; }
    // exchange the device storage of two vectors (what LinearAlgebra::distributed::Vector::swap does)
    void swap(GpuSolutionVec& other) { std::swap(ctx_, other.ctx_); std::swap(id_, other.id_); }
    bool bound() const { return (bool)ctx_; }
    const std::shared_ptr<GpuContext>& context() const { return ctx_; }

   private:
    void bind(std::shared_ptr<GpuContext> ctx) {
        ctx_ = std::move(ctx);
        id_ = ctx_->acquire_vector();
        check(warpii_gpu_zero_state(ctx_->get(), id_));
    }
    std::shared_ptr<GpuContext> ctx_;
    int id_ = -1;
};

// q5 = f(x[dim], t): conserved inflow state, the Function<dim> handed to EulerBCMap::set_inflow_boundary (bc_helper.h:52-64)
using InflowFunction = std::function<void(const double* x, double t, double* q5)>;

class GpuFluidFluxESDGSEMOperator {
   public:
    explicit GpuFluidFluxESDGSEMOperator(std::shared_ptr<GpuContext> ctx) : ctx_(std::move(ctx)) {}

    // dst = beta*dst + alpha*(u + dt*M^-1 R(u)).  sol_registers is accepted for signature parity; the reference
    // only uses them as zero-initialised scratch (fluid_flux_es_dgsem_operator.h:135-137), which the fused kernel
    // does not need.  A ZeroOutPolicy other than the default has no effect for the same reason.
    void perform_forward_euler_step(GpuSolutionVec& dst, const GpuSolutionVec& u, std::vector<GpuSolutionVec>& /*sol_registers*/,
                                    const double dt, const double t, const double alpha = 1.0, const double beta = 0.0,
                                    const ZeroOutPolicy /*zero_out_policy*/ = DO_NOT_ZERO_DST_VECTOR) {
        refresh_inflow(t);   // set_time(t) on every inflow function, :139-144
        // the second SSPRK2 stage is the one whose result recommend_dt is asked about next: fuse the CFL sweep there
        const int flags = (beta != 0.0) ? WARPII_FUSE_CFL : 0;
        check(warpii_gpu_forward_euler_step_ex(ctx_->get(), dst.id(), u.id(), dt, t, alpha, beta, flags));
    }

    // One low-storage Runge-Kutta stage, the interface LowStorageRungeKuttaIntegrator::perform_time_step drives (rk.h:53-71;
    // deal.II step-67's EulerOperator::perform_stage, tutorial-67.cc:880-899):  k = M^-1 R(current_ri);
    // next_ri = solution + factor_ai * k;  solution += factor_solution * k  (next_ri from the solution BEFORE its update).  vec_ki is the third register: the fused kernel never stores k, but it must not
    // write into the vector whose neighbour traces it is reading, so the register is used to avoid exactly that and the
    // vectors trade storage (swap) to end up where the caller expects them.
    void perform_stage(const double current_time, const double factor_solution, const double factor_ai,
                       GpuSolutionVec& current_ri, GpuSolutionVec& vec_ki, GpuSolutionVec& solution, GpuSolutionVec& next_ri) {
        refresh_inflow(current_time);
        warpii_gpu_ctx* c = ctx_->get();
        if (&current_ri == &solution) {            // first stage: (solution, vec_ri, solution, vec_ri) -- vec_ki IS next_ri here
            GpuSolutionVec& spare = (&vec_ki != &next_ri) ? vec_ki : lsrk_register(solution);
            check(warpii_gpu_lsrk_stage(c, spare.id(), next_ri.id(), solution.id(), solution.id(), factor_solution, factor_ai, current_time));
            solution.swap(spare);
        } else if (&current_ri == &next_ri) {      // later stages: (vec_ri, vec_ki, solution, vec_ri)
            check(warpii_gpu_lsrk_stage(c, solution.id(), vec_ki.id(), solution.id(), current_ri.id(), factor_solution, factor_ai, current_time));
            next_ri.swap(vec_ki);
        } else {
            check(warpii_gpu_lsrk_stage(c, solution.id(), next_ri.id(), solution.id(), current_ri.id(), factor_solution, factor_ai, current_time));
        }
    }

    // the reference also takes the MatrixFree object; its role is played by the context
    double recommend_dt(const GpuSolutionVec& sol) {
        double dt = 0.0;
        check(warpii_gpu_recommend_dt(ctx_->get(), sol.id(), &dt));
        return dt;
    }

    void set_inflow(int species, int boundary_id, const double q[5]) {
        check(warpii_gpu_set_inflow(ctx_->get(), species, boundary_id, q));
        constants_.push_back({species, boundary_id, {{q[0], q[1], q[2], q[3], q[4]}}});
        if (species >= 0 && species < (int)tables_.size() && !tables_[species].empty()) fill_constant(tables_[species], constants_.back());
    }

    // Two-fluid source terms (north_star kernel 4; not in the reference operator, off by default): see warpii_gpu_set_sources.
    // Perfectly hyperbolic Maxwell fluxes for the field components (not in the reference, off by default): warpii_gpu_set_maxwell.
    void set_maxwell(bool enabled, double light_speed, double chi, double gamma) {
        check(warpii_gpu_set_maxwell(ctx_->get(), enabled ? 1 : 0, light_speed, chi, gamma));
    }
    void set_sources(bool enabled, double epsilon0, double chi, const std::vector<double>& charge_over_mass) {
        check(warpii_gpu_set_sources(ctx_->get(), enabled ? 1 : 0, epsilon0, chi, charge_over_mass.data()));
    }

    // Where the inflow functions are evaluated: xyz[face][point][dim] of this rank's boundary quadrature points and the
    // boundary id of every face, both in the order of warpii_gpu_mesh.boundary_face_* (the solver provides them).
    void set_boundary_points(std::vector<double> xyz, std::vector<int32_t> face_boundary_id, int dim, int n_species) {
        bxyz_ = std::move(xyz);
        bid_ = std::move(face_boundary_id);
        dim_ = dim;
        tables_.assign(n_species, {});
    }
    // Space/time-dependent inflow (EulerBCMap::set_inflow_boundary).  time_dependent = false tabulates the function
    // once, at the first stage; true re-tabulates it at every stage time and rules out the device-resident time loop.
    void set_inflow_function(int species, int boundary_id, InflowFunction f, bool time_dependent = true) {
        if (species < 0 || species >= (int)tables_.size()) throw std::invalid_argument("set_inflow_function: species out of range");
        inflow_.push_back({species, boundary_id, std::move(f), time_dependent});
        stale_ = true;
    }
    bool has_time_dependent_inflow() const {
        for (const auto& e : inflow_)
            if (e.time_dependent) return true;
        return false;
    }
    // Tabulate the inflow functions at time t and upload (no-op without functions, or when nothing changed).
    void refresh_inflow(double t) {
        if (inflow_.empty()) return;
        if (!stale_ && (!has_time_dependent_inflow() || t == table_time_)) return;
        const size_t n_faces = bid_.size();
        if (n_faces == 0) { stale_ = false; table_time_ = t; return; }
        const size_t nq = bxyz_.size() / (n_faces * dim_);
        std::vector<bool> touched(tables_.size(), false);
        for (const auto& e : inflow_) {
            if (!stale_ && !e.time_dependent) continue;
            std::vector<double>& tab = tables_[e.species];
            if (tab.empty()) {
                tab.assign(n_faces * nq * 5, 0.0);
                for (const auto& k : constants_)
                    if (k.species == e.species) fill_constant(tab, k);
            }
            for (size_t f = 0; f < n_faces; f++) {
                if (bid_[f] != e.boundary_id) continue;
                for (size_t q = 0; q < nq; q++) e.f(&bxyz_[(f * nq + q) * dim_], t, &tab[(f * nq + q) * 5]);
            }
            touched[e.species] = true;
        }
        for (size_t s = 0; s < tables_.size(); s++)
            if (touched[s]) check(warpii_gpu_set_inflow_table(ctx_->get(), (int)s, tables_[s].data()));
        stale_ = false;
        table_time_ = t;
    }

   private:
    // a register of the operator's own for the first low-storage stage (HBM is taken when it is first used)
    GpuSolutionVec& lsrk_register(const GpuSolutionVec& like) {
        if (!lsrk_register_) {
            lsrk_register_ = std::make_unique<GpuSolutionVec>();
            lsrk_register_->reinit(like);
        }
        return *lsrk_register_;
    }
    std::unique_ptr<GpuSolutionVec> lsrk_register_;
    struct InflowEntry {
        int species, boundary_id;
        InflowFunction f;
        bool time_dependent;
    };
    struct InflowConstant {
        int species, boundary_id;
        std::array<double, 5> q;
    };
    void fill_constant(std::vector<double>& tab, const InflowConstant& k) const {
        const size_t n_faces = bid_.size(), nq = n_faces ? tab.size() / (n_faces * 5) : 0;
        for (size_t f = 0; f < n_faces; f++)
            if (bid_[f] == k.boundary_id)
                for (size_t q = 0; q < nq; q++)
                    for (int c = 0; c < 5; c++) tab[(f * nq + q) * 5 + c] = k.q[c];
    }
    std::shared_ptr<GpuContext> ctx_;
    std::vector<InflowEntry> inflow_;
    std::vector<InflowConstant> constants_;
    std::vector<double> bxyz_;
    std::vector<int32_t> bid_;
    std::vector<std::vector<double>> tables_;   // per species [face][point][5]
    int dim_ = 1;
    bool stale_ = false;
    double table_time_ = 0.0;
};

// rk.h:10-77.  The reference takes the coefficients from deal.II's TimeStepping::LowStorageRungeKutta::get_coefficients;
// deal.II is not available here, so the published coefficients (Kennedy, Carpenter & Lewis 2000; Tselios & Simos 2007) are
// restated and pinned by what defines them: every order condition of the scheme's order holds to round-off (8 conditions
// for order 4, 17 for order 5) and the observed order in time matches (tests/test_lsrk_cpu.py).  The last weight of the
// 7-stage scheme is written as 1 - sum(b_i), which is what consistency demands of it.
enum LowStorageRungeKuttaScheme {
    stage_3_order_3, /* Kennedy, Carpenter, Lewis, 2000 */
    stage_5_order_4, /* Kennedy, Carpenter, Lewis, 2000 */
    stage_7_order_4, /* Tselios, Simos, 2007 */
    stage_9_order_5, /* Kennedy, Carpenter, Lewis, 2000 */
};

class LowStorageRungeKuttaIntegrator {
   public:
    explicit LowStorageRungeKuttaIntegrator(const LowStorageRungeKuttaScheme scheme) {
        switch (scheme) {
            case stage_3_order_3:
                bi = {0.245170287303492, 0.184896052186740, 0.569933660509768};
                ai = {0.755726351946097, 0.386954477304099};
                break;
            case stage_5_order_4:
                bi = {1153189308089. / 22510343858157., 1772645290293. / 4653164025191., -1672844663538. / 4480602732383.,
                      2114624349019. / 3568978502595., 5198255086312. / 14908931495163.};
                ai = {970286171893. / 4311952581923., 6584761158862. / 12103376702013., 2251764453980. / 15575788980749.,
                      26877169314380. / 34165994151039.};
                break;
            case stage_7_order_4: {
                bi = {0.0941840925477795334, 0.149683694803496998, 0.285204742060440058, -0.122201846148053668,
                      0.0605151571191401122, 0.345986987898399296, 0.0};
                bi[6] = 1.0 - (((((bi[0] + bi[1]) + bi[2]) + bi[3]) + bi[4]) + bi[5]);
                const double gi[6] = {0.241566650129646868, 0.0423866513027719953, 0.215602732678803776,
                                      0.232328007537583987, 0.256223412574146438, 0.0978694102142697230};
                ai.resize(6);
                for (int i = 0; i < 6; i++) ai[i] = gi[i] + bi[i];
                break;
            }
            case stage_9_order_5:
                bi = {2274579626619. / 23610510767302., 693987741272. / 12394497460941., -347131529483. / 15096185902911.,
                      1144057200723. / 32081666971178., 1562491064753. / 11797114684756., 13113619727965. / 44346030145118.,
                      393957816125. / 7825732611452., 720647959663. / 6565743875477., 3559252274877. / 14424734981077.};
                ai = {1107026461565. / 5417078080134., 38141181049399. / 41724347789894., 493273079041. / 11940823631197.,
                      1851571280403. / 6147804934346., 11782306865191. / 62590030070788., 9452544825720. / 13648368537481.,
                      4435885630781. / 26285702406235., 2357909744247. / 11371140753790.};
                break;
            default:
                throw std::runtime_error("ExcNotImplemented");   // rk.h:43-44
        }
        // c_i = row sums of the 2-register Butcher tableau: A[i][j] = b_j (j < i-1), A[i][i-1] = a_{i-1}
        ci.assign(bi.size(), 0.0);
        for (size_t i = 1; i < bi.size(); i++) {
            double c = ai[i - 1];
            for (size_t j = 0; j + 1 < i; j++) c += bi[j];
            ci[i] = c;
        }
    }

    unsigned int n_stages() const { return (unsigned int)bi.size(); }
    const std::vector<double>& get_bi() const { return bi; }
    const std::vector<double>& get_ai() const { return ai; }
    const std::vector<double>& get_ci() const { return ci; }

    template <typename VectorType, typename Operator>
    void perform_time_step(Operator& pde_operator, const double current_time, const double time_step, VectorType& solution,
                           VectorType& vec_ri, VectorType& vec_ki) const {
        pde_operator.perform_stage(current_time, bi[0] * time_step, ai[0] * time_step, solution, vec_ri, solution, vec_ri);
        for (unsigned int stage = 1; stage < bi.size(); ++stage) {
            const double c_i = ci[stage];
            pde_operator.perform_stage(current_time + c_i * time_step, bi[stage] * time_step,
                                       (stage == bi.size() - 1 ? 0 : ai[stage] * time_step), vec_ri, vec_ki, solution, vec_ri);
        }
    }

   private:
    std::vector<double> bi;
    std::vector<double> ai;
    std::vector<double> ci;
};

// rk.h:79-117
template <typename Number, typename SolutionVec, typename Operator>
class SSPRK2Integrator {
   public:
    SSPRK2Integrator() {}

    void evolve_one_time_step(Operator& forward_euler_operator, SolutionVec& solution, const double dt, const double t) {
        forward_euler_operator.perform_forward_euler_step(f_1, solution, sol_registers, dt, t);
        forward_euler_operator.perform_forward_euler_step(solution, f_1, sol_registers, dt, t + dt, 0.5, 0.5);
    }

    // The GPU operator needs no scratch registers, so sol_register_count device vectors are NOT allocated:
    // at C4 size each would cost 10.5 GB of HBM for nothing.
    void reinit(const SolutionVec& sol, int /*sol_register_count*/) { f_1.reinit(sol); }
    const SolutionVec& stage_vector() const { return f_1; }

   private:
    SolutionVec f_1;
    std::vector<SolutionVec> sol_registers;
};

// timestepper.h:9-21
struct TimestepCallback {
    TimestepCallback(double interval, std::function<void(double)> callback, bool perform_zeroth = true, bool perform_final = true)
        : interval(interval), callback(std::move(callback)), perform_zeroth(perform_zeroth), perform_final(perform_final) {}
    double interval;
    std::function<void(double t)> callback;
    bool perform_zeroth;   // fire at t = 0
    bool perform_final;    // fire at t_end even if not scheduled
};

// The outer structure of advance() (timestepper.cc:6-56): run to the next due callback or t_end, fire, repeat.
// run_to(t, stop) advances the state from t to stop (within 1e-12) and returns the time it reached.
inline void advance_segments(const std::function<double(double t, double stop)>& run_to, double t_end,
                             std::vector<TimestepCallback>& callbacks) {
    const double slack = 1e-12;
    double t = 0.0;
    std::vector<double> due(callbacks.size());
    for (size_t i = 0; i < callbacks.size(); i++) {
        if (callbacks[i].perform_zeroth) callbacks[i].callback(0.0);
        due[i] = t + callbacks[i].interval;
    }
    while (t < t_end - slack) {
        size_t next = 0;
        for (size_t i = 1; i < due.size(); i++)
            if (due[i] < due[next]) next = i;
        const double next_due = callbacks.empty() ? t_end : due[next];
        const bool fire = next_due < t_end && std::fabs(next_due - t_end) > slack;
        const double stop = std::fmin(next_due, t_end);
        t = run_to(t, stop);
        if (fire) {
            callbacks[next].callback(t);
            due[next] = t + callbacks[next].interval;
        }
    }
    for (size_t i = 0; i < callbacks.size(); i++)
        if (std::fabs(t_end - due[i]) < slack || callbacks[i].perform_final) callbacks[i].callback(t_end);
}

// timestepper.cc:6-56: adaptive dt clipped to the next callback / end time, 1e-12 of slack against stutter steps
inline void advance(std::function<bool(double t, double dt)> step, double t_end, std::function<double()> recommend_dt,
                    std::vector<TimestepCallback>& callbacks) {
    const double slack = 1e-12;
    advance_segments(
        [&](double t, double stop) {
            while (t < stop - slack) {
                const double dt = std::fmin(recommend_dt(), stop - t);
                if (step(t, dt)) t += dt;
            }
            return t;
        },
        t_end, callbacks);
}

}  // namespace warpii_b200
