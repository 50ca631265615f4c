// ============================================================================
// oracle/dgsem_oracle.cc -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
// See dgsem_oracle.h for scope, parity status and layout conventions.
//
// Every function cites the reference file:line whose arithmetic it restates
// (paths relative to the reference root).  Operation ORDER follows the
// reference expression by expression so that round-off is as close to the
// reference's as a deal.II-free program can get; compile with
// -ffp-contract=off so no FMA contraction sneaks in.
// ============================================================================
#include "dgsem_oracle.h"
#include "det_log.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// Point physics.  State q = [rho, mx, my, mz, E] (always three momenta).
// ---------------------------------------------------------------------------

// The reference calls std::log here.  1 = the deterministic logarithm of oracle/det_log.h (default: it is what makes
// a 1e-12 comparison with another platform meaningful, see that header), 0 = this platform's libm.
int g_log_impl = 1;
// liboracle_fma.so (oracle/Makefile) is this file compiled with -ffp-contract=fast, i.e. with the fused multiply-adds a
// default GCC build of the reference would contain, EXCEPT in the two places the parity runs pin bit for bit: the
// logarithm and the q -> p chain feeding ln_avg.  It exists to MEASURE how far two legitimate FP64 builds of the same
// algorithm are apart (tests/test_oracle_golden.py::test_oracle_rounding_floor); nothing is checked against it.
#ifdef ORACLE_FMA_VARIANT
#define ORACLE_PINNED __attribute__((noinline, optimize("fp-contract=off")))
#else
#define ORACLE_PINNED inline
#endif
ORACLE_PINNED double olog(double x) { return g_log_impl ? detlog::det_log(x) : std::log(x); }

// src/five_moment/euler.h:118-125
inline double ln_avg(double a, double b) {
    double diff_log = std::fabs(olog(b) - olog(a));
    const double C = 1e6;
    // std::max, not fmax: a NaN logarithm (negative pressure) must poison the flux as it does in the reference
    double lhs = std::max(C * std::fabs(b - a), b + a);
    double denom = std::max(C * diff_log, 2.0);
    return lhs / denom;
}

// src/five_moment/euler.h:32-44
ORACLE_PINNED double pressure(const double* q, double gamma) {
    const double rho = q[0];
    double squared_momentum = 0.0;
    for (int d = 0; d < 3; d++) squared_momentum += q[d + 1] * q[d + 1];
    const double kinetic_energy = squared_momentum / (2. * rho);
    return (gamma - 1) * (q[4] - kinetic_energy);
}

// src/five_moment/euler.h:14-25
template <int N>
inline void velocity(const double* q, double* v) {
    const double inverse_density = 1. / q[0];
    for (int d = 0; d < N; d++) v[d] = q[d + 1] * inverse_density;
}

// src/five_moment/euler.h:46-63.  F[c][d], d < dim.
template <int dim>
inline void euler_flux(const double* q, double gamma, double F[5][dim]) {
    double v[dim];
    velocity<dim>(q, v);
    const double p = pressure(q, gamma);
    for (int d = 0; d < dim; d++) {
        F[0][d] = q[d + 1];
        for (int e = 0; e < 3; e++) F[e + 1][d] = q[e + 1] * v[d];
        F[d + 1][d] += p;
        F[4][d] = v[d] * (q[4] + p);
    }
}

template <int dim>
inline void flux_dot(const double F[5][dim], const double* n, double* out) {
    // tensor_utils.h:9-17 (Tensor<1,dim> * Tensor<1,dim> accumulates in index order)
    for (int c = 0; c < 5; c++) {
        double s = F[c][0] * n[0];
        for (int d = 1; d < dim; d++) s += F[c][d] * n[d];
        out[c] = s;
    }
}

// src/five_moment/euler.h:68-89 (Lax-Friedrichs, boundary faces only)
template <int dim>
inline void lf_flux(const double* qin, const double* qout, const double* n, double gamma, double* out) {
    double uin[dim], uout[dim];
    velocity<dim>(qin, uin);
    velocity<dim>(qout, uout);
    const double pin = pressure(qin, gamma), pout = pressure(qout, gamma);
    double Fin[5][dim], Fout[5][dim];
    euler_flux<dim>(qin, gamma, Fin);
    euler_flux<dim>(qout, gamma, Fout);
    double nsq_in = 0, nsq_out = 0;
    for (int d = 0; d < dim; d++) { nsq_in += uin[d] * uin[d]; nsq_out += uout[d] * uout[d]; }
    const double lambda = 0.5 * std::sqrt(std::max(nsq_out + gamma * pout / qout[0], nsq_in + gamma * pin / qin[0]));
    double fin[5], fout[5];
    flux_dot<dim>(Fin, n, fin);
    flux_dot<dim>(Fout, n, fout);
    for (int c = 0; c < 5; c++) out[c] = 0.5 * (fin[c] + fout[c]) + 0.5 * lambda * (qin[c] - qout[c]);
}

// src/five_moment/euler.h:186-228 (Chandrashekar entropy-conserving two-point flux)
template <int dim>
inline void ec_flux(const double* qj, const double* ql, double gamma, double F[5][dim]) {
    const double p_j = pressure(qj, gamma);
    const double beta_j = qj[0] / (2.0 * p_j);
    const double p_l = pressure(ql, gamma);
    const double beta_l = ql[0] / (2.0 * p_l);
    const double beta_avg = 0.5 * (beta_j + beta_l);
    const double beta_ln = ln_avg(beta_j, beta_l);
    const double rho_j = qj[0], rho_l = ql[0];
    double u_j[3], u_l[3];
    velocity<3>(qj, u_j);
    velocity<3>(ql, u_l);
    const double rho_ln = ln_avg(rho_j, rho_l);
    const double rho_avg = 0.5 * (rho_j + rho_l);
    double u_avg[3], u2_avg[3], u_avg_2[3];
    for (int e = 0; e < 3; e++) {
        u_avg[e] = 0.5 * (u_j[e] + u_l[e]);
        u2_avg[e] = 0.5 * (u_j[e] * u_j[e] + u_l[e] * u_l[e]);
        u_avg_2[e] = u_avg[e] * u_avg[e];
    }
    double sum_u2_avg = 0.0, sum_u_avg_2 = 0.0;
    for (int e = 0; e < 3; e++) { sum_u2_avg += u2_avg[e]; sum_u_avg_2 += u_avg_2[e]; }
    const double p_hat = rho_avg / (2.0 * beta_avg);
    const double h_hat = 1.0 / (2.0 * beta_ln * (gamma - 1.0)) - 0.5 * sum_u2_avg + p_hat / rho_ln + sum_u_avg_2;
    for (int d = 0; d < dim; d++) {
        F[0][d] = rho_ln * u_avg[d];
        for (int e = 0; e < 3; e++) F[e + 1][d] = rho_ln * u_avg[d] * u_avg[e];
        F[d + 1][d] += p_hat;
        F[4][d] = rho_ln * u_avg[d] * h_hat;
    }
}

// src/five_moment/euler.h:232-284 (EC flux + matrix-free scalar dissipation); j = inside, l = outside
template <int dim>
inline void es_flux(const double* qj, const double* ql, const double* n, double gamma, double* flux) {
    const double p_j = pressure(qj, gamma);
    const double beta_j = qj[0] / (2.0 * p_j);
    const double p_l = pressure(ql, gamma);
    const double beta_l = ql[0] / (2.0 * p_l);
    const double beta_ln = ln_avg(beta_j, beta_l);
    const double rho_j = qj[0], rho_l = ql[0];
    double u_j[3], u_l[3];
    velocity<3>(qj, u_j);
    velocity<3>(ql, u_l);
    const double rho_avg = 0.5 * (rho_j + rho_l);
    double u_avg[3], u2_j[3], u2_l[3];
    for (int e = 0; e < 3; e++) {
        u_avg[e] = 0.5 * (u_j[e] + u_l[e]);
        u2_j[e] = u_j[e] * u_j[e];
        u2_l[e] = u_l[e] * u_l[e];
    }
    double F[5][dim];
    ec_flux<dim>(qj, ql, gamma, F);
    flux_dot<dim>(F, n, flux);

    const double rho_jump = rho_l - rho_j;
    double rho_u_jump[3], u_jump[3];
    for (int e = 0; e < 3; e++) {
        rho_u_jump[e] = rho_l * u_l[e] - rho_j * u_j[e];
        u_jump[e] = u_l[e] - u_j[e];
    }
    const double c_j = std::sqrt(gamma * p_j / rho_j);
    const double c_l = std::sqrt(gamma * p_l / rho_l);
    double s_j = 0.0, s_l = 0.0;
    for (int e = 0; e < 3; e++) { s_j += u2_j[e]; s_l += u2_l[e]; }
    const double lambda_max = std::max(std::sqrt(s_j) + c_j, std::sqrt(s_l) + c_l);
    const double beta_inv_jump = 1.0 / beta_l - 1.0 / beta_j;
    double sum_u_prod = 0.0, sum_jump_avg = 0.0;
    for (int e = 0; e < 3; e++) sum_u_prod += u_j[e] * u_l[e];
    for (int e = 0; e < 3; e++) sum_jump_avg += u_jump[e] * u_avg[e];
    const double energy_stab = (1.0 / (2.0 * (gamma - 1.0) * beta_ln) + 0.5 * sum_u_prod) * rho_jump +
                               rho_avg * sum_jump_avg + rho_avg / (2.0 * (gamma - 1.0)) * beta_inv_jump;
    flux[0] -= 0.5 * lambda_max * rho_jump;
    for (int e = 0; e < 3; e++) flux[e + 1] -= 0.5 * lambda_max * rho_u_jump[e];
    flux[4] -= 0.5 * lambda_max * energy_stab;
}

// src/five_moment/euler.h:135-169
inline double specific_entropy(const double* q, double gamma) {
    return std::log(pressure(q, gamma)) - gamma * std::log(q[0]);
}

// ---------------------------------------------------------------------------
// Reference element (deal.II QGaussLobatto / QGauss / FE_DGQ::shape_grad on [0,1]);
// computed in long double and rounded once.  SURVEY.md 9.1.
// ---------------------------------------------------------------------------
long double legendre_ld(int k, long double x) {
    if (k == 0) return 1.0L;
    long double p0 = 1.0L, p1 = x;
    for (int n = 1; n < k; n++) {
        long double p2 = ((2 * n + 1) * x * p1 - n * p0) / (n + 1);
        p0 = p1;
        p1 = p2;
    }
    return p1;
}

void gll_ld(int N, std::vector<long double>& x, std::vector<long double>& w) {
    x.assign(N, 0);
    w.assign(N, 0);
    const int n = N - 1;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < N; i++) {
        long double xi = -std::cos(pi * i / n);
        if (i == 0) xi = -1.0L;
        else if (i == n) xi = 1.0L;
        else {
            for (int it = 0; it < 100; it++) {
                long double pn = legendre_ld(n, xi), pnm1 = legendre_ld(n - 1, xi);
                long double dx = (xi * pn - pnm1) / ((n + 1) * pn);
                xi -= dx;
                if (std::fabs((double)dx) < 1e-19) break;
            }
        }
        long double pn = legendre_ld(n, xi);
        x[i] = xi;
        w[i] = 2.0L / (n * (n + 1) * pn * pn);
    }
    // symmetrise, then map [-1,1] -> [0,1]
    for (int i = 0; i < N / 2; i++) {
        long double a = 0.5L * (x[N - 1 - i] - x[i]);
        x[i] = -a;
        x[N - 1 - i] = a;
        long double ww = 0.5L * (w[i] + w[N - 1 - i]);
        w[i] = w[N - 1 - i] = ww;
    }
    if (N % 2) x[N / 2] = 0.0L;
    for (int i = 0; i < N; i++) {
        x[i] = 0.5L * (x[i] + 1.0L);
        w[i] *= 0.5L;
    }
}

void gauss_ld(int N, std::vector<long double>& x, std::vector<long double>& w) {
    x.assign(N, 0);
    w.assign(N, 0);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < N; i++) {
        long double xi = -std::cos(pi * (i + 0.75L) / (N + 0.5L));
        long double dp = 0;
        for (int it = 0; it < 100; it++) {
            long double pn = legendre_ld(N, xi), pnm1 = legendre_ld(N - 1, xi);
            dp = N * (xi * pn - pnm1) / (xi * xi - 1.0L);
            long double dx = pn / dp;
            xi -= dx;
            if (std::fabs((double)dx) < 1e-19) break;
        }
        long double pn = legendre_ld(N, xi), pnm1 = legendre_ld(N - 1, xi);
        dp = N * (xi * pn - pnm1) / (xi * xi - 1.0L);
        x[i] = xi;
        w[i] = 2.0L / ((1.0L - xi * xi) * dp * dp);
    }
    for (int i = 0; i < N / 2; i++) {
        long double a = 0.5L * (x[N - 1 - i] - x[i]);
        x[i] = -a;
        x[N - 1 - i] = a;
        long double ww = 0.5L * (w[i] + w[N - 1 - i]);
        w[i] = w[N - 1 - i] = ww;
    }
    if (N % 2) x[N / 2] = 0.0L;
    for (int i = 0; i < N; i++) {
        x[i] = 0.5L * (x[i] + 1.0L);
        w[i] *= 0.5L;
    }
}

struct Basis {
    int Np = 0;
    std::vector<double> x, w;    // GLL(Np) on [0,1]
    std::vector<double> D, Q;    // D[j*Np+l] = l_l'(x_j);  Q = diag(w) D   (split_form_volume_flux.h:36-45, subcell_finite_volume_flux.h:36-52)
    int Ng = 0;
    std::vector<double> xg, wg;  // Gauss(p+2) on [0,1]  (nodal_dg_discretization.cc:12-13, quadrature 0)
    std::vector<double> Ig;      // Ig[q*Np+i] = l_i(xg_q)
    std::vector<double> V;       // Legendre analysis, V[k*Np+q]  

    void init(int Np_) {
        Np = Np_;
        std::vector<long double> xl, wl;
        gll_ld(Np, xl, wl);
        x.resize(Np);
        w.resize(Np);
        for (int i = 0; i < Np; i++) { x[i] = (double)xl[i]; w[i] = (double)wl[i]; }
        // barycentric weights
        std::vector<long double> lam(Np, 1.0L);
        for (int j = 0; j < Np; j++)
            for (int l = 0; l < Np; l++)
                if (l != j) lam[j] /= (xl[j] - xl[l]);
        D.assign(Np * Np, 0.0);
        Q.assign(Np * Np, 0.0);
        for (int j = 0; j < Np; j++) {
            long double diag = 0.0L;
            for (int l = 0; l < Np; l++) {
                if (l == j) continue;
                long double d = (lam[l] / lam[j]) / (xl[j] - xl[l]);
                D[j * Np + l] = (double)d;
                diag -= d;
            }
            D[j * Np + j] = (double)diag;
            // interior diagonal entries of the GLL differentiation matrix vanish identically
            if (j != 0 && j != Np - 1) D[j * Np + j] = 0.0;
        }
        for (int j = 0; j < Np; j++)
            for (int l = 0; l < Np; l++) Q[j * Np + l] = w[j] * D[j * Np + l];
        Ng = Np + 1;  // fe_degree + 2
        std::vector<long double> xgl, wgl;
        gauss_ld(Ng, xgl, wgl);
        xg.resize(Ng);
        wg.resize(Ng);
        for (int i = 0; i < Ng; i++) { xg[i] = (double)xgl[i]; wg[i] = (double)wgl[i]; }
        Ig.assign(Ng * Np, 0.0);
        for (int q = 0; q < Ng; q++)
            for (int i = 0; i < Np; i++) {
                long double v = 1.0L;
                for (int m = 0; m < Np; m++)
                    if (m != i) v *= (xgl[q] - xl[m]) / (xl[i] - xl[m]);
                Ig[q * Np + i] = (double)v;
            }
        V.assign(Np * Np, 0.0);
        for (int k = 0; k < Np; k++)
            for (int q = 0; q < Np; q++)
                V[k * Np + q] = (double)((k + 0.5L) * wl[q] * std::sqrt(2.0L) * legendre_ld(k, 2.0L * xl[q] - 1.0L));
    }
};

// ---------------------------------------------------------------------------
// Operator context on a Cartesian box (src/grid_descriptions.cc:51-74).
// ---------------------------------------------------------------------------
struct Ctx {
    int dim = 1, p = 1, Np = 2, NN = 2, nsp = 1, nc = 5, nfaceN = 1, nfaceG = 1;
    double gamma = 5.0 / 3.0;
    int nx[3] = {1, 1, 1};
    double left[3] = {0, 0, 0}, right[3] = {1, 1, 1}, h[3] = {1, 1, 1};
    int periodic[3] = {1, 1, 1};
    int64_t nelem = 1;
    int nthreads = 1;
    Basis B;
    std::vector<int> bc_kind;     // [nsp][2*dim]
    std::vector<double> inflow;   // [nsp][2*dim][5]
    // space/time-dependent inflow: the Function<dim> of EulerBCMap::get_inflow, evaluated at the boundary quadrature
    // points with the stage time (fluid_flux_es_dgsem_operator.h:139-144, 381-384)
    struct InflowFn { orc_inflow_fn fn = nullptr; void* user = nullptr; };
    std::vector<InflowFn> inflow_fn;            // [nsp][2*dim]
    std::vector<int64_t> bface_of;              // [nelem][2*dim] -> boundary face number or -1
    int64_t n_bfaces = 0;
    mutable std::vector<double> inflow_vals;    // [nsp][n_bfaces][nG][5], refreshed at the start of every RHS
    // two-fluid source terms (off by default: the reference has none)
    bool sources_on = false;
    double epsilon0 = 1.0, chi = 0.0;
    std::vector<double> charge_over_mass;       // [nsp]
    // perfectly hyperbolic Maxwell fluxes for the 8 field components (off by default: the reference evolves nothing there)
    bool maxwell_on = false;
    double light_speed = 1.0, mx_chi = 0.0, mx_gamma = 0.0;
    // Cartesian geometry, formed the way the reference forms it from inverse_jacobian(q)
    double Jinv[3] = {1, 1, 1};   // diagonal of J^{-T}
    double Jdet = 1;              // jacobian_utils.h:12-18: 1/det(Jinv)
    double Ja[3] = {1, 1, 1};     // jacobian_utils.h:32-40: Jdet * Jinv[d][d]  (d-th comp of Ja^d)
    double face_area[3] = {1, 1, 1};
    std::vector<double> wN;       // tensor GLL weights per node (prod over dims)
    std::vector<double> wF;       // tensor GLL weights per face node (dim-1)
    std::vector<double> wG;       // tensor Gauss weights per face Gauss point (dim-1)
    std::vector<double> Lfull;    // tensor Legendre analysis matrix [NN][NN] (mode k, node q)
    std::vector<int> shell;       // per mode: 1 if max_d k_d == Np-1
    double max_eig = 1;           // power-iteration result of :487-502 (constant on a Cartesian box)
    int nbnd = 2;                 // number of boundary ids (2*dim on a box: grid_descriptions.cc:55-57 colorize)
    // ---- general geometry (curved / unstructured elements: what MappingQ(p) + MatrixFree hand to the operator) -------
    // Supplied per node / per face point by orc_create_general; the operator code below reads geometry only through
    // jdet() / Jai() / jinv_mat(), which on a box return the constants above (bit-identical to the Cartesian path).
    bool general = false;
    std::vector<double> gJinv;    // [nelem][NN][dim][dim]  J^{-T} = FEEvaluation::inverse_jacobian(q)
    std::vector<double> gJdet;    // [nelem][NN]            jacobian_utils.h:12-18
    std::vector<double> geig;     // [nelem][NN]            power-iteration value of :487-502 (depends on geometry only)
    std::vector<int64_t> gnbr;    // [nelem][2*dim]         neighbour element, or -1 - (boundary face number)
    std::vector<int> gnbrf;       // [nelem][2*dim]         neighbour's local face + 8 * (tangential order reversed)
    std::vector<int> gbf_id;      // [n_bfaces]             boundary id of every boundary face
    std::vector<double> gfn;      // [nelem][2*dim][nF][dim] unit outward normal, FEFaceEvaluation::normal_vector(q), quadrature 1
    std::vector<double> gfJ;      // [nelem][2*dim][nF]      surface Jacobian: face JxW = gfJ * (tensor GLL weight)
    std::vector<double> gbn;      // [n_bfaces][nG][dim]     unit outward normal at the Gauss(p+2) points of boundary faces
    std::vector<double> gbJ;      // [n_bfaces][nG]          surface Jacobian there

    double jdet(int64_t e, int q) const { return general ? gJdet[(size_t)e * NN + q] : Jdet; }
    // J^{-T} at node q as a full matrix K[r][c]
    void jinv_mat(int64_t e, int q, double K[3][3]) const {
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) K[r][cc] = 0.0;
        if (general) {
            const double* g = &gJinv[((size_t)e * NN + q) * dim * dim];
            for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) K[r][cc] = g[r * dim + cc];
        } else {
            for (int d = 0; d < dim; d++) K[d][d] = Jinv[d];
        }
    }
    // scaled contravariant basis vector Ja^d = Jdet * column d of J^{-T}  (jacobian_utils.h:32-40)
    void Jai(int64_t e, int q, int d, double* out) const {
        if (general) {
            const double* g = &gJinv[((size_t)e * NN + q) * dim * dim];
            const double J = gJdet[(size_t)e * NN + q];
            for (int r = 0; r < dim; r++) out[r] = J * g[r * dim + d];
        } else {
            for (int r = 0; r < dim; r++) out[r] = (r == d) ? Ja[d] : 0.0;
        }
    }

    int64_t elem_index(const int* idx) const {
        int64_t e = 0;
        for (int d = dim - 1; d >= 0; d--) e = e * nx[d] + idx[d];
        return e;
    }
    void elem_coords(int64_t e, int* idx) const {
        for (int d = 0; d < dim; d++) { idx[d] = (int)(e % nx[d]); e /= nx[d]; }
        for (int d = dim; d < 3; d++) idx[d] = 0;
    }
    // neighbour across face f = 2*d+side; returns -1 - boundary_id on a non-periodic boundary
    int64_t neighbor(int64_t e, int f) const {
        if (general) return gnbr[(size_t)e * 2 * dim + f];
        int idx[3];
        elem_coords(e, idx);
        int d = f / 2, side = f % 2;
        int i = idx[d] + (side ? 1 : -1);
        if (i < 0 || i >= nx[d]) {
            if (!periodic[d]) return -1 - f;
            i = (i + nx[d]) % nx[d];
        }
        idx[d] = i;
        return elem_index(idx);
    }
    int stride(int d) const { return d == 0 ? 1 : (d == 1 ? Np : Np * Np); }
    // node index of face node t on face (d, side); tangential dims in increasing order
    int face_node(int d, int side, int t) const {
        int idx[3] = {0, 0, 0};
        for (int e = 0; e < dim; e++) {
            if (e == d) continue;
            idx[e] = t % Np;
            t /= Np;
        }
        idx[d] = side ? Np - 1 : 0;
        return idx[0] + Np * (idx[1] + Np * idx[2]);
    }
};

int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

void setup_geometry(Ctx& c) {
    double det = 1.0;
    for (int d = 0; d < c.dim; d++) {
        c.h[d] = (c.right[d] - c.left[d]) / c.nx[d];
        c.Jinv[d] = 1.0 / c.h[d];
    }
    // tensor_utils.h:70-82 (1x1, 2x2 determinants; the diagonal 3x3 case is the obvious extension)
    if (c.dim == 1) det = c.Jinv[0];
    else if (c.dim == 2) det = c.Jinv[0] * c.Jinv[1] - 0.0 * 0.0;
    else det = c.Jinv[0] * c.Jinv[1] * c.Jinv[2];
    c.Jdet = 1.0 / det;
    for (int d = 0; d < c.dim; d++) {
        c.Ja[d] = c.Jdet * c.Jinv[d];
        double a = 1.0;
        for (int e = 0; e < c.dim; e++) if (e != d) a *= c.h[e];
        c.face_area[d] = a;
    }
    const int Np = c.Np;
    c.wN.assign(c.NN, 1.0);
    for (int j = 0; j < c.NN; j++) {
        int t = j;
        double w = 1.0;
        for (int d = 0; d < c.dim; d++) { w *= c.B.w[t % Np]; t /= Np; }
        c.wN[j] = w;
    }
    c.nfaceN = ipow(Np, c.dim - 1);
    c.nfaceG = ipow(c.B.Ng, c.dim - 1);
    c.wF.assign(c.nfaceN, 1.0);
    for (int j = 0; j < c.nfaceN; j++) {
        int t = j;
        double w = 1.0;
        for (int d = 0; d < c.dim - 1; d++) { w *= c.B.w[t % Np]; t /= Np; }
        c.wF[j] = w;
    }
    c.wG.assign(c.nfaceG, 1.0);
    for (int j = 0; j < c.nfaceG; j++) {
        int t = j;
        double w = 1.0;
        for (int d = 0; d < c.dim - 1; d++) { w *= c.B.wg[t % c.B.Ng]; t /= c.B.Ng; }
        c.wG[j] = w;
    }
    // tensor Legendre analysis (persson_peraire_shock_indicator.h:12-23 + deal.II FESeries::Legendre)
    c.Lfull.assign((size_t)c.NN * c.NN, 0.0);
    c.shell.assign(c.NN, 0);
    for (int k = 0; k < c.NN; k++) {
        int kk[3] = {0, 0, 0}, t = k, kmax = 0;
        for (int d = 0; d < c.dim; d++) { kk[d] = t % Np; t /= Np; kmax = std::max(kmax, kk[d]); }
        c.shell[k] = (kmax == Np - 1);
        for (int q = 0; q < c.NN; q++) {
            int qq[3] = {0, 0, 0};
            t = q;
            for (int d = 0; d < c.dim; d++) { qq[d] = t % Np; t /= Np; }
            double v = 1.0;
            for (int d = 0; d < c.dim; d++) v *= c.B.V[kk[d] * Np + qq[d]];
            c.Lfull[(size_t)k * c.NN + q] = v;
        }
    }
    // fluid_flux_es_dgsem_operator.h:487-502: 5 power iterations on Jinv^T Jinv, start (1,..,1)
    {
        double ev[3] = {1, 1, 1};
        for (int it = 0; it < 5; it++) {
            double nrm = 0;
            for (int d = 0; d < c.dim; d++) { ev[d] = c.Jinv[d] * (c.Jinv[d] * ev[d]); nrm = std::max(nrm, std::fabs(ev[d])); }
            for (int d = 0; d < c.dim; d++) ev[d] /= nrm;
        }
        double num = 0, den = 0;
        for (int d = 0; d < c.dim; d++) { double j = c.Jinv[d] * ev[d]; num += j * j; den += ev[d] * ev[d]; }
        c.max_eig = std::sqrt(num / den);
    }
}

// ---------------------------------------------------------------------------
// Shock indicator: src/dgsem/persson_peraire_shock_indicator.h:44-123
// ---------------------------------------------------------------------------
double shock_indicator(const Ctx& c, const double* v /*[NN] = p*rho at nodes*/) {
    const int NN = c.NN, Np = c.Np;
    double g0 = 0.0, g1 = 0.0;   // sum of squares: modes below the shell / on the max-degree shell
    for (int k = 0; k < NN; k++) {
        double ck = 0.0;
        const double* L = &c.Lfull[(size_t)k * NN];
        for (int q = 0; q < NN; q++) ck += L[q] * v[q];
        if (c.shell[k]) g1 += ck * ck; else g0 += ck * ck;
    }
    // deal.II FESeries::process_coefficients(L2_norm, smallest_abs_coefficient = 1e-10): a group is
    // kept only if its norm exceeds the threshold; the indicator then squares the norm again (:92-95).
    double n0 = std::sqrt(g0), n1 = std::sqrt(g1);
    double total_energy = 0., total_energy_minus_1 = 0., top_mode = 0., top_mode_minus_1 = 0.;
    if (n0 > 1e-10) { double e = n0 * n0; total_energy += e; total_energy_minus_1 += e; }
    if (n1 > 1e-10) { double e = n1 * n1; top_mode_minus_1 += e; total_energy_minus_1 += e; total_energy += e; }
    // group "max_degree == Np" is unreachable (:78-80) => top_mode stays 0
    double E = std::max(top_mode / total_energy, top_mode_minus_1 / total_energy_minus_1);
    double T = 0.5 * std::pow(10.0, -1.8 * std::pow((double)Np, 0.25));
    double s = 9.21024;
    double alpha = 1.0 / (1.0 + std::exp(-s / T * (E - T)));
    double alpha_max = 0.5;
    if (alpha < 1e-3) alpha = 0.0;
    else if (alpha > alpha_max) alpha = alpha_max;
    return alpha;
}

template <int dim>
double cell_alpha(const Ctx& c, const double* ue /* [5][NN] of one species */) {
    // fluid_flux_es_dgsem_operator.h:282-293
    std::vector<double> v(c.NN);
    for (int j = 0; j < c.NN; j++) {
        double q[5];
        for (int k = 0; k < 5; k++) q[k] = ue[k * c.NN + j];
        v[j] = pressure(q, c.gamma) * q[0];
    }
    return shock_indicator(c, v.data());
}

// ---------------------------------------------------------------------------
// Cell residual: split_form_volume_flux.h:61-99 and subcell_finite_volume_flux.h:68-159
// R[5][NN] receives the INTEGRATED residual (value * JxW), as integrate_scatter does.
// ---------------------------------------------------------------------------
template <int dim>
void cell_residual(const Ctx& c, int64_t e_geo, const double* ue, double alpha, double* R) {
    const int Np = c.Np, NN = c.NN;
    const double gamma = c.gamma;
    const double* D = c.B.D.data();
    const double* Q = c.B.Q.data();
    const double* w1 = c.B.w.data();
    auto state = [&](int q, double* s) { for (int k = 0; k < 5; k++) s[k] = ue[k * NN + q]; };

    // ---- split-form volume term
    for (int d = 0; d < dim; d++) {
        const int st = c.stride(d);
        for (int qj = 0; qj < NN; qj++) {
            const double Jdet_j = c.jdet(e_geo, qj);
            double Jai_j[dim];
            c.Jai(e_geo, qj, d, Jai_j);
            const int j = (qj / st) % Np;
            const int base = qj - j * st;
            double uj[5];
            state(qj, uj);
            double flux_j[5] = {0, 0, 0, 0, 0};
            for (int l = 0; l < Np; l++) {
                const int ql = base + st * l;
                double Jai_avg[dim], Jai_l[dim];
                c.Jai(e_geo, ql, d, Jai_l);
                for (int e = 0; e < dim; e++) Jai_avg[e] = 0.5 * (Jai_j[e] + Jai_l[e]);
                double ul[5];
                state(ql, ul);
                const double d_jl = D[j * Np + l];
                double F[5][dim];
                ec_flux<dim>(uj, ul, gamma, F);
                const double s = 2.0 * d_jl;
                for (int k = 0; k < 5; k++) {
                    double acc = (s * F[k][0]) * Jai_avg[0];
                    for (int e = 1; e < dim; e++) acc += (s * F[k][e]) * Jai_avg[e];
                    flux_j[k] -= acc;
                }
            }
            const double JxW = c.jdet(e_geo, qj) * c.wN[qj];
            for (int k = 0; k < 5; k++) R[k * NN + qj] += ((1.0 - alpha) * flux_j[k] / Jdet_j) * JxW;
        }
    }

    // ---- subcell finite-volume term (computed for every cell, also when alpha == 0)
    std::vector<double> fd((size_t)Np * 5);
    for (int d = 0; d < dim; d++) {
        const int st = c.stride(d);
        for (int ps = 0; ps < NN; ps++) {
            if ((ps / st) % Np != 0) continue;   // pencil starts: nodes whose d-index is 0 (dof_utils.cc:77-96)
            std::fill(fd.begin(), fd.end(), 0.0);
            double n[dim], Jad_m[dim];
            c.Jai(e_geo, ps, d, n);
            const double Jdet_0 = c.jdet(e_geo, ps);
            double s0[5], F0[5][dim], f0[5];
            state(ps, s0);
            euler_flux<dim>(s0, gamma, F0);
            flux_dot<dim>(F0, n, f0);
            for (int k = 0; k < 5; k++) fd[0 * 5 + k] += alpha * f0[k] / w1[0] / Jdet_0;
            for (int i = 0; i < Np - 1; i++) {
                const int qi = ps + st * i;
                for (int m = 0; m < Np; m++) {
                    c.Jai(e_geo, ps + st * m, d, Jad_m);
                    for (int e = 0; e < dim; e++) n[e] += Q[i * Np + m] * Jad_m[e];
                }
                const double Jdet_i = c.jdet(e_geo, qi), Jdet_i_plus_1 = c.jdet(e_geo, qi + st);
                double nsq = 0;
                for (int e = 0; e < dim; e++) nsq += n[e] * n[e];
                const double nn = std::sqrt(nsq);
                double nhat[dim];
                for (int e = 0; e < dim; e++) nhat[e] = n[e] / nn;
                double sl[5], sr[5], fl[5];
                state(qi, sl);
                state(qi + st, sr);
                es_flux<dim>(sl, sr, nhat, gamma, fl);
                for (int k = 0; k < 5; k++) {
                    const double fdn = fl[k] * nn;
                    fd[i * 5 + k] += (-alpha * fdn / w1[i] / Jdet_i);
                    fd[(i + 1) * 5 + k] += alpha * fdn / w1[i + 1] / Jdet_i_plus_1;
                }
            }
            for (int m = 0; m < Np; m++) {
                c.Jai(e_geo, ps + st * m, d, Jad_m);
                for (int e = 0; e < dim; e++) n[e] += Q[(Np - 1) * Np + m] * Jad_m[e];
            }
            const int qN = ps + st * (Np - 1);
            const double Jdet_Np = c.jdet(e_geo, qN);
            double sN[5], FN[5][dim], fN[5];
            state(qN, sN);
            euler_flux<dim>(sN, gamma, FN);
            flux_dot<dim>(FN, n, fN);
            for (int k = 0; k < 5; k++) fd[(Np - 1) * 5 + k] += (-alpha * fN[k] / w1[Np - 1] / Jdet_Np);
            for (int i = 0; i < Np; i++) {
                const int q = ps + st * i;
                const double JxW = c.jdet(e_geo, q) * c.wN[q];
                for (int k = 0; k < 5; k++) R[k * NN + q] += fd[i * 5 + k] * JxW;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Faces in gather form: element e adds its own side of every face.
// Interior/periodic: fluid_flux_es_dgsem_operator.h:301-342 (the "plus" side contribution
// (f* - f(u_p).n) equals (f(u_p).n_p - f*(u_p,u_m,n_p)) bit for bit because the ES flux is
// exactly antisymmetric under (j<->l, n -> -n)).  Boundary: :344-440.
// ---------------------------------------------------------------------------
template <int dim>
void face_residual(const Ctx& c, const double* u, int64_t e, int sp, double* R, double* bif_rate) {
    const int Np = c.Np, NN = c.NN, nF = c.nfaceN;
    const double gamma = c.gamma;
    const double* ue = u + ((size_t)e * c.nc + 5 * sp) * NN;
    for (int f = 0; f < 2 * dim; f++) {
        const int d = f / 2, side = f % 2;
        double n[dim];
        for (int k = 0; k < dim; k++) n[k] = 0.0;
        n[d] = side ? 1.0 : -1.0;
        const int64_t nb = c.neighbor(e, f);
        if (nb >= 0) {
            const double* un = u + ((size_t)nb * c.nc + 5 * sp) * NN;
            // the neighbour's matching face: the opposite one on a box; on a general mesh any of its faces, possibly
            // traversed in the opposite tangential order (2D)
            const int nf_code = c.general ? c.gnbrf[(size_t)e * 2 * dim + f] : (f ^ 1);
            const int nfa = nf_code & 7, flip = nf_code >> 3;
            for (int t = 0; t < nF; t++) {
                const int qm = c.face_node(d, side, t);
                const int qp = c.face_node(nfa / 2, nfa % 2, flip ? nF - 1 - t : t);
                if (c.general)
                    for (int k = 0; k < dim; k++) n[k] = c.gfn[(((size_t)e * 2 * dim + f) * nF + t) * dim + k];
                double sm[5], spl[5];
                for (int k = 0; k < 5; k++) { sm[k] = ue[k * NN + qm]; spl[k] = un[k * NN + qp]; }
                double Fm[5][dim], fm[5], fs[5];
                euler_flux<dim>(sm, gamma, Fm);
                flux_dot<dim>(Fm, n, fm);
                es_flux<dim>(sm, spl, n, gamma, fs);
                const double JxW = (c.general ? c.gfJ[((size_t)e * 2 * dim + f) * nF + t] : c.face_area[d]) * c.wF[t];
                for (int k = 0; k < 5; k++) R[k * NN + qm] += (fm[k] - fs[k]) * JxW;
            }
        } else {
            // boundary face, Gauss(p+2) quadrature (quad_no 0, :355-358)
            const int64_t bface = c.general ? -1 - nb : c.bface_of[(size_t)e * 2 * dim + f];
            const int bid = c.general ? c.gbf_id[bface] : f;
            // (an unknown id, fluid_flux_es_dgsem_operator.h:406-411, is rejected by orc_create_general)
            const int kind = c.bc_kind[(size_t)sp * c.nbnd + bid];
            const int Ng = c.B.Ng, nG = c.nfaceG;
            std::vector<double> val((size_t)nG * 5);   // submitted values at Gauss points
            double bsum[5] = {0, 0, 0, 0, 0};
            for (int g = 0; g < nG; g++) {
                // interpolate the nodal trace to this Gauss point
                int gi[2] = {g % Ng, (g / Ng) % Ng};
                double wm[5] = {0, 0, 0, 0, 0};
                for (int t = 0; t < nF; t++) {
                    int ti[2] = {t % Np, (t / Np) % Np};
                    double phi = 1.0;
                    for (int a = 0; a < dim - 1; a++) phi *= c.B.Ig[gi[a] * Np + ti[a]];
                    const int qm = c.face_node(d, side, t);
                    for (int k = 0; k < 5; k++) wm[k] += phi * ue[k * NN + qm];
                }
                if (c.general)
                    for (int k = 0; k < dim; k++) n[k] = c.gbn[((size_t)bface * nG + g) * dim + k];
                double rho_u_dot_n = wm[1] * n[0];
                for (int a = 1; a < dim; a++) rho_u_dot_n += wm[1 + a] * n[a];
                double wp[5];
                if (kind == ORC_BC_INFLOW || kind == ORC_BC_SUBSONIC_OUTFLOW) {
                    if (!c.general && c.inflow_fn[(size_t)sp * c.nbnd + bid].fn) {
                        const int64_t bfn = bface;
                        for (int k = 0; k < 5; k++) wp[k] = c.inflow_vals[(((size_t)sp * c.n_bfaces + bfn) * nG + g) * 5 + k];
                    } else {
                        for (int k = 0; k < 5; k++) wp[k] = c.inflow[((size_t)sp * c.nbnd + bid) * 5 + k];
                    }
                    if (kind == ORC_BC_SUBSONIC_OUTFLOW) {
                        // fluid_flux_es_dgsem_operator.h:385-390: w_p = w_m; w_p[4] = get_subsonic_outflow_energy(id) component 4
                        for (int k = 0; k < 4; k++) wp[k] = wm[k];
                    }
                } else if (kind == ORC_BC_OUTFLOW) {
                    for (int k = 0; k < 5; k++) wp[k] = wm[k];
                } else {  // wall (:394-405)
                    wp[0] = wm[0];
                    for (int a = 0; a < dim; a++) wp[a + 1] = wm[a + 1] - 2.0 * rho_u_dot_n * n[a];
                    for (int a = dim; a < 3; a++) wp[a + 1] = 0.0;
                    wp[4] = wm[4];
                }
                double Fm[5][dim], fm[5], fs[5];
                euler_flux<dim>(wm, gamma, Fm);
                flux_dot<dim>(Fm, n, fm);
                lf_flux<dim>(wm, wp, n, gamma, fs);
                const double JxW = (c.general ? c.gbJ[(size_t)bface * nG + g] : c.face_area[d]) * c.wG[g];
                for (int k = 0; k < 5; k++) {
                    val[(size_t)g * 5 + k] = (fm[k] - fs[k]) * JxW;
                    bsum[k] += fs[k] * JxW;
                }
            }
            // integrate against the nodal face basis and scatter
            for (int t = 0; t < nF; t++) {
                int ti[2] = {t % Np, (t / Np) % Np};
                const int qm = c.face_node(d, side, t);
                for (int g = 0; g < nG; g++) {
                    int gi[2] = {g % Ng, (g / Ng) % Ng};
                    double phi = 1.0;
                    for (int a = 0; a < dim - 1; a++) phi *= c.B.Ig[gi[a] * Np + ti[a]];
                    for (int k = 0; k < 5; k++) R[k * NN + qm] += phi * val[(size_t)g * 5 + k];
                }
            }
            if (bif_rate) for (int k = 0; k < 5; k++) bif_rate[bid * 5 + k] += bsum[k];
        }
    }
}

// Two-fluid source terms of the five-moment system (north_star kernel 4; NOT in the reference, which carries the field
// components through unchanged -- SURVEY 8(c) "new-physics note"; parity unpinned by construction, switched off by
// default).  Per node, with the 8 field components [Ex,Ey,Ez,Bx,By,Bz,phi,psi] of five_moment.h:131-137:
//   d(m_s)/dt  += (q/m)_s (rho_s E + m_s x B)          Lorentz force on species s
//   d(E_s)/dt  += (q/m)_s m_s . E                      work done by the electric field
//   dE/dt      += -J / eps0,   J     = sum_s (q/m)_s m_s
//   dphi/dt    += chi rho_c / eps0,  rho_c = sum_s (q/m)_s rho_s      (perfectly hyperbolic Maxwell correction source)
// The curl / divergence-cleaning fluxes of Maxwell's equations are not part of this operator.
inline void add_sources(const Ctx& c, const double* u, int64_t e, double* dudt) {
    const int NN = c.NN;
    const double* uf = u + ((size_t)e * c.nc + 5 * c.nsp) * NN;
    double* df = dudt + ((size_t)e * c.nc + 5 * c.nsp) * NN;
    for (int j = 0; j < NN; j++) {
        const double Ex = uf[0 * NN + j], Ey = uf[1 * NN + j], Ez = uf[2 * NN + j];
        const double Bx = uf[3 * NN + j], By = uf[4 * NN + j], Bz = uf[5 * NN + j];
        double Jx = 0, Jy = 0, Jz = 0, rc = 0;
        for (int sp = 0; sp < c.nsp; sp++) {
            const double qm = c.charge_over_mass[sp];
            const double* us = u + ((size_t)e * c.nc + 5 * sp) * NN;
            double* ds = dudt + ((size_t)e * c.nc + 5 * sp) * NN;
            const double rho = us[0 * NN + j], mx = us[1 * NN + j], my = us[2 * NN + j], mz = us[3 * NN + j];
            ds[1 * NN + j] += qm * (rho * Ex + (my * Bz - mz * By));
            ds[2 * NN + j] += qm * (rho * Ey + (mz * Bx - mx * Bz));
            ds[3 * NN + j] += qm * (rho * Ez + (mx * By - my * Bx));
            ds[4 * NN + j] += qm * (mx * Ex + my * Ey + mz * Ez);
            Jx += qm * mx; Jy += qm * my; Jz += qm * mz; rc += qm * rho;
        }
        df[0 * NN + j] += -Jx / c.epsilon0;
        df[1 * NN + j] += -Jy / c.epsilon0;
        df[2 * NN + j] += -Jz / c.epsilon0;
        df[6 * NN + j] += c.chi * rc / c.epsilon0;
    }
}

// Perfectly hyperbolic Maxwell (PHM) system for the field components F = [Ex,Ey,Ez,Bx,By,Bz,phi,psi] of
// five_moment.h:131-137 (north_star kernel 4, BASELINE config 5; NOT in the reference, which only allocates them --
// SURVEY 8(c) "new-physics note"; parity unpinned by construction, off by default):
//   dE/dt   - c^2 curl B + chi c^2 grad phi = -J/eps0        dB/dt   + curl E + gamma grad psi = 0
//   dphi/dt + chi div E = chi rho_c/eps0                     dpsi/dt + gamma c^2 div B = 0
// (sources: add_sources).  As a conservation law dF/dt + sum_d d f_d(F)/dx_d = S with the LINEAR fluxes
//   f_d(E) = -c^2 (e_d x B) + chi c^2 phi e_d,   f_d(B) = e_d x E + gamma psi e_d,   f_d(phi) = chi E_d,   f_d(psi) = gamma c^2 B_d.
// Discretised like the fluid: collocated DGSEM on the same Gauss-Lobatto nodes, volume term in split form with the
// arithmetic mean as two-point flux (for a linear flux this is the plain strong form: rows of D sum to zero), local
// Lax-Friedrichs (Rusanov) numerical flux with the fastest PHM speed lambda = c max(1, chi, gamma), diagonal mass.  On a
// non-periodic boundary the outside state is the inside state (zero-gradient).  Cartesian boxes only.
inline void maxwell_flux(const Ctx& c, int d, const double* F, double* f) {
    const double c2 = c.light_speed * c.light_speed;
    const double* E = F;
    const double* B = F + 3;
    // (e_d x V)_i for V = B, E
    const int i1 = (d + 1) % 3, i2 = (d + 2) % 3;
    double xB[3] = {0, 0, 0}, xE[3] = {0, 0, 0};
    xB[i1] = -B[i2]; xB[i2] = B[i1];
    xE[i1] = -E[i2]; xE[i2] = E[i1];
    for (int i = 0; i < 3; i++) {
        f[i] = -c2 * xB[i];
        f[3 + i] = xE[i];
    }
    f[d] += c.mx_chi * c2 * F[6];
    f[3 + d] += c.mx_gamma * F[7];
    f[6] = c.mx_chi * E[d];
    f[7] = c.mx_gamma * c2 * B[d];
}

template <int dim>
void add_maxwell(const Ctx& c, const double* u, int64_t e, double* dudt) {
    const int NN = c.NN, Np = c.Np, nF = c.nfaceN;
    const double* uf = u + ((size_t)e * c.nc + 5 * c.nsp) * NN;
    double* df = dudt + ((size_t)e * c.nc + 5 * c.nsp) * NN;
    const double lam = c.light_speed * std::max(1.0, std::max(c.mx_chi, c.mx_gamma));
    // volume: -(1/h_d) sum_l D[j_d][l] f_d(F_l)   (D on [0,1])
    for (int j = 0; j < NN; j++) {
        for (int d = 0; d < dim; d++) {
            const int st = c.stride(d), jd = (j / st) % Np;
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int l = 0; l < Np; l++) {
                const int q = j + (l - jd) * st;
                double F[8], f[8];
                for (int k = 0; k < 8; k++) F[k] = uf[k * NN + q];
                maxwell_flux(c, d, F, f);
                for (int k = 0; k < 8; k++) acc[k] += c.B.D[jd * Np + l] * f[k];
            }
            for (int k = 0; k < 8; k++) df[k * NN + j] -= acc[k] / c.h[d];
        }
    }
    // faces: + (f(F_m).n - f*) / (h_d w_0),  f* = 1/2 (f(F_m) + f(F_p)).n - 1/2 lambda (F_p - F_m)
    for (int f = 0; f < 2 * dim; f++) {
        const int d = f / 2, side = f % 2;
        const double sgn = side ? 1.0 : -1.0;
        const int64_t nb = c.neighbor(e, f);
        const double* un = nb >= 0 ? u + ((size_t)nb * c.nc + 5 * c.nsp) * NN : nullptr;
        for (int t = 0; t < nF; t++) {
            const int qm = c.face_node(d, side, t), qp = c.face_node(d, 1 - side, t);
            double dF[8], fn[8];
            for (int k = 0; k < 8; k++) dF[k] = un ? un[k * NN + qp] - uf[k * NN + qm] : 0.0;
            maxwell_flux(c, d, dF, fn);
            const double cf = 1.0 / (2.0 * c.h[d] * c.B.w[0]);
            for (int k = 0; k < 8; k++) df[k * NN + qm] += cf * (lam * dF[k] - sgn * fn[k]);
        }
    }
}

// Evaluate the inflow functions at every boundary quadrature point for stage time t (serially: the callbacks may be
// Python).  Point of Gauss index g on face f of element e: the face's coordinate in d, Gauss abscissae in the others.
template <int dim>
void refresh_inflow(const Ctx& c, double t) {
    bool any = false;
    for (const auto& f : c.inflow_fn) any = any || f.fn;
    if (!any) return;
    const int Ng = c.B.Ng, nG = c.nfaceG;
    c.inflow_vals.assign((size_t)c.nsp * c.n_bfaces * nG * 5, 0.0);
    for (int64_t e = 0; e < c.nelem; e++) {
        int idx[3];
        c.elem_coords(e, idx);
        for (int f = 0; f < 2 * dim; f++) {
            const int64_t bfn = c.bface_of[(size_t)e * 2 * dim + f];
            if (bfn < 0) continue;
            const int d = f / 2, side = f % 2;
            for (int sp = 0; sp < c.nsp; sp++) {
                const auto& fn = c.inflow_fn[(size_t)sp * 2 * dim + f];
                if (!fn.fn || (c.bc_kind[(size_t)sp * 2 * dim + f] != ORC_BC_INFLOW && c.bc_kind[(size_t)sp * 2 * dim + f] != ORC_BC_SUBSONIC_OUTFLOW)) continue;
                for (int g = 0; g < nG; g++) {
                    int gi[2] = {g % Ng, (g / Ng) % Ng};
                    double x[3] = {0, 0, 0};
                    int k = 0;
                    for (int a = 0; a < dim; a++) {
                        if (a == d) x[a] = c.left[a] + (idx[a] + side) * c.h[a];
                        else x[a] = c.left[a] + (idx[a] + c.B.xg[gi[k++]]) * c.h[a];
                    }
                    fn.fn(x, t, &c.inflow_vals[(((size_t)sp * c.n_bfaces + bfn) * nG + g) * 5], fn.user);
                }
            }
        }
    }
}

// dudt = M^-1 R(u):  mf.loop + inverse mass (fluid_flux_es_dgsem_operator.h:183-240)
template <int dim>
void rhs_impl(const Ctx& c, const double* u, double t, double* dudt, double* bif_rate, double* alpha_out) {
    const int NN = c.NN, nb5 = 5 * c.nbnd;
    if (!c.general) refresh_inflow<dim>(c, t);
    const int nthreads = std::max(1, c.nthreads);
    std::vector<std::vector<double>> bif_local(nthreads, std::vector<double>(nb5, 0.0));
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t e = 0; e < c.nelem; e++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        std::vector<double> R((size_t)5 * NN);
        for (int sp = 0; sp < c.nsp; sp++) {
            std::fill(R.begin(), R.end(), 0.0);
            const double* ue = u + ((size_t)e * c.nc + 5 * sp) * NN;
            const double alpha = cell_alpha<dim>(c, ue);
            if (alpha_out) alpha_out[(size_t)e * c.nsp + sp] = alpha;
            if (dudt) {
                cell_residual<dim>(c, e, ue, alpha, R.data());
                face_residual<dim>(c, u, e, sp, R.data(), bif_rate ? bif_local[tid].data() : nullptr);
                double* de = dudt + ((size_t)e * c.nc + 5 * sp) * NN;
                for (int k = 0; k < 5; k++)
                    for (int j = 0; j < NN; j++) de[k * NN + j] = R[k * NN + j] / (c.jdet(e, j) * c.wN[j]);   // :216-240 (diagonal mass)
            }
        }
        if (dudt) {
            for (int k = 5 * c.nsp; k < c.nc; k++)
                for (int j = 0; j < NN; j++) dudt[((size_t)e * c.nc + k) * NN + j] = 0.0;
            if (c.maxwell_on && c.nc >= 5 * c.nsp + 8) add_maxwell<dim>(c, u, e, dudt);
            if (c.sources_on && c.nc >= 5 * c.nsp + 8) add_sources(c, u, e, dudt);
        }
    }
    if (bif_rate) {
        for (int i = 0; i < nb5; i++) {
            double s = 0;
            for (int t = 0; t < nthreads; t++) s += bif_local[t][i];
            bif_rate[i] += s;
        }
    }
}

void rhs_dispatch(const Ctx& c, const double* u, double t, double* dudt, double* bif_rate, double* alpha_out) {
    if (c.dim == 1) rhs_impl<1>(c, u, t, dudt, bif_rate, alpha_out);
    else if (c.dim == 2) rhs_impl<2>(c, u, t, dudt, bif_rate, alpha_out);
    else rhs_impl<3>(c, u, t, dudt, bif_rate, alpha_out);
}

// fluid_flux_es_dgsem_operator.h:127-214
void forward_euler(const Ctx& c, double* dst, const double* u, double dt, double t, double a, double beta,
                   double* bif_dst, const double* bif_u) {
    const size_t N = (size_t)c.nelem * c.nc * c.NN;
    std::vector<double> dudt(N);
    const int nb5 = 5 * c.nbnd;
    std::vector<double> rate(nb5, 0.0);
    rhs_dispatch(c, u, t, dudt.data(), rate.data(), nullptr);
    const int nthreads = std::max(1, c.nthreads);
    const size_t blockN = (size_t)c.nc * c.NN;
    const size_t fluidN = (size_t)5 * c.nsp * c.NN;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t e = 0; e < c.nelem; e++) {
        // the reference's post-loop lambda (:195-205) runs over every DoF, fields included (dudt = 0 there)
        for (size_t i = 0; i < blockN; i++) {
            const size_t g = (size_t)e * blockN + i;
            const double dudt_i = (i < fluidN || c.sources_on || c.maxwell_on) ? dudt[g] : 0.0;
            dst[g] = beta * dst[g] + a * (u[g] + dt * dudt_i);
        }
    }
    if (bif_dst) {
        for (int i = 0; i < nb5; i++) {
            bif_dst[i] = beta * bif_dst[i] + (a * dt) * rate[i];   // sadd(beta, alpha*dt, rate)  :208-209
            bif_dst[i] = 1.0 * bif_dst[i] + a * (bif_u ? bif_u[i] : 0.0);   // sadd(1, alpha, u.bif)  :210-211
        }
    }
}

// fluid_flux_es_dgsem_operator.h:450-514
double max_transport_speed(const Ctx& c, const double* u) {
    double max_transport = 0;
    const int NN = c.NN, dim = c.dim;
    for (int sp = 0; sp < c.nsp; sp++) {
        double m = 0;
        const int nthreads = std::max(1, c.nthreads);
#pragma omp parallel for num_threads(nthreads) reduction(max : m) schedule(static)
        for (int64_t e = 0; e < c.nelem; e++) {
            const double* ue = u + ((size_t)e * c.nc + 5 * sp) * NN;
            for (int j = 0; j < NN; j++) {
                double q[5];
                for (int k = 0; k < 5; k++) q[k] = ue[k * NN + j];
                const double inv = 1. / q[0];
                const double pr = pressure(q, c.gamma);
                double conv = 0, eig = c.max_eig;
                if (c.general) {
                    // convective_speed = inverse_jacobian * velocity (:476-481), with the full matrix
                    double K[3][3];
                    c.jinv_mat(e, j, K);
                    for (int r = 0; r < dim; r++) {
                        double s = 0;
                        for (int cc = 0; cc < dim; cc++) s += K[r][cc] * (q[cc + 1] * inv);
                        conv = std::max(conv, std::fabs(s));
                    }
                    eig = c.geig[(size_t)e * NN + j];
                } else {
                    for (int d = 0; d < dim; d++) conv = std::max(conv, std::fabs(c.Jinv[d] * (q[d + 1] * inv)));
                }
                const double cs = std::sqrt(c.gamma * pr * (1. / q[0]));
                m = std::max(m, eig * cs + conv);
            }
        }
        max_transport = std::max(max_transport, m);
    }
    if (c.maxwell_on && c.nc >= 5 * c.nsp + 8) {
        // the field system: fastest PHM wave through the same metric factor as the sound speed, and -- when the two-fluid
        // sources couple it to the fluids -- the plasma and cyclotron frequencies, limited to omega dt <= 0.1 (SSPRK2 has no
        // stability interval on the imaginary axis: |R(i y)|^2 = 1 + y^4/4).  With dt = 0.5 / (vmax Np^2) that bound is the
        // equivalent speed 5 omega / Np^2.
        max_transport = std::max(max_transport, c.max_eig * c.light_speed * std::max(1.0, std::max(c.mx_chi, c.mx_gamma)));
        if (c.sources_on) {
            double m = 0;
            const int nthreads = std::max(1, c.nthreads);
#pragma omp parallel for num_threads(nthreads) reduction(max : m) schedule(static)
            for (int64_t e = 0; e < c.nelem; e++) {
                const double* uf = u + ((size_t)e * c.nc + 5 * c.nsp) * NN;
                for (int j = 0; j < NN; j++) {
                    double wp2 = 0, qmax = 0;
                    for (int sp = 0; sp < c.nsp; sp++) {
                        const double qm = c.charge_over_mass[sp];
                        wp2 += qm * qm * u[((size_t)e * c.nc + 5 * sp) * NN + j] / c.epsilon0;
                        qmax = std::max(qmax, std::fabs(qm));
                    }
                    const double Bx = uf[3 * NN + j], By = uf[4 * NN + j], Bz = uf[5 * NN + j];
                    const double omega = std::max(std::sqrt(wp2), qmax * std::sqrt(Bx * Bx + By * By + Bz * Bz));
                    m = std::max(m, 5.0 * omega / (double)(c.Np * c.Np));
                }
            }
            max_transport = std::max(max_transport, m);
        }
    }
    return max_transport;
}

void ssprk2(const Ctx& c, double* u, double* f1, double dt, double t, double* bif, double* bif_f1) {
    // rk.h:97-106.  f_1's previous contents are multiplied by beta = 0.
    const size_t N = (size_t)c.nelem * c.nc * c.NN;
    std::fill(f1, f1 + N, 0.0);
    if (bif_f1) std::fill(bif_f1, bif_f1 + 5 * c.nbnd, 0.0);
    forward_euler(c, f1, u, dt, t, 1.0, 0.0, bif_f1, bif);
    forward_euler(c, u, f1, dt, t + dt, 0.5, 0.5, bif, bif_f1);
}

// src/timestepper.cc:6-56
struct Callback {
    double interval;
    std::function<void(double)> fn;
    bool perform_zeroth, perform_final;
};
void advance(const std::function<bool(double, double)>& step, double t_end, const std::function<double()>& recommend_dt,
             std::vector<Callback>& callbacks) {
    double t = 0.0;
    double dt;
    std::vector<double> callback_times;
    for (auto& cb : callbacks) {
        if (cb.perform_zeroth) cb.fn(0.0);
        callback_times.push_back(t + cb.interval);
    }
    while (t < t_end - 1e-12) {
        size_t next_idx = 0;
        for (size_t i = 1; i < callback_times.size(); i++)
            if (callback_times[i] < callback_times[next_idx]) next_idx = i;
        double next_cb_time = callbacks.empty() ? t_end : callback_times[next_idx];
        bool perform_cb = next_cb_time < t_end && std::fabs(next_cb_time - t_end) > 1e-12;
        double next_stop = std::fmin(next_cb_time, t_end);
        while (t < next_stop - 1e-12) {
            dt = std::fmin(recommend_dt(), next_stop - t);
            bool ok = step(t, dt);
            if (!ok) continue;
            t += dt;
        }
        if (perform_cb) {
            auto& cb = callbacks[next_idx];
            cb.fn(t);
            callback_times[next_idx] = t + cb.interval;
        }
    }
    for (size_t i = 0; i < callbacks.size(); i++) {
        auto& cb = callbacks[i];
        if (std::fabs(t_end - callback_times[i]) < 1e-12 || cb.perform_final) cb.fn(t_end);
    }
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

double orc_ln_avg(double a, double b) { return ln_avg(a, b); }
void orc_set_log_impl(int deterministic) { g_log_impl = deterministic ? 1 : 0; }
int orc_get_log_impl(void) { return g_log_impl; }
double orc_det_log(double x) { return detlog::det_log(x); }
double orc_pressure(const double q[5], double gamma) { return pressure(q, gamma); }

#define DIM_DISPATCH(dim, CALL1, CALL2, CALL3) \
    do { if ((dim) == 1) { CALL1; } else if ((dim) == 2) { CALL2; } else { CALL3; } } while (0)

void orc_euler_flux(int dim, const double q[5], double gamma, double* F) {
    DIM_DISPATCH(dim, euler_flux<1>(q, gamma, (double(*)[1])F), euler_flux<2>(q, gamma, (double(*)[2])F),
                 euler_flux<3>(q, gamma, (double(*)[3])F));
}
void orc_lf_flux(int dim, const double qin[5], const double qout[5], const double* n, double gamma, double out[5]) {
    DIM_DISPATCH(dim, lf_flux<1>(qin, qout, n, gamma, out), lf_flux<2>(qin, qout, n, gamma, out),
                 lf_flux<3>(qin, qout, n, gamma, out));
}
void orc_ec_flux(int dim, const double qj[5], const double ql[5], double gamma, double* F) {
    DIM_DISPATCH(dim, ec_flux<1>(qj, ql, gamma, (double(*)[1])F), ec_flux<2>(qj, ql, gamma, (double(*)[2])F),
                 ec_flux<3>(qj, ql, gamma, (double(*)[3])F));
}
void orc_es_flux(int dim, const double qj[5], const double ql[5], const double* n, double gamma, double out[5]) {
    DIM_DISPATCH(dim, es_flux<1>(qj, ql, n, gamma, out), es_flux<2>(qj, ql, n, gamma, out),
                 es_flux<3>(qj, ql, n, gamma, out));
}
double orc_mathematical_entropy(const double q[5], double gamma) {
    double s = specific_entropy(q, gamma);   // euler.h:143-148
    return -s * q[0] / (gamma - 1.0);
}
void orc_entropy_variables(const double q[5], double gamma, double w[5]) {
    // euler.h:150-169
    double beta = q[0] / (2.0 * pressure(q, gamma));
    double s = specific_entropy(q, gamma);
    double u[3];
    velocity<3>(q, u);
    double u2 = 0.0;
    for (int d = 0; d < 3; d++) u2 += u[d] * u[d];
    w[0] = (gamma - s) / (gamma - 1.0) - beta * u2;
    for (int d = 0; d < 3; d++) w[d + 1] = 2 * beta * u[d];
    w[4] = -2 * beta;
}
void orc_entropy_flux(int dim, const double q[5], double gamma, double* out) {
    // euler.h:171-181
    double s = specific_entropy(q, gamma);
    double inv = 1. / q[0];
    for (int d = 0; d < dim; d++) out[d] = -q[0] * (q[d + 1] * inv) * s / (gamma - 1.0);
}
void orc_primitive_to_conserved(const double prim[5], double gamma, double cons[5]) {
    // species_func.cc:15-28
    double rho = prim[0];
    cons[0] = rho;
    for (int c = 1; c <= 3; c++) cons[c] = rho * prim[c];
    double kinetic_energy = 0.0;
    for (int d = 0; d < 3; d++) kinetic_energy += 0.5 * rho * prim[d + 1] * prim[d + 1];
    cons[4] = kinetic_energy + prim[4] / (gamma - 1);
}

// ---- index maps: src/dof_utils.cc ------------------------------------------
unsigned orc_pencil_stride(unsigned Np, unsigned d) { return d == 0 ? 1 : (d == 1 ? Np : Np * Np); }
unsigned orc_pencil_base(int dim, unsigned q, unsigned Np, unsigned d) {
    if (dim == 1) return 0;
    // node q = (i0, i1, i2); the pencil through q along d starts where i_d = 0
    unsigned st = orc_pencil_stride(Np, d);
    unsigned id = (q / st) % Np;
    return q - id * st;
}
unsigned orc_quadrature_point_neighbor(int dim, unsigned q, unsigned k, unsigned Np, unsigned d) {
    if (dim == 1) return k;
    return orc_pencil_base(dim, q, Np, d) + orc_pencil_stride(Np, d) * k;
}
unsigned orc_quad_point_1d_index(int dim, unsigned q, unsigned Np, unsigned d) {
    if (dim == 1) return q;
    return (q / orc_pencil_stride(Np, d)) % Np;
}
int orc_pencil_starts(int dim, unsigned Np, unsigned d, unsigned* out) {
    if (dim == 1) { out[0] = 0; return 1; }
    unsigned NN = 1;
    for (int i = 0; i < dim; i++) NN *= Np;
    int n = 0;
    unsigned st = orc_pencil_stride(Np, d);
    for (unsigned q = 0; q < NN; q++)
        if ((q / st) % Np == 0) out[n++] = q;
    return n;
}

void orc_advance(orc_step_fn step, double t_end, orc_dt_fn recommend_dt, int n_callbacks, const double* intervals,
                 const int* perform_zeroth, const int* perform_final, orc_cb_fn cb, void* user) {
    std::vector<Callback> cbs;
    for (int i = 0; i < n_callbacks; i++)
        cbs.push_back(Callback{intervals[i], [=](double t) { cb(t, i, user); }, perform_zeroth[i] != 0, perform_final[i] != 0});
    advance([&](double t, double dt) { return step(t, dt, user) != 0; }, t_end, [&]() { return recommend_dt(user); }, cbs);
}

void orc_gll(int Np, double* x, double* w) {
    Basis b;
    b.init(Np);
    for (int i = 0; i < Np; i++) { x[i] = b.x[i]; w[i] = b.w[i]; }
}
void orc_gauss(int n, double* x, double* w) {
    std::vector<long double> xl, wl;
    gauss_ld(n, xl, wl);
    for (int i = 0; i < n; i++) { x[i] = (double)xl[i]; w[i] = (double)wl[i]; }
}
void orc_diff_matrix(int Np, double* D) {
    Basis b;
    b.init(Np);
    std::copy(b.D.begin(), b.D.end(), D);
}
void orc_legendre_analysis_1d(int Np, double* V) {
    Basis b;
    b.init(Np);
    std::copy(b.V.begin(), b.V.end(), V);
}

void* orc_create(int dim, int fe_degree, int n_species, int fields_enabled, double gamma, const int* nx,
                 const double* left, const double* right, const int* periodic, const int* bc_kinds) {
    if (dim < 1 || dim > 3 || fe_degree < 1 || n_species < 1) return nullptr;
    Ctx* c = new Ctx();
    c->dim = dim;
    c->p = fe_degree;
    c->Np = fe_degree + 1;
    c->NN = ipow(c->Np, dim);
    c->nsp = n_species;
    c->nc = 5 * n_species + (fields_enabled ? 8 : 0);   // five_moment.h:175-176
    c->gamma = gamma;
    c->nelem = 1;
    for (int d = 0; d < dim; d++) {
        c->nx[d] = nx[d];
        c->left[d] = left[d];
        c->right[d] = right[d];
        c->periodic[d] = periodic[d];
        c->nelem *= nx[d];
    }
    c->B.init(c->Np);
    c->nbnd = 2 * dim;
    setup_geometry(*c);
    c->bc_kind.assign((size_t)n_species * 2 * dim, ORC_BC_WALL);   // species.cc:17 default "Wall"
    if (bc_kinds) for (size_t i = 0; i < c->bc_kind.size(); i++) c->bc_kind[i] = bc_kinds[i];
    c->inflow.assign((size_t)n_species * 2 * dim * 5, 0.0);
    c->inflow_fn.assign((size_t)n_species * 2 * dim, Ctx::InflowFn());
    c->bface_of.assign((size_t)c->nelem * 2 * dim, -1);
    for (int64_t e = 0; e < c->nelem; e++)
        for (int f = 0; f < 2 * dim; f++)
            if (c->neighbor(e, f) < 0) c->bface_of[(size_t)e * 2 * dim + f] = c->n_bfaces++;
    return c;
}
// General geometry: the operator on an arbitrary conforming mesh of (curved) quadrilaterals / hexahedra, fed with what
// the reference reads from deal.II per quadrature point: inverse_jacobian(q) (fluid_flux_es_dgsem_operator.h:476,
// jacobian_utils.h:14,36), normal_vector(q) and the face JxW (:317-339, :361-437).  Box-only services (node_coords,
// inflow functions) are not available on such a context.
void* orc_create_general(int dim, int fe_degree, int n_species, int fields_enabled, double gamma, int64_t n_elems,
                         int n_boundaries, const int64_t* face_neighbor, const int32_t* neighbor_face, int64_t n_bfaces,
                         const int32_t* bf_id, const int* bc_kinds, const double* inverse_jacobian,
                         const double* face_normal, const double* face_jacobian, const double* boundary_normal,
                         const double* boundary_jacobian) {
    if (dim < 1 || dim > 3 || fe_degree < 1 || n_species < 1 || n_elems < 0 || !face_neighbor || !inverse_jacobian ||
        !face_normal || !face_jacobian)
        return nullptr;
    for (int64_t b = 0; b < n_bfaces; b++)
        if (bf_id[b] < 0 || bf_id[b] >= n_boundaries) return nullptr;   // "Unknown boundary id" (:406-411)
    Ctx* c = new Ctx();
    c->dim = dim;
    c->p = fe_degree;
    c->Np = fe_degree + 1;
    c->NN = ipow(c->Np, dim);
    c->nsp = n_species;
    c->nc = 5 * n_species + (fields_enabled ? 8 : 0);
    c->gamma = gamma;
    c->nelem = n_elems;
    for (int d = 0; d < dim; d++) { c->nx[d] = 1; c->left[d] = 0; c->right[d] = 1; c->periodic[d] = 0; }
    c->B.init(c->Np);
    setup_geometry(*c);   // reference-element tables (weights, Legendre analysis); the box constants are not used
    c->general = true;
    c->nbnd = n_boundaries;
    c->n_bfaces = n_bfaces;
    const int nf = 2 * dim, NN = c->NN, nF = c->nfaceN, nG = c->nfaceG;
    c->gnbr.assign(face_neighbor, face_neighbor + (size_t)n_elems * nf);
    c->gnbrf.resize((size_t)n_elems * nf);
    for (size_t i = 0; i < c->gnbrf.size(); i++) c->gnbrf[i] = neighbor_face ? neighbor_face[i] : (int)((i % nf) ^ 1);
    c->gbf_id.assign(bf_id, bf_id + n_bfaces);
    c->gJinv.assign(inverse_jacobian, inverse_jacobian + (size_t)n_elems * NN * dim * dim);
    c->gfn.assign(face_normal, face_normal + (size_t)n_elems * nf * nF * dim);
    c->gfJ.assign(face_jacobian, face_jacobian + (size_t)n_elems * nf * nF);
    if (n_bfaces > 0) {
        c->gbn.assign(boundary_normal, boundary_normal + (size_t)n_bfaces * nG * dim);
        c->gbJ.assign(boundary_jacobian, boundary_jacobian + (size_t)n_bfaces * nG);
    }
    c->gJdet.resize((size_t)n_elems * NN);
    c->geig.resize((size_t)n_elems * NN);
    for (int64_t e = 0; e < n_elems; e++)
        for (int q = 0; q < NN; q++) {
            double K[3][3];
            c->jinv_mat(e, q, K);
            // tensor_utils.h:70-82 (1x1, 2x2); the 3x3 cofactor expansion is the extension the reference lacks
            double det;
            if (dim == 1) det = K[0][0];
            else if (dim == 2) det = K[0][0] * K[1][1] - K[0][1] * K[1][0];
            else det = K[0][0] * (K[1][1] * K[2][2] - K[1][2] * K[2][1]) - K[0][1] * (K[1][0] * K[2][2] - K[1][2] * K[2][0]) +
                       K[0][2] * (K[1][0] * K[2][1] - K[1][1] * K[2][0]);
            c->gJdet[(size_t)e * NN + q] = 1.0 / det;   // jacobian_utils.h:15-16
            // fluid_flux_es_dgsem_operator.h:487-502: 5 power iterations on K^T K from (1,...,1)
            double ev[3] = {1, 1, 1};
            for (int it = 0; it < 5; it++) {
                double Kv[3] = {0, 0, 0}, w[3] = {0, 0, 0}, nrm = 0;
                for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) Kv[r] += K[r][cc] * ev[cc];
                for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) w[r] += K[cc][r] * Kv[cc];
                for (int r = 0; r < dim; r++) nrm = std::max(nrm, std::fabs(w[r]));
                for (int r = 0; r < dim; r++) ev[r] = w[r] / nrm;
            }
            double Kv[3] = {0, 0, 0}, num = 0, den = 0;
            for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) Kv[r] += K[r][cc] * ev[cc];
            for (int r = 0; r < dim; r++) { num += Kv[r] * Kv[r]; den += ev[r] * ev[r]; }
            c->geig[(size_t)e * NN + q] = std::sqrt(num / den);
        }
    c->bc_kind.assign((size_t)n_species * std::max(1, n_boundaries), ORC_BC_WALL);
    if (bc_kinds) for (size_t i = 0; i < (size_t)n_species * n_boundaries; i++) c->bc_kind[i] = bc_kinds[i];
    c->inflow.assign((size_t)n_species * std::max(1, n_boundaries) * 5, 0.0);
    c->inflow_fn.assign((size_t)n_species * std::max(1, n_boundaries), Ctx::InflowFn());
    return c;
}
void orc_destroy(void* h) { delete (Ctx*)h; }
void orc_set_threads(void* h, int n) { ((Ctx*)h)->nthreads = std::max(1, n); }
int64_t orc_n_elems(void* h) { return ((Ctx*)h)->nelem; }
int64_t orc_n_dofs(void* h) { Ctx* c = (Ctx*)h; return c->nelem * c->nc * c->NN; }
int orc_n_components(void* h) { return ((Ctx*)h)->nc; }
int orc_nodes_per_elem(void* h) { return ((Ctx*)h)->NN; }
int orc_n_boundaries(void* h) { return ((Ctx*)h)->nbnd; }
void orc_node_coords(void* h, double* xyz) {
    Ctx& c = *(Ctx*)h;
    for (int64_t e = 0; e < c.nelem; e++) {
        int idx[3];
        c.elem_coords(e, idx);
        for (int j = 0; j < c.NN; j++) {
            int t = j;
            for (int d = 0; d < c.dim; d++) {
                xyz[((size_t)e * c.NN + j) * c.dim + d] = c.left[d] + (idx[d] + c.B.x[t % c.Np]) * c.h[d];
                t /= c.Np;
            }
        }
    }
}
void orc_set_inflow(void* h, int species, int boundary_id, const double q[5]) {
    Ctx& c = *(Ctx*)h;
    for (int k = 0; k < 5; k++) c.inflow[((size_t)species * c.nbnd + boundary_id) * 5 + k] = q[k];
}
void orc_set_sources(void* h, int enabled, double epsilon0, double chi, const double* charge_over_mass) {
    Ctx& c = *(Ctx*)h;
    c.sources_on = enabled != 0;
    c.epsilon0 = epsilon0;
    c.chi = chi;
    c.charge_over_mass.assign(c.nsp, 0.0);
    if (charge_over_mass) for (int s = 0; s < c.nsp; s++) c.charge_over_mass[s] = charge_over_mass[s];
}
int orc_set_maxwell(void* h, int enabled, double light_speed, double chi, double gamma) {
    Ctx& c = *(Ctx*)h;
    if (enabled && (c.general || c.nc < 5 * c.nsp + 8 || !(light_speed > 0.0))) return 1;
    c.maxwell_on = enabled != 0;
    c.light_speed = light_speed;
    c.mx_chi = chi;
    c.mx_gamma = gamma;
    return 0;
}
void orc_set_inflow_function(void* h, int species, int boundary_id, orc_inflow_fn fn, void* user) {
    Ctx& c = *(Ctx*)h;
    c.inflow_fn[(size_t)species * 2 * c.dim + boundary_id].fn = fn;
    c.inflow_fn[(size_t)species * 2 * c.dim + boundary_id].user = user;
}
void orc_rhs(void* h, const double* u, double t, double* dudt, double* bif_rate) {
    Ctx& c = *(Ctx*)h;
    if (bif_rate) std::fill(bif_rate, bif_rate + 5 * c.nbnd, 0.0);
    rhs_dispatch(c, u, t, dudt, bif_rate, nullptr);
}
void orc_cell_residual(void* h, const double* ue, double alpha, double* R) {
    Ctx& c = *(Ctx*)h;
    std::fill(R, R + 5 * c.NN, 0.0);
    if (c.dim == 1) cell_residual<1>(c, 0, ue, alpha, R);
    else if (c.dim == 2) cell_residual<2>(c, 0, ue, alpha, R);
    else cell_residual<3>(c, 0, ue, alpha, R);
}
double orc_shock_indicator(void* h, const double* v) { return shock_indicator(*(Ctx*)h, v); }
void orc_alpha(void* h, const double* u, double* alpha) { rhs_dispatch(*(Ctx*)h, u, 0.0, nullptr, nullptr, alpha); }
void orc_forward_euler_step(void* h, double* dst, const double* u, double dt, double t, double a, double beta,
                            double* bif_dst, const double* bif_u) {
    forward_euler(*(Ctx*)h, dst, u, dt, t, a, beta, bif_dst, bif_u);
}
// One low-storage Runge-Kutta stage as LowStorageRungeKuttaIntegrator::perform_time_step asks of the operator (rk.h:53-71;
// tutorial-67.cc:880-899): k = M^-1 R(r_in); with s = sol: sol = s + factor_solution * k; r_out = s + factor_ai * k
// (r_out not written when factor_ai == 0).  r_out may alias r_in (k is formed first), as in the reference's calls.
void orc_lsrk_stage(void* h, double* sol, double* r_out, const double* r_in, double factor_solution, double factor_ai, double t) {
    Ctx& c = *(Ctx*)h;
    const size_t N = (size_t)c.nelem * c.nc * c.NN;
    std::vector<double> k(N, 0.0);
    rhs_dispatch(c, r_in, t, k.data(), nullptr, nullptr);
    const size_t blockN = (size_t)c.nc * c.NN, fluidN = (size_t)5 * c.nsp * c.NN;
    for (size_t g = 0; g < N; g++) {
        const double k_i = ((g % blockN) < fluidN || c.sources_on) ? k[g] : 0.0;
        const double sol_i = sol[g];
        sol[g] = sol_i + factor_solution * k_i;
        if (factor_ai != 0.0) r_out[g] = sol_i + factor_ai * k_i;
    }
}
double orc_max_transport_speed(void* h, const double* u) { return max_transport_speed(*(Ctx*)h, u); }
double orc_recommend_dt(void* h, const double* u) {
    Ctx& c = *(Ctx*)h;   // fluid_flux_es_dgsem_operator.h:442-448
    double v = max_transport_speed(c, u);
    return 0.5 / (v * (c.p + 1) * (c.p + 1));
}
void orc_ssprk2_step(void* h, double* u, double* f1, double dt, double t, double* bif, double* bif_f1) {
    ssprk2(*(Ctx*)h, u, f1, dt, t, bif, bif_f1);
}
int64_t orc_solve(void* h, double* u, double t_end, double* bif, int64_t max_steps, double fixed_dt) {
    // five_moment/dg_solver.cc:23-38: step = SSPRK2, recommend_dt every step, no callbacks
    Ctx& c = *(Ctx*)h;
    const size_t N = (size_t)c.nelem * c.nc * c.NN;
    std::vector<double> f1(N, 0.0), bif_f1(5 * c.nbnd, 0.0);
    int64_t steps = 0;
    std::vector<Callback> cbs;
    bool stop = false;
    advance(
        [&](double t, double dt) {
            if (!stop) {
                ssprk2(c, u, f1.data(), dt, t, bif, bif ? bif_f1.data() : nullptr);
                steps++;
                if (max_steps > 0 && steps >= max_steps) stop = true;
            }
            return true;
        },
        t_end, [&]() { return stop ? 1e300 : (fixed_dt > 0 ? fixed_dt : orc_recommend_dt(h, u)); }, cbs);
    return steps;
}
void orc_global_integral(void* h, const double* u, int species, double out[5]) {
    Ctx& c = *(Ctx*)h;   // dg_solution_helper.cc:71-98
    for (int k = 0; k < 5; k++) out[k] = 0.0;
    for (int64_t e = 0; e < c.nelem; e++) {
        const double* ue = u + ((size_t)e * c.nc + 5 * species) * c.NN;
        double cell[5] = {0, 0, 0, 0, 0};
        for (int j = 0; j < c.NN; j++)
            for (int k = 0; k < 5; k++) cell[k] += ue[k * c.NN + j] * (c.jdet(e, j) * c.wN[j]);
        for (int k = 0; k < 5; k++) out[k] += cell[k];
    }
}

}  // extern "C"
