// 1-D reference-element tables of the DGSEM discretisation on [0,1] (product code, host side).
//
// Stands in for what the reference pulls out of deal.II at construction time:
//   QGaussLobatto<1>(p+1), QGauss<1>(p+2)            nodal_dg_discretization.cc:12-13
//   D(j,l) = FE_DGQ::shape_grad(l, x_j)[0]            split_form_volume_flux.h:36-45
//   FESeries::Legendre coefficients on GLL quadrature persson_peraire_shock_indicator.h:12-23
// Everything is evaluated in long double and rounded once.
#pragma once
#include <cmath>
#include <vector>

namespace warpii_b200 {

struct ReferenceElement {
    int Np = 0;                 // GLL nodes per direction (fe_degree + 1)
    int Ng = 0;                 // Gauss points per direction on boundary faces (fe_degree + 2)
    std::vector<double> x, w;   // GLL nodes / weights on [0,1]
    std::vector<double> D;      // [Np*Np] derivative matrix on [0,1]
    std::vector<double> V;      // [Np*Np] Legendre analysis: V[k*Np+q] = (k+1/2) w_q sqrt2 P_k(2 x_q - 1)
    std::vector<double> xg, wg; // Gauss(Ng) on [0,1]
    std::vector<double> Ig;     // [Ng*Np] Ig[q*Np+i] = l_i(xg_q)

    explicit ReferenceElement(int fe_degree) { build(fe_degree + 1); }

   private:
    using ld = long double;

    // P_n and P_n' at x by the three-term recurrence
    static void legendre(int n, ld x, ld& P, ld& dP) {
        ld p0 = 1, p1 = x, d0 = 0, d1 = 1;
        if (n == 0) { P = p0; dP = d0; return; }
        for (int k = 2; k <= n; k++) {
            const ld p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
            const ld d2 = d0 + (2 * k - 1) * p1;
            p0 = p1; p1 = p2; d0 = d1; d1 = d2;
        }
        P = p1; dP = d1;
    }

    void build(int n_nodes) {
        Np = n_nodes;
        Ng = Np + 1;
        const int n = Np - 1;   // polynomial degree
        const ld pi = 3.141592653589793238462643383279502884L;
        std::vector<ld> xs(Np), ws(Np);
        // interior GLL nodes are the roots of P_n'(x): Newton on q(x) = P_n'(x), q' from Legendre's ODE
        for (int i = 0; i <= n; i++) {
            ld xi = -std::cos(pi * i / n);
            if (i == 0) xi = -1;
            else if (i == n) xi = 1;
            else {
                for (int it = 0; it < 60; it++) {
                    ld P, dP;
                    legendre(n, xi, P, dP);
                    const ld d2P = (2 * xi * dP - n * (n + 1) * P) / (1 - xi * xi);
                    const ld dx = dP / d2P;
                    xi -= dx;
                    if (std::fabs((double)dx) < 1e-20) break;
                }
            }
            ld P, dP;
            legendre(n, xi, P, dP);
            xs[i] = xi;
            ws[i] = 2 / (n * (n + 1) * P * P);
        }
        for (int i = 0; i < Np / 2; i++) {   // enforce symmetry exactly
            const ld a = (xs[Np - 1 - i] - xs[i]) / 2, b = (ws[i] + ws[Np - 1 - i]) / 2;
            xs[i] = -a; xs[Np - 1 - i] = a; ws[i] = ws[Np - 1 - i] = b;
        }
        if (Np % 2) xs[Np / 2] = 0;
        std::vector<ld> x01(Np), w01(Np);
        x.resize(Np); w.resize(Np);
        for (int i = 0; i < Np; i++) {
            x01[i] = (xs[i] + 1) / 2;
            w01[i] = ws[i] / 2;
            x[i] = (double)x01[i];
            w[i] = (double)w01[i];
        }
        // derivative matrix from the product form of the Lagrange basis
        D.assign(Np * Np, 0.0);
        for (int j = 0; j < Np; j++) {
            ld rowsum = 0;
            for (int l = 0; l < Np; l++) {
                if (l == j) continue;
                ld num = 1, den = 1;
                for (int m = 0; m < Np; m++) {
                    if (m != l) den *= (x01[l] - x01[m]);
                    if (m != l && m != j) num *= (x01[j] - x01[m]);
                }
                const ld d = num / den;
                D[j * Np + l] = (double)d;
                rowsum += d;
            }
            D[j * Np + j] = (j == 0 || j == Np - 1) ? (double)(-rowsum) : 0.0;
        }
        V.assign(Np * Np, 0.0);
        for (int k = 0; k < Np; k++)
            for (int q = 0; q < Np; q++) {
                ld P, dP;
                legendre(k, 2 * x01[q] - 1, P, dP);
                V[k * Np + q] = (double)((k + 0.5L) * w01[q] * std::sqrt((ld)2) * P);
            }
        // Gauss(Ng)
        std::vector<ld> g(Ng), gw(Ng);
        for (int i = 0; i < Ng; i++) {
            ld xi = -std::cos(pi * (i + 0.75L) / (Ng + 0.5L));
            for (int it = 0; it < 60; it++) {
                ld P, dP;
                legendre(Ng, xi, P, dP);
                const ld dx = P / dP;
                xi -= dx;
                if (std::fabs((double)dx) < 1e-20) break;
            }
            ld P, dP;
            legendre(Ng, xi, P, dP);
            g[i] = xi;
            gw[i] = 2 / ((1 - xi * xi) * dP * dP);
        }
        for (int i = 0; i < Ng / 2; i++) {
            const ld a = (g[Ng - 1 - i] - g[i]) / 2, b = (gw[i] + gw[Ng - 1 - i]) / 2;
            g[i] = -a; g[Ng - 1 - i] = a; gw[i] = gw[Ng - 1 - i] = b;
        }
        if (Ng % 2) g[Ng / 2] = 0;
        xg.resize(Ng); wg.resize(Ng);
        std::vector<ld> g01(Ng);
        for (int i = 0; i < Ng; i++) {
            g01[i] = (g[i] + 1) / 2;
            xg[i] = (double)g01[i];
            wg[i] = (double)(gw[i] / 2);
        }
        Ig.assign(Ng * Np, 0.0);
        for (int q = 0; q < Ng; q++)
            for (int i = 0; i < Np; i++) {
                ld v = 1;
                for (int m = 0; m < Np; m++)
                    if (m != i) v *= (g01[q] - x01[m]) / (x01[i] - x01[m]);
                Ig[q * Np + i] = (double)v;
            }
    }

   public:
    // QGauss<1>(n) on [0,1] and the values of this element's Lagrange basis there: I[q*Np+i] = l_i(x_q)
    // (for quadratures other than the two the operator uses, e.g. VectorTools::integrate_difference with QGauss(fe_degree))
    void gauss_rule(int n, std::vector<double>& xq, std::vector<double>& wq, std::vector<double>& I) const {
        const ld pi = std::acos((ld)-1);
        std::vector<ld> g(n), gw(n);
        for (int i = 0; i < n; i++) {
            ld xi = -std::cos(pi * (i + 0.75L) / (n + 0.5L));
            ld P, dP;
            for (int it = 0; it < 60; it++) {
                legendre(n, xi, P, dP);
                const ld dx = P / dP;
                xi -= dx;
                if (std::fabs((double)dx) < 1e-20) break;
            }
            legendre(n, xi, P, dP);
            g[i] = (xi + 1) / 2;
            gw[i] = 1 / ((1 - xi * xi) * dP * dP);
        }
        xq.resize(n); wq.resize(n);
        for (int i = 0; i < n; i++) { xq[i] = (double)g[i]; wq[i] = (double)gw[i]; }
        I.assign((size_t)n * Np, 0.0);
        for (int q = 0; q < n; q++)
            for (int i = 0; i < Np; i++) {
                ld v = 1;
                for (int m = 0; m < Np; m++)
                    if (m != i) v *= (g[q] - (ld)x[m]) / ((ld)x[i] - (ld)x[m]);
                I[(size_t)q * Np + i] = (double)v;
            }
    }
};

}  // namespace warpii_b200
