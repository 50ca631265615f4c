// Minimal stand-in for the part of deal.II's matrix-free layer that the reference's volume and subcell drivers touch
// (src/five_moment/fluxes/split_form_volume_flux.h:61-99, subcell_finite_volume_flux.h:68-159, jacobian_utils.h): ONE cell,
// one SIMD lane, collocated Gauss-Lobatto nodes (dof index == quadrature index, which the reference relies on,
// fluid_flux_es_dgsem_operator.h:286-287).  It serves values and geometry and implements integrate_scatter the way
// deal.II does for a nodal basis collocated with the quadrature: dst_i += value_i * JxW_i.  No physics lives here.
#pragma once
#include <cmath>
#include <memory>
#include <vector>

#include <deal.II/base/tensor.h>
#include <deal.II/base/vectorization.h>

namespace dealii {

namespace EvaluationFlags { enum EvaluationFlags { nothing = 0, values = 1, gradients = 2 }; }
enum UpdateFlags { update_default = 0, update_values = 1 };

template <int dim, typename Number = double>
class Point : public Tensor<1, dim, Number> {};

template <typename Number>
class FullMatrix {
   public:
    FullMatrix(unsigned r = 0, unsigned c = 0) : rows(r), cols(c), a((size_t)r * c, Number()) {}
    Number& operator()(unsigned i, unsigned j) { return a[(size_t)i * cols + j]; }
    const Number& operator()(unsigned i, unsigned j) const { return a[(size_t)i * cols + j]; }
    unsigned rows, cols;
    std::vector<Number> a;
};

namespace LinearAlgebra { namespace distributed {
template <typename Number>
class Vector {
   public:
    explicit Vector(size_t n = 0) : a(n, Number()) {}
    Number& operator[](size_t i) { return a[i]; }
    const Number& operator[](size_t i) const { return a[i]; }
    std::vector<Number> a;
};
}}  // namespace LinearAlgebra::distributed

// Gauss-Lobatto nodes / weights on [0,1] (QGaussLobatto<1>), by Newton's method on (1 - x^2) P'_{n-1}(x) in long double
inline void shim_gll(unsigned n, std::vector<double>& x, std::vector<double>& w) {
    const unsigned N = n - 1;
    std::vector<long double> xs(n), ws(n);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (unsigned i = 0; i < n; i++) {
        long double t = -std::cos(pi * i / N);
        for (int it = 0; it < 100; it++) {
            long double p0 = 1.0L, p1 = t;
            for (unsigned k = 2; k <= N; k++) { const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
            // p1 = P_N(t), p0 = P_{N-1}(t); q(t) = t P_N - P_{N-1} has the GLL nodes as roots
            const long double q = t * p1 - p0, dq = (N + 1) * p1;
            const long double dt = q / dq;
            t -= dt;
            if (std::fabs((double)dt) < 1e-20) break;
        }
        if (i == 0) t = -1.0L;
        if (i == N) t = 1.0L;
        long double p0 = 1.0L, p1 = t;
        for (unsigned k = 2; k <= N; k++) { const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
        if (N == 1) p1 = t;
        xs[i] = t;
        ws[i] = 2.0L / (N * (N + 1) * p1 * p1);
    }
    x.resize(n); w.resize(n);
    for (unsigned i = 0; i < n; i++) { x[i] = (double)((xs[i] + 1.0L) / 2.0L); w[i] = (double)(ws[i] / 2.0L); }
}

template <int dim>
class Quadrature {
   public:
    Quadrature() {}
    const std::vector<double>& get_weights() const { return weights; }
    Point<dim> point(unsigned q) const { return points[q]; }
    std::vector<Point<dim>> points;
    std::vector<double> weights;
};
// tensor Gauss-Lobatto rule, x fastest
template <int dim>
class QGaussLobatto : public Quadrature<dim> {
   public:
    explicit QGaussLobatto(unsigned n) {
        std::vector<double> x, w;
        shim_gll(n, x, w);
        unsigned total = 1;
        for (int d = 0; d < dim; d++) total *= n;
        for (unsigned q = 0; q < total; q++) {
            Point<dim> p;
            double wq = 1.0;
            unsigned r = q;
            for (int d = 0; d < dim; d++) { p[d] = x[r % n]; wq *= w[r % n]; r /= n; }
            this->points.push_back(p);
            this->weights.push_back(wq);
        }
    }
};

// FE_DGQ(p)^5 as far as shape_grad(l, point) for l < Np at points of the x-axis goes: shape function l (first component,
// base index l) is l_l(x) * l_0(y) [* l_0(z)] with l the Lagrange polynomials through the Gauss-Lobatto nodes
template <int dim>
class FiniteElementShim {
   public:
    explicit FiniteElementShim(unsigned degree) { std::vector<double> w; shim_gll(degree + 1, nodes, w); }
    double lagrange(unsigned i, double x) const {
        long double v = 1.0L;
        for (unsigned m = 0; m < nodes.size(); m++) if (m != i) v *= ((long double)x - nodes[m]) / ((long double)nodes[i] - nodes[m]);
        return (double)v;
    }
    double lagrange_derivative(unsigned i, double x) const {
        long double s = 0.0L;
        for (unsigned k = 0; k < nodes.size(); k++) {
            if (k == i) continue;
            long double t = 1.0L / ((long double)nodes[i] - nodes[k]);
            for (unsigned m = 0; m < nodes.size(); m++) if (m != i && m != k) t *= ((long double)x - nodes[m]) / ((long double)nodes[i] - nodes[m]);
            s += t;
        }
        return (double)s;
    }
    Tensor<1, dim, double> shape_grad(unsigned l, const Point<dim>& p) const {
        const unsigned n = (unsigned)nodes.size();
        unsigned idx[3] = {l % n, (l / n) % n, l / (n * n)};
        Tensor<1, dim, double> g;
        for (int d = 0; d < dim; d++) {
            double v = 1.0;
            for (int e = 0; e < dim; e++) v *= (e == d) ? lagrange_derivative(idx[e], p[e]) : lagrange(idx[e], p[e]);
            g[d] = v;
        }
        return g;
    }
    std::vector<double> nodes;
};
template <int dim> class MappingShim {};

template <int dim>
class FEValues {
   public:
    FEValues(const MappingShim<dim>&, const FiniteElementShim<dim>& fe, const Quadrature<dim>& q, UpdateFlags) : fe(fe), quad(q) {}
    const Quadrature<dim>& get_quadrature() const { return quad; }
    const FiniteElementShim<dim>& get_fe() const { return fe; }
   private:
    FiniteElementShim<dim> fe;
    Quadrature<dim> quad;
};

// One cell of FEEvaluation<dim, -1, 0, n_components, Number>
template <int dim, int fe_degree, int n_q_points_1d, int n_components, typename Number>
class FEEvaluation {
   public:
    typedef VectorizedArray<Number> VA;
    FEEvaluation(unsigned Np, const Number* u /*[n_components][NN]*/, const Number* jinv /*[NN][dim][dim] = J^{-T}*/) : Np(Np) {
        NN = 1;
        for (int d = 0; d < dim; d++) NN *= Np;
        QGaussLobatto<dim> q(Np);
        w = q.get_weights();
        vals.assign(u, u + (size_t)n_components * NN);
        Jinv.assign(jinv, jinv + (size_t)NN * dim * dim);
        submitted.assign((size_t)n_components * NN, Number());
    }
    std::vector<unsigned> quadrature_point_indices() const { std::vector<unsigned> v(NN); for (unsigned i = 0; i < NN; i++) v[i] = i; return v; }
    Tensor<1, n_components, VA> get_value(unsigned q) const {
        Tensor<1, n_components, VA> t;
        for (int c = 0; c < n_components; c++) t[c] = VA(vals[(size_t)c * NN + q]);
        return t;
    }
    Tensor<2, dim, VA> inverse_jacobian(unsigned q) const {
        Tensor<2, dim, VA> t;
        for (int r = 0; r < dim; r++) for (int c = 0; c < dim; c++) t[TableIndices<2>(r, c)] = VA(Jinv[((size_t)q * dim + r) * dim + c]);
        return t;
    }
    Number JxW(unsigned q) const {   // |det J| * weight, det J = 1 / det(J^{-T})
        Number det;
        const Number* K = &Jinv[(size_t)q * dim * dim];
        if (dim == 1) det = K[0];
        else if (dim == 2) det = K[0] * K[3] - K[1] * K[2];
        else det = K[0] * (K[4] * K[8] - K[5] * K[7]) - K[1] * (K[3] * K[8] - K[5] * K[6]) + K[2] * (K[3] * K[7] - K[4] * K[6]);
        return (Number(1) / det) * w[q];
    }
    void submit_value(const Tensor<1, n_components, VA>& v, unsigned q) { for (int c = 0; c < n_components; c++) submitted[(size_t)c * NN + q] = v[c].v; }
    void integrate_scatter(EvaluationFlags::EvaluationFlags, LinearAlgebra::distributed::Vector<Number>& dst) {
        for (int c = 0; c < n_components; c++)
            for (unsigned q = 0; q < NN; q++) dst[(size_t)c * NN + q] += submitted[(size_t)c * NN + q] * JxW(q);
    }
    unsigned Np, NN;
    std::vector<Number> vals, Jinv, submitted, w;
};

}  // namespace dealii
