#pragma once
#include <cmath>
#include <vector>
#include <deal.II/base/point.h>
namespace dealii {
template <int dim>
class Quadrature {
   public:
    unsigned size() const { return (unsigned)points.size(); }
    const Point<dim>& point(unsigned q) const { return points[q]; }
    double weight(unsigned q) const { return weights[q]; }
    std::vector<Point<dim>> points;
    std::vector<double> weights;
};
// Gauss-Legendre on [0,1]
template <int dim>
class QGauss : public Quadrature<dim> {
   public:
    explicit QGauss(unsigned n) {
        static_assert(dim == 1, "the stub provides the 1-D rule");
        const long double pi = 3.14159265358979323846264338327950288L;
        for (unsigned i = 0; i < n; i++) {
            long double t = -std::cos(pi * (i + 0.75L) / (n + 0.5L)), dp = 1;
            for (int it = 0; it < 100; it++) {
                long double p0 = 1, p1 = t;
                for (unsigned k = 2; k <= n; k++) { const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
                if (n == 1) { p0 = 1; p1 = t; }
                dp = n * (t * p1 - p0) / (t * t - 1);
                const long double dt = p1 / dp;
                t -= dt;
                if (std::fabs((double)dt) < 1e-20) break;
            }
            Point<dim> p;
            p[0] = (double)((t + 1) / 2);
            this->points.push_back(p);
            this->weights.push_back((double)(1 / ((1 - t * t) * dp * dp)));
        }
    }
};
}  // namespace dealii
