// C entry points for the reference's volume and subcell-finite-volume drivers, compiled from the reference sources by
// `make -C oracle ref` (see README.md in this directory).  The two class templates are taken, at build time, from
//   /root/reference/src/five_moment/fluxes/split_form_volume_flux.h  and  subcell_finite_volume_flux.h
// (everything from `namespace warpii {` on, i.e. the files without their #include block, written to oracle/_ref/gen/ which
// is git-ignored: nothing of the reference enters this repository) because their include blocks pull in the whole
// application (ParameterHandler, Triangulation, MatrixFree, ...), while the drivers themselves need only FEEvaluation,
// VectorizedArray, FullMatrix and the discretisation's degree / element.  This file contains no physics.
#include <memory>
#include <vector>

#include <deal.II/matrix_free/fe_evaluation.h>

#include "src/dof_utils.h"
#include "src/five_moment/euler.h"
#include "src/five_moment/fluxes/jacobian_utils.h"

namespace warpii {
// what the two drivers ask of NodalDGDiscretization<dim> (src/dgsem/nodal_dg_discretization.h:22-88)
template <int dim>
class NodalDGDiscretization {
   public:
    explicit NodalDGDiscretization(unsigned fe_degree) : fe_degree(fe_degree), fe(fe_degree) {}
    unsigned get_fe_degree() const { return fe_degree; }
    const dealii::MappingShim<dim>& get_mapping() const { return mapping; }
    const dealii::FiniteElementShim<dim>& get_fe() const { return fe; }
   private:
    unsigned fe_degree;
    dealii::MappingShim<dim> mapping;
    dealii::FiniteElementShim<dim> fe;
};
}  // namespace warpii

#include "gen/split_form_volume_flux.body.h"
#include "gen/subcell_finite_volume_flux.body.h"

using namespace dealii;
using namespace warpii;
using namespace warpii::five_moment;

namespace {
template <int dim>
void cell_residual(unsigned Np, double gamma, const double* u, const double* jinv, double alpha, int which, double* R) {
    unsigned NN = 1;
    for (int d = 0; d < dim; d++) NN *= Np;
    auto disc = std::make_shared<NodalDGDiscretization<dim>>(Np - 1);
    FEEvaluation<dim, -1, 0, 5, double> phi(Np, u, jinv), phi_reader(Np, u, jinv);
    LinearAlgebra::distributed::Vector<double> dst((size_t)5 * NN);
    if (which & 1) {
        SplitFormVolumeFlux<dim> volume(disc, gamma);
        volume.calculate_flux(dst, phi, phi_reader, VectorizedArray<double>(alpha), false);
    }
    if (which & 2) {
        SubcellFiniteVolumeFlux<dim> fv(*disc, gamma);
        fv.calculate_flux(dst, phi, phi_reader, VectorizedArray<double>(alpha), false);
    }
    for (size_t i = 0; i < (size_t)5 * NN; i++) R[i] = dst[i];
}
}  // namespace

extern "C" {
// Integrated cell residual (value * JxW, as integrate_scatter leaves it) of ONE cell: which = 1 volume
// (SplitFormVolumeFlux::calculate_flux), 2 subcell FV (SubcellFiniteVolumeFlux::calculate_flux), 3 both.
// u[5][Np^dim]; jinv[Np^dim][dim][dim] = inverse_jacobian(q) (J^{-T}); dim = 1 or 2 (the reference instantiates no 3-D).
int ref_cell_residual(int dim, int Np, double gamma, const double* u, const double* jinv, double alpha, int which, double* R) {
    if (dim == 1) cell_residual<1>((unsigned)Np, gamma, u, jinv, alpha, which, R);
    else if (dim == 2) cell_residual<2>((unsigned)Np, gamma, u, jinv, alpha, which, R);
    else return 1;
    return 0;
}
// the matrices the drivers build from the finite element: D(j,l) = shape_grad(l, x_j)[0]
void ref_diff_matrix(int Np, double* D) {
    FiniteElementShim<1> fe((unsigned)Np - 1);
    QGaussLobatto<1> q((unsigned)Np);
    for (int j = 0; j < Np; j++) for (int l = 0; l < Np; l++) D[j * Np + l] = fe.shape_grad((unsigned)l, q.point((unsigned)j))[0];
}
}
