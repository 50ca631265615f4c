// Metric terms of a mesh of curved (or merely non-rectangular) quadrilaterals / hexahedra.
//
// Stands in for what the reference obtains from deal.II's MappingQ(fe_degree) + MatrixFree mapping info
// (nodal_dg_discretization.h:80, nodal_dg_discretization.cc:15-28) and then reads per quadrature point:
//   FEEvaluation::inverse_jacobian(q)              -> jacobian_utils.h:14,36, fluid_flux_es_dgsem_operator.h:476
//   FEFaceEvaluation::normal_vector(q), face JxW   -> fluid_flux_es_dgsem_operator.h:317-339 (GLL face nodes, quadrature 1)
//                                                     and :361-437 (Gauss(p+2) points of boundary faces, quadrature 0)
// The geometry of an element is the degree-p tensor polynomial through its (p+1)^dim Gauss-Lobatto support points
// (MappingQ's support points), so every Jacobian is that interpolant's derivative:
//   J[a][b] = d x_a / d xi_b = sum_l l_l'(xi_b) x_a(.., l, ..),        K = J^{-T},   Jdet = det J,
//   scaled face normal N = +-Jdet * column d of K  (= +-Ja^d, the contravariant vector of jacobian_utils.h:32-40),
//   unit normal N / |N|, surface Jacobian |N|.
// Interior faces get ONE normal, computed on the element with the lower index (deal.II's "interior" cell of the face
// loop) and mirrored bit for bit to the other side, so that both sides evaluate the same numerical flux and the scheme
// stays conservative to round-off.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "reference_element.hpp"

namespace warpii_b200 {

struct MappedMeshMetrics {
    int dim = 1, Np = 2, NN = 2, NF = 1, NG = 1;
    std::vector<double> inverse_jacobian;    // [n_elems][NN][dim][dim]
    std::vector<double> face_normal;         // [n_elems][2*dim][NF][dim]
    std::vector<double> face_jacobian;       // [n_elems][2*dim][NF]
    std::vector<double> boundary_normal;     // [n_bfaces][NG][dim]
    std::vector<double> boundary_jacobian;   // [n_bfaces][NG]
    std::vector<double> boundary_points;     // [n_bfaces][NG][dim] physical coordinates of the Gauss points
};

namespace detail {
inline int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

// K = J^{-T} and det J of a dim x dim matrix J[a][b]
inline double invert_transpose(int dim, const double J[3][3], double K[3][3]) {
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[r][c] = 0.0;
    if (dim == 1) {
        K[0][0] = 1.0 / J[0][0];
        return J[0][0];
    }
    if (dim == 2) {
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double inv = 1.0 / det;
        // J^{-1} = inv * [[J11, -J01], [-J10, J00]];  K[r][c] = J^{-1}[c][r]
        K[0][0] = J[1][1] * inv;
        K[1][0] = -J[0][1] * inv;
        K[0][1] = -J[1][0] * inv;
        K[1][1] = J[0][0] * inv;
        return det;
    }
    double C[3][3];   // cofactors: C[a][b] = cofactor of J[a][b];  J^{-1} = C^T / det  =>  K = J^{-T} = C / det
    C[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    C[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    C[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    C[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    C[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    C[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    C[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * C[0][0] + J[0][1] * C[0][1] + J[0][2] * C[0][2];
    const double inv = 1.0 / det;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[r][c] = C[r][c] * inv;
    return det;
}
}  // namespace detail

// xyz[n_elems][NN][dim]: the elements' GLL support points, nodes lexicographic with x fastest (the layout of the state).
// face_neighbor / neighbor_face / boundary tables as in warpii_gpu_mesh / warpii_gpu_geometry (include/warpii_gpu.h);
// entries >= n_elems of face_neighbor (ghost traces of a sharded run) take the element's own normal.
inline MappedMeshMetrics build_mapped_metrics(int dim, int fe_degree, int64_t n_elems, const double* xyz,
                                              const int32_t* face_neighbor, const int32_t* neighbor_face, int64_t n_bfaces,
                                              const int32_t* bf_elem, const int32_t* bf_side) {
    if (dim < 1 || dim > 3) throw std::invalid_argument("n_dims must be 1, 2, or 3");
    ReferenceElement re(fe_degree);
    MappedMeshMetrics M;
    const int Np = re.Np, Ng = re.Ng;
    M.dim = dim;
    M.Np = Np;
    M.NN = detail::ipow(Np, dim);
    M.NF = detail::ipow(Np, dim - 1);
    M.NG = detail::ipow(Ng, dim - 1);
    const int NN = M.NN, NF = M.NF, NG = M.NG, nf = 2 * dim;
    auto stride = [&](int d) { return d == 0 ? 1 : (d == 1 ? Np : Np * Np); };
    auto face_node = [&](int d, int side, int t) {
        int idx[3] = {0, 0, 0};
        for (int a = 0; a < dim; a++) {
            if (a == d) continue;
            idx[a] = t % Np;
            t /= Np;
        }
        idx[d] = side ? Np - 1 : 0;
        return idx[0] + Np * (idx[1] + Np * idx[2]);
    };

    M.inverse_jacobian.assign((size_t)n_elems * NN * dim * dim, 0.0);
    M.face_normal.assign((size_t)n_elems * nf * NF * dim, 0.0);
    M.face_jacobian.assign((size_t)n_elems * nf * NF, 0.0);
    std::vector<double> jdet((size_t)n_elems * NN, 0.0);
    for (int64_t e = 0; e < n_elems; e++) {
        const double* xe = xyz + (size_t)e * NN * dim;
        for (int q = 0; q < NN; q++) {
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, K[3][3];
            for (int b = 0; b < dim; b++) {
                const int st = stride(b), jb = (q / st) % Np, base = q - jb * st;
                for (int l = 0; l < Np; l++) {
                    const double dl = re.D[jb * Np + l];
                    for (int a = 0; a < dim; a++) J[a][b] += dl * xe[(size_t)(base + l * st) * dim + a];
                }
            }
            const double det = detail::invert_transpose(dim, J, K);
            if (!(det > 0.0)) throw std::invalid_argument("mapped mesh: non-positive Jacobian determinant (inverted element)");
            jdet[(size_t)e * NN + q] = det;
            for (int r = 0; r < dim; r++)
                for (int c = 0; c < dim; c++) M.inverse_jacobian[(((size_t)e * NN + q) * dim + r) * dim + c] = K[r][c];
        }
    }
    // face normals at the GLL face nodes
    for (int64_t e = 0; e < n_elems; e++)
        for (int f = 0; f < nf; f++) {
            const int64_t v = face_neighbor[(size_t)e * nf + f];
            const int code = neighbor_face ? neighbor_face[(size_t)e * nf + f] : (f ^ 1);
            const int nfa = code & 7, flip = code >> 3;
            // the face is owned by the lower (element, face) of its two sides; boundary and ghost faces by the element
            const bool mirrored = v >= 0 && v < n_elems && (v < e || (v == e && nfa < f));
            if (mirrored) continue;
            const int d = f / 2, side = f % 2;
            for (int t = 0; t < NF; t++) {
                const int q = face_node(d, side, t);
                const double* K = &M.inverse_jacobian[((size_t)e * NN + q) * dim * dim];
                const double J = jdet[(size_t)e * NN + q];
                double N[3] = {0, 0, 0}, n2 = 0.0;
                for (int r = 0; r < dim; r++) {
                    N[r] = (side ? 1.0 : -1.0) * (J * K[r * dim + d]);
                    n2 += N[r] * N[r];
                }
                const double len = std::sqrt(n2);
                double* out = &M.face_normal[(((size_t)e * nf + f) * NF + t) * dim];
                for (int r = 0; r < dim; r++) out[r] = N[r] / len;
                M.face_jacobian[((size_t)e * nf + f) * NF + t] = len;
                if (v >= 0 && v < n_elems) {
                    const int tp = flip ? NF - 1 - t : t;
                    double* o2 = &M.face_normal[(((size_t)v * nf + nfa) * NF + tp) * dim];
                    for (int r = 0; r < dim; r++) o2[r] = -out[r];
                    M.face_jacobian[((size_t)v * nf + nfa) * NF + tp] = len;
                }
            }
        }
    // boundary faces at the Gauss(p+2) points: derivative of the interpolant at non-nodal points
    M.boundary_normal.assign((size_t)n_bfaces * NG * dim, 0.0);
    M.boundary_jacobian.assign((size_t)n_bfaces * NG, 0.0);
    M.boundary_points.assign((size_t)n_bfaces * NG * dim, 0.0);
    std::vector<double> dIg((size_t)Ng * Np, 0.0);   // dIg[g*Np+i] = l_i'(xg_g)
    for (int g = 0; g < Ng; g++)
        for (int i = 0; i < Np; i++) {
            long double s = 0.0L;
            for (int m = 0; m < Np; m++) {
                if (m == i) continue;
                long double prod = 1.0L / ((long double)re.x[i] - re.x[m]);
                for (int k = 0; k < Np; k++)
                    if (k != i && k != m) prod *= ((long double)re.xg[g] - re.x[k]) / ((long double)re.x[i] - re.x[k]);
                s += prod;
            }
            dIg[(size_t)g * Np + i] = (double)s;
        }
    for (int64_t b = 0; b < n_bfaces; b++) {
        const int64_t e = bf_elem[b];
        const int f = bf_side[b], d = f / 2, side = f % 2, end = side ? Np - 1 : 0;
        const double* xe = xyz + (size_t)e * NN * dim;
        for (int g = 0; g < NG; g++) {
            int gi[2] = {g % Ng, (g / Ng) % Ng};
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, K[3][3], X[3] = {0, 0, 0};
            for (int q = 0; q < NN; q++) {
                int idx[3] = {q % Np, (q / Np) % Np, q / (Np * Np)};
                // value and derivative factors of the 1-D basis functions at this point, per direction
                double val[3] = {1, 1, 1}, der[3] = {0, 0, 0};
                int k = 0;
                for (int a = 0; a < dim; a++) {
                    if (a == d) {
                        val[a] = (idx[a] == end) ? 1.0 : 0.0;
                        der[a] = re.D[end * Np + idx[a]];
                    } else {
                        val[a] = re.Ig[gi[k] * Np + idx[a]];
                        der[a] = dIg[(size_t)gi[k] * Np + idx[a]];
                        k++;
                    }
                }
                double phi = 1.0;
                for (int a = 0; a < dim; a++) phi *= val[a];
                for (int a = 0; a < dim; a++) X[a] += phi * xe[(size_t)q * dim + a];
                for (int bdir = 0; bdir < dim; bdir++) {
                    double dphi = 1.0;
                    for (int a = 0; a < dim; a++) dphi *= (a == bdir) ? der[a] : val[a];
                    if (dphi == 0.0) continue;
                    for (int a = 0; a < dim; a++) J[a][bdir] += dphi * xe[(size_t)q * dim + a];
                }
            }
            const double det = detail::invert_transpose(dim, J, K);
            double N[3] = {0, 0, 0}, n2 = 0.0;
            for (int r = 0; r < dim; r++) {
                N[r] = (side ? 1.0 : -1.0) * (det * K[r][d]);
                n2 += N[r] * N[r];
            }
            const double len = std::sqrt(n2);
            for (int r = 0; r < dim; r++) {
                M.boundary_normal[((size_t)b * NG + g) * dim + r] = N[r] / len;
                M.boundary_points[((size_t)b * NG + g) * dim + r] = X[r];
            }
            M.boundary_jacobian[(size_t)b * NG + g] = len;
        }
    }
    return M;
}

}  // namespace warpii_b200
