// Meshes that are not Cartesian boxes: what the reference gets from `GridType = Extension` (src/extensions/extension.h:24-45,
// src/grid.h: a GridExtension fills a dealii::Triangulation<dim>) and from MappingQ(fe_degree) on it.
//
//   Triangulation2D     stands in for dealii::Triangulation<2>: vertices, cells (deal.II vertex order v00, v10, v01, v11),
//                       boundary ids of boundary faces (default 0, as in deal.II); what an extension populates
//   GridExtension       the reference's GridExtension<dim> interface (same method names and argument meaning)
//   GeneralMesh         flat tables for the C ABI: face pairing with (face, orientation) codes, boundary faces, and the
//                       Gauss-Lobatto support points of every element (straight-sided elements: the bilinear image of the
//                       reference nodes, which is what MappingQ produces without a manifold; or any mapped box)
// The metric terms follow from the support points (mapped_mesh.hpp) and go to warpii_gpu_set_geometry.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "box_mesh.hpp"
#include "parameter_file.hpp"
#include "reference_element.hpp"

namespace warpii_b200 {

struct Triangulation2D {
    std::vector<std::array<double, 2>> vertices;
    std::vector<std::array<int, 4>> cells;                    // (v00, v10, v01, v11)
    std::map<std::pair<int, int>, int> boundary_ids;          // (cell, local face) -> id; unset boundary faces have id 0

    void clear() { vertices.clear(); cells.clear(); boundary_ids.clear(); }
    // local faces 0..3 = x-low, x-high, y-low, y-high; their vertices in face-node order
    static std::pair<int, int> face_vertices(int f) {
        static const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
        return {fv[f][0], fv[f][1]};
    }
    std::array<double, 2> face_center(int cell, int f) const {
        const auto fv = face_vertices(f);
        const auto &a = vertices[cells[cell][fv.first]], &b = vertices[cells[cell][fv.second]];
        return {{0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1])}};
    }
    // subdivided_hyper_rectangle without colorize, minus the cells whose centre satisfies `removed`
    // (GridGenerator::create_triangulation_with_removed_cells): enough for L-shaped and stepped channels
    void subdivided_rectangle(int nx, int ny, double x0, double y0, double x1, double y1,
                              const std::function<bool(double, double)>& removed = nullptr) {
        clear();
        std::vector<int> vid((size_t)(nx + 1) * (ny + 1), -1);
        const double hx = (x1 - x0) / nx, hy = (y1 - y0) / ny;
        auto vertex = [&](int i, int j) {
            int& v = vid[(size_t)j * (nx + 1) + i];
            if (v < 0) {
                v = (int)vertices.size();
                vertices.push_back({{x0 + i * hx, y0 + j * hy}});
            }
            return v;
        };
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                if (removed && removed(x0 + (i + 0.5) * hx, y0 + (j + 0.5) * hy)) continue;
                cells.push_back({{vertex(i, j), vertex(i + 1, j), vertex(i, j + 1), vertex(i + 1, j + 1)}});
            }
    }
};

// src/extensions/extension.h:24-45
class GridExtension {
   public:
    virtual ~GridExtension() = default;
    // prm is scoped to the `geometry` subsection
    virtual void declare_geometry_parameters(ParameterFile& /*prm*/) {}
    virtual void populate_triangulation(Triangulation2D& /*tria*/, const ParameterFile& /*prm*/) {}
};

struct GeneralMesh {
    int dim = 2, fe_degree = 1;
    int64_t n_elems = 0;
    std::vector<int32_t> face_neighbor;   // [n_elems][2*dim]
    std::vector<int32_t> neighbor_face;   // [n_elems][2*dim]; empty = opposite face, same order
    std::vector<int32_t> bf_elem, bf_side, bf_id;
    std::vector<double> xyz;              // [n_elems][NN][dim]
    // element-sharded mapped boxes only (empty / zero otherwise): ghost trace slots and the halo lists of this rank
    int rank = 0, n_ranks = 1;
    int64_t n_ghost_faces = 0, n_interface = 0;
    std::vector<int64_t> local_to_global, send_offset, recv_offset;
    std::vector<int32_t> peer_rank, send_elem, send_side;
    void fill(warpii_gpu_halo& hl) const {
        hl = warpii_gpu_halo{};
        hl.n_peers = (int32_t)peer_rank.size();
        hl.peer_rank = peer_rank.data();
        hl.send_offset = send_offset.data();
        hl.send_elem = send_elem.data();
        hl.send_side = send_side.data();
        hl.recv_offset = recv_offset.data();
        hl.n_interface_elems = n_interface;
    }

    // connectivity of a Triangulation2D + bilinear support points
    static GeneralMesh from_triangulation(const Triangulation2D& tria, int fe_degree) {
        GeneralMesh m;
        m.dim = 2;
        m.fe_degree = fe_degree;
        m.n_elems = (int64_t)tria.cells.size();
        const int64_t n = m.n_elems;
        struct Side { int64_t cell; int face; int va, vb; };
        std::map<std::pair<int, int>, std::vector<Side>> edges;
        for (int64_t e = 0; e < n; e++)
            for (int f = 0; f < 4; f++) {
                const auto fv = Triangulation2D::face_vertices(f);
                const int va = tria.cells[e][fv.first], vb = tria.cells[e][fv.second];
                if (va < 0 || vb < 0 || va >= (int)tria.vertices.size() || vb >= (int)tria.vertices.size())
                    throw std::invalid_argument("triangulation: cell refers to a vertex that does not exist");
                edges[{std::min(va, vb), std::max(va, vb)}].push_back({e, f, va, vb});
            }
        m.face_neighbor.assign((size_t)n * 4, 0);
        m.neighbor_face.assign((size_t)n * 4, 0);
        for (int64_t e = 0; e < n; e++)
            for (int f = 0; f < 4; f++) {
                const auto fv = Triangulation2D::face_vertices(f);
                const int va = tria.cells[e][fv.first], vb = tria.cells[e][fv.second];
                const std::vector<Side>& sides = edges[{std::min(va, vb), std::max(va, vb)}];
                if (sides.size() == 1) {
                    m.face_neighbor[(size_t)e * 4 + f] = -1 - (int32_t)m.bf_elem.size();
                    m.neighbor_face[(size_t)e * 4 + f] = f ^ 1;
                    m.bf_elem.push_back((int32_t)e);
                    m.bf_side.push_back(f);
                    const auto it = tria.boundary_ids.find({(int)e, f});
                    m.bf_id.push_back(it == tria.boundary_ids.end() ? 0 : it->second);
                    continue;
                }
                if (sides.size() != 2) throw std::invalid_argument("triangulation: an edge is shared by more than two cells");
                const Side& o = (sides[0].cell == e && sides[0].face == f) ? sides[1] : sides[0];
                m.face_neighbor[(size_t)e * 4 + f] = (int32_t)o.cell;
                m.neighbor_face[(size_t)e * 4 + f] = o.face + ((o.va == vb && o.vb == va) ? 8 : 0);
            }
        const ReferenceElement re(fe_degree);
        const int Np = re.Np, NN = Np * Np;
        m.xyz.assign((size_t)n * NN * 2, 0.0);
        for (int64_t e = 0; e < n; e++) {
            const auto &v00 = tria.vertices[tria.cells[e][0]], &v10 = tria.vertices[tria.cells[e][1]],
                       &v01 = tria.vertices[tria.cells[e][2]], &v11 = tria.vertices[tria.cells[e][3]];
            // counter-clockwise cells only: a clockwise cell has a negative Jacobian (deal.II rejects it as well)
            const double area2 = (v10[0] - v00[0]) * (v01[1] - v00[1]) - (v10[1] - v00[1]) * (v01[0] - v00[0]);
            if (!(area2 > 0.0)) throw std::invalid_argument("triangulation: cell with non-positive orientation");
            for (int j = 0; j < NN; j++) {
                const double xi = re.x[j % Np], eta = re.x[j / Np];
                for (int a = 0; a < 2; a++)
                    m.xyz[((size_t)e * NN + j) * 2 + a] = (1 - xi) * (1 - eta) * v00[a] + xi * (1 - eta) * v10[a] +
                                                          (1 - xi) * eta * v01[a] + xi * eta * v11[a];
            }
        }
        return m;
    }

    // the connectivity of a box (this rank's slab of it), support points pushed through x' = mapping(x): curved elements.
    // On a sharded run every rank evaluates the normals of its interface faces from its own element; the two sides agree
    // to round-off (they are functions of the shared face nodes), so the run conserves to round-off like a single-GPU
    // one, without being bit-identical to it.
    static GeneralMesh mapped_box(const BoxDescription& box, int fe_degree, int elems_per_block,
                                  const std::function<void(const double* x, double* x_out)>& mapping, int rank = 0,
                                  int n_ranks = 1) {
        GeneralMesh m;
        const BoxMeshTables t(box, rank, n_ranks, elems_per_block);
        m.rank = rank;
        m.n_ranks = n_ranks;
        m.n_ghost_faces = t.n_ghost_faces();
        m.n_interface = t.n_interface();
        m.local_to_global = t.local_to_global();
        m.peer_rank = t.peer_rank();
        m.send_offset = t.send_offset();
        m.recv_offset = t.recv_offset();
        m.send_elem = t.send_elem();
        m.send_side = t.send_side();
        m.dim = box.dim;
        m.fe_degree = fe_degree;
        m.n_elems = t.n_local();
        m.face_neighbor = t.face_neighbor();
        m.bf_elem = t.boundary_face_elem();
        m.bf_side = t.boundary_face_side();
        m.bf_id = t.boundary_face_id();
        const ReferenceElement re(fe_degree);
        const int Np = re.Np;
        int NN = 1;
        for (int d = 0; d < box.dim; d++) NN *= Np;
        m.xyz.assign((size_t)m.n_elems * NN * box.dim, 0.0);
        for (int64_t l = 0; l < m.n_elems; l++) {
            int idx[3];
            t.elem_multi_index(t.local_to_global()[l], idx);
            for (int j = 0; j < NN; j++) {
                double x[3] = {0, 0, 0}, y[3] = {0, 0, 0};
                int tt = j;
                for (int d = 0; d < box.dim; d++) { x[d] = box.left[d] + (idx[d] + re.x[tt % Np]) * t.h(d); tt /= Np; }
                if (mapping) mapping(x, y);
                else for (int d = 0; d < box.dim; d++) y[d] = x[d];
                for (int d = 0; d < box.dim; d++) m.xyz[((size_t)l * NN + j) * box.dim + d] = y[d];
            }
        }
        return m;
    }
};

}  // namespace warpii_b200
