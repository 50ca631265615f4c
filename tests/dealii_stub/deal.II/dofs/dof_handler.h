#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <deal.II/base/point.h>
#include <deal.II/base/types.h>
namespace dealii {
// A structured box standing in for Triangulation + DoFHandler: nx cells per direction between left and right, boundary ids
// 2d / 2d+1 (GridGenerator::subdivided_hyper_rectangle with colorize = true), periodic pairing per direction -- what
// HyperRectangleDescription::reinit (src/grid_descriptions.cc:51-74) builds -- with the cell-iterator calls the adapter uses.
class CellId {
   public:
    explicit CellId(std::int64_t i = -1) : i(i) {}
    bool operator<(const CellId& o) const { return i < o.i; }
    bool operator==(const CellId& o) const { return i == o.i; }
   private:
    std::int64_t i;
};
template <int dim>
class DoFHandler {
   public:
    struct Face { types::boundary_id bid; types::boundary_id boundary_id() const { return bid; } };
    class cell_accessor {
       public:
        cell_accessor(const DoFHandler* dh, std::int64_t e) : dh(dh), e(e) {}
        CellId id() const { return CellId(e); }
        bool at_boundary(unsigned f) const { const int d = f / 2; const int i = idx(d) + ((f % 2) ? 1 : -1); return i < 0 || i >= dh->nx[d]; }
        bool has_periodic_neighbor(unsigned f) const { return at_boundary(f) && dh->periodic[f / 2]; }
        cell_accessor neighbor_or_periodic_neighbor(unsigned f) const {
            const int d = f / 2;
            int i = idx(d) + ((f % 2) ? 1 : -1);
            i = (i + dh->nx[d]) % dh->nx[d];
            std::int64_t stride = 1;
            for (int a = 0; a < d; a++) stride *= dh->nx[a];
            return cell_accessor(dh, e + (std::int64_t)(i - idx(d)) * stride);
        }
        std::shared_ptr<Face> face(unsigned f) const { return std::make_shared<Face>(Face{(types::boundary_id)f}); }
        double extent_in_direction(unsigned d) const { return (dh->right[d] - dh->left[d]) / dh->nx[d]; }
        Point<dim> center() const {
            Point<dim> p;
            for (int d = 0; d < dim; d++) p[d] = dh->left[d] + (idx(d) + 0.5) * extent_in_direction(d);
            return p;
        }
        Point<dim> vertex(unsigned v) const {
            Point<dim> p;
            for (int d = 0; d < dim; d++) p[d] = dh->left[d] + (idx(d) + ((v >> d) & 1)) * extent_in_direction(d);
            return p;
        }
        void get_dof_indices(std::vector<types::global_dof_index>& out) const {
            for (std::size_t k = 0; k < out.size(); k++) out[k] = (types::global_dof_index)e * out.size() + k;
        }
        const cell_accessor* operator->() const { return this; }
       private:
        int idx(int d) const { std::int64_t r = e; for (int a = 0; a < d; a++) r /= dh->nx[a]; return (int)(r % dh->nx[d]); }
        const DoFHandler* dh;
        std::int64_t e;
    };
    typedef cell_accessor active_cell_iterator;
    DoFHandler(const int* nx_, const double* left_, const double* right_, const int* periodic_) {
        n_cells = 1;
        for (int d = 0; d < dim; d++) { nx[d] = nx_[d]; left[d] = left_[d]; right[d] = right_[d]; periodic[d] = periodic_[d] != 0; n_cells *= nx[d]; }
    }
    std::vector<active_cell_iterator> active_cell_iterators() const {
        std::vector<active_cell_iterator> v;
        for (std::int64_t e = 0; e < n_cells; e++) v.emplace_back(this, e);
        return v;
    }
    int nx[3] = {1, 1, 1};
    double left[3] = {0, 0, 0}, right[3] = {1, 1, 1};
    bool periodic[3] = {true, true, true};
    std::int64_t n_cells = 1;
};
}  // namespace dealii
