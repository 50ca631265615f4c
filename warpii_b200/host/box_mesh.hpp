// Flat mesh tables for a structured box, built once on the host and uploaded to the GPU.
//
// Restates the semantics of WarpII's HyperRectangle grid (src/grid_descriptions.cc:51-74:
// GridGenerator::subdivided_hyper_rectangle(colorize = true) + collect_periodic_faces): nx[d] cells between
// left[d] and right[d], boundary ids 2d / 2d+1 on the low / high face of dimension d, periodic pairing per
// dimension.  On top of that it shards the elements over ranks (slabs along the last dimension), orders each
// rank's elements interface-first and emits the face-pair / boundary / halo tables of include/warpii_gpu.h.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "warpii_gpu.h"

namespace warpii_b200 {

struct BoxDescription {
    int dim = 1;
    std::array<int, 3> nx{{1, 1, 1}};
    std::array<double, 3> left{{0, 0, 0}}, right{{1, 1, 1}};
    std::array<bool, 3> periodic{{true, true, true}};   // periodic_dimensions default "x,y,z"
};

class BoxMeshTables {
   public:
    // elems_per_block: warpii_gpu_elems_per_block(dim, fe_degree); owned elements are numbered patch by patch
    // (patches of that many elements, as cubic as powers of two allow) so that most faces are block-internal.
    BoxMeshTables(const BoxDescription& box, int rank = 0, int n_ranks = 1, int elems_per_block = 1)
        : box_(box), rank_(rank), n_ranks_(n_ranks), group_(elems_per_block < 1 ? 1 : elems_per_block) {
        if (box.dim < 1 || box.dim > 3) throw std::invalid_argument("n_dims must be 1, 2, or 3");
        for (int d = 0; d < box.dim; d++) {
            if (box.nx[d] < 1) throw std::invalid_argument("nx must be positive");
            if (!(box.right[d] > box.left[d])) throw std::invalid_argument("right must exceed left");
        }
        if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::invalid_argument("bad rank / n_ranks");
        if (n_ranks > box.nx[box.dim - 1]) throw std::invalid_argument("more ranks than element layers in the sharded dimension");
        build();
    }

    const BoxDescription& box() const { return box_; }
    int64_t n_local() const { return (int64_t)local_to_global_.size(); }
    int64_t n_global() const { int64_t n = 1; for (int d = 0; d < box_.dim; d++) n *= box_.nx[d]; return n; }
    int64_t n_ghost_faces() const { return n_ghost_; }
    int64_t n_interface() const { return n_interface_; }
    double h(int d) const { return (box_.right[d] - box_.left[d]) / box_.nx[d]; }

    const std::vector<int64_t>& local_to_global() const { return local_to_global_; }
    const std::vector<int32_t>& face_neighbor() const { return face_neighbor_; }
    const std::vector<int32_t>& boundary_face_elem() const { return bf_elem_; }
    const std::vector<int32_t>& boundary_face_side() const { return bf_side_; }
    const std::vector<int32_t>& boundary_face_id() const { return bf_id_; }
    const std::vector<int32_t>& peer_rank() const { return peer_rank_; }
    const std::vector<int64_t>& send_offset() const { return send_offset_; }
    const std::vector<int64_t>& recv_offset() const { return recv_offset_; }
    const std::vector<int32_t>& send_elem() const { return send_elem_; }
    const std::vector<int32_t>& send_side() const { return send_side_; }
    // for tests: global element index and side of the remote face behind every ghost slot
    const std::vector<int64_t>& ghost_global_elem() const { return ghost_gelem_; }
    const std::vector<int32_t>& ghost_side() const { return ghost_side_; }

    void elem_multi_index(int64_t g, int idx[3]) const {
        for (int d = 0; d < 3; d++) idx[d] = 0;
        for (int d = 0; d < box_.dim; d++) { idx[d] = (int)(g % box_.nx[d]); g /= box_.nx[d]; }
    }
    int64_t elem_global_index(const int idx[3]) const {
        int64_t g = 0;
        for (int d = box_.dim - 1; d >= 0; d--) g = g * box_.nx[d] + idx[d];
        return g;
    }
    // owner of a global element: slabs of the last dimension, layers [r*n/R, (r+1)*n/R)
    int owner(int64_t g) const {
        int idx[3];
        elem_multi_index(g, idx);
        const int layer = idx[box_.dim - 1], n = box_.nx[box_.dim - 1];
        int r = (int)(((int64_t)layer * n_ranks_) / n);
        while (r + 1 < n_ranks_ && layer >= layer_begin(r + 1)) r++;
        while (r > 0 && layer < layer_begin(r)) r--;
        return r;
    }
    int layer_begin(int r) const { return (int)(((int64_t)r * box_.nx[box_.dim - 1]) / n_ranks_); }

    // neighbour across face f = 2*d + side in GLOBAL numbering; -1 - boundary_id on a non-periodic boundary
    int64_t global_neighbor(int64_t g, int f) const {
        int idx[3];
        elem_multi_index(g, idx);
        const int d = f / 2, side = f % 2;
        int i = idx[d] + (side ? 1 : -1);
        if (i < 0 || i >= box_.nx[d]) {
            if (!box_.periodic[d]) return -1 - f;   // colorize = true: boundary id 2d + side
            i = (i + box_.nx[d]) % box_.nx[d];
        }
        idx[d] = i;
        return elem_global_index(idx);
    }

    // fill the C struct (pointers stay owned by this object)
    void fill(warpii_gpu_mesh& m, int fe_degree, int n_species, bool fields_enabled, double gas_gamma,
              int n_boundaries, const std::vector<int32_t>& bc_kind, int n_vectors = 2) const {
        m = warpii_gpu_mesh{};
        m.dim = box_.dim;
        m.fe_degree = fe_degree;
        m.n_species = n_species;
        m.fields_enabled = fields_enabled ? 1 : 0;
        m.gas_gamma = gas_gamma;
        m.n_elems = n_local();
        m.n_ghost_faces = n_ghost_;
        m.n_boundary_faces = (int64_t)bf_elem_.size();
        m.n_boundaries = n_boundaries;
        for (int d = 0; d < 3; d++) m.h[d] = d < box_.dim ? h(d) : 1.0;
        m.face_neighbor = face_neighbor_.data();
        m.boundary_face_elem = bf_elem_.data();
        m.boundary_face_side = bf_side_.data();
        m.boundary_face_id = bf_id_.data();
        m.bc_kind = bc_kind.empty() ? nullptr : bc_kind.data();
        m.n_vectors = n_vectors;
    }
    void fill(warpii_gpu_halo& hl) const {
        hl = warpii_gpu_halo{};
        hl.n_peers = (int32_t)peer_rank_.size();
        hl.peer_rank = peer_rank_.data();
        hl.send_offset = send_offset_.data();
        hl.send_elem = send_elem_.data();
        hl.send_side = send_side_.data();
        hl.recv_offset = recv_offset_.data();
        hl.n_interface_elems = n_interface_;
    }

   private:
    void build() {
        const int nf = 2 * box_.dim;
        const int64_t ng = n_global();
        // owned elements, interface ones (any face owned by another rank) first
        std::vector<int64_t> iface, inner;
        for (int64_t g = 0; g < ng; g++) {
            if (owner(g) != rank_) continue;
            bool touches = false;
            for (int f = 0; f < nf; f++) {
                const int64_t nb = global_neighbor(g, f);
                if (nb >= 0 && owner(nb) != rank_) touches = true;
            }
            (touches ? iface : inner).push_back(g);
        }
        // patch-major numbering inside each group: key = (patch index, index within the patch), x fastest in both
        int shape[3] = {1, 1, 1};
        for (int g = group_, d = 0; g > 1; g /= 2, d = (d + 1) % box_.dim) shape[d] *= 2;
        auto patch_key = [&](int64_t g) {
            int idx[3];
            elem_multi_index(g, idx);
            int64_t patch = 0, within = 0;
            for (int d = box_.dim - 1; d >= 0; d--) {
                patch = patch * ((box_.nx[d] + shape[d] - 1) / shape[d]) + idx[d] / shape[d];
                within = within * shape[d] + idx[d] % shape[d];
            }
            return patch * group_ + within;
        };
        auto sort_by_patch = [&](std::vector<int64_t>& v) {
            std::vector<std::pair<int64_t, int64_t>> keyed(v.size());
            for (size_t i = 0; i < v.size(); i++) keyed[i] = {patch_key(v[i]), v[i]};
            std::sort(keyed.begin(), keyed.end());
            for (size_t i = 0; i < v.size(); i++) v[i] = keyed[i].second;
        };
        if (group_ > 1) {
            sort_by_patch(iface);
            sort_by_patch(inner);
        }
        n_interface_ = (int64_t)iface.size();
        local_to_global_ = iface;
        local_to_global_.insert(local_to_global_.end(), inner.begin(), inner.end());
        std::vector<int32_t> g2l((size_t)ng, -1);
        for (size_t l = 0; l < local_to_global_.size(); l++) g2l[(size_t)local_to_global_[l]] = (int32_t)l;

        // ghost faces: (peer, remote global elem, remote side) sorted => slot order both sides can derive
        struct Ghost { int peer; int64_t gelem; int side; int32_t lelem; int lface; };
        std::vector<Ghost> ghosts;
        for (int64_t l = 0; l < n_interface_; l++) {
            const int64_t g = local_to_global_[l];
            for (int f = 0; f < nf; f++) {
                const int64_t nb = global_neighbor(g, f);
                if (nb >= 0 && owner(nb) != rank_) ghosts.push_back({owner(nb), nb, f ^ 1, (int32_t)l, f});
            }
        }
        std::sort(ghosts.begin(), ghosts.end(), [](const Ghost& a, const Ghost& b) {
            if (a.peer != b.peer) return a.peer < b.peer;
            if (a.gelem != b.gelem) return a.gelem < b.gelem;
            return a.side < b.side;
        });
        n_ghost_ = (int64_t)ghosts.size();

        face_neighbor_.assign((size_t)n_local() * nf, 0);
        for (int64_t l = 0; l < n_local(); l++) {
            const int64_t g = local_to_global_[l];
            for (int f = 0; f < nf; f++) {
                const int64_t nb = global_neighbor(g, f);
                if (nb < 0) {
                    face_neighbor_[(size_t)l * nf + f] = -1 - (int32_t)bf_elem_.size();
                    bf_elem_.push_back((int32_t)l);
                    bf_side_.push_back(f);
                    bf_id_.push_back((int32_t)(-1 - nb));
                } else if (owner(nb) == rank_) {
                    face_neighbor_[(size_t)l * nf + f] = g2l[(size_t)nb];
                }
            }
        }
        peer_rank_.clear();
        recv_offset_.assign(1, 0);
        for (size_t s = 0; s < ghosts.size(); s++) {
            const Ghost& gh = ghosts[s];
            if (peer_rank_.empty() || peer_rank_.back() != gh.peer) {
                if (!peer_rank_.empty()) recv_offset_.push_back((int64_t)s);
                peer_rank_.push_back(gh.peer);
            }
            face_neighbor_[(size_t)gh.lelem * nf + gh.lface] = (int32_t)(n_local() + (int64_t)s);
            ghost_gelem_.push_back(gh.gelem);
            ghost_side_.push_back(gh.side);
        }
        if (!peer_rank_.empty()) recv_offset_.push_back(n_ghost_);

        // send lists: my faces that are ghosts of peer q, in the order q sorts them: (my global elem, my side)
        struct Send { int peer; int64_t gelem; int side; int32_t lelem; };
        std::vector<Send> sends;
        for (int64_t l = 0; l < n_interface_; l++) {
            const int64_t g = local_to_global_[l];
            for (int f = 0; f < nf; f++) {
                const int64_t nb = global_neighbor(g, f);
                if (nb >= 0 && owner(nb) != rank_) sends.push_back({owner(nb), g, f, (int32_t)l});
            }
        }
        std::sort(sends.begin(), sends.end(), [](const Send& a, const Send& b) {
            if (a.peer != b.peer) return a.peer < b.peer;
            if (a.gelem != b.gelem) return a.gelem < b.gelem;
            return a.side < b.side;
        });
        send_offset_.assign(1, 0);
        size_t pi = 0;
        for (size_t s = 0; s < sends.size(); s++) {
            while (pi < peer_rank_.size() && peer_rank_[pi] != sends[s].peer) { pi++; send_offset_.push_back((int64_t)s); }
            send_elem_.push_back(sends[s].lelem);
            send_side_.push_back(sends[s].side);
        }
        while (send_offset_.size() < peer_rank_.size() + 1) send_offset_.push_back((int64_t)sends.size());
        if (peer_rank_.empty()) { send_offset_.assign(1, 0); recv_offset_.assign(1, 0); }
    }

    BoxDescription box_;
    int rank_, n_ranks_, group_;
    int64_t n_ghost_ = 0, n_interface_ = 0;
    std::vector<int64_t> local_to_global_;
    std::vector<int32_t> face_neighbor_, bf_elem_, bf_side_, bf_id_;
    std::vector<int32_t> peer_rank_, send_elem_, send_side_, ghost_side_;
    std::vector<int64_t> send_offset_, recv_offset_, ghost_gelem_;
};

}  // namespace warpii_b200
