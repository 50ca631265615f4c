"""Test meshes for the general-geometry path (curved / unstructured elements): test infrastructure, numpy only.

Connectivity conventions are those of include/warpii_gpu.h (and of deal.II's hypercube numbering): local face
f = 2*d + side; 2D cell vertices (v00, v10, v01, v11); a face's nodes run along the remaining coordinate(s) in
increasing order.
"""
import numpy as np

import oracle


def gll_nodes(fe_degree):
    x, _ = oracle.gll(fe_degree + 1)
    return np.asarray(x)


def ref_nodes(dim, fe_degree):
    """Reference coordinates of the element's nodes, [NN][dim], x fastest."""
    x = gll_nodes(fe_degree)
    Np = x.size
    out = np.zeros((Np ** dim, dim))
    for j in range(Np ** dim):
        t = j
        for d in range(dim):
            out[j, d] = x[t % Np]
            t //= Np
    return out


# ---------------------------------------------------------------------------------------------------------------
# mapped box: the connectivity of a Cartesian box, node coordinates pushed through a smooth mapping
# ---------------------------------------------------------------------------------------------------------------
def box_connectivity(dim, nx, periodic):
    """Lexicographic (x fastest) element order.  face_neighbor[e][f] = neighbour or -1 - boundary face number."""
    nx = list(nx) + [1] * (3 - dim)
    n = int(np.prod(nx[:dim]))
    nbr = np.zeros((n, 2 * dim), dtype=np.int64)
    bf_elem, bf_side, bf_id = [], [], []
    for e in range(n):
        idx = [e % nx[0], (e // nx[0]) % nx[1], e // (nx[0] * nx[1])]
        for f in range(2 * dim):
            d, side = f // 2, f % 2
            i = idx[d] + (1 if side else -1)
            if i < 0 or i >= nx[d]:
                if not periodic[d]:
                    nbr[e, f] = -1 - len(bf_elem)
                    bf_elem.append(e)
                    bf_side.append(f)
                    bf_id.append(f)
                    continue
                i %= nx[d]
            j = list(idx)
            j[d] = i
            nbr[e, f] = j[0] + nx[0] * (j[1] + nx[1] * j[2])
    return {"face_neighbor": nbr, "neighbor_face": None, "bf_elem": np.array(bf_elem, dtype=np.int32),
            "bf_side": np.array(bf_side, dtype=np.int32), "bf_id": np.array(bf_id, dtype=np.int32)}


def box_node_coords(dim, fe_degree, nx, left, right):
    nx = list(nx) + [1] * (3 - dim)
    n = int(np.prod(nx[:dim]))
    ref = ref_nodes(dim, fe_degree)
    h = [(right[d] - left[d]) / nx[d] for d in range(dim)]
    xyz = np.zeros((n, ref.shape[0], dim))
    for e in range(n):
        idx = [e % nx[0], (e // nx[0]) % nx[1], e // (nx[0] * nx[1])]
        for d in range(dim):
            xyz[e, :, d] = left[d] + (idx[d] + ref[:, d]) * h[d]
    return xyz


def mapped_box(dim, fe_degree, nx, left, right, periodic, mapping=None):
    """mesh dict + xyz[n_elems][NN][dim]; mapping(xyz) -> xyz' is applied to the Cartesian node positions, so the
    geometry is the degree-p interpolant of the mapping (what MappingQ(p) does with a manifold description)."""
    mesh = box_connectivity(dim, nx, periodic)
    xyz = box_node_coords(dim, fe_degree, nx, left, right)
    if mapping is not None:
        xyz = np.ascontiguousarray(mapping(xyz))
    return mesh, xyz


def wavy(left, right, amp=0.06):
    """A smooth mapping of the box onto itself whose perturbation is periodic (and vanishes on the boundary planes'
    normal displacement), the usual curved-mesh test of DGSEM papers."""
    left, right = np.asarray(left, dtype=float), np.asarray(right, dtype=float)
    L = right - left

    def f(xyz):
        s = (xyz - left) / L   # unit coordinates
        dim = xyz.shape[-1]
        out = xyz.copy()
        for d in range(dim):
            bump = np.sin(2 * np.pi * s[..., d])
            for o in range(dim):
                if o != d:
                    bump = bump * np.cos(2 * np.pi * s[..., o] + 0.3 * (o + 1))
            # displacement along d vanishes at s_d = 0, 1 and is periodic in every coordinate
            out[..., d] = xyz[..., d] + amp * L[d] * bump * (1.0 if dim > 1 else 1.0)
        return out

    return f


def rotation2d(theta):
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s], [s, c]])
    return (lambda xyz: xyz @ R.T), R


# ---------------------------------------------------------------------------------------------------------------
# unstructured 2D quadrilaterals
# ---------------------------------------------------------------------------------------------------------------
_FACE_VERTS = [(0, 2), (1, 3), (0, 1), (2, 3)]   # local vertex pairs of faces 0..3, in face-node order


def quad_mesh(vertices, cells, fe_degree, boundary_id, warp=None):
    """vertices [nv][2], cells [nc][4] in (v00, v10, v01, v11) order.  boundary_id(midpoint) -> id.
    Returns mesh dict (with neighbor_face codes: neighbour's local face + 8 if its nodes run the other way) and xyz."""
    vertices = np.asarray(vertices, dtype=float)
    cells = np.asarray(cells, dtype=np.int64)
    n = cells.shape[0]
    edges = {}
    for e in range(n):
        for f, (a, b) in enumerate(_FACE_VERTS):
            va, vb = int(cells[e, a]), int(cells[e, b])
            edges.setdefault(frozenset((va, vb)), []).append((e, f, va, vb))
    nbr = np.zeros((n, 4), dtype=np.int64)
    nbf = np.zeros((n, 4), dtype=np.int32)
    bf_elem, bf_side, bf_id = [], [], []
    for e in range(n):
        for f, (a, b) in enumerate(_FACE_VERTS):
            va, vb = int(cells[e, a]), int(cells[e, b])
            sides = edges[frozenset((va, vb))]
            if len(sides) == 1:
                nbr[e, f] = -1 - len(bf_elem)
                nbf[e, f] = f ^ 1
                bf_elem.append(e)
                bf_side.append(f)
                bf_id.append(boundary_id(0.5 * (vertices[va] + vertices[vb])))
                continue
            assert len(sides) == 2, "non-manifold edge"
            (e2, f2, wa, wb) = sides[0] if sides[1][0] == e and sides[1][1] == f else sides[1]
            nbr[e, f] = e2
            nbf[e, f] = f2 + (8 if (wa, wb) == (vb, va) else 0)
    ref = ref_nodes(2, fe_degree)
    xi, eta = ref[:, 0][:, None], ref[:, 1][:, None]
    xyz = np.zeros((n, ref.shape[0], 2))
    for e in range(n):
        v00, v10, v01, v11 = (vertices[cells[e, k]] for k in range(4))
        xyz[e] = (1 - xi) * (1 - eta) * v00 + xi * (1 - eta) * v10 + (1 - xi) * eta * v01 + xi * eta * v11
    if warp is not None:
        xyz = np.ascontiguousarray(warp(xyz))
    mesh = {"face_neighbor": nbr, "neighbor_face": nbf, "bf_elem": np.array(bf_elem, dtype=np.int32),
            "bf_side": np.array(bf_side, dtype=np.int32), "bf_id": np.array(bf_id, dtype=np.int32)}
    return mesh, xyz


def hexagon_blocks(n, rotate_cells=True):
    """Three n x n blocks of quadrilaterals around the centre of a regular hexagon (three cells meet at the centre: no
    global (i, j) structure), every cell's local frame rotated by a cell-dependent multiple of 90 degrees, so that all
    pairings of local faces and both tangential orientations occur.  Returns vertices, cells."""
    hexv = [np.array([np.cos(k * np.pi / 3), np.sin(k * np.pi / 3)]) for k in range(6)]
    c = np.zeros(2)
    verts, index, cells = [], {}, []

    def vid(p):
        key = (round(float(p[0]), 9), round(float(p[1]), 9))
        if key not in index:
            index[key] = len(verts)
            verts.append(np.array(p, dtype=float))
        return index[key]

    for k in range(3):
        v00, v10, v11, v01 = c, hexv[2 * k], hexv[2 * k + 1], hexv[(2 * k + 2) % 6]
        P = lambda s, t: (1 - s) * (1 - t) * v00 + s * (1 - t) * v10 + (1 - s) * t * v01 + s * t * v11
        for j in range(n):
            for i in range(n):
                s0, s1, t0, t1 = i / n, (i + 1) / n, j / n, (j + 1) / n
                ccw = [vid(P(s0, t0)), vid(P(s1, t0)), vid(P(s1, t1)), vid(P(s0, t1))]   # counter-clockwise
                r = (len(cells) * 7 + k) % 4 if rotate_cells else 0
                ccw = ccw[r:] + ccw[:r]
                cells.append([ccw[0], ccw[1], ccw[3], ccw[2]])   # (v00, v10, v01, v11)
    return np.array(verts), np.array(cells, dtype=np.int64)


def hexagon_boundary_id(mid):
    """Three boundary ids by polar angle of the face midpoint."""
    ang = np.arctan2(mid[1], mid[0]) % (2 * np.pi)
    return int(ang // (2 * np.pi / 3))


def swirl_warp(amp=0.04):
    def f(xyz):
        x, y = xyz[..., 0], xyz[..., 1]
        out = xyz.copy()
        out[..., 0] = x + amp * np.sin(2.0 * y + 0.3) * np.cos(1.5 * x)
        out[..., 1] = y + amp * np.sin(2.5 * x - 0.2) * np.cos(1.0 * y)
        return out
    return f


def smooth_state(gamma, dim):
    """A smooth, non-symmetric primitive state as a function of physical position: [...,dim] -> [...,5]."""
    def f(xyz):
        x = xyz[..., 0]
        y = xyz[..., 1] if dim > 1 else 0.0 * x
        z = xyz[..., 2] if dim > 2 else 0.0 * x
        out = np.zeros(xyz.shape[:-1] + (5,))
        out[..., 0] = 1.0 + 0.2 * np.sin(1.3 * x + 0.5 * y + 0.7 * z)
        out[..., 1] = 0.4 + 0.1 * np.cos(0.9 * y + 0.2 * z)
        out[..., 2] = -0.2 + 0.1 * np.sin(1.1 * x) if dim > 1 else 0.05 + 0.0 * x
        out[..., 3] = 0.1 * np.cos(0.8 * x - 0.6 * y + 0.5 * z)
        out[..., 4] = 1.0 + 0.1 * np.cos(0.7 * x - 1.2 * y + 0.4 * z)
        return out
    return f


def periodic_state(gamma, left, right, dim):
    """Smooth primitive state that is periodic on the box [left, right] (evaluated at REFERENCE box coordinates)."""
    left, right = np.asarray(left, dtype=float), np.asarray(right, dtype=float)

    def f(xyz_box):
        s = 2 * np.pi * (xyz_box - left) / (right - left)
        a = s[..., 0]
        b = s[..., 1] if dim > 1 else 0.0 * a
        c = s[..., 2] if dim > 2 else 0.0 * a
        out = np.zeros(xyz_box.shape[:-1] + (5,))
        out[..., 0] = 1.0 + 0.2 * np.sin(a + 0.4) * np.cos(b) * np.cos(c + 0.2)
        out[..., 1] = 0.4 + 0.1 * np.cos(b + 0.1) * np.sin(a)
        out[..., 2] = -0.2 + 0.1 * np.sin(a - 0.3) if dim > 1 else 0.05 + 0.0 * a
        out[..., 3] = 0.1 * np.cos(a + b + c)
        out[..., 4] = 1.0 + 0.1 * np.cos(a - b + 0.5 * np.sin(c))
        return out
    return f


def add_kinks(prim, every=3, factor=1.6):
    """Density (and pressure) jump INSIDE every `every`-th element: under-resolved there, so the shock indicator
    switches the subcell finite-volume blend on (alpha > 0) in those elements only."""
    NN = prim.shape[1]
    prim[::every, NN // 2:, 0] *= factor
    prim[::every, NN // 2:, 4] *= 1.0 + 0.5 * (factor - 1.0)
    return prim


def to_conserved(prim, gamma):
    return oracle.primitive_to_conserved(prim, gamma)


def state_from(prim_vals, gamma, n_species=1, fields=False):
    """[n_elems][NN][5] primitive -> device layout [n_elems][nc][NN] (same state for every species)."""
    cons = np.transpose(to_conserved(prim_vals, gamma), (0, 2, 1))
    nc = 5 * n_species + (8 if fields else 0)
    u = np.zeros((cons.shape[0], nc, cons.shape[2]))
    for s in range(n_species):
        u[:, 5 * s:5 * s + 5, :] = cons * (1.0 + 0.1 * s)
    return u
