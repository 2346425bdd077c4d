"""Structured meshes generated directly in array form, with the reference's node/element/equation order.

The reference builds these through Python triple loops and a kd-tree de-duplication
(/root/reference/pyoomph/meshes/simplemeshes.py:224-305 RectangularQuadMesh, :456-540 CuboidBrickMesh,
src/meshtemplate.cpp:1378-1398 add_node_unique), then converts every C1 template element to C2 in element
order (src/meshtemplate.cpp:401-409 quads, :579-620 bricks).  That is O(minutes-hours) at the BASELINE
sizes, so the same *ordering rules* are applied here with vectorised first-occurrence numbering:

* elements: x outermost ... last coordinate innermost;
* all vertex nodes precede all mid nodes; within each class nodes are numbered at first touch, visiting
  elements in order and, inside an element, in the order the reference creates them;
* element-local node layout = oomph tensor-product order (first local coordinate fastest);
* equations: nodes in order; per node position dofs first, then nodal values by index, pinned ones skipped
  (oomph-lib mesh.cc:686-708, nodes.cc:896-927, :3652-3659).  C1 fields own a (pinned, dummy) slot on
  non-vertex nodes as well (src/elements.cpp:2235-2271).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np

from .codegen import FiniteElementCode


@dataclasses.dataclass
class StructuredMesh:
    dim: int
    N: Tuple[int, ...]
    elem_nodes: np.ndarray           # [n_elem, nnode] int32, oomph local order
    node_pos: np.ndarray             # [n_node, dim] float64
    node_lattice: np.ndarray         # [n_node, dim] int32 position on the (2N+1)^dim lattice
    boundaries: Dict[str, np.ndarray]  # name -> node indices
    element_type: str

    @property
    def n_elem(self) -> int:
        return self.elem_nodes.shape[0]

    @property
    def n_node(self) -> int:
        return self.node_pos.shape[0]

    def is_vertex(self) -> np.ndarray:
        return np.all(self.node_lattice % 2 == 0, axis=1)

    def element_patches(self, width=None) -> np.ndarray:
        """Locality hint for the GPU schedule: id of the compact patch (8x8 quads / 4x4x4 bricks) of every element.
        width: one int or one per dimension (rectangular patches, e.g. (8, 4))."""
        import os
        w = width or os.environ.get("PB2_PATCH_WIDTH", "8" if self.dim == 2 else "4")      # "8" or per dimension "8x4"
        if isinstance(w, str):
            w = [int(x) for x in w.split("x")]
            w = w[0] if len(w) == 1 else w
        ws = [int(w)] * self.dim if np.isscalar(w) else [int(x) for x in w]
        grids = np.meshgrid(*[np.arange(n, dtype=np.int64) // ws[d] for d, n in enumerate(self.N)], indexing="ij")
        npd = [(n + ws[d] - 1) // ws[d] for d, n in enumerate(self.N)]
        pid = grids[0].ravel()
        for d in range(1, self.dim):
            pid = pid * npd[d] + grids[d].ravel()
        if self.element_type == "Tri2dC2":
            pid = np.repeat(pid, 2)          # two triangles per lattice cell, stored one after the other
        if self.element_type == "Tetra3dC2":
            pid = np.repeat(pid, 6)          # six tetrahedra per lattice cell
        return pid.astype(np.int32)


def _first_occurrence_ids(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """ids by first occurrence in the flattened key stream; returns (ids shaped like keys, first index per id)."""
    flat = keys.ravel()
    uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv].reshape(keys.shape), first[order]


def _structured(N: Sequence[int], size: Sequence[float], lower_left: Sequence[float]) -> StructuredMesh:
    dim = len(N)
    N = tuple(int(n) for n in N)
    # element index grid, first coordinate outermost (simplemeshes.py:225/:233, :512-514)
    grids = np.meshgrid(*[np.arange(n, dtype=np.int64) for n in N], indexing="ij")
    eidx = np.stack([g.ravel() for g in grids], axis=1)          # [n_elem, dim]
    n_elem = eidx.shape[0]
    L = [2 * n + 1 for n in N]                                   # lattice extents

    def lattice_key(lat):  # lat [..., dim]
        k = lat[..., 0]
        for d in range(1, dim):
            k = k * L[d] + lat[..., d]
        return k

    # creation order inside an element, as lattice offsets in {0,1,2}^dim (x fastest in the oomph local index)
    if dim == 2:
        vert_order = [(0, 0), (2, 0), (0, 2), (2, 2)]            # n00,n10,n01,n11 (simplemeshes.py:234-240)
        mid_order = [(1, 0), (0, 1), (1, 1), (2, 1), (1, 2)]     # meshtemplate.cpp:403-407
    else:
        vert_order = [(0, 0, 0), (2, 0, 0), (0, 2, 0), (2, 2, 0), (0, 0, 2), (2, 0, 2), (0, 2, 2), (2, 2, 2)]
        # meshtemplate.cpp:583-617: local indices 1,3,4,5,7, 9,11,15,17, 19,21,22,23,25, 10,12,14,16,13
        loc = [1, 3, 4, 5, 7, 9, 11, 15, 17, 19, 21, 22, 23, 25, 10, 12, 14, 16, 13]
        mid_order = [(l % 3, (l // 3) % 3, l // 9) for l in loc]
    base = 2 * eidx                                              # lattice origin of each element
    vkeys = np.stack([lattice_key(base + np.array(o)) for o in vert_order], axis=1)
    vid, vfirst = _first_occurrence_ids(vkeys)
    n_vert = vfirst.size
    mkeys = np.stack([lattice_key(base + np.array(o)) for o in mid_order], axis=1)
    mid, mfirst = _first_occurrence_ids(mkeys)
    n_node = n_vert + mfirst.size

    # oomph local order: index = sum_d off_d * 3^d
    nnode = 3 ** dim
    elem_nodes = np.empty((n_elem, nnode), dtype=np.int32)
    for k, o in enumerate(vert_order):
        elem_nodes[:, sum(o[d] * 3 ** d for d in range(dim))] = vid[:, k]
    for k, o in enumerate(mid_order):
        elem_nodes[:, sum(o[d] * 3 ** d for d in range(dim))] = n_vert + mid[:, k]

    # lattice coordinate of every node
    node_lat = np.empty((n_node, dim), dtype=np.int32)
    for k, o in enumerate(vert_order):
        node_lat[vid[:, k]] = base + np.array(o)
    for k, o in enumerate(mid_order):
        node_lat[n_vert + mid[:, k]] = base + np.array(o)

    # coordinates: vertices (i*size)/N + lower_left (simplemeshes.py:234); mid nodes by averaging the bracketing
    # vertices the way the creating element does (0.5*(a+b), 0.25*(a+b+c+d))
    def vcoord(i, d):
        return (i * size[d]) / N[d] + lower_left[d]

    node_pos = np.empty((n_node, dim), dtype=np.float64)
    for d in range(dim):
        lat = node_lat[:, d].astype(np.int64)
        lo, hi = lat // 2, (lat + 1) // 2
        a, b = vcoord(lo.astype(np.float64), d), vcoord(hi.astype(np.float64), d)
        # odd lattice coordinate: average of the two bracketing vertices.  (The reference averages 4 vertices for
        # centre nodes, 0.25*(a+b+c+d); on these uniform lattices that is the same number up to one ulp.)
        node_pos[:, d] = np.where(lat % 2 == 1, 0.5 * (a + b), a)
    names = [("left", "right"), ("bottom", "top"), ("back", "front")]
    boundaries = {}
    for d in range(dim):
        boundaries[names[d][0]] = np.nonzero(node_lat[:, d] == 0)[0].astype(np.int64)
        boundaries[names[d][1]] = np.nonzero(node_lat[:, d] == 2 * N[d])[0].astype(np.int64)
    return StructuredMesh(dim, N, elem_nodes, node_pos, node_lat, boundaries, "Quad2dC2" if dim == 2 else "Brick3dC2")


def RectangularQuadMesh(N=10, size=1.0, lower_left=(0.0, 0.0)) -> StructuredMesh:
    """Q9 mesh of N[0] x N[1] elements (simplemeshes.py:111)."""
    N = (N, N) if np.isscalar(N) else tuple(N)
    size = (size, size) if np.isscalar(size) else tuple(size)
    return _structured(N, [float(s) for s in size], [float(v) for v in lower_left])


def RectangularTriangleMesh(N=10, size=1.0, lower_left=(0.0, 0.0)) -> StructuredMesh:
    """Six-node triangles (BulkElementTri2dC2 = oomph TElement<2,3>, src/elements.hpp:990) on the node set of the Q9 mesh of N[0] x N[1]
    cells: every cell is cut along its lower-left -> upper-right diagonal into two counter-clockwise triangles; the cell's centre node
    becomes the mid-side node of the diagonal.  Local node order of TElementShape<2,3> (Telements.h:575-621): vertices 0, 1, 2, then
    the mid-side nodes 3 (between 0 and 1), 4 (1 and 2), 5 (2 and 0).  Node numbering, positions, boundaries and equation numbering are
    those of the quad mesh (vertices before mid nodes); a stand-in for the gmsh triangle meshes of the reference's droplet
    scripts, whose element order is the generator's."""
    q = RectangularQuadMesh(N, size, lower_left)
    en = q.elem_nodes
    a = en[:, [0, 2, 8, 1, 5, 4]]          # lower-right triangle: ll, lr, ur | mids ll-lr, lr-ur, ur-ll (the cell centre)
    b = en[:, [0, 8, 6, 4, 7, 3]]          # upper-left triangle:  ll, ur, ul | mids ll-ur (the cell centre), ur-ul, ul-ll
    tri = np.empty((2 * en.shape[0], 6), dtype=np.int32)
    tri[0::2], tri[1::2] = a, b
    m = StructuredMesh(2, q.N, np.ascontiguousarray(tri), q.node_pos, q.node_lattice, q.boundaries, "Tri2dC2")
    return m


def boundary_face_mesh(mesh: StructuredMesh, names: Sequence[str]) -> InterfaceMesh:
    """The boundary edges of a RectangularQuadMesh as FACES OF THEIR BULK ELEMENTS (element type QuadFace2dC2): per edge the nine nodes
    of the bulk element, rotated (orientation preserved) so that the edge is the local face s1 = -1 traversed in the direction of
    increasing s0; the outer normal is then (t_y, -t_x)/|t|.  Rotations of the 3x3 node grid: right face (i, j) <- (2 - j, i), top face
    (i, j) <- (2 - i, 2 - j), left face (i, j) <- (j, 2 - i)."""
    if mesh.dim != 2 or mesh.element_type != "Quad2dC2":
        raise NotImplementedError("bulk faces: Q9 meshes only")
    Nx, Ny = mesh.N
    rot = {"bottom": [i + 3 * j for j in range(3) for i in range(3)],
           "right": [(2 - j) + 3 * i for j in range(3) for i in range(3)],
           "top": [(2 - i) + 3 * (2 - j) for j in range(3) for i in range(3)],
           "left": [j + 3 * (2 - i) for j in range(3) for i in range(3)]}
    elems = {"bottom": [ex * Ny for ex in range(Nx)], "top": [ex * Ny + Ny - 1 for ex in range(Nx)],
             "left": list(range(Ny)), "right": [(Nx - 1) * Ny + ey for ey in range(Ny)]}
    en, be, fi = [], [], []
    face_id = {"left": -1, "right": 1, "bottom": -2, "top": 2}
    for name in names:
        for e in elems[name]:
            en.append(mesh.elem_nodes[e][rot[name]])
            be.append(e)
            fi.append(face_id[name])
    en = np.ascontiguousarray(np.stack(en), dtype=np.int32)
    return InterfaceMesh(2, mesh.N, en, mesh.node_pos, mesh.node_lattice, mesh.boundaries, "QuadFace2dC2", np.array(be, dtype=np.int64),
                         np.array(fi, dtype=np.int32), mesh.is_vertex())


def CuboidTetraMesh(N=4, size=1.0, lower_left=(0.0, 0.0, 0.0)) -> StructuredMesh:
    """Ten-node tetrahedra (BulkElementTetra3dC2 = oomph TElement<3,3>, src/elements.cpp:11397) on the node set of the Q27 mesh: every
    cell is cut into the six Kuhn tetrahedra around its main diagonal (one per order in which the three axes are walked from the
    lower corner to the upper one); the edge, face-diagonal and body-diagonal midpoints are exactly the cell's remaining lattice nodes.
    Local node order of TElementShape<3,3> (Telements.h:2051-2127): vertices 0-3 = (1,0,0), (0,1,0), (0,0,1), (0,0,0) in local
    coordinates, mid-edge nodes 4: 0-1, 5: 0-2, 6: 0-3, 7: 1-2, 8: 2-3, 9: 1-3.  All tetrahedra positively oriented."""
    import itertools
    b = CuboidBrickMesh(N, size, lower_left)
    loc = lambda o: o[0] + 3 * o[1] + 9 * o[2]
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=int)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] += 2
            v.append(nxt)
        # oomph local vertices 0, 1, 2 span the tetrahedron from vertex 3: take v[3] (the upper corner) as local 3
        quad = [v[0], v[1], v[2], v[3]]
        e1, e2, e3 = quad[0] - quad[3], quad[1] - quad[3], quad[2] - quad[3]
        if np.linalg.det(np.array([e1, e2, e3], dtype=float)) < 0:
            quad[0], quad[1] = quad[1], quad[0]
        mids = [(0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (1, 3)]
        tets.append([loc(q) for q in quad] + [loc((quad[i] + quad[j]) // 2) for i, j in mids])
    tets = np.array(tets, dtype=np.int64)                       # [6, 10] local brick-node indices
    en = b.elem_nodes[:, tets].reshape(-1, 10)
    return StructuredMesh(3, b.N, np.ascontiguousarray(en, dtype=np.int32), b.node_pos, b.node_lattice, b.boundaries, "Tetra3dC2")


@dataclasses.dataclass
class InterfaceMesh:
    """Line elements (InterfaceElementLine1dC2, src/elements.hpp:1435-2298) on boundary edges of a 2D bulk mesh: they live on the bulk
    mesh's nodes (same node numbers, positions and nodal values) and remember the bulk element and face they were built from."""
    dim: int
    N: Tuple[int, ...]
    elem_nodes: np.ndarray           # [n_elem, 3] bulk node numbers, ordered so that (-t_y, t_x) is the OUTER normal
    node_pos: np.ndarray
    node_lattice: np.ndarray
    boundaries: Dict[str, np.ndarray]
    element_type: str
    bulk_element: np.ndarray         # [n_elem] bulk element the edge belongs to
    face_index: np.ndarray           # [n_elem] oomph face index of that edge (-1: s0=-1, 1: s0=+1, -2: s1=-1, 2: s1=+1)
    is_vertex_mask: np.ndarray

    @property
    def n_elem(self) -> int:
        return self.elem_nodes.shape[0]

    @property
    def n_node(self) -> int:
        return self.node_pos.shape[0]

    def is_vertex(self) -> np.ndarray:
        return self.is_vertex_mask


def boundary_line_mesh(mesh: StructuredMesh, names: Sequence[str]) -> InterfaceMesh:
    """The Q9 edges on the named boundaries ("left", "right", "bottom", "top") of a RectangularQuadMesh as three-node line elements.
    BulkElementBase::get_normal_at_s gives a line element the normal (-t_y, t_x)/|t| (src/elements.cpp:1730-1752); the reference's
    interface elements orient it with FaceElement::normal_sign so that it leaves the bulk (outer_unit_normal, src/elements.hpp:1630).
    Here the node order of every edge is chosen so that (-t_y, t_x) IS the outer normal: bottom edges run right -> left, right edges
    top -> bottom, top edges left -> right, left edges bottom -> top (pinned against the compiled oomph-lib FaceElement in
    tests/test_oracle_ref.py)."""
    if mesh.dim != 2 or mesh.element_type != "Quad2dC2":
        raise ValueError("boundary_line_mesh needs a 2D Q9 mesh")
    nx, ny = mesh.N
    e_of = lambda ix, iy: ix * ny + iy
    local = {"bottom": ((2, 1, 0), -2), "right": ((8, 5, 2), 1), "top": ((6, 7, 8), 2), "left": ((0, 3, 6), -1)}
    en, be, fi = [], [], []
    for nm in names:
        loc, face = local[nm]
        if nm == "bottom":
            els = [e_of(ix, 0) for ix in range(nx)]
        elif nm == "top":
            els = [e_of(ix, ny - 1) for ix in range(nx)]
        elif nm == "left":
            els = [e_of(0, iy) for iy in range(ny)]
        else:
            els = [e_of(nx - 1, iy) for iy in range(ny)]
        for e in els:
            en.append([int(mesh.elem_nodes[e, l]) for l in loc])
            be.append(e)
            fi.append(face)
    return InterfaceMesh(2, mesh.N, np.ascontiguousarray(np.array(en, dtype=np.int32).reshape(-1, 3)), mesh.node_pos, mesh.node_lattice,
                         mesh.boundaries, "Line1dC2", np.array(be, dtype=np.int32), np.array(fi, dtype=np.int32), mesh.is_vertex())


def CuboidBrickMesh(N=4, size=1.0, lower_left=(0.0, 0.0, 0.0)) -> StructuredMesh:
    """Q27 mesh of N[0] x N[1] x N[2] elements (simplemeshes.py:456)."""
    N = (N, N, N) if np.isscalar(N) else tuple(N)
    size = (size, size, size) if np.isscalar(size) else tuple(size)
    return _structured(N, [float(s) for s in size], [float(v) for v in lower_left])


@dataclasses.dataclass
class HangingNodes:
    """Hanging-node constraints per interpolation space (oomph-lib HangInfo per value index; pyoomph keeps one set for the C2 / position
    values and one for the C1 values of a node, src/elements.cpp:812-1160): node -> (master nodes, weights); the value at the node is
    sum_k weight_k * value(master_k) and it has no equation of its own."""
    C2: Dict[int, Tuple[np.ndarray, np.ndarray]]
    C1: Dict[int, Tuple[np.ndarray, np.ndarray]]

    def of_space(self, space: str) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
        return self.C1 if space == "C1" else self.C2


@dataclasses.dataclass
class RefinedQuadMesh:
    """A Q9 mesh after one level of quadtree refinement of some elements (RefineableQElement<2> sons SW, SE, NW, NE): conforming
    through hanging nodes on the edges between refined and unrefined elements."""
    dim: int
    elem_nodes: np.ndarray
    node_pos: np.ndarray
    node_lattice: np.ndarray          # position on the (4N+1)^2 lattice of the refined level
    boundaries: Dict[str, np.ndarray]
    element_type: str
    vertex_mask: np.ndarray
    hanging: HangingNodes
    parent_element: np.ndarray        # [n_elem] element of the unrefined mesh

    @property
    def n_elem(self) -> int:
        return self.elem_nodes.shape[0]

    @property
    def n_node(self) -> int:
        return self.node_pos.shape[0]

    def is_vertex(self) -> np.ndarray:
        return self.vertex_mask


def _lag3(s: float) -> np.ndarray:
    return np.array([0.5 * s * (s - 1.0), 1.0 - s * s, 0.5 * s * (s + 1.0)])


def refine_quad_mesh(mesh: StructuredMesh, refine: np.ndarray) -> RefinedQuadMesh:
    """Split the elements flagged in `refine` ([n_elem] bool, element order of `mesh`) into their four sons.  New nodes are placed by the
    father's Q9 mapping (so a distorted mesh is refined on its own geometry); on an edge between a refined and an unrefined element the
    two quarter-point nodes hang on the three nodes of the coarse edge with the quadratic weights psi(-1/2), psi(+1/2) (C2 / position
    values) and the coarse mid-edge node -- a vertex of the sons -- hangs on the two end vertices with weights 1/2 for the C1 values
    (oomph-lib refineable_quad_element.cc, quad_hang_helper)."""
    if mesh.dim != 2 or mesh.element_type != "Quad2dC2":
        raise NotImplementedError("refinement with hanging nodes: Q9 meshes only")
    refine = np.asarray(refine, dtype=bool)
    Nx, Ny = mesh.N
    assert refine.shape == (mesh.n_elem,)
    key_of = {}
    lat = mesh.node_lattice.astype(np.int64) * 2
    for n in range(mesh.n_node):
        key_of[(int(lat[n, 0]), int(lat[n, 1]))] = n
    pos = [mesh.node_pos[n].copy() for n in range(mesh.n_node)]
    flat = [tuple(int(v) for v in lat[n]) for n in range(mesh.n_node)]
    elems, parent = [], []
    psi1 = {k: _lag3(0.5 * k - 1.0) for k in range(5)}
    for e in range(mesh.n_elem):
        en = mesh.elem_nodes[e]
        if not refine[e]:
            elems.append(en.copy())
            parent.append(e)
            continue
        ex, ey = e // Ny, e % Ny
        X = mesh.node_pos[en]                      # [9, 2], local index i + 3 j
        grid = np.empty((5, 5), dtype=np.int64)
        for fj in range(5):
            for fi in range(5):
                key = (4 * ex + fi, 4 * ey + fj)
                if key not in key_of:
                    w = np.outer(psi1[fj], psi1[fi]).ravel()        # psi_{i + 3 j}(s) = L_i(s0) L_j(s1)
                    key_of[key] = len(pos)
                    pos.append(w @ X)
                    flat.append(key)
                grid[fi, fj] = key_of[key]
        for b in range(2):              # sons SW, SE, NW, NE
            for a in range(2):
                elems.append(np.array([grid[2 * a + i, 2 * b + j] for j in range(3) for i in range(3)], dtype=np.int32))
                parent.append(e)
    hang_C2, hang_C1 = {}, {}
    rgrid = refine.reshape(Nx, Ny)
    for ex in range(Nx):
        for ey in range(Ny):
            if not rgrid[ex, ey]:
                continue
            # (neighbour offset, fine keys of the five nodes along the shared edge)
            edges = [((-1, 0), [(4 * ex, 4 * ey + k) for k in range(5)]), ((1, 0), [(4 * ex + 4, 4 * ey + k) for k in range(5)]),
                     ((0, -1), [(4 * ex + k, 4 * ey) for k in range(5)]), ((0, 1), [(4 * ex + k, 4 * ey + 4) for k in range(5)])]
            for (dx, dy), keys in edges:
                nx, ny = ex + dx, ey + dy
                if not (0 <= nx < Nx and 0 <= ny < Ny) or rgrid[nx, ny]:
                    continue
                nodes = [key_of[k] for k in keys]
                masters = np.array([nodes[0], nodes[2], nodes[4]], dtype=np.int64)
                hang_C2[nodes[1]] = (masters, _lag3(-0.5))
                hang_C2[nodes[3]] = (masters, _lag3(0.5))
                hang_C1[nodes[2]] = (np.array([nodes[0], nodes[4]], dtype=np.int64), np.array([0.5, 0.5]))
    elem_nodes = np.ascontiguousarray(np.stack(elems), dtype=np.int32)
    node_pos = np.ascontiguousarray(np.stack(pos), dtype=np.float64)
    # positions of hanging nodes are the interpolation of their masters (BulkElementBase::interpolate_hang_values, src/elements.cpp:448)
    for n, (m, w) in hang_C2.items():
        node_pos[n] = w @ node_pos[m]
    node_lat = np.array(flat, dtype=np.int32)
    vertex = np.zeros(node_pos.shape[0], dtype=bool)
    vertex[elem_nodes[:, [0, 2, 6, 8]].ravel()] = True
    boundaries = {"left": np.nonzero(node_lat[:, 0] == 0)[0], "right": np.nonzero(node_lat[:, 0] == 4 * Nx)[0],
                  "bottom": np.nonzero(node_lat[:, 1] == 0)[0], "top": np.nonzero(node_lat[:, 1] == 4 * Ny)[0]}
    return RefinedQuadMesh(2, elem_nodes, node_pos, node_lat, boundaries, "Quad2dC2", vertex, HangingNodes(hang_C2, hang_C1), np.array(parent, dtype=np.int64))


def refine_brick_mesh(mesh: StructuredMesh, refine: np.ndarray) -> RefinedQuadMesh:
    """One level of octree refinement of the flagged Q27 elements (RefineableQElement<3>: eight sons).  New nodes are placed by the father's
    Q27 mapping.  A coarse FACE or EDGE that is shared by at least one refined and one unrefined element is constrained: the new nodes on
    it hang on the coarse nodes of that entity -- the nine nodes of a face with the weights L_i(s_a) L_j(s_b), the three nodes of an edge
    with L_i(s) (oomph-lib refineable_brick_element.cc, oc_hang_helper; zero weights dropped).  Only the C2 / position hang infos are
    generated (element classes with C1 fields are refused by the equation numbering on these meshes)."""
    if mesh.dim != 3 or mesh.element_type != "Brick3dC2":
        raise NotImplementedError("octree refinement: Q27 meshes only")
    refine = np.asarray(refine, dtype=bool)
    Nx, Ny, Nz = mesh.N
    assert refine.shape == (mesh.n_elem,)
    rgrid = refine.reshape(Nx, Ny, Nz)
    lat = mesh.node_lattice.astype(np.int64) * 2
    key_of = {tuple(int(v) for v in lat[n]): n for n in range(mesh.n_node)}
    pos = [mesh.node_pos[n].copy() for n in range(mesh.n_node)]
    flat = [tuple(int(v) for v in lat[n]) for n in range(mesh.n_node)]
    psi1 = {k: _lag3(0.5 * k - 1.0) for k in range(5)}
    elems, parent = [], []
    for e in range(mesh.n_elem):
        en = mesh.elem_nodes[e]
        if not refine[e]:
            elems.append(en.copy())
            parent.append(e)
            continue
        ex, ey, ez = e // (Ny * Nz), (e // Nz) % Ny, e % Nz
        X = mesh.node_pos[en]                      # [27, 3], local index i + 3 j + 9 k
        grid = np.empty((5, 5, 5), dtype=np.int64)
        for fk in range(5):
            for fj in range(5):
                for fi in range(5):
                    key = (4 * ex + fi, 4 * ey + fj, 4 * ez + fk)
                    if key not in key_of:
                        w = np.einsum("k,j,i->kji", psi1[fk], psi1[fj], psi1[fi]).ravel()
                        key_of[key] = len(pos)
                        pos.append(w @ X)
                        flat.append(key)
                    grid[fi, fj, fk] = key_of[key]
        for c in range(2):
            for b in range(2):
                for a in range(2):
                    elems.append(np.array([grid[2 * a + i, 2 * b + j, 2 * c + k] for k in range(3) for j in range(3) for i in range(3)], dtype=np.int32))
                    parent.append(e)

    def is_refined(cx, cy, cz):
        return None if not (0 <= cx < Nx and 0 <= cy < Ny and 0 <= cz < Nz) else bool(rgrid[cx, cy, cz])
    hang_C2 = {}
    for key, n in key_of.items():
        if all(k % 2 == 0 for k in key):
            continue                                   # a node of the coarse level (vertex, mid-edge, mid-face, centre): never C2-hanging
        on = [k % 4 == 0 for k in key]                 # lies on a coarse lattice plane in that direction
        if sum(on) == 0:
            continue                                   # interior of a coarse element
        # the coarse elements sharing the entity (face: 2, edge: 4)
        cells = [[]]
        for d in range(3):
            opts = [key[d] // 4 - 1, key[d] // 4] if on[d] else [key[d] // 4]
            cells = [c + [o] for c in cells for o in opts]
        states = [is_refined(*c) for c in cells]
        if not (any(s_ is False for s_ in states) and any(s_ is True for s_ in states)):
            continue
        # masters: the coarse nodes of the entity; local coordinate of the node in every free direction
        free = [d for d in range(3) if not on[d]]
        base = [key[d] - (key[d] % 4) if not on[d] else key[d] for d in range(3)]
        masters, weights = [], []
        import itertools
        for idx in itertools.product(range(3), repeat=len(free)):
            mk = list(base)
            wgt = 1.0
            for d, i in zip(free, idx):
                mk[d] = base[d] + 2 * i
                wgt *= psi1[key[d] - base[d]][i]
            if wgt != 0.0:
                masters.append(key_of[tuple(mk)])
                weights.append(wgt)
        hang_C2[n] = (np.array(masters, dtype=np.int64), np.array(weights))
    elem_nodes = np.ascontiguousarray(np.stack(elems), dtype=np.int32)
    node_pos = np.ascontiguousarray(np.stack(pos), dtype=np.float64)
    for n, (m, w) in hang_C2.items():
        node_pos[n] = w @ node_pos[m]
    node_lat = np.array(flat, dtype=np.int32)
    vertex = np.zeros(node_pos.shape[0], dtype=bool)
    vertex[elem_nodes[:, [0, 2, 6, 8, 18, 20, 24, 26]].ravel()] = True
    names = [("left", "right"), ("bottom", "top"), ("back", "front")]
    boundaries = {}
    for d, nd in enumerate((Nx, Ny, Nz)):
        boundaries[names[d][0]] = np.nonzero(node_lat[:, d] == 0)[0]
        boundaries[names[d][1]] = np.nonzero(node_lat[:, d] == 4 * nd)[0]
    return RefinedQuadMesh(3, elem_nodes, node_pos, node_lat, boundaries, "Brick3dC2", vertex, HangingNodes(hang_C2, {}), np.array(parent, dtype=np.int64))


@dataclasses.dataclass
class DofMap:
    node_eqn: np.ndarray             # [n_node, nval] int32, -1 pinned
    pos_eqn: Optional[np.ndarray]    # [n_node, dim] int32 or None
    n_dof: int


def assign_equation_numbers(mesh: StructuredMesh, code: FiniteElementCode,
                            pinned: Optional[Dict[str, Iterable[int]]] = None,
                            pinned_positions: Optional[Dict[str, Iterable[int]]] = None) -> DofMap:
    """Global equation numbers in oomph's order (mesh.cc:686-708): node by node; positions first.  Hanging values (mesh.hanging) are
    constrained: no equation (oomph-lib Data::Is_constrained)."""
    nval = code.n_nodal_values
    n_node, dim = mesh.n_node, mesh.dim
    free = np.ones((n_node, nval), dtype=bool)
    vertex = mesh.is_vertex()
    for f in code.nodal_fields():
        if f.space == "C1":
            free[~vertex, f.index] = False     # dummy values on non-vertex nodes (src/elements.cpp:2235-2271)
    hanging = getattr(mesh, "hanging", None)
    if hanging is not None:
        for f in code.nodal_fields():
            if f.space == "C1" and mesh.dim == 3:
                raise NotImplementedError("C1 fields on octree-refined meshes (only the C2 hang infos are generated in 3D)")
            hn = np.fromiter(hanging.of_space(f.space).keys(), dtype=np.int64)
            if hn.size:
                free[hn, f.index] = False
    for name, nodes in (pinned or {}).items():
        free[np.asarray(list(nodes) if not isinstance(nodes, np.ndarray) else nodes, dtype=np.int64), code.fields[name].index] = False
    if code.coordinates_as_dofs:
        pfree = np.ones((n_node, dim), dtype=bool)
        if hanging is not None and hanging.C2:
            pfree[np.fromiter(hanging.C2.keys(), dtype=np.int64), :] = False       # hanging positions follow their masters
        for name, nodes in (pinned_positions or {}).items():
            d = "xyz".index(name[-1])
            pfree[np.asarray(list(nodes) if not isinstance(nodes, np.ndarray) else nodes, dtype=np.int64), d] = False
        allfree = np.concatenate([pfree, free], axis=1)
    else:
        pfree = None
        allfree = free
    flat = allfree.ravel()
    eq = np.where(flat, np.cumsum(flat) - 1, -1).astype(np.int32).reshape(allfree.shape)
    n_dof = int(flat.sum())
    if pfree is not None:
        return DofMap(np.ascontiguousarray(eq[:, dim:]), np.ascontiguousarray(eq[:, :dim]), n_dof)
    return DofMap(np.ascontiguousarray(eq), None, n_dof)


def spatial_patches(node_pos: np.ndarray, elem_nodes: np.ndarray, patch_elems: int = 64) -> np.ndarray:
    """Locality hint (`pb2_mesh_desc.elem_patch`) for meshes that come without one -- external generators, refined or relabelled
    meshes: elements are ranked along a Morton (Z-order) curve through their centroids and cut into patches of `patch_elems`
    consecutive ranks.  Patches are then compact, touch few other patches (few patch colours = few tile gates per assembly) and
    their interior CSR rows are completed while resident in L2.  Returns the patch id of every element (int32); the element order
    itself is not changed (the engine keeps its own schedule order)."""
    pos = np.asarray(node_pos, dtype=np.float64)
    cen = pos[np.asarray(elem_nodes)].mean(axis=1)                       # [n_elem, dim]
    lo, hi = cen.min(axis=0), cen.max(axis=0)
    bits = 20 if pos.shape[1] == 3 else 30
    q = np.clip(((cen - lo) / np.maximum(hi - lo, 1e-300) * ((1 << bits) - 1)).astype(np.uint64), 0, (1 << bits) - 1)
    key = np.zeros(cen.shape[0], dtype=np.uint64)
    dim = pos.shape[1]
    for b in range(bits):
        for d in range(dim):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * dim + d)
    rank = np.empty(cen.shape[0], dtype=np.int64)
    rank[np.argsort(key, kind="stable")] = np.arange(cen.shape[0])
    return (rank // max(1, int(patch_elems))).astype(np.int32)
