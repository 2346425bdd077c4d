"""Full-size parity through oracle windows: the rows of nodes interior to a w x w element window of a large structured
mesh depend only on the elements of that window, so they can be checked against the CPU oracle assembled on a small mesh
carrying the same positions and nodal values (BASELINE-size parity without a BASELINE-size oracle run)."""
import numpy as np

from pyoomph_b200.meshes import CuboidBrickMesh, RectangularQuadMesh, assign_equation_numbers


def check_windows(pb, make_oracle, indptr, indices, jac, res, windows, w=4, tol=1e-12, new_of_old=None, row_begin=0, row_end=None,
                  stats=None):
    """pb: the large problem (2D structured); (indptr, indices, jac, res): its assembled CSR Jacobian and residual;
    windows: list of element offsets (one per dimension).  Returns the largest row-scaled error seen.

    Row-block form (multi-GPU): (indptr, indices, jac, res) are ONE rank's owned rows [row_begin, row_end) of the matrix in the
    partition-aligned numbering `new_of_old` (old equation -> new equation), `indices` holding new global columns; rows of a window
    that another rank owns are skipped (that rank checks them).  `stats`, if given, receives the number of rows compared."""
    mesh, dm, code = pb["mesh"], pb["dofmap"], pb["code"]
    dim = mesh.dim
    L = [2 * n + 1 for n in mesh.N]

    def key(lat):
        k = lat[:, 0].astype(np.int64)
        for d in range(1, dim):
            k = k * L[d] + lat[:, d]
        return k
    lut = np.full(int(np.prod(L)), -1, dtype=np.int64)
    lut[key(mesh.node_lattice)] = np.arange(mesh.n_node)
    worst = 0.0
    for off in windows:
        small = RectangularQuadMesh(w) if dim == 2 else CuboidBrickMesh(w)
        big_of_small = lut[key(small.node_lattice.astype(np.int64) + 2 * np.asarray(off, dtype=np.int64)[None, :])]
        assert (big_of_small >= 0).all()
        small.node_pos[:] = mesh.node_pos[big_of_small]
        sdm = assign_equation_numbers(small, code, {})
        spb = dict(pb)
        spb.update(mesh=small, dofmap=sdm, vals=pb["vals"][:, big_of_small, :].copy(), pos_hist=None)
        op = make_oracle(spb)
        r_ref, mats = op.assemble(flag=1)
        op.close()
        rs, ci, va = mats[0]
        # small equation -> big equation (-1: pinned in the big problem)
        big_eq = np.full(sdm.n_dof, -1, dtype=np.int64)
        m = sdm.node_eqn >= 0
        big_eq[sdm.node_eqn[m]] = dm.node_eqn[big_of_small][m]
        interior = np.all((small.node_lattice > 0) & (small.node_lattice < 2 * w), axis=1)
        for n in np.nonzero(interior)[0]:
            for f in range(sdm.node_eqn.shape[1]):
                se = sdm.node_eqn[n, f]
                if se < 0:
                    continue
                be = big_eq[se]
                if be < 0:
                    continue                      # pinned in the large problem: the row does not exist there
                cols_s, vals_s = ci[rs[se]:rs[se + 1]], va[rs[se]:rs[se + 1]]
                cols_big = big_eq[cols_s]
                if new_of_old is not None:
                    be = int(new_of_old[be])
                    if be < row_begin or (row_end is not None and be >= row_end):
                        continue                  # owned by another rank
                    be -= row_begin
                    cols_big = np.where(cols_big >= 0, new_of_old[np.maximum(cols_big, 0)], -1)
                if stats is not None:
                    stats["rows"] = stats.get("rows", 0) + 1
                ref = {}
                for c, v in zip(cols_big, vals_s):
                    if c >= 0:
                        ref[int(c)] = ref.get(int(c), 0.0) + v
                gc, gv = indices[indptr[be]:indptr[be + 1]], jac[indptr[be]:indptr[be + 1]]
                got = dict(zip(gc.tolist(), gv.tolist()))
                scale = max(max(abs(v) for v in ref.values()), 1e-300)
                for c, v in ref.items():
                    assert c in got, ("missing entry", off, be, c)
                    worst = max(worst, abs(got[c] - v) / scale)
                for c, v in got.items():
                    if c not in ref:
                        worst = max(worst, abs(v) / scale)   # structural-only entries must be zero on the row scale
                rscale = max(np.abs(r_ref).max(), 1e-300)
                worst = max(worst, abs(res[be] - r_ref[se]) / rscale)
    assert worst <= tol, worst
    return worst
