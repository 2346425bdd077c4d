"""The host side of pb2_problem_create without a GPU (pattern-only problems, device = -1): CSR pattern, schedule and the compressed
element -> CSR position maps with their first-touch flags, against an independent numpy construction; every compute entry point
refuses a pattern-only problem.  (The maps feed the scatter of every kernel: a wrong offset or first-touch bit is a wrong matrix.)"""
import numpy as np
import pytest

from problems import make_problem
from pyoomph_b200.assembly import B200Assembly
from pyoomph_b200.distributed import element_dof_table, structural_pattern


def _expected_maps(pb, perm, indptr, indices, extra=None):
    code, mesh, dm = pb["code"], pb["mesh"], pb["dofmap"]
    ed = element_dof_table(code, mesh, dm, np.arange(mesh.n_elem))[perm]          # scheduled order
    ne, nd = ed.shape
    rowstart = np.where(ed >= 0, indptr[np.maximum(ed, 0)], -1)
    off = np.full((ne, nd, nd), -1, dtype=np.int64)
    first = np.zeros((ne, nd, nd), dtype=bool)
    touched = set()
    res = np.full((ne, nd), np.iinfo(np.int32).min, dtype=np.int64)
    rtouched = set()
    for q in range(ne):
        for i in range(nd):
            r = ed[q, i]
            if r < 0:
                continue
            res[q, i] = r if r in rtouched else ~r
            rtouched.add(r)
            cols = indices[indptr[r]:indptr[r + 1]]
            for j in range(nd):
                c = ed[q, j]
                if c < 0:
                    continue
                o = int(np.searchsorted(cols, c))
                assert cols[o] == c
                off[q, i, j] = o
                pos = indptr[r] + o
                if pos not in touched:
                    touched.add(pos)
                    first[q, i, j] = True
    return rowstart, off, first, res


@pytest.mark.parametrize("kind,N,unstructured", [("ns", 5, False), ("ale", 4, True), ("heat3d", 2, False), ("poisson", 7, True), ("ns_axi_swirl", 4, False)])
def test_pattern_and_position_maps_against_numpy(kind, N, unstructured):
    pb = make_problem(kind, N, unstructured=unstructured)
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, device=-1)
    n = pb["dofmap"].n_dof
    ed = element_dof_table(pb["code"], pb["mesh"], pb["dofmap"], np.arange(pb["mesh"].n_elem))
    ip, ix = structural_pattern(ed, n)
    assert np.array_equal(asm.indptr, ip) and np.array_equal(asm.indices, ix)
    assert asm.indptr.dtype == np.int32 and asm.indices.dtype == np.int32
    perm, rowstart, off, res = asm.host_maps()
    assert sorted(perm.tolist()) == list(range(pb["mesh"].n_elem))
    e_rowstart, e_off, e_first, e_res = _expected_maps(pb, perm, ip, ix)
    assert np.array_equal(rowstart, e_rowstart)
    bits = 8 * off.dtype.itemsize
    skip, firstbit = (1 << bits) - 1, 1 << (bits - 1)
    live = e_off >= 0
    assert np.all(off[~live] == skip)
    assert np.array_equal((off[live].astype(np.int64) & (firstbit - 1)), e_off[live])
    assert np.array_equal((off[live].astype(np.int64) & firstbit) != 0, e_first[live])
    assert np.array_equal(res.astype(np.int64), e_res)
    # a pattern-only problem computes nothing
    with pytest.raises(RuntimeError, match="pattern-only"):
        asm.assemble(flag=1)
    with pytest.raises(RuntimeError, match="pattern-only"):
        asm.set_nodal_values(0, pb["vals"][0])
    asm.close()


def test_pattern_with_extra_entries_and_element_subset():
    """the multi-GPU form: a subset of the elements plus extra (row, col) entries other ranks contribute"""
    pb = make_problem("ns", 6)
    n = pb["dofmap"].n_dof
    elements = np.arange(10, 25)
    ed = element_dof_table(pb["code"], pb["mesh"], pb["dofmap"], elements)
    touched = np.unique(ed[ed >= 0])
    rng = np.random.default_rng(0)
    ex_r = rng.choice(touched, 40).astype(np.int32)
    ex_c = rng.integers(0, n, 40).astype(np.int32)
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, device=-1, elements=elements, extra_pattern=(ex_r, ex_c))
    ip, ix = structural_pattern(ed, n, (ex_r, ex_c))
    assert np.array_equal(asm.indptr, ip) and np.array_equal(asm.indices, ix)
    asm.close()


@pytest.mark.parametrize("kind,N", [("robin_if", 6), ("freesurf_if", 5)])
def test_child_problem_maps_point_into_the_parents_pattern(kind, N):
    """pb2_problem_create_child: an interface element class on the bulk class's nodes and equations scatters into the PARENT's CSR
    pattern -- every position map entry of the child names the parent's entry (row of dof i, column of dof j), nothing is a
    first-touch store (the child adds to what the parent wrote), and a parent whose pattern lacks the child's entries is refused."""
    pb = make_problem(kind, N, distortion=0.1)
    bulk = B200Assembly(pb["bulk_code"], pb["bulk_mesh"], pb["dofmap"], name=pb["bulk_code"].name, device=-1)
    child = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=bulk)
    assert child in bulk.children
    assert np.array_equal(child.indptr, bulk.indptr) and np.array_equal(child.indices, bulk.indices)
    perm, rowstart, off, res = child.host_maps()
    ne = pb["mesh"].n_elem
    assert sorted(perm.tolist()) == list(range(ne))
    ed = element_dof_table(pb["code"], pb["mesh"], pb["dofmap"], np.arange(ne))[perm]
    bits = 8 * off.dtype.itemsize
    skip, firstbit = (1 << bits) - 1, 1 << (bits - 1)
    live = ed >= 0
    assert live.any() and (~live).any()
    assert np.array_equal(rowstart[live], bulk.indptr[ed[live]]) and np.all(rowstart[~live] == -1)
    assert np.array_equal(res[live], ed[live])                      # plain row numbers: always added, never the ~row store form
    assert np.all(res[~live] == np.iinfo(np.int32).min)
    for q in range(ne):
        for i in np.nonzero(live[q])[0]:
            for j in range(ed.shape[1]):
                if ed[q, j] < 0:
                    assert off[q, i, j] == skip
                else:
                    assert (int(off[q, i, j]) & firstbit) == 0
                    assert bulk.indices[rowstart[q, i] + int(off[q, i, j])] == ed[q, j]
    # a parent that does not hold the child's entries (bulk elements far from the interface only)
    far = B200Assembly(pb["bulk_code"], pb["bulk_mesh"], pb["dofmap"], name=pb["bulk_code"].name, device=-1, elements=np.arange(3))
    with pytest.raises(RuntimeError, match="lacks"):
        B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=far)
    # a child of a child, or a class with another nodal record, is refused
    with pytest.raises(RuntimeError, match="root problem"):
        B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=child)
    child.close()
    bulk.close()
    far.close()


def test_parent_is_not_released_under_its_children():
    """pb2_problem_free on a parent with live children does nothing (the children alias its buffers) and says so"""
    pb = make_problem("robin_if", 4, distortion=0.1)
    bulk = B200Assembly(pb["bulk_code"], pb["bulk_mesh"], pb["dofmap"], name=pb["bulk_code"].name, device=-1)
    child = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=bulk)
    bulk.lib.pb2_problem_free(bulk.prob)                    # refused
    assert b"child problem" in bulk.lib.pb2_last_error()
    assert np.array_equal(child.indptr, bulk.indptr)         # still usable
    bulk.close()                                             # releases the child first, then itself
    assert child.prob is None and bulk.prob is None
