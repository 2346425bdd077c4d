"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: partition, partition-aligned numbering, ghost
columns, exchange lists and the deterministic interface-row exchange.  The local per-rank assembler is the CPU oracle
(allowed here: tests/), so the test checks  sum over ranks == single-process assembly  up to the symmetric permutation."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class _SubMesh:
    def __init__(self, mesh, elements):
        self.dim, self.N, self.node_pos, self.node_lattice = mesh.dim, mesh.N, mesh.node_pos, mesh.node_lattice
        self.elem_nodes = np.ascontiguousarray(mesh.elem_nodes[elements])
        self.n_node, self.n_elem = mesh.n_node, len(elements)


class OracleLocalAssembler:
    """CPU stand-in for B200Assembly behind DistributedAssembly: oracle values on the structural local pattern"""

    def __init__(self, pb, elements, local_dofmap, extra):
        import torch
        from scipy.sparse import csr_matrix
        from oracle import OracleProblem
        from problems import TIME
        from pyoomph_b200.distributed import element_dof_table, structural_pattern
        self.torch = torch
        sub = _SubMesh(pb["mesh"], elements)
        self.op = OracleProblem(pb["code"], sub, local_dofmap, pb["vals"], node_pos_hist=pb["pos_hist"], name=pb["code"].name)
        if pb["unsteady"]:
            self.op.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
        else:
            self.op.set_steady()
        self.n_dof = local_dofmap.n_dof
        ed = element_dof_table(pb["code"], pb["mesh"], local_dofmap, elements)
        self.indptr, self.indices = structural_pattern(ed, self.n_dof, extra)
        self.nnz = self.indices.size
        self._csr = csr_matrix

    def _embed(self, rs, ci, va):
        A = self._csr((va, ci, rs), shape=(self.n_dof, self.n_dof)).tocoo()
        big = self.n_dof
        rows = np.repeat(np.arange(self.n_dof), np.diff(self.indptr))
        keys = rows.astype(np.int64) * big + self.indices
        pos = np.searchsorted(keys, A.row.astype(np.int64) * big + A.col)
        assert np.all(keys[pos] == A.row.astype(np.int64) * big + A.col)
        out = np.zeros(self.nnz)
        np.add.at(out, pos, A.data)
        return self.torch.from_numpy(out)

    def assemble(self, flag=1, **kw):
        r, mats = self.op.assemble(flag=flag)
        self._res = self.torch.from_numpy(r.copy())
        self._jac = self._embed(*mats[0]) if flag >= 1 else None
        self._mass = self._embed(*mats[1]) if flag >= 2 else None

    def evaluate_integral_expressions(self):
        return self.op.evaluate_integral_expressions()

    def residual_tensor(self): return self._res
    def jacobian_tensor(self): return self._jac
    def mass_tensor(self): return self._mass


def _worker(rank, world, port, kind, N, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch.distributed as dist
    from problems import make_problem
    from pyoomph_b200.distributed import DistributedAssembly
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    pb = make_problem(kind, N)
    da = DistributedAssembly.create(pb["code"], pb["mesh"], pb["dofmap"], rank, world,
                                    lambda el, dm, extra: OracleLocalAssembler(pb, el, dm, extra), dist=dist)
    da.assemble(flag=2)
    rb, re, ip, gc, jv, mv, res = da.owned_block(want_mass=True)
    h_res, h_jac, h_mass = da.assemble_host(None, 2)           # the host-facing call of one rank returns the same owned block
    assert np.array_equal(h_res.numpy(), res) and np.array_equal(h_jac.numpy(), jv) and np.array_equal(h_mass.numpy(), mv)
    if kind == "ns_unsteady":
        # the row-block form of the window check bench.py runs on N GPUs (tests/windows.py): every owned row of the windows that
        # straddle the partition interface against the oracle assembled on the window alone
        from problems import make_oracle
        from windows import check_windows
        st = {}
        check_windows(pb, make_oracle, ip, gc, jv, res, [(0, 0), (1, 2), (2, 1), (2, 2)], w=4, tol=1e-12, new_of_old=da.part.new_of_old,
                      row_begin=rb, row_end=re, stats=st)
        assert st["rows"] > 0
    obs = da.evaluate_integral_expressions() if pb["code"].integral_expressions else {}
    np.savez(os.path.join(outdir, "r%d.npz" % rank), rb=rb, re=re, ip=ip, gc=gc, jv=jv, mv=mv, res=res, new_of_old=da.part.new_of_old,
             xbytes=da.exchange_bytes, obs=np.array([obs[k] for k in obs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,N,world", [("ns_unsteady", 6, 2), ("ale", 5, 3), ("heat3d", 3, 2), ("ale_axi_obs", 5, 3)])
def test_row_block_assembly_matches_single_process(kind, N, world):
    import torch.multiprocessing as mp
    from scipy.sparse import csr_matrix
    sys.path.insert(0, HERE)
    from problems import csr_to_sorted, make_oracle, make_problem
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, kind, N, d), nprocs=world, join=True)
        blocks = [np.load(os.path.join(d, "r%d.npz" % r)) for r in range(world)]
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    r_ref, mats = op.assemble(flag=2)
    n = pb["dofmap"].n_dof
    p = blocks[0]["new_of_old"]
    # contiguous row blocks that tile [0, n)
    assert blocks[0]["rb"] == 0 and blocks[-1]["re"] == n
    for a, b in zip(blocks[:-1], blocks[1:]):
        assert a["re"] == b["rb"]
    assert sum(int(b["xbytes"]) for b in blocks) > 0
    for key, ref in (("jv", mats[0]), ("mv", mats[1])):
        A_ref = csr_to_sorted(n, *ref).tocoo()
        A_perm = csr_matrix((A_ref.data, (p[A_ref.row], p[A_ref.col])), shape=(n, n))
        rows = []
        for b in blocks:
            nloc = int(b["re"] - b["rb"])
            rows.append(csr_matrix((b[key], b["gc"], b["ip"]), shape=(nloc, n)))
        from scipy.sparse import vstack
        A = vstack(rows).tocsr()
        assert abs(A - A_perm).max() <= 1e-13 * abs(A_perm).max()
    res = np.concatenate([b["res"] for b in blocks])
    r_perm = np.empty(n); r_perm[p] = r_ref
    assert np.abs(res - r_perm).max() <= 1e-13 * np.abs(r_ref).max()
    if pb["code"].integral_expressions:       # element blocks partition the mesh: rank-ordered sums equal the serial integrals
        ref = np.array(list(op.evaluate_integral_expressions().values()))
        for b in blocks:
            assert np.array_equal(b["obs"], blocks[0]["obs"]) and np.abs(b["obs"] - ref).max() <= 1e-13 * np.abs(ref).max()
