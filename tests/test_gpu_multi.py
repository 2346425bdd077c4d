"""2-GPU parity test (NCCL): row-block assembly over two B200s == single-GPU assembly under the symmetric permutation.
Skipped when fewer than two devices are visible."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ndev():
    try:
        import ctypes
        from pyoomph_b200.assembly import load_library
        n = ctypes.c_int(0)
        return n.value if load_library().pb2_device_count(ctypes.byref(n)) else n.value
    except Exception:
        return 0


def _worker(rank, world, port, kind, N, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    from problems import TIME, make_problem
    from pyoomph_b200.assembly import B200Assembly
    from pyoomph_b200.distributed import DistributedAssembly, GPULocalAssembler
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pb = make_problem(kind, N)

    def make_local(el, dm, extra):
        a = B200Assembly(pb["code"], pb["mesh"], dm, name=pb["code"].name, device=rank, elements=el, extra_pattern=extra)
        for t in range(pb["vals"].shape[0]):
            a.set_nodal_values(t, pb["vals"][t])
        if pb["pos_hist"] is not None:
            for t in range(pb["pos_hist"].shape[0]):
                a.set_nodal_positions(t, pb["pos_hist"][t])
        if pb["unsteady"]:
            a.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
        return GPULocalAssembler(a, rank)
    da = DistributedAssembly.create(pb["code"], pb["mesh"], pb["dofmap"], rank, world, make_local, dist=dist, device="cuda:%d" % rank)
    out = []
    for rep in range(2):                      # twice: the entries only neighbours write must restart from zero
        da.assemble(flag=2)
        out.append(da.owned_block(want_mass=True))
    for a, b in zip(out[0][4:], out[1][4:]):
        assert np.array_equal(a, b)
    rb, re, ip, gc, jv, mv, res = out[1]
    np.savez(os.path.join(outdir, "r%d.npz" % rank), rb=rb, re=re, ip=ip, gc=gc, jv=jv, mv=mv, res=res, new_of_old=da.part.new_of_old)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N", [("ns_unsteady", 10), ("ale", 8)])
def test_two_gpu_row_blocks_match_oracle(kind, N):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from scipy.sparse import csr_matrix, vstack
    sys.path.insert(0, HERE)
    from problems import csr_to_sorted, make_oracle, make_problem
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, kind, N, d), nprocs=world, join=True)
        blocks = [np.load(os.path.join(d, "r%d.npz" % r)) for r in range(world)]
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    r_ref, mats = op.assemble(flag=2)
    n = pb["dofmap"].n_dof
    p = blocks[0]["new_of_old"]
    for key, ref in (("jv", mats[0]), ("mv", mats[1])):
        A_ref = csr_to_sorted(n, *ref).tocoo()
        A_perm = csr_matrix((A_ref.data, (p[A_ref.row], p[A_ref.col])), shape=(n, n))
        A = vstack([csr_matrix((b[key], b["gc"], b["ip"]), shape=(int(b["re"] - b["rb"]), n)) for b in blocks]).tocsr()
        rowmax = np.maximum(abs(A_perm).max(axis=1).toarray().ravel(), 1e-300)
        err = abs(A - A_perm).max(axis=1).toarray().ravel() / rowmax
        assert err.max() <= 1e-12
    res = np.concatenate([b["res"] for b in blocks])
    r_perm = np.empty(n); r_perm[p] = r_ref
    assert np.abs(res - r_perm).max() <= 1e-12 * np.abs(r_ref).max()
