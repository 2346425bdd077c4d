"""Multi-GPU assembly: element blocks per GPU, contiguous CSR row block per GPU, interface rows exchanged over NCCL.

Mirrors what oomph-lib does for the reference under MPI (/root/reference/src/thirdparty/oomph-lib/include/problem.cc):
``Problem::parallel_sparse_assemble`` (:6543) lets every rank assemble a contiguous element range
(``First_el_for_assembly``, :5354-5358) and ships contributions to rows owned by another rank to their owner
(:6970-7121); after ``Problem::distribute`` equations are numbered rank by rank, so each rank owns a contiguous row
block (``LinearAlgebraDistribution``).  Here:

* elements are split into contiguous blocks in mesh order (strips/slabs on the structured meshes);
* a node belongs to the lowest rank that has an element touching it; its dofs get partition-aligned global numbers
  (rank 0's dofs first, ...), i.e. the matrix is the single-GPU matrix under a symmetric permutation;
* every rank assembles its elements into a LOCAL CSR (owned rows first, then the halo rows of interface nodes owned by a
  neighbour) with exactly the single-GPU kernels -- no collective on that path;
* one exchange step per assembly: halo-row values and halo residual entries are packed with precomputed position lists,
  sent to the owners (``isend``/``irecv``; NCCL on GPUs, gloo in the CPU tests) and added there in increasing source-rank
  order, so the sums are deterministic.

The local assembler is pluggable (anything with ``indptr``, ``indices``, ``n_dof``, ``assemble(flag)`` and tensor views of
its outputs), which is how the CPU tests drive the host-side logic without a GPU.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .meshes import DofMap


@dataclasses.dataclass
class Partition:
    rank: int
    world: int
    elements: np.ndarray           # element ids of this rank (contiguous block in mesh order)
    new_of_old: np.ndarray         # [n_dof] partition-aligned global equation number of every original equation
    row_offsets: np.ndarray        # [world+1] row block boundaries in the new numbering
    local_dofmap: DofMap           # local numbering: owned rows first, then halo rows; -1 elsewhere
    l2g: np.ndarray                # [n_local] local row -> new global equation
    n_owned: int

    @property
    def row_begin(self) -> int:
        return int(self.row_offsets[self.rank])

    @property
    def row_end(self) -> int:
        return int(self.row_offsets[self.rank + 1])


def _renumber(eq: np.ndarray, new_of_old: np.ndarray) -> np.ndarray:
    out = np.full(eq.shape, -1, dtype=np.int64)
    m = eq >= 0
    out[m] = new_of_old[eq[m]]
    return out


def element_blocks(n_elem: int, world: int) -> List[Tuple[int, int]]:
    """Static element-range split, First_el_for_assembly style (problem.cc:5354-5358)."""
    return [(n_elem * r // world, n_elem * (r + 1) // world) for r in range(world)]


def build_partition(mesh, dofmap: DofMap, rank: int, world: int) -> Partition:
    ne, nn = mesh.elem_nodes.shape
    blocks = element_blocks(ne, world)
    # node owner = lowest rank with an element touching it
    owner = np.full(mesh.n_node, world, dtype=np.int32)
    for r in range(world - 1, -1, -1):
        lo, hi = blocks[r]
        owner[np.unique(mesh.elem_nodes[lo:hi])] = r
    # dof owner = node owner; new numbering: by (owner, old equation)
    eqs = [dofmap.node_eqn]
    if dofmap.pos_eqn is not None:
        eqs.insert(0, dofmap.pos_eqn)
    alleq = np.concatenate(eqs, axis=1)                     # [n_node, ndofs per node]
    dof_owner = np.full(dofmap.n_dof, -1, dtype=np.int32)
    valid = alleq >= 0
    dof_owner[alleq[valid]] = np.broadcast_to(owner[:, None], alleq.shape)[valid]
    order = np.lexsort((np.arange(dofmap.n_dof), dof_owner))   # sort by owner, then old equation
    new_of_old = np.empty(dofmap.n_dof, dtype=np.int64)
    new_of_old[order] = np.arange(dofmap.n_dof)
    counts = np.bincount(dof_owner, minlength=world)
    row_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    # local numbering of this rank
    lo, hi = blocks[rank]
    my_nodes = np.unique(mesh.elem_nodes[lo:hi])
    touched_old = alleq[my_nodes]
    touched_old = np.unique(touched_old[touched_old >= 0])
    touched_new = new_of_old[touched_old]
    own_mask = (touched_new >= row_offsets[rank]) & (touched_new < row_offsets[rank + 1])
    owned_new = np.sort(touched_new[own_mask])
    n_owned = int(row_offsets[rank + 1] - row_offsets[rank])
    if owned_new.size != n_owned:
        raise RuntimeError("ownership is inconsistent: a dof owned by rank %d is not touched by its elements" % rank)
    halo_new = np.sort(touched_new[~own_mask])
    l2g = np.concatenate([owned_new, halo_new]).astype(np.int64)
    loc = np.full(dofmap.n_dof, -1, dtype=np.int64)
    loc[l2g] = np.arange(l2g.size)                             # indexed by NEW global number
    def localise(eq):
        out = np.full(eq.shape, -1, dtype=np.int32)
        m = eq >= 0
        out[m] = loc[new_of_old[eq[m]]]
        return out
    node_eqn_l = localise(dofmap.node_eqn)
    pos_eqn_l = None if dofmap.pos_eqn is None else localise(dofmap.pos_eqn)
    # dofs of nodes this rank never touches stay -1 (they are never referenced by its elements)
    return Partition(rank, world, np.arange(lo, hi), new_of_old, row_offsets,
                     DofMap(node_eqn_l, pos_eqn_l, int(l2g.size)), l2g, n_owned)


def element_dof_table(code, mesh, dofmap: DofMap, elements: np.ndarray) -> np.ndarray:
    """[n_elem, ndof_el] equation of every local dof in the class's dof layout (pyoomph_b200.codegen.dof_layout), -1 pinned"""
    en = mesh.elem_nodes[elements]
    cols = []
    for (field, lnode) in code.dof_layout():
        fld = code.fields[field]
        node = en[:, code.space_nodes(fld.space)[lnode]]
        cols.append(dofmap.pos_eqn[node, fld.index] if fld.space == "Pos" else dofmap.node_eqn[node, fld.index])
    return np.stack(cols, axis=1)


def structural_pattern(edofs: np.ndarray, n_rows: int, extra: Optional[Tuple[np.ndarray, np.ndarray]] = None):
    """CSR pattern (ascending columns) of the element-dof-complete matrix, plus extra (row, col) pairs"""
    nd = edofs.shape[1]
    rows = np.repeat(edofs, nd, axis=1).ravel()
    cols = np.tile(edofs, (1, nd)).ravel()
    m = (rows >= 0) & (cols >= 0)
    rows, cols = rows[m].astype(np.int64), cols[m].astype(np.int64)
    if extra is not None and len(extra[0]):
        rows = np.concatenate([rows, np.asarray(extra[0], dtype=np.int64)])
        cols = np.concatenate([cols, np.asarray(extra[1], dtype=np.int64)])
    key = np.unique(rows * n_rows + cols)
    r, c = key // n_rows, key % n_rows
    indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n_rows))]).astype(np.int32)
    return indptr, c.astype(np.int32)


class DistributedAssembly:
    """Owns one rank's local assembler and the exchange lists; `dist` is torch.distributed (already initialised)."""

    @classmethod
    def create(cls, code, mesh, dofmap: DofMap, rank: int, world: int, make_local: Callable, dist=None, device="cpu"):
        """make_local(elements, local_dofmap, extra_pattern) -> local assembler.  Collective over all ranks."""
        part = build_partition(mesh, dofmap, rank, world)
        # halo rows: the (row, col) pairs my elements produce in rows owned by somebody else (structural)
        gmap = DofMap(_renumber(dofmap.node_eqn, part.new_of_old), None if dofmap.pos_eqn is None else _renumber(dofmap.pos_eqn, part.new_of_old), dofmap.n_dof)
        ed = element_dof_table(code, mesh, gmap, part.elements).astype(np.int64)          # new global numbers
        halo = (ed >= 0) & ((ed < part.row_begin) | (ed >= part.row_end))
        sel = np.nonzero(halo.any(axis=1))[0]
        nd = ed.shape[1]
        rr = np.repeat(ed[sel], nd, axis=1).ravel()
        cc = np.tile(ed[sel], (1, nd)).ravel()
        hm = np.repeat(halo[sel], nd, axis=1).ravel() & (cc >= 0)
        key = np.unique(rr[hm] * dofmap.n_dof + cc[hm])
        hr, hc = key // dofmap.n_dof, key % dofmap.n_dof
        owner = np.searchsorted(part.row_offsets, hr, side="right") - 1
        outbox = [None] * world
        for q in np.unique(owner):
            outbox[int(q)] = (hr[owner == q], hc[owner == q])
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, outbox)
        else:
            gathered = [outbox]
        # ghost columns: dofs neighbours couple to my rows that none of my elements touches
        inc_r = [gathered[src][rank][0] for src in range(world) if src != rank and gathered[src][rank] is not None]
        inc_c = [gathered[src][rank][1] for src in range(world) if src != rank and gathered[src][rank] is not None]
        inc_r = np.concatenate(inc_r) if inc_r else np.zeros(0, np.int64)
        inc_c = np.concatenate(inc_c) if inc_c else np.zeros(0, np.int64)
        ghosts = np.setdiff1d(np.unique(inc_c), part.l2g)
        l2g = np.concatenate([part.l2g, ghosts]).astype(np.int64)
        loc = np.full(dofmap.n_dof, -1, dtype=np.int64)
        loc[l2g] = np.arange(l2g.size)
        extra = (loc[inc_r].astype(np.int32), loc[inc_c].astype(np.int32))
        part = dataclasses.replace(part, l2g=l2g, local_dofmap=DofMap(part.local_dofmap.node_eqn, part.local_dofmap.pos_eqn, int(l2g.size)))
        local = make_local(part.elements, part.local_dofmap, extra)
        return cls(part, local, dist=dist, device=device, sent_pairs=outbox)

    def __init__(self, part: Partition, local, dist=None, device="cpu", sent_pairs=None):
        import torch
        self.torch = torch
        self.part, self.local, self.dist, self.device = part, local, dist, device
        self.rank, self.world = part.rank, part.world
        self.indptr = np.asarray(local.indptr)
        self.indices = np.asarray(local.indices)
        self.n_owned = part.n_owned
        self.gcols = part.l2g[self.indices]                   # global (new) column of every local entry
        self._setup_exchange()

    # ---- setup: which of my entries go where ------------------------------------------------------
    def _setup_exchange(self):
        part, torch = self.part, self.torch
        n_loc = part.l2g.size
        # halo rows = local rows after the owned block that have entries (ghost rows are empty)
        cand = np.arange(self.n_owned, n_loc)
        halo_rows = cand[(self.indptr[cand + 1] - self.indptr[cand]) > 0]
        halo_g = part.l2g[halo_rows]
        halo_owner = np.searchsorted(part.row_offsets, halo_g, side="right") - 1
        send: Dict[int, dict] = {}
        for q in np.unique(halo_owner):
            rows = halo_rows[halo_owner == q]
            cnt = (self.indptr[rows + 1] - self.indptr[rows]).astype(np.int64)
            pos = np.repeat(self.indptr[rows].astype(np.int64) - (np.cumsum(cnt) - cnt), cnt) + np.arange(int(cnt.sum()), dtype=np.int64)
            rr = np.repeat(part.l2g[rows], self.indptr[rows + 1] - self.indptr[rows])
            send[int(q)] = dict(rows_local=rows, rows_global=part.l2g[rows], pos=pos.astype(np.int64), ent_row=rr, ent_col=self.gcols[pos])
        if self.world > 1:
            outbox = [None] * self.world
            for q, s_ in send.items():
                outbox[q] = (s_["rows_global"], s_["ent_row"], s_["ent_col"])
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, outbox)
        else:
            gathered = [[None]]
        self.send = send
        self.recv: Dict[int, dict] = {}
        big = int(part.row_offsets[-1]) + 1
        for src in range(self.world):
            if src == self.rank or gathered[src] is None or gathered[src][self.rank] is None:
                continue
            rows_g, ent_row, ent_col = gathered[src][self.rank]
            rows_l = rows_g - part.row_begin
            if rows_l.size and (rows_l.min() < 0 or rows_l.max() >= self.n_owned):
                raise RuntimeError("received a row this rank does not own")
            # (row, global column) -> position, searched only among the entries of the rows that receive something
            # (sorting the whole owned block, 2e8 entries per GPU at config 2, took minutes)
            rr = np.unique(ent_row - part.row_begin).astype(np.int64)
            cnt = (self.indptr[rr + 1] - self.indptr[rr]).astype(np.int64)
            start = np.repeat(self.indptr[rr].astype(np.int64) - (np.cumsum(cnt) - cnt), cnt)
            pos_all = start + np.arange(int(cnt.sum()), dtype=np.int64)
            keys = np.repeat(rr, cnt) * big + self.gcols[pos_all]
            srt = np.argsort(keys)
            skeys = keys[srt]
            want = (ent_row - part.row_begin).astype(np.int64) * big + ent_col
            idx = np.searchsorted(skeys, want)
            if idx.size and (idx.max() >= srt.size or np.any(skeys[np.minimum(idx, srt.size - 1)] != want)):
                raise RuntimeError("interface entry missing in the owner's pattern")
            self.recv[src] = dict(rows_local=rows_l.astype(np.int64), pos=pos_all[srt[idx]].astype(np.int64))
        dev = self.device
        for d in list(self.send.values()) + list(self.recv.values()):
            d["pos_t"] = torch.as_tensor(d["pos"], device=dev)
            d["rows_t"] = torch.as_tensor(d["rows_local"].astype(np.int64), device=dev)
        self.exchange_bytes = sum(8 * (d["pos"].size + d["rows_local"].size) for d in self.send.values())

    # ---- per assembly ----------------------------------------------------------------------------
    def assemble(self, flag: int = 1, **kw):
        """local kernels (no collective), then ONE exchange step for the interface rows and the residual"""
        self.local.assemble(flag=flag, **kw)
        self.exchange(flag)

    def exchange(self, flag: int):
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return
        tensors = [self.local.residual_tensor()]
        if flag >= 1:
            tensors.append(self.local.jacobian_tensor())
        if flag >= 2:
            tensors.append(self.local.mass_tensor())
        native = hasattr(self.local, "pack_rows")       # GPU: one pack kernel per neighbour, one grouped NCCL send/recv, one add kernel
        ops, recvbufs = [], {}
        for q, s in sorted(self.send.items()):
            n = s["rows_local"].size + flag * s["pos"].size
            if native:
                buf = self._buffer(("s", q, flag), n)
                self.local.pack_rows(s["rows_t"], s["pos_t"], flag, buf)
            else:
                buf = torch.cat([tensors[0][s["rows_t"]]] + [t[s["pos_t"]] for t in tensors[1:]]).contiguous()
            ops.append(dist.P2POp(dist.isend, buf, q))
        for src, r in sorted(self.recv.items()):
            n = r["rows_local"].size + (len(tensors) - 1) * r["pos"].size
            recvbufs[src] = self._buffer(("r", src, flag), n)
            ops.append(dist.P2POp(dist.irecv, recvbufs[src], src))
        if ops:
            for rq in dist.batch_isend_irecv(ops):      # one grouped send/recv (ncclGroupStart/End under NCCL)
                rq.wait()
        # deterministic: add sources in increasing rank order
        for src in sorted(recvbufs):
            r, buf = self.recv[src], recvbufs[src]
            nr, npos = r["rows_local"].size, r["pos"].size
            if native:
                self.local.unpack_add(r["rows_t"], r["pos_t"], flag, buf)
                continue
            tensors[0].index_add_(0, r["rows_t"], buf[:nr])
            for k, t in enumerate(tensors[1:]):
                t.index_add_(0, r["pos_t"], buf[nr + k * npos: nr + (k + 1) * npos])

    def _buffer(self, key, n: int):
        """persistent exchange buffers (one per neighbour, direction and flag)"""
        bufs = self.__dict__.setdefault("_bufs", {})
        if key not in bufs or bufs[key].numel() != n:
            bufs[key] = self.torch.empty(n, dtype=self.torch.float64, device=self.device)
        return bufs[key]

    def assemble_host(self, local_dofs: Optional[np.ndarray], flag: int = 1, out=None, **kw):
        """The reference-facing call of one rank: host dof values of its LOCAL rows in (owned | halo | ghost order, `part.l2g`),
        host residual / CSR values of its OWNED row block out.  Host<->device copies, the local kernels and the interface
        exchange are all inside the call; `out` = (residual[n_owned], jac[nnz_owned] or None, mass[nnz_owned] or None) torch
        CPU tensors (pinned for full PCIe speed), allocated when omitted.  Every rank moves only its own block, so N GPUs use
        N host links."""
        torch = self.torch
        if local_dofs is not None:
            self.local.set_dofs(local_dofs)
        self.assemble(flag=flag, **kw)
        nnz = int(self.indptr[self.n_owned])
        if out is None:
            out = (torch.empty(self.n_owned, dtype=torch.float64), torch.empty(nnz, dtype=torch.float64) if flag >= 1 else None,
                   torch.empty(nnz, dtype=torch.float64) if flag >= 2 else None)
        out[0].copy_(self.local.residual_tensor()[:self.n_owned], non_blocking=True)
        if flag >= 1:
            out[1].copy_(self.local.jacobian_tensor()[:nnz], non_blocking=True)
        if flag >= 2:
            out[2].copy_(self.local.mass_tensor()[:nnz], non_blocking=True)
        if str(self.device).startswith("cuda"):
            torch.cuda.synchronize(self.device)
        return out

    def evaluate_integral_expressions(self) -> Dict[str, float]:
        """integral expressions over the whole mesh: every rank integrates over its own elements (they partition the mesh), the
        per-rank values are gathered and added in rank order (deterministic, identical on all ranks)"""
        local = self.local.evaluate_integral_expressions()
        names = list(local)
        mine = self.torch.tensor([local[n] for n in names], dtype=self.torch.float64, device=self.device)
        if self.world == 1:
            return dict(local)
        parts = [self.torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine)
        total = parts[0].clone()
        for q in range(1, self.world):
            total += parts[q]
        return {n: float(v) for n, v in zip(names, total.cpu().tolist())}

    # ---- results ---------------------------------------------------------------------------------
    def owned_block(self, want_mass: bool = False):
        """(row_begin, row_end, indptr, global column indices, values[, mass values], residual) of the owned row block"""
        nnz = int(self.indptr[self.n_owned])
        jac = self.local.jacobian_tensor()[:nnz].cpu().numpy()
        res = self.local.residual_tensor()[:self.n_owned].cpu().numpy()
        out = [self.part.row_begin, self.part.row_end, self.indptr[:self.n_owned + 1].copy(), self.gcols[:nnz].copy(), jac]
        if want_mass:
            out.append(self.local.mass_tensor()[:nnz].cpu().numpy())
        out.append(res)
        return tuple(out)


class _DevicePointerArray:
    """zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class GPULocalAssembler:
    """adapter: B200Assembly -> the local-assembler protocol of DistributedAssembly"""

    def __init__(self, asm, device_index: int):
        import torch
        self.asm, self.torch, self.dev = asm, torch, torch.device("cuda", device_index)
        self.indptr, self.indices, self.n_dof = asm.indptr, asm.indices, asm.n_dof
        self._views = {}

    def assemble(self, flag=1, **kw):
        self.asm.assemble(flag=flag, **kw)
        self._views = {}

    def _view(self, k, n):
        if k not in self._views:
            ptrs = self.asm.device_outputs()
            self._views[k] = self.torch.as_tensor(_DevicePointerArray(ptrs[k], n), device=self.dev)
        return self._views[k]

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def evaluate_integral_expressions(self):
        return self.asm.evaluate_integral_expressions()

    def set_dofs(self, dofs):
        """host values of the LOCAL dofs (owned | halo | ghost) -> nodal storage on the device (pb2_problem_set_dofs)"""
        self.asm.set_dofs(dofs.numpy() if hasattr(dofs, "numpy") else dofs)

    def pack_rows(self, rows_t, pos_t, flag: int, buf):
        self.asm.pack_rows(rows_t.data_ptr(), rows_t.numel(), pos_t.data_ptr(), pos_t.numel(), flag, buf.data_ptr(), self._stream())

    def unpack_add(self, rows_t, pos_t, flag: int, buf):
        self.asm.unpack_add(rows_t.data_ptr(), rows_t.numel(), pos_t.data_ptr(), pos_t.numel(), flag, buf.data_ptr(), self._stream())

    def residual_tensor(self):
        return self._view(0, self.asm.n_dof)

    def jacobian_tensor(self):
        return self._view(1, self.asm.nnz)

    def mass_tensor(self):
        return self._view(2, self.asm.nnz)
