"""Mesh partitioning for the multi-GPU engine: one process per GPU, each holding the
*local mesh* of its rank.

The reference gets its domain decomposition implicitly from Firedrake's DMPlex partition, PyOP2
halos and PETSc's parallel Mat/Vec (the only MPI symbols in the reference are `COMM_WORLD` at
`hdg_imex.py:110` and `conforming_implicit.py:86`).  Here the decomposition is explicit:

* every cell has an owner rank (`cell_rank`); a facet belongs to the owner of its first adjacent
  cell; a P1 vertex to the smallest rank among the cells that touch it; a coarse P1 vertex to the
  owner of the fine vertex it is injected from;
* the local mesh of rank r = owned cells + one *vertex-adjacent* ghost layer, with all facets and
  vertices of those cells.  Owned entities are numbered first (in global order), ghosts follow,
  grouped by owner rank (then global order), so that every receive lands in one contiguous block;
* a :class:`HaloPlan` per entity kind lists, per peer, which owned entities to pack and send and
  where the received block goes.  Send lists are derived from the *peer's* local set, which every
  rank can compute because the partition is a deterministic function of (mesh, cell_rank).

Engine kernels run over all local entities, reductions over owned entities only, and a vector's
ghost entries are refreshed (`hdg_halo_exchange_dev`) before a kernel that reads neighbours.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .mesh import Mesh

__all__ = ["strip_partition", "block_partition", "HaloPlan", "LocalMesh", "partition_mesh", "build_plan",
           "LocalHierarchy", "partition_hierarchy", "exchange_host", "cg_plan"]


# ------------------------------------------------------------------------------------------------
# cell -> rank maps
# ------------------------------------------------------------------------------------------------
def strip_partition(mesh: Mesh, nranks: int) -> np.ndarray:
    """contiguous blocks of cells in mesh order (rows of squares for the structured generators, so
    the cut is nx horizontal edges per interface: SURVEY.md §8e); balanced to +-1 row of squares"""
    nc = mesh.nc
    meta = mesh.meta
    if "nx" in meta and "ny" in meta and nc == 2 * meta["nx"] * meta["ny"]:
        ny, per_row = meta["ny"], 2 * meta["nx"]
        rows = np.arange(ny)
        row_rank = np.minimum(rows * nranks // ny, nranks - 1)
        return np.repeat(row_rank, per_row).astype(np.int32)
    return np.minimum(np.arange(nc, dtype=np.int64) * nranks // nc, nranks - 1).astype(np.int32)


def block_partition(mesh: Mesh, px: int, py: int) -> np.ndarray:
    """px x py blocks by cell centroid (any mesh); rank = jy * px + jx"""
    c = mesh.cell_xy.mean(axis=1)
    lo, hi = c.min(axis=0), c.max(axis=0)
    ij = np.minimum(((c - lo) / (hi - lo + 1e-300) * np.array([px, py])).astype(np.int64), np.array([px - 1, py - 1]))
    return (ij[:, 1] * px + ij[:, 0]).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# halo plans
# ------------------------------------------------------------------------------------------------
@dataclass
class HaloPlan:
    """exchange plan of one entity kind on one rank (all index arrays int32, local numbering)"""
    n_owned: int
    n_local: int
    peers: np.ndarray      # [npeers] ranks this rank talks to (sorted)
    send_ptr: np.ndarray   # [npeers+1] CSR pointer into send_idx
    send_idx: np.ndarray   # owned local ids to pack for each peer (in the peer's ghost order)
    recv_off: np.ndarray   # [npeers] first local id of the ghost block received from the peer
    recv_cnt: np.ndarray   # [npeers]
    local_gid: np.ndarray  # [n_local] global id of every local entity (int64)

    @property
    def owned_gid(self):
        return self.local_gid[: self.n_owned]


def _order_local(owner: np.ndarray, local_gids: np.ndarray, rank: int):
    """owned first (global order), then ghosts sorted by (owner, gid)"""
    local_gids = np.unique(local_gids)
    own = owner[local_gids] == rank
    owned = local_gids[own]
    ghost = local_gids[~own]
    ghost = ghost[np.lexsort((ghost, owner[ghost]))]
    return owned, ghost


def build_plan(owner: np.ndarray, local_sets: list, rank: int) -> HaloPlan:
    """owner[g]: owning rank of global entity g; local_sets[q]: global ids held by rank q"""
    nranks = len(local_sets)
    owned, ghost = _order_local(owner, local_sets[rank], rank)
    local_gid = np.concatenate([owned, ghost]).astype(np.int64)
    n_owned = owned.size
    g2l = {}
    # vectorised global->local for owned ids
    srt = owned  # already sorted
    gowner = owner[ghost]
    recv = {}
    off = n_owned
    for q in np.unique(gowner):
        cnt = int((gowner == q).sum())
        recv[int(q)] = (off, cnt)
        off += cnt
    send = {}
    for q in range(nranks):
        if q == rank:
            continue
        lq = np.unique(local_sets[q])
        need = lq[owner[lq] == rank]  # sorted by gid == the order of q's ghost block from this rank
        if need.size:
            loc = np.searchsorted(srt, need)
            assert np.all(srt[loc] == need), "peer needs an entity this rank does not hold"
            send[q] = loc.astype(np.int32)
    peers = sorted(set(recv) | set(send))
    send_ptr = [0]
    send_idx = []
    recv_off, recv_cnt = [], []
    for q in peers:
        s = send.get(q, np.zeros(0, dtype=np.int32))
        send_idx.append(s)
        send_ptr.append(send_ptr[-1] + s.size)
        o, c = recv.get(q, (n_owned, 0))
        recv_off.append(o)
        recv_cnt.append(c)
    del g2l
    return HaloPlan(
        n_owned=int(n_owned), n_local=int(local_gid.size), peers=np.asarray(peers, dtype=np.int32),
        send_ptr=np.asarray(send_ptr, dtype=np.int32),
        send_idx=(np.concatenate(send_idx) if send_idx else np.zeros(0)).astype(np.int32),
        recv_off=np.asarray(recv_off, dtype=np.int32), recv_cnt=np.asarray(recv_cnt, dtype=np.int32),
        local_gid=local_gid)


def _global_to_local(local_gid: np.ndarray, n_global: int) -> np.ndarray:
    g2l = np.full(n_global, -1, dtype=np.int64)
    g2l[local_gid] = np.arange(local_gid.size)
    return g2l


# ------------------------------------------------------------------------------------------------
# local meshes
# ------------------------------------------------------------------------------------------------
@dataclass
class LocalMesh:
    """the part of a partitioned mesh one rank holds; `mesh` is a complete :class:`Mesh` of the
    local cells (owned first) that the engine treats like any other mesh"""
    mesh: Mesh
    rank: int
    nranks: int
    cells: HaloPlan
    facets: HaloPlan
    verts: HaloPlan
    global_nc: int
    global_nf: int
    global_nv: int
    global_volume: float
    cell_rank: np.ndarray = field(repr=False, default=None)

    @property
    def nc_owned(self):
        return self.cells.n_owned

    @property
    def nf_owned(self):
        return self.facets.n_owned


def _entity_owners(mesh: Mesh, cell_rank: np.ndarray):
    facet_owner = cell_rank[mesh.facet_cell[:, 0]].astype(np.int32)
    vert_owner = np.full(mesh.nv, np.iinfo(np.int32).max, dtype=np.int32)
    np.minimum.at(vert_owner, mesh.cell_vert.ravel(), np.repeat(cell_rank, 3))
    return facet_owner, vert_owner


def _local_cells(mesh: Mesh, cell_rank: np.ndarray, q: int) -> np.ndarray:
    """owned cells of rank q plus every cell sharing a vertex with one of them"""
    vmask = np.zeros(mesh.nv, dtype=bool)
    vmask[mesh.cell_vert[cell_rank == q].ravel()] = True
    return np.nonzero(vmask[mesh.cell_vert].any(axis=1))[0]


def partition_mesh(mesh: Mesh, cell_rank: np.ndarray, rank: int, nranks: int | None = None) -> LocalMesh:
    cell_rank = np.ascontiguousarray(cell_rank, dtype=np.int32)
    nranks = int(cell_rank.max()) + 1 if nranks is None else nranks
    assert cell_rank.shape == (mesh.nc,) and 0 <= rank < nranks
    facet_owner, vert_owner = _entity_owners(mesh, cell_rank)
    lc, lf, lv = [], [], []
    for q in range(nranks):
        c = _local_cells(mesh, cell_rank, q)
        lc.append(c)
        lf.append(np.unique(mesh.cell_facet[c].ravel()))
        lv.append(np.unique(mesh.cell_vert[c].ravel()))
    cells = build_plan(cell_rank, lc, rank)
    facets = build_plan(facet_owner, lf, rank)
    verts = build_plan(vert_owner, lv, rank)
    c2l = _global_to_local(cells.local_gid, mesh.nc)
    f2l = _global_to_local(facets.local_gid, mesh.nf)
    v2l = _global_to_local(verts.local_gid, mesh.nv)
    cg, fg = cells.local_gid, facets.local_gid
    # facet -> cell: drop non-local neighbours; a facet whose *first* cell is not local keeps its second
    fc = mesh.facet_cell[fg].astype(np.int64)
    fl = mesh.facet_local[fg].astype(np.int64)
    fc_loc = np.where(fc >= 0, c2l[np.maximum(fc, 0)], -1)
    fl_loc = np.where(fc_loc >= 0, fl, -1)
    swap = fc_loc[:, 0] < 0
    fc_loc[swap] = fc_loc[swap][:, ::-1]
    fl_loc[swap] = fl_loc[swap][:, ::-1]
    assert np.all(fc_loc[:, 0] >= 0)
    m = Mesh(
        cell_xy=np.ascontiguousarray(mesh.cell_xy[cg]),
        cell_vert=np.ascontiguousarray(v2l[mesh.cell_vert[cg]], dtype=np.int32),
        cell_facet=np.ascontiguousarray(f2l[mesh.cell_facet[cg]], dtype=np.int32),
        cell_flip=np.ascontiguousarray(mesh.cell_flip[cg], dtype=np.int32),
        facet_cell=np.ascontiguousarray(fc_loc, dtype=np.int32),
        facet_local=np.ascontiguousarray(fl_loc, dtype=np.int32),
        facet_vert=np.ascontiguousarray(v2l[mesh.facet_vert[fg]], dtype=np.int32),
        nv=int(verts.n_local),
        name=f"{mesh.name}[rank {rank}/{nranks}]",
    )
    m.meta.update(partitioned=True, periodic=mesh.meta.get("periodic", False))
    return LocalMesh(mesh=m, rank=rank, nranks=nranks, cells=cells, facets=facets, verts=verts, global_nc=mesh.nc,
                     global_nf=mesh.nf, global_nv=mesh.nv, global_volume=mesh.volume, cell_rank=cell_rank)


# ------------------------------------------------------------------------------------------------
# continuous Lagrange dofs (velocity projection of the passive-tracer path, cgspace.py)
# ------------------------------------------------------------------------------------------------
def cg_plan(mesh: Mesh, lm: LocalMesh, degree: int):
    """halo plan of the CG_degree dofs of rank `lm.rank` and the renumbering that makes it usable.

    A CG dof belongs to the owner of the entity it sits on (vertex, facet, cell).  ``cgspace`` numbers the
    dofs of the *local* mesh as [vertices | facet interiors | cell interiors] in local entity order, which
    interleaves owned and ghost dofs; the engine's exchange wants owned dofs first and the ghosts of every
    peer in one block.  Returns ``(plan, perm)`` with ``perm[l]`` = position of local cgspace dof ``l`` in
    that order (pass it to ``cgspace.build_cg_space(..., perm=perm)``); ``plan.local_gid`` are the ids in the
    global numbering of ``cgspace`` on the complete mesh, so both ends of an exchange agree on the order.
    """
    d = int(degree)
    ne, nint = d - 1, (d - 1) * (d - 2) // 2
    NV, NF = mesh.nv, mesh.nf
    cell_rank = lm.cell_rank
    facet_owner, vert_owner = _entity_owners(mesh, cell_rank)
    owner = np.concatenate([vert_owner, np.repeat(facet_owner, ne), np.repeat(cell_rank, nint)]).astype(np.int32)

    def gids(v, f, c):  # global dof ids of global entity id arrays
        parts = [np.asarray(v, dtype=np.int64)]
        if ne:
            parts.append((NV + np.asarray(f, dtype=np.int64)[:, None] * ne + np.arange(ne)).ravel())
        if nint:
            parts.append((NV + NF * ne + np.asarray(c, dtype=np.int64)[:, None] * nint + np.arange(nint)).ravel())
        return np.concatenate(parts)

    local_sets = []
    for q in range(lm.nranks):
        c = _local_cells(mesh, cell_rank, q)
        local_sets.append(gids(np.unique(mesh.cell_vert[c].ravel()), np.unique(mesh.cell_facet[c].ravel()), c))
    plan = build_plan(owner, local_sets, lm.rank)
    # global id of every dof in the local cgspace numbering (local entity order)
    gid_local = gids(lm.verts.local_gid, lm.facets.local_gid, lm.cells.local_gid)
    order = np.argsort(plan.local_gid)
    pos = np.searchsorted(plan.local_gid[order], gid_local)
    assert np.array_equal(plan.local_gid[order][pos], gid_local), "local CG dofs and the plan disagree"
    perm = order[pos].astype(np.int64)
    return plan, perm


# ------------------------------------------------------------------------------------------------
# multigrid hierarchy
# ------------------------------------------------------------------------------------------------
@dataclass
class LocalHierarchy:
    """rank-local view of a :class:`multigrid.Hierarchy`.

    Levels l < repl are row-distributed (A[l], P[l], R[l] hold the owned rows, columns in the local
    numbering of `plans[l]`); levels l >= repl are replicated on every rank in global numbering.  At
    the interface the restriction produces the owned rows of level `repl`, which are all-gathered
    (`gather_gid[q]` = global ids of rank q's rows, in order); the prolongation from level `repl`
    reads the replicated vector directly.
    """
    T: sp.csr_matrix
    Tt: sp.csr_matrix
    A: list
    P: list
    R: list
    pinv: np.ndarray
    lmax: list
    repl: int
    plans: list            # HaloPlan per distributed level (l < repl)
    n_owned: list          # per level: owned rows on this rank (= n for replicated levels)
    gather_counts: np.ndarray  # [nranks] owned rows of level `repl` per rank
    gather_gid: np.ndarray     # concatenated global ids, rank-major

    @property
    def nlevels(self):
        return len(self.A)


def _coarse_owner(P: sp.csr_matrix, fine_owner: np.ndarray) -> np.ndarray:
    """owner of coarse dof j = owner of the fine dof with the largest weight in column j"""
    Pc = P.tocsc()
    nnz_per_col = np.diff(Pc.indptr)
    assert np.all(nnz_per_col > 0), "empty prolongation column"
    col = np.repeat(np.arange(Pc.shape[1]), nnz_per_col)
    order = np.lexsort((-np.abs(Pc.data), col))
    first = order[np.concatenate([[0], np.cumsum(nnz_per_col)[:-1]])]
    return fine_owner[Pc.indices[first]].astype(np.int32)


def _rows_cols(M: sp.csr_matrix, rows: np.ndarray) -> np.ndarray:
    return np.unique(M[rows].indices) if rows.size else np.zeros(0, dtype=np.int64)


def _take(M: sp.csr_matrix, rows: np.ndarray, col_g2l: np.ndarray | None, ncols: int) -> sp.csr_matrix:
    sub = M[rows].tocsr()
    if col_g2l is not None:
        idx = col_g2l[sub.indices]
        assert np.all(idx >= 0), "matrix row references a column outside the local set"
        sub = sp.csr_matrix((sub.data, idx, sub.indptr), shape=(rows.size, ncols))
    sub.sort_indices()
    return sub


def partition_hierarchy(H, mesh: Mesh, lm: LocalMesh, k: int, repl_threshold: int = 100_000) -> LocalHierarchy:
    """split the global hierarchy `H` (multigrid.build_hierarchy on the *global* mesh) for rank lm.rank"""
    rank, nranks = lm.rank, lm.nranks
    nl = H.nlevels
    sizes = [a.shape[0] for a in H.A]
    repl = next((l for l, n in enumerate(sizes) if n <= repl_threshold), nl - 1)
    repl = min(repl, nl - 1)
    # ownership per level
    _, vert_owner = _entity_owners(mesh, lm.cell_rank)
    owners = [vert_owner]
    for l in range(nl - 1):
        owners.append(_coarse_owner(H.P[l], owners[-1]))
    owned = [[np.nonzero(owners[l] == q)[0] for q in range(nranks)] for l in range(nl)]
    # local sets of every rank on the distributed levels
    plans = []
    Rg = [p.T.tocsr() for p in H.P]
    for l in range(repl):
        sets = []
        for q in range(nranks):
            parts = [owned[l][q], _rows_cols(H.A[l], owned[l][q])]
            if l == 0:
                c = _local_cells(mesh, lm.cell_rank, q)
                parts.append(np.unique(mesh.cell_vert[c].ravel()))
            else:
                parts.append(_rows_cols(H.P[l - 1], owned[l - 1][q]))
            parts.append(_rows_cols(Rg[l], owned[l + 1][q]))
            sets.append(np.unique(np.concatenate(parts)))
        plans.append(build_plan(owners[l], sets, rank))
    if repl > 0:
        # level-0 numbering must coincide with the local mesh's vertex numbering (T acts on mesh vertices)
        assert np.array_equal(plans[0].local_gid, lm.verts.local_gid)
    g2l = [_global_to_local(p.local_gid, sizes[l]) for l, p in enumerate(plans)]
    A, P, R, n_owned = [], [], [], []
    for l in range(nl):
        if l < repl:
            rows = owned[l][rank]
            A.append(_take(H.A[l], rows, g2l[l], plans[l].n_local))
            n_owned.append(int(rows.size))
        else:
            A.append(H.A[l].tocsr())
            n_owned.append(sizes[l])
    for l in range(nl - 1):
        if l < repl:
            rows_f = owned[l][rank]
            if l + 1 < repl:
                P.append(_take(H.P[l], rows_f, g2l[l + 1], plans[l + 1].n_local))
            else:
                P.append(_take(H.P[l], rows_f, None, sizes[l + 1]))
            R.append(_take(Rg[l], owned[l + 1][rank], g2l[l], plans[l].n_local))
        else:
            P.append(H.P[l].tocsr())
            R.append(Rg[l])
    # trace transfers: rows/cols in the local SoA trace numbering mode * nf_local + facet
    b = k + 1
    nf_g, nf_l = mesh.nf, lm.facets.n_local
    fg = lm.facets.local_gid
    trace_rows = (np.arange(b)[:, None] * nf_g + fg[None, :]).ravel()  # local dof (m, f) -> global dof
    Tg = H.T.tocsr()
    if repl > 0:
        T = _take(Tg, trace_rows, g2l[0], plans[0].n_local)
        Tt_full = T.T.tocsr()  # [n_local verts, b nf_l]
        Tt = Tt_full[: plans[0].n_owned].tocsr()
    else:
        T = _take(Tg, trace_rows, None, sizes[0])
        # restriction rows = owned level-0 vertices (global order); all-gathered afterwards
        Tt = T.T.tocsr()[owned[0][rank]].tocsr()
    Tt.sort_indices()
    counts = np.array([owned[repl][q].size for q in range(nranks)], dtype=np.int32)
    gid = np.concatenate([owned[repl][q] for q in range(nranks)]).astype(np.int32)
    return LocalHierarchy(T=T, Tt=Tt, A=A, P=P, R=R, pinv=H.pinv, lmax=list(H.lmax), repl=repl, plans=plans,
                          n_owned=n_owned, gather_counts=counts, gather_gid=gid)


# ------------------------------------------------------------------------------------------------
# host-side exchange (tests and CPU emulation; the engine does the same with NCCL)
# ------------------------------------------------------------------------------------------------
def exchange_host(plan: HaloPlan, field_soa: np.ndarray, rank: int, dist=None):
    """refresh the ghost entries of a host SoA field [ndof, n_local] over torch.distributed (gloo)"""
    import torch
    import torch.distributed as td

    dist = td if dist is None else dist
    assert field_soa.shape[-1] == plan.n_local
    f2 = field_soa.reshape(-1, plan.n_local)
    ops, bufs = [], []
    for j, q in enumerate(plan.peers):
        s = plan.send_idx[plan.send_ptr[j]:plan.send_ptr[j + 1]]
        if s.size:
            sb = torch.from_numpy(np.ascontiguousarray(f2[:, s]))
            ops.append(dist.P2POp(dist.isend, sb, int(q)))
        if plan.recv_cnt[j]:
            rb = torch.empty((f2.shape[0], int(plan.recv_cnt[j])), dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, rb, int(q)))
            bufs.append((int(plan.recv_off[j]), rb))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for off, rb in bufs:
        f2[:, off:off + rb.shape[1]] = rb.numpy()
    return field_soa
