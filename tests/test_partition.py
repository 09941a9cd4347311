"""CPU tests of the multi-GPU host logic (partition.py): local meshes, halo plans, the split
multigrid hierarchy, and a world_size-2 `gloo` run of a distributed block-Jacobi CG on the condensed
trace system that follows exactly the engine's scheme (kernels over all local entities, reductions
over owned entities + all-reduce, ghost refresh before every neighbour read).

The reference has no partitioning code of its own (Firedrake/PETSc do it implicitly, SURVEY.md
§2.3), so these tests pin the explicit decomposition against the undecomposed problem.
"""
import os
import socket

import numpy as np
import pytest

from incompressibleeulerhdg_b200 import multigrid as MG
from incompressibleeulerhdg_b200 import partition as PT
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh


def _inproc_exchange(plans, fields):
    """what the NCCL exchange does, for all ranks at once: fields[r] is [ndof, n_local]"""
    for r, pl in enumerate(plans):
        for j, q in enumerate(pl.peers):
            s = pl.send_idx[pl.send_ptr[j]:pl.send_ptr[j + 1]]
            if s.size == 0:
                continue
            pq = plans[q]
            jj = list(pq.peers).index(r)
            assert pq.recv_cnt[jj] == s.size
            fields[q][:, pq.recv_off[jj]:pq.recv_off[jj] + s.size] = fields[r][:, s]


CASES = [
    (lambda: UnitSquareMesh(8, perturb=0.1), 2, "strip"),
    (lambda: UnitSquareMesh(12), 4, "strip"),
    (lambda: UnitSquareMesh(16), 4, "block"),
    (lambda: UnitDiskMesh(3), 3, "block"),
    (lambda: PeriodicSquareMesh(6), 2, "strip"),
    (lambda: UnitSquareMesh(6), 1, "strip"),
]


def _cell_rank(mesh, N, kind):
    if kind == "strip":
        return PT.strip_partition(mesh, N)
    return PT.block_partition(mesh, 2, 2) if N == 4 else PT.block_partition(mesh, N, 1)


@pytest.mark.parametrize("mesh_fn,N,kind", CASES)
def test_local_meshes_and_plans(mesh_fn, N, kind):
    mesh = mesh_fn()
    cr = _cell_rank(mesh, N, kind)
    lms = [PT.partition_mesh(mesh, cr, r, N) for r in range(N)]
    for attr, ng in (("cells", mesh.nc), ("facets", mesh.nf), ("verts", mesh.nv)):
        plans = [getattr(l, attr) for l in lms]
        # owned sets partition the global set
        assert np.array_equal(np.sort(np.concatenate([p.owned_gid for p in plans])), np.arange(ng))
        fields = []
        for p in plans:
            f = np.full((2, p.n_local), -1.0)
            f[0, :p.n_owned] = p.owned_gid
            f[1, :p.n_owned] = 2.0 * p.owned_gid + 1
            fields.append(f)
        _inproc_exchange(plans, fields)
        for p, f in zip(plans, fields):
            assert np.array_equal(f[0], p.local_gid) and np.array_equal(f[1], 2.0 * p.local_gid + 1)
            # ghost blocks contiguous, ordered by peer, covering the ghost range (what hdg_set_halo_plan checks)
            off = p.n_owned
            for o, c in zip(p.recv_off, p.recv_cnt):
                if c:
                    assert o == off
                    off += c
            assert off == p.n_local
    for l in lms:
        m = l.mesh
        f = np.arange(m.nf)
        assert np.array_equal(m.cell_facet[m.facet_cell[:, 0], m.facet_local[:, 0]], f)
        two = m.facet_cell[:, 1] >= 0
        assert np.array_equal(m.cell_facet[m.facet_cell[two, 1], m.facet_local[two, 1]], f[two])
        # geometry, orientation and topology of the local mesh are the global ones
        assert np.array_equal(m.cell_xy, mesh.cell_xy[l.cells.local_gid])
        assert np.array_equal(m.cell_flip, mesh.cell_flip[l.cells.local_gid])
        assert np.array_equal(l.facets.local_gid[m.cell_facet], mesh.cell_facet[l.cells.local_gid])
        # owned facets and the facets of owned cells see both of their cells, in the global slot order
        nfo = l.facets.n_owned
        gfc = mesh.facet_cell[l.facets.local_gid]
        interior = gfc[:, 1] >= 0
        need = np.zeros(m.nf, dtype=bool)
        need[:nfo] = True
        need[m.cell_facet[: l.cells.n_owned].ravel()] = True
        assert np.all(m.facet_cell[need & interior, 1] >= 0)
        assert np.array_equal(l.cells.local_gid[m.facet_cell[:nfo, 0]], gfc[:nfo, 0])
        assert abs(l.global_volume - mesh.volume) < 1e-14


@pytest.mark.parametrize("thr", [10 ** 6, 300, 100, 30])
@pytest.mark.parametrize("N", [2, 3])
def test_split_hierarchy_is_the_global_one(N, thr):
    """every distributed operator applied to a ghost-refreshed local vector reproduces the owned rows
    of the global operator applied to the global vector"""
    k = 2
    mesh = UnitSquareMesh(32, perturb=0.1)
    H = MG.build_hierarchy(mesh, k)
    cr = PT.strip_partition(mesh, N)
    rng = np.random.default_rng(1)
    xs = [rng.standard_normal(a.shape[0]) for a in H.A]
    xt = rng.standard_normal(H.T.shape[0])
    gathered = None
    for r in range(N):
        lm = PT.partition_mesh(mesh, cr, r, N)
        LH = PT.partition_hierarchy(H, mesh, lm, k, repl_threshold=thr)
        assert 0 <= LH.repl <= H.nlevels - 1
        loc = lambda l: xs[l][LH.plans[l].local_gid] if l < LH.repl else xs[l]
        own = lambda l: LH.plans[l].owned_gid if l < LH.repl else np.arange(H.A[l].shape[0])
        for l in range(H.nlevels):
            assert np.allclose(LH.A[l] @ loc(l), (H.A[l] @ xs[l])[own(l)], rtol=1e-13, atol=1e-13)
        for l in range(H.nlevels - 1):
            assert np.allclose(LH.P[l] @ loc(l + 1), (H.P[l] @ xs[l + 1])[own(l)], rtol=1e-13, atol=1e-13)
            rows = own(l + 1)
            if l + 1 == LH.repl and LH.repl > 0:
                ptr = np.concatenate([[0], np.cumsum(LH.gather_counts)])
                rows = LH.gather_gid[ptr[r]:ptr[r + 1]]
            assert np.allclose(LH.R[l] @ loc(l), (H.P[l].T @ xs[l])[rows], rtol=1e-13, atol=1e-13)
        # trace transfers in the local SoA numbering mode * nf_local + facet
        b, fg = k + 1, lm.facets.local_gid
        tl = (np.arange(b)[:, None] * mesh.nf + fg[None, :]).ravel()
        assert np.allclose(LH.T @ loc(0), (H.T @ xs[0])[tl], rtol=1e-13, atol=1e-13)
        ptr = np.concatenate([[0], np.cumsum(LH.gather_counts)])
        rows0 = own(0) if LH.repl > 0 else LH.gather_gid[ptr[r]:ptr[r + 1]]
        # restriction to the owned P1 vertices needs every facet around them: all local
        full = H.T.T @ xt
        assert np.allclose(LH.Tt @ xt[tl], full[rows0], rtol=1e-13, atol=1e-13)
        g = np.zeros(H.A[LH.repl].shape[0], dtype=np.int64) if gathered is None else gathered
        g[LH.gather_gid[ptr[r]:ptr[r + 1]]] += 1
        gathered = g
    assert np.all(gathered == 1)  # the all-gather of the first replicated level covers it exactly once


# ---- world_size 2 over gloo -------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dist_cg_worker(rank, world, port, k, nx, out_dir):
    import torch
    import torch.distributed as dist

    from oracle.hdg_oracle import HDGOracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = UnitSquareMesh(nx, perturb=0.1)
        b = k + 1
        lm = PT.partition_mesh(mesh, PT.strip_partition(mesh, world), rank, world)
        nfo, nfl, fg = lm.nf_owned, lm.facets.n_local, lm.facets.local_gid
        # global problem (every rank can afford it at this size): P = -S, rhs in the range of P
        og = HDGOracle(mesh, k)
        Pg = (-og.assemble_trace_matrix()).tocsr()
        xtrue = np.random.default_rng(5).standard_normal((mesh.nf, b))
        bg = (Pg @ xtrue.ravel()).reshape(mesh.nf, b)
        # local operator from the *local mesh only*: rows of owned facets are complete
        Pl = (-HDGOracle(lm.mesh, k).assemble_trace_matrix()).tocsr()
        D = Pl.diagonal().reshape(nfl, b)

        def allsum(v):
            t = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        def dot(u, v):  # owned entries only, then summed over ranks (the engine's reduction rule)
            return allsum(float(np.sum(u[:nfo] * v[:nfo])))

        def refresh(v):  # v: [nfl, b] AoS -> exchange works on SoA [b, nfl]
            soa = np.ascontiguousarray(v.T)
            PT.exchange_host(lm.facets, soa, rank)
            v[:] = soa.T

        r = bg[fg].copy()
        x = np.zeros((nfl, b))
        z = r / D
        p = z.copy()
        rz = dot(r, z)
        rz0 = rz
        its = 0
        while rz > 1e-26 * rz0 and its < 5000:
            refresh(p)
            q = (Pl @ p.ravel()).reshape(nfl, b)
            alpha = rz / dot(p, q)
            x += alpha * p
            r -= alpha * q
            z = r / D
            rz_new = dot(r, z)
            p = z + (rz_new / rz) * p
            rz = rz_new
            its += 1
        # compare with the true solution modulo the constant null vector (mode 0)
        d = x[:nfo] - xtrue[fg[:nfo]]
        shift = allsum(float(d[:, 0].sum())) / mesh.nf
        d[:, 0] -= shift
        err = allsum(float(np.sum(d * d))) ** 0.5 / np.linalg.norm(xtrue)
        np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([err, its]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_distributed_cg_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_dist_cg_worker, args=(2, port, 1, 8, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        err, its = np.load(tmp_path / f"res{r}.npy")
        assert its < 5000
        assert err < 1e-9, err
