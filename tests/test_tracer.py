"""Passive-tracer path, CPU part: the topological CG_{k+1} space, the quadrature
tables and a numpy mirror of the device kernels' arithmetic (`csrc/hdg_tracer.cuh`) against the oracle
(`oracle/tracer.py`, which finds the CG space by coordinate matching and uses sparse direct solves)."""
import numpy as np
import pytest

from incompressibleeulerhdg_b200 import cgspace
from incompressibleeulerhdg_b200 import refelem as R
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.tracer import TracerOracle

MESHES = {"square": lambda: UnitSquareMesh(5, perturb=0.15), "disk": lambda: UnitDiskMesh(1)}


def tg_velocity(x, y):
    return (-np.cos((x - 0.5) * np.pi) * np.sin((y - 0.5) * np.pi), np.sin((x - 0.5) * np.pi) * np.cos((y - 0.5) * np.pi))


def tracer0(x, y):
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)


# ---- numpy mirrors of the kernels (same tables, same formulas, vectorised over cells) ------------------
def mirror_mass_apply(mesh, sp_, x):
    """k_cgp_cellop + k_cgp_gather<1>:  y = sum_K G^T detJ W^T W G x"""
    detJ = 2.0 * mesh.cell_area()
    xl = x[sp_.cellmap]  # [nc, nloc]
    yK = detJ[:, None] * ((xl @ sp_.W.T) @ sp_.W)
    flat = yK.T.ravel()  # SoA [nloc][nc]: index j*nc + cell
    y = np.zeros(sp_.ndof)
    for g in range(sp_.ndof):
        y[g] = flat[sp_.inc_idx[sp_.inc_ptr[g]:sp_.inc_ptr[g + 1]]].sum()
    return y


def mirror_project(mesh, sp_, Q, rtol=1e-14, maxit=500):
    """run_project_cg: load vector, Jacobi-PCG on the matrix-free mass matrix, back to the cells"""
    detJ = 2.0 * mesh.cell_area()
    out = np.empty_like(Q)
    its = []
    for c in range(2):
        yK = detJ[:, None] * (Q[:, c, :] @ sp_.W)  # yK[cell][j] = detJ sum_i W[i][j] U[i]
        b = np.bincount(sp_.cellmap.ravel(), weights=yK.ravel(), minlength=sp_.ndof)
        dinv = 1.0 / sp_.diag
        x = np.zeros_like(b)
        r = b.copy()
        z = dinv * r
        p = z.copy()
        rz = rz0 = r @ z
        for it in range(maxit):
            Ap = mirror_mass_apply(mesh, sp_, p)
            al = rz / (p @ Ap)
            x += al * p
            r -= al * Ap
            z = dinv * r
            rz_new = r @ z
            p = z + (rz_new / rz) * p
            rz = rz_new
            if rz <= rtol ** 2 * rz0:
                break
        its.append(it + 1)
        out[:, c, :] = x[sp_.cellmap] @ sp_.W.T
    return out, its


def mirror_advection(mesh, o, k, U, q):
    """k_tracer_adv with c0 = 0, c1 = 1"""
    tab_cell, tab_facet = cgspace.tracer_tables(k, o.nq_facet)
    NP, NQ1 = R.ncell(k), R.ncell(k + 1)
    Ji = o.Jinv  # Ji[n, d, c]
    res = np.zeros((mesh.nc, NP))
    for t in tab_cell:
        w, chi, d0c, d1c = t[0], t[1:1 + NP], t[1 + NP:1 + 2 * NP], t[1 + 2 * NP:1 + 3 * NP]
        psi, d0p, d1p = (t[1 + 3 * NP + m * NQ1:1 + 3 * NP + (m + 1) * NQ1] for m in range(3))
        qv = q @ chi
        uv = U @ psi  # [nc, 2]
        du0, du1 = U @ d0p, U @ d1p  # d u_c / d xi_0, d xi_1
        b0 = Ji[:, 0, 0] * uv[:, 0] + Ji[:, 0, 1] * uv[:, 1]
        b1 = Ji[:, 1, 0] * uv[:, 0] + Ji[:, 1, 1] * uv[:, 1]
        divu = Ji[:, 0, 0] * du0[:, 0] + Ji[:, 1, 0] * du1[:, 0] + Ji[:, 0, 1] * du0[:, 1] + Ji[:, 1, 1] * du1[:, 1]
        res += (w * qv)[:, None] * (b0[:, None] * d0c + b1[:, None] * d1c + divu[:, None] * chi)
    nqf = tab_facet.shape[1]
    for e in range(3):
        nb, ne = o.nbr[:, e], o.nbr_e[:, e]
        has = nb >= 0
        for qf in range(nqf):
            t = tab_facet[e, qf]
            tn = tab_facet[np.maximum(ne, 0), nqf - 1 - qf]  # [nc, SF]
            un = (o.normal[:, e, 0, None] * U[:, 0, :] + o.normal[:, e, 1, None] * U[:, 1, :]) @ t[1 + NP:]
            qin = q @ t[1:1 + NP]
            qout = np.einsum("na,na->n", q[np.maximum(nb, 0)], tn[:, 1:1 + NP])
            flux = np.maximum(un, 0) * qin + np.minimum(un, 0) * qout
            wf = np.where(has, -t[0] * o.elen[:, e] / o.detJ * flux, 0.0)
            res += wf[:, None] * t[1:1 + NP]
    return res


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("name", list(MESHES))
def test_cg_space_matches_coordinate_matching(k, name):
    mesh = MESHES[name]()
    t = TracerOracle(HDGOracle(mesh, k))
    sp_ = cgspace.build_cg_space(mesh, k + 1)
    assert sp_.ndof == t.ndof
    pairs = set(zip(t.cellmap.ravel().tolist(), sp_.cellmap.ravel().tolist()))
    assert len(pairs) == sp_.ndof  # the two numberings are a bijection of each other
    perm = np.empty(sp_.ndof, dtype=np.int64)
    for a, b in pairs:
        perm[b] = a
    assert np.abs(t.M.diagonal()[perm] - sp_.diag).max() < 1e-14
    # the incidence CSR is the transpose of the cell map
    nc = mesh.nc
    for g in (0, sp_.ndof // 2, sp_.ndof - 1):
        idx = sp_.inc_idx[sp_.inc_ptr[g]:sp_.inc_ptr[g + 1]]
        assert np.all(sp_.cellmap[idx % nc, idx // nc] == g)


def test_cg_space_periodic_euler_characteristic():
    """on the torus V - E + F = 0, so CG_d has nv + nf (d-1) + nc nint dofs with nv = nx^2"""
    mesh = PeriodicSquareMesh(4, L=2 * np.pi)
    sp_ = cgspace.build_cg_space(mesh, 3)
    assert mesh.nv - mesh.nf + mesh.nc == 0
    assert sp_.ndof == mesh.nv + 2 * mesh.nf + mesh.nc
    assert np.bincount(sp_.cellmap.ravel(), minlength=sp_.ndof).min() >= 1


@pytest.mark.parametrize("k", [1, 2])
def test_matrix_free_mass_matches_assembled(k):
    mesh = MESHES["square"]()
    t = TracerOracle(HDGOracle(mesh, k))
    sp_ = cgspace.build_cg_space(mesh, k + 1)
    x = np.random.default_rng(3).standard_normal(sp_.ndof)
    perm = np.empty(sp_.ndof, dtype=np.int64)
    perm[sp_.cellmap.ravel()] = t.cellmap.ravel()
    xo = np.zeros(sp_.ndof)
    xo[perm] = x
    y = mirror_mass_apply(mesh, sp_, x)
    assert np.abs(y - (t.M @ xo)[perm]).max() < 1e-13 * np.abs(y).max()


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("name", list(MESHES))
def test_kernel_mirror_matches_oracle(k, name):
    mesh = MESHES[name]()
    o = HDGOracle(mesh, k)
    t = TracerOracle(o)
    sp_ = cgspace.build_cg_space(mesh, k + 1)
    rng = np.random.default_rng(5)
    Q = o.interpolate_cell(tg_velocity, "Q") + 0.05 * rng.standard_normal((mesh.nc, 2, o.nQ1))  # discontinuous
    U, its = mirror_project(mesh, sp_, Q)
    Uo = t.project_cg(Q)
    assert max(its) < 100
    assert np.abs(U - Uo).max() < 1e-11 * np.abs(Uo).max()
    q = o.interpolate_cell(tracer0, "p") + 0.1 * rng.standard_normal((mesh.nc, o.np_))
    adv = mirror_advection(mesh, o, k, Uo, q)
    advo = t.advection(q, Uo)
    assert np.abs(adv - advo).max() < 1e-11 * np.abs(advo).max()


def test_advection_is_conservative_and_consistent():
    """sum_K int adv = int q div u_cg (chi = 1 has no jump); a constant tracer in a discretely
    divergence-free, boundary-tangential velocity stays put up to div u_cg"""
    mesh = UnitSquareMesh(8)
    k = 2
    o = HDGOracle(mesh, k)
    t = TracerOracle(o)
    U = t.project_cg(o.interpolate_cell(tg_velocity, "Q"))
    q = o.interpolate_cell(tracer0, "p")
    adv = t.advection(q, U)
    qv = np.einsum("na,aq->nq", q, o.phiP)
    gpsi = np.einsum("ndc,iqd->niqc", o.Jinv, o.dphiQ)
    divu = np.einsum("nci,niqc->nq", U, gpsi)
    assert abs(t.total_mass(adv) - np.einsum("n,q,nq,nq->", o.detJ, o.wq, qv, divu)) < 1e-13
    one = o.interpolate_cell(lambda x, y: 1.0 + 0 * x, "p")
    adv1 = t.advection(one, U)
    # M^-1 adv(1, u) = L2 projection of div u (facet fluxes of a constant cancel the boundary term of the
    # integration by parts only where u.n is continuous, which it is for the CG velocity)
    assert np.abs(adv1).max() < 5e-3


def test_compiled_vinv_table_is_the_refelem_map():
    """the projection kernels carry VINV (csrc/hdg_tables.inc, tools/gen_tables.py) as compile-time
    constants; it must be the modal <- nodal map that cgspace.build_cg_space hands to hdg_tracer_setup"""
    import os
    import re

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "incompressibleeulerhdg_b200", "csrc",
                        "hdg_tables.inc")
    txt = open(path).read()
    for k in (1, 2, 3, 4):
        ns = txt[txt.index(f"namespace hdg_tab_k{k} {{"):]
        m = re.search(r"VINV\[(\d+)\]\[(\d+)\] = (\{.*?\});", ns)
        arr = np.array(eval(m.group(3).replace("{", "[").replace("}", "]")))
        W = R.nodal_to_modal_cell(k + 1)
        assert arr.shape == W.shape == (R.ncell(k + 1),) * 2
        assert np.abs(arr - W).max() < 1e-13 * np.abs(W).max()


# ---- partitioned mesh: CG-dof halo plan and the distributed projection (host logic of the multi-GPU path) --------
@pytest.mark.parametrize("world", [2, 3])
def test_cg_plan_matches_global_space(world):
    from incompressibleeulerhdg_b200 import partition as PT

    mesh = UnitSquareMesh(6, perturb=0.1)
    cr = PT.strip_partition(mesh, world)
    G = cgspace.build_cg_space(mesh, 3)
    cover = np.zeros(G.ndof, dtype=int)
    for r in range(world):
        lm = PT.partition_mesh(mesh, cr, r, world)
        plan, perm = PT.cg_plan(mesh, lm, 3)
        L = cgspace.build_cg_space(lm.mesh, 3, perm=perm)
        assert L.ndof == plan.n_local and np.array_equal(np.sort(perm), np.arange(L.ndof))
        # the renumbered local cell map names the same global dofs as the global space (incl. facet direction)
        assert np.array_equal(plan.local_gid[L.cellmap], G.cellmap[lm.cells.local_gid])
        own = plan.local_gid[:plan.n_owned]
        assert np.allclose(L.diag[:plan.n_owned], G.diag[own], rtol=1e-14)  # owned rows see all their cells
        # ghost blocks: contiguous per peer, in the order the owner packs them
        off = plan.n_owned
        for j in range(len(plan.peers)):
            if plan.recv_cnt[j]:
                assert plan.recv_off[j] == off
                off += plan.recv_cnt[j]
        assert off == plan.n_local
        cover[own] += 1
    assert np.all(cover == 1)


def _dist_project_worker(rank, world, port, out_dir):
    import os

    import torch
    import torch.distributed as dist

    from incompressibleeulerhdg_b200 import partition as PT

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = 1
        mesh = UnitSquareMesh(6, perturb=0.1)
        lm = PT.partition_mesh(mesh, PT.strip_partition(mesh, world), rank, world)
        plan, perm = PT.cg_plan(mesh, lm, k + 1)
        L = cgspace.build_cg_space(lm.mesh, k + 1, perm=perm)
        no = plan.n_owned
        og = HDGOracle(mesh, k)
        Qg = og.interpolate_cell(tg_velocity, "Q") + 0.05 * np.random.default_rng(9).standard_normal((mesh.nc, 2, og.nQ1))
        Ug = TracerOracle(og).project_cg(Qg)  # global reference, every rank can afford it at this size
        Ql = Qg[lm.cells.local_gid]
        detJ = 2.0 * lm.mesh.cell_area()

        def allsum(v):
            t = torch.tensor(v, dtype=torch.float64)
            dist.all_reduce(t)
            return t.numpy()

        def dots(u, v):  # owned dofs only, summed over the ranks: [2]
            return allsum(np.sum(u[:, :no] * v[:, :no], axis=1))

        def gather(yK):  # yK [2, nc, nloc] -> [2, ndof] (ghost rows incomplete, like k_cgp_gather)
            return np.stack([np.bincount(L.cellmap.ravel(), weights=yK[c].ravel(), minlength=L.ndof) for c in range(2)])

        def apply_mass(p):  # k_cgp_cellop + k_cgp_gather<1> after the halo exchange of p
            PT.exchange_host(plan, p, rank)
            xl = p[:, L.cellmap]  # [2, nc, nloc]
            return gather(detJ[None, :, None] * ((xl @ L.W.T) @ L.W))

        b = gather(detJ[None, :, None] * (np.swapaxes(Ql, 0, 1) @ L.W))
        dinv = 1.0 / L.diag
        x = np.zeros_like(b)
        r = b.copy()
        z = dinv * r
        p = z.copy()
        rz = rz0 = dots(r, z)
        its = 0
        while np.any(rz > 1e-28 * rz0) and its < 300:
            Ap = apply_mass(p)
            al = rz / dots(p, Ap)
            x += al[:, None] * p
            r -= al[:, None] * Ap
            z = dinv * r
            rz_new = dots(r, z)
            p = z + (rz_new / rz)[:, None] * p
            rz = rz_new
            its += 1
        PT.exchange_host(plan, x, rank)  # before k_cgp_tocell
        Ul = np.swapaxes(x[:, L.cellmap] @ L.W.T, 0, 1)  # [nc_local, 2, nloc]
        nco = lm.nc_owned
        err = np.abs(Ul[:nco] - Ug[lm.cells.local_gid[:nco]]).max() / np.abs(Ug).max()
        np.save(os.path.join(out_dir, f"proj{rank}.npy"), np.array([err, its]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_distributed_cg_projection_gloo_world2(tmp_path):
    """the multi-GPU projection algorithm (owned-first dofs, owned-only dots, ghost refresh of p and x) on two
    gloo ranks reproduces the single-domain oracle projection on every owned cell"""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dist_project_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        err, its = np.load(tmp_path / f"proj{r}.npy")
        assert its < 300 and err < 1e-11, (err, its)
