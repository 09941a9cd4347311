"""Array-level Firedrake marshalling (`incompressibleeulerhdg_b200/firedrake_adapter.py`) on
fabricated nodal layouts: arbitrary local node order per cell, arbitrary global numbering, two node variants.
Firedrake itself is not installable here; the Firedrake-facing wrapper is import-guarded."""
import numpy as np
import pytest

from incompressibleeulerhdg_b200 import refelem as R
from incompressibleeulerhdg_b200.firedrake_adapter import NodalCellLayout, NodalFacetLayout, SCPC, mesh_from_arrays
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh

MESHES = {"square": lambda: UnitSquareMesh(4, perturb=0.15), "disk": lambda: UnitDiskMesh(1),
          "periodic": lambda: PeriodicSquareMesh(4, L=2 * np.pi)}
VARIANTS = {"equispaced": lambda n: n, "interior": lambda n: 1.0 / 3.0 + 0.8 * (n - 1.0 / 3.0)}


def fabricate_cells(mesh, m, variant, rng):
    """a DG_m nodal layout with a random local order in every cell and a random global numbering"""
    nodes = VARIANTS[variant](R.lagrange_nodes_cell(m))
    nc, nloc = mesh.nc, nodes.shape[0]
    perm = np.array([rng.permutation(nloc) for _ in range(nc)])
    cell_nodes = rng.permutation(nc * nloc).reshape(nc, nloc)
    x = mesh.cell_xy
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)
    xy_loc = x[:, None, 0, :] + np.einsum("ncd,nqd->nqc", J, nodes[perm])
    node_xy = np.empty((nc * nloc, 2))
    node_xy[cell_nodes] = xy_loc
    return cell_nodes, node_xy


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("m", [1, 2, 3])
def test_cell_layout_round_trip_and_exactness(name, variant, m):
    mesh = MESHES[name]()
    rng = np.random.default_rng(7)
    cell_nodes, node_xy = fabricate_cells(mesh, m, variant, rng)
    lay = NodalCellLayout(mesh, m, cell_nodes, node_xy)
    # a polynomial of degree m, sampled at the nodes, is reproduced by the modal coefficients
    poly = lambda x, y: 1.0 + 0.5 * x - 0.25 * y + (x * y if m >= 2 else 0) + (x ** 3 - y ** 2 * x if m >= 3 else 0)
    x = mesh.cell_xy
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)
    # evaluate in the frame of each cell (periodic meshes: nodes were generated unwrapped)
    xi = rng.dirichlet(np.ones(3), size=5)[:, 1:]
    xp = x[:, None, 0, :] + np.einsum("ncd,qd->nqc", J, xi)
    data = poly(node_xy[:, 0], node_xy[:, 1])
    coef = lay.to_modal(data)
    vals = np.einsum("ni,iq->nq", coef, R.dubiner(m, xi))
    assert np.abs(vals - poly(xp[..., 0], xp[..., 1])).max() < 1e-10 * max(1.0, np.abs(data).max())
    # vector-valued round trip
    d2 = rng.standard_normal((node_xy.shape[0], 2))
    c2 = lay.to_modal(d2)
    assert c2.shape == (mesh.nc, 2, R.ncell(m))
    assert np.abs(lay.from_modal(c2, np.zeros_like(d2)) - d2).max() < 1e-11
    # duality: <r, u> is the same number in both bases
    r = rng.standard_normal(d2.shape)
    assert abs(np.sum(r * d2) - np.sum(lay.dual_to_modal(r) * c2)) < 1e-9 * np.abs(r).sum()
    assert np.abs(lay.dual_from_modal(lay.dual_to_modal(r), np.zeros_like(r)) - r).max() < 1e-10


def fabricate_facets(mesh, k, nodes01, rng):
    """a DGT_k layout: per cell the three facet groups in random order, nodes inside a group in random
    direction, global node ids shared by the two cells of a facet"""
    k1 = k + 1
    nc, nf = mesh.nc, mesh.nf
    ids = rng.permutation(nf * k1).reshape(nf, k1)  # node j of facet f sits at global parameter nodes01[j]
    cell_nodes = np.empty((nc, 3 * k1), dtype=np.int64)
    node_xy = np.full((nf * k1, 2), np.nan)
    x = mesh.cell_xy
    for c in range(nc):
        for g, e in enumerate(rng.permutation(3)):
            f = mesh.cell_facet[c, e]
            a, b = x[c, (e + 1) % 3], x[c, (e + 2) % 3]
            s_glob = nodes01
            s_loc = 1.0 - s_glob if mesh.cell_flip[c, e] else s_glob
            order = rng.permutation(k1)
            cell_nodes[c, g * k1:(g + 1) * k1] = ids[f, order]
            pos = a[None, :] + s_loc[order, None] * (b - a)[None, :]
            if np.isnan(node_xy[ids[f, 0], 0]):  # the first cell that sees the facet places its nodes
                node_xy[ids[f, order]] = pos
    return cell_nodes, node_xy, ids


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("k", [0, 1, 2, 3])
@pytest.mark.parametrize("variant", ["equispaced", "gauss"])
def test_facet_layout(name, k, variant):
    mesh = MESHES[name]()
    rng = np.random.default_rng(11)
    nodes01 = R.lagrange_nodes_facet(k) if variant == "equispaced" else np.sort(R.gauss_legendre(k + 1)[0])
    cell_nodes, node_xy, ids = fabricate_facets(mesh, k, nodes01, rng)
    lay = NodalFacetLayout(mesh, k, cell_nodes, node_xy)
    assert np.array_equal(lay.facet_nodes, ids)  # nodes recovered facet by facet, ordered along the global direction
    assert np.abs(lay.s - nodes01[None, :]).max() < 1e-10
    # a polynomial of degree k in the global facet parameter is reproduced by the Legendre coefficients
    coefp = rng.standard_normal((mesh.nf, k + 1))
    data = np.zeros(mesh.nf * (k + 1))
    data[ids] = sum(coefp[:, [j]] * nodes01[None, :] ** j for j in range(k + 1))
    lam = lay.to_modal(data)
    st = np.array([0.1, 0.45, 0.8])
    vals = lam @ R.legendre01(k, st)
    assert np.abs(vals - sum(coefp[:, [j]] * st[None, :] ** j for j in range(k + 1))).max() < 1e-10
    assert np.abs(lay.from_modal(lam, np.zeros_like(data)) - data).max() < 1e-11
    r = rng.standard_normal(data.shape)
    assert abs(np.sum(r * data) - np.sum(lay.dual_to_modal(r) * lam)) < 1e-9 * np.abs(r).sum()


def test_mesh_from_arrays_and_pc_protocol():
    m0 = UnitSquareMesh(3)
    vert_xy = np.zeros((m0.nv, 2))
    vert_xy[m0.cell_vert] = m0.cell_xy
    m = mesh_from_arrays(m0.cell_vert, vert_xy=vert_xy)
    assert m.nc == m0.nc and m.nf == m0.nf and np.array_equal(m.cell_facet, m0.cell_facet)
    # the PC exposes what hdg_imex.py:265-271 reads, and fails loudly without Firedrake
    for name in ("initialize", "update", "apply", "applyTranspose", "view"):
        assert callable(getattr(SCPC, name))
    pc = SCPC()
    with pytest.raises(ImportError, match="needs Firedrake"):
        pc.initialize(None)
