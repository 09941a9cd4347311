"""The driver mirror (`incompressibleeulerhdg_b200/driver.py`) against the reference's CLI
(`src/driver.py:24-178`): option names, choices and defaults on CPU; complete runs on the GPU."""
import io
import os

import numpy as np
import pytest

from incompressibleeulerhdg_b200 import driver

# (option, default, choices) of the reference's parser, `src/driver.py:26-176`
REFERENCE_OPTIONS = [
    ("problem", "taylorgreen", ["taylorgreen", "kelvinhelmholtz", "shear"]),
    ("nx", 8, None),
    ("refinement", 2, None),
    ("degree", 1, None),
    ("tfinal", 1.0, None),
    ("kappa", 0.5, None),
    ("dt", 0.04, None),
    ("discretisation", "hdg", ["conforming", "dg", "hdg"]),
    ("use_projection_method", False, None),
    ("richardson", 2, None),
    ("flux", "upwind", ["upwind", "centered"]),
    ("timestepper", "imex_ssp2_332",
     ["implicit", "imex_implicit", "imex_ars2_232", "imex_ars3_443", "imex_ssp2_332", "imex_ssp3_433"]),
    ("forcing", "exponential", ["exponential", "constant"]),
    ("test_pressure_solver", False, None),
    ("warmup", False, None),
    ("animation", False, None),
    ("tracer_advection", False, None),
]


def test_parser_matches_reference_cli():
    parser = driver.build_parser()
    args = parser.parse_args([])
    actions = {a.dest: a for a in parser._actions}
    for name, default, choices in REFERENCE_OPTIONS:
        assert getattr(args, name) == default, name
        if choices is not None:
            assert list(actions[name].choices) == choices, name
    assert parser.prog == "Mesh specifications and polynomial degree"
    a = parser.parse_args("--nx 16 --degree 2 --timestepper imex_ars3_443 --use_projection_method --flux centered".split())
    assert (a.nx, a.degree, a.timestepper, a.use_projection_method, a.flux) == (16, 2, "imex_ars3_443", True, "centered")
    with pytest.raises(SystemExit):
        parser.parse_args(["--timestepper", "rk4"])


def test_mesh_selection_follows_reference():
    p = driver.build_parser()
    m = driver.build_mesh(p.parse_args(["--nx", "4"]))
    assert m.nc == 32 and abs(m.volume - 1.0) < 1e-14
    m = driver.build_mesh(p.parse_args(["--problem", "shear", "--nx", "4"]))
    assert m.nc == 32 and abs(m.volume - (2 * np.pi) ** 2) < 1e-12 and len(m.boundary_facets) == 0
    m = driver.build_mesh(p.parse_args(["--problem", "kelvinhelmholtz", "--refinement", "1"]))
    assert abs(m.volume - np.pi) < 0.5 and len(m.boundary_facets) > 0


def test_timestepper_dispatch_and_scope():
    from incompressibleeulerhdg_b200 import timesteppers as T

    assert driver.timestepper_class("implicit") is T.IncompressibleEulerHDGImplicit
    assert driver.timestepper_class("imex_ssp3_433") is T.IncompressibleEulerHDGIMEXSSP3_433
    args = driver.build_parser().parse_args(["--discretisation", "dg"])
    with pytest.raises(RuntimeError, match="outside the scope"):
        driver.build_timestepper(args, None, None, 0)


def test_header_matches_reference_layout():
    class _TS:
        label = "HDG IMEX SSP2(3,3,2)"

    out = io.StringIO()
    driver.print_header(driver.build_parser().parse_args([]), _TS(), file=out)
    lines = out.getvalue().splitlines()
    assert lines[1] == "! timesteppers for incompressible Euler equations !"
    assert "mesh size = 8 x 8" in lines and "kappa = 0.5" in lines and "advect tracer = False" in lines
    assert lines[-2] == "timestepping method = HDG IMEX SSP2(3,3,2)"


def test_vtk_writer(tmp_path):
    """.pvd collection + one ASCII .vtu per write, discontinuous P1 sampling at the cell vertices"""
    from incompressibleeulerhdg_b200 import refelem as R
    from incompressibleeulerhdg_b200.auxilliary.callbacks import cell_vorticity_at_vertices
    from incompressibleeulerhdg_b200.auxilliary.vtk import VTKFile
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh

    mesh = UnitSquareMesh(3)

    class _Space:
        def __init__(self, name, degree):
            self.name, self.degree = name, degree

        def mesh(self):
            return mesh

    class _Fn:
        def __init__(self, space, coef, name):
            self.space, self.coef, self.name = space, coef, name

        def function_space(self):
            return self.space

        def to_host(self):
            return self.coef

    # Q = (-y, x) (rigid rotation, vorticity 2) and p = x + 2 y as modal coefficients
    nodes = R.lagrange_nodes_cell(2)
    Vinv = R.nodal_to_modal_cell(2, nodes)
    x0 = mesh.cell_xy[:, 0]
    J = np.stack([mesh.cell_xy[:, 1] - x0, mesh.cell_xy[:, 2] - x0], axis=-1)
    xp = x0[:, None, :] + np.einsum("ncd,qd->nqc", J, nodes)
    Qc = np.einsum("iq,nqc->nci", Vinv, np.stack([-xp[..., 1], xp[..., 0]], axis=-1))
    n1 = R.lagrange_nodes_cell(1)
    xp1 = x0[:, None, :] + np.einsum("ncd,qd->nqc", J, n1)
    pc = np.einsum("aq,nq->na", R.nodal_to_modal_cell(1, n1), xp1[..., 0] + 2 * xp1[..., 1])
    Q, p = _Fn(_Space("Q", 2), Qc, "velocity"), _Fn(_Space("p", 1), pc, "pressure")
    assert np.allclose(cell_vorticity_at_vertices(Q), 2.0)
    f = VTKFile(str(tmp_path / "out.pvd"))
    f.write(Q, p, ("vorticity", cell_vorticity_at_vertices(Q)), time=0.0)
    f.write(Q, p, time=0.5)
    pvd = (tmp_path / "out.pvd").read_text()
    assert pvd.count("<DataSet") == 2 and 'timestep="0.5"' in pvd and "out_1.vtu" in pvd
    vtu = (tmp_path / "out_0.vtu").read_text()
    assert f'NumberOfPoints="{3 * mesh.nc}"' in vtu and 'Name="velocity"' in vtu and 'Name="vorticity"' in vtu
    # the pressure samples are x + 2y at the vertices
    block = vtu.split('Name="pressure"')[1].split("</DataArray>")[0].split("\n", 1)[1]
    vals = np.array(block.split(), dtype=float)
    assert np.allclose(vals, (mesh.cell_xy[..., 0] + 2 * mesh.cell_xy[..., 1]).ravel())


# ---- complete runs (GPU) ------------------------------------------------------------------------------
@pytest.mark.gpu
def test_driver_imex_run_reports_small_errors(tmp_path):
    out = io.StringIO()
    res = driver.main(["--nx", "8", "--degree", "1", "--dt", "0.05", "--tfinal", "0.1", "--use_projection_method",
                       "--output", str(tmp_path)], file=out)
    text = out.getvalue()
    assert "timestepping method = " in text and "velocity error = " in text and "pressure error = " in text
    assert res["velocity_error"] < 2e-2 and res["pressure_error"] < 5e-2
    assert res["divergence_norm"] < 1.0
    assert os.path.exists(tmp_path / "solution.pvd") and os.path.exists(tmp_path / "solution_0.vtu")


@pytest.mark.gpu
def test_driver_chorin_with_tracer_and_animation(tmp_path):
    out = io.StringIO()
    res = driver.main(["--nx", "8", "--degree", "2", "--dt", "0.02", "--tfinal", "0.06", "--timestepper", "implicit",
                       "--use_projection_method", "--tracer_advection", "--animation", "--output", str(tmp_path)],
                      file=out)
    assert res["q_tracer"] is not None and np.isfinite(res["q_tracer"].to_host()).all()
    assert res["velocity_error"] < 2e-2
    pvd = (tmp_path / "evolution.pvd").read_text()
    assert pvd.count("<DataSet") == 4  # t = 0 and three steps
    assert 'Name="tracer"' in (tmp_path / "evolution_3.vtu").read_text()


@pytest.mark.gpu
def test_driver_pressure_solver_test_and_warmup(tmp_path):
    out = io.StringIO()
    res = driver.main(["--nx", "16", "--degree", "2", "--test_pressure_solver"], file=out)
    seconds, its = res["pressure_solver"]
    assert "=== Testing pressure solver" in out.getvalue() and 0 < its < 500 and seconds > 0
    out = io.StringIO()
    res = driver.main(["--nx", "8", "--warmup", "--timestepper", "imex_ars2_232", "--use_projection_method",
                       "--output", "none"], file=out)
    assert "WARNING: performing a single timestep only!" in out.getvalue()
    assert "velocity_error" not in res


@pytest.mark.parametrize("m", [2, 3])
def test_vorticity_projector_matches_reference_form(m):
    """`callbacks.py:44-69`: tau xi dx == -eps:(grad tau (x) Q) dx + tau eps:(n (x) Q) ds on CG_m.  For a
    continuous velocity whose curl lies in CG_m the weak projection is the curl itself; for a discontinuous
    velocity it differs from the cell-wise curl (interior jumps are not in the reference's form)."""
    from incompressibleeulerhdg_b200 import refelem as R
    from incompressibleeulerhdg_b200.auxilliary.callbacks import VorticityProjector
    from incompressibleeulerhdg_b200.mesh import UnitDiskMesh

    mesh = UnitDiskMesh(1)
    nodes = R.lagrange_nodes_cell(m)
    Vinv = R.nodal_to_modal_cell(m, nodes)
    x0 = mesh.cell_xy[:, 0]
    J = np.stack([mesh.cell_xy[:, 1] - x0, mesh.cell_xy[:, 2] - x0], axis=-1)
    xp = x0[:, None, :] + np.einsum("ncd,qd->nqc", J, nodes)

    def coef(f):
        return np.einsum("iq,nqc->nci", Vinv, np.stack(f(xp[..., 0], xp[..., 1]), axis=-1))

    P = VorticityProjector(mesh, m)
    assert np.abs(P.at_vertices(coef(lambda x, y: (-y, x))) - 2.0).max() < 1e-11
    w = P.at_vertices(coef(lambda x, y: (x * y, x ** 2)))  # curl = x
    assert np.abs(w - mesh.cell_xy[..., 0]).max() < 1e-11
    # the mass matrix is the CG mass matrix: constants integrate to the area
    one = np.ones(P.cg.ndof)
    rows = np.bincount(P.cg.cellmap.ravel(), weights=(P.detJ[:, None] * (P.cg.W.T @ P.cg.W @ np.ones(P.cg.nloc))[None, :]).ravel(),
                       minlength=P.cg.ndof)
    assert abs(one @ rows - mesh.volume) < 1e-12
