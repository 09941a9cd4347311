"""Minimal ParaView output: the stand-in for ``firedrake.output.VTKFile`` that the reference driver
and its ``AnimationCallback`` write to (`driver.py:384-385`, `auxilliary/callbacks.py:40,85`).

A ``.pvd`` collection points at one ASCII ``.vtu`` file per ``write`` call.  Every triangle is written
with its own three points (discontinuous), and every field is sampled at the cell vertices from its
modal coefficients -- the same "interpolate to discontinuous P1 for output" that Firedrake's writer
performs for DG functions.  Host-side and off the hot path: fields are downloaded once per write.
"""

from __future__ import annotations

import os

import numpy as np

from .. import refelem as R

__all__ = ["VTKFile"]


def _vertex_values(function):
    """values of a cell-wise field at the three vertices of every cell: [nc, 3] or [nc, 3, 2]"""
    space = function.function_space()
    coef = function.to_host()
    tab = R.dubiner(space.degree, R.REF_VERTS)  # [ndof, 3]
    if space.name == "Q":
        return np.einsum("nci,iv->nvc", coef, tab)
    return np.einsum("na,av->nv", coef, tab)


class VTKFile:
    def __init__(self, filename: str, mode: str = "w"):
        assert filename.endswith(".pvd"), "the collection file must be a .pvd"
        self.filename = filename
        self.base = filename[:-4]
        self.entries = []
        if mode == "w" and os.path.exists(filename):
            os.remove(filename)

    def _write_vtu(self, path, mesh, fields):
        nc = mesh.nc
        xy = mesh.cell_xy.reshape(nc * 3, 2)
        with open(path, "w") as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
            fh.write(f'<UnstructuredGrid>\n<Piece NumberOfPoints="{3 * nc}" NumberOfCells="{nc}">\n')
            fh.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
            np.savetxt(fh, np.column_stack([xy, np.zeros(3 * nc)]), fmt="%.17g")
            fh.write("</DataArray>\n</Points>\n<Cells>\n")
            fh.write('<DataArray type="Int32" Name="connectivity" format="ascii">\n')
            np.savetxt(fh, np.arange(3 * nc).reshape(nc, 3), fmt="%d")
            fh.write('</DataArray>\n<DataArray type="Int32" Name="offsets" format="ascii">\n')
            np.savetxt(fh, 3 * np.arange(1, nc + 1), fmt="%d")
            fh.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
            np.savetxt(fh, np.full(nc, 5), fmt="%d")  # VTK_TRIANGLE
            fh.write("</DataArray>\n</Cells>\n<PointData>\n")
            for name, vals in fields:
                if vals.ndim == 3:  # vectors are padded to three components
                    v = np.concatenate([vals.reshape(3 * nc, 2), np.zeros((3 * nc, 1))], axis=1)
                    fh.write(f'<DataArray type="Float64" Name="{name}" NumberOfComponents="3" format="ascii">\n')
                else:
                    v = vals.reshape(3 * nc, 1)
                    fh.write(f'<DataArray type="Float64" Name="{name}" NumberOfComponents="1" format="ascii">\n')
                np.savetxt(fh, v, fmt="%.17g")
                fh.write("</DataArray>\n")
            fh.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")

    def write(self, *functions, time=None):
        """``outfile.write(Q, p, ..., time=t)``; fields are named by ``Function.rename``.  Plain
        ``(name, ndarray[nc,3(,2)])`` pairs of vertex values are accepted too (vorticity)."""
        fields, mesh = [], None
        for n, f in enumerate(functions):
            if isinstance(f, tuple):
                fields.append(f)
                continue
            mesh = f.function_space().mesh() if mesh is None else mesh
            fields.append((f.name or f"function_{n}", _vertex_values(f)))
        assert mesh is not None, "at least one Function is needed to define the mesh"
        idx = len(self.entries)
        path = f"{self.base}_{idx}.vtu"
        self._write_vtu(path, mesh, fields)
        self.entries.append((float(idx) if time is None else float(time), os.path.basename(path)))
        with open(self.filename, "w") as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
            for t, name in self.entries:
                fh.write(f'<DataSet timestep="{t:.17g}" part="0" file="{name}"/>\n')
            fh.write("</Collection>\n</VTKFile>\n")
        return path
