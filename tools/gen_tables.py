#!/usr/bin/env python
"""Generate the reference-element constant tables used by the CUDA kernels.

Writes ``incompressibleeulerhdg_b200/csrc/hdg_tables.inc`` (committed; regenerate with
``python tools/gen_tables.py``).  All integrals are evaluated in 80-bit long double with
Gauss rules that are exact for the polynomial integrands, then rounded to FP64; entries that are
zero by orthogonality are snapped to exactly 0.0 so that the fully unrolled kernels drop them at
compile time.

Tables, for pressure degree k (velocity P_{k+1}, pressure P_k, trace P_k per facet), all on the
reference triangle with the orthonormal Dubiner / Legendre bases of ``refelem.py``:

  D[d][a][i]      = int_T^ psi_a d_d phi_i                       (discrete divergence, B-block)
  E[e][m][i]      = int_0^1 phi_i(x_e(s)) l_m(s) ds              (normal-flux coupling, E-block)
  F[e][m][a]      = int_0^1 psi_a(x_e(s)) l_m(s) ds              (trace-pressure coupling)
  KK[0..2][a][b]  = D0 D0^T,  D0 D1^T + D1 D0^T,  D1 D1^T         (B M^-1 B^T building blocks)
  TT[e][a][b]     = F_e^T F_e                                    (tau-stabilisation on facet e)
  LL[e][d][m][a]  = E_e D_d^T                                    (E M^-1 B^T building blocks)
  NN[e][f][m][n]  = E_e E_f^T                                    (E M^-1 E^T building blocks)
  BF[e][j][i]     = int_0^1 phi_i(x_e(s)) l_j(s) ds, j <= k+1    (facet normal-moment functionals)
  GG[e][f][j][l]  = BF_e BF_f^T                                  (N M^-1 N^T building blocks)

plus tabulations at quadrature points for the advection operator f_impl (`hdg_imex.py:313-331`)
and the BDM projection (`common.py:91-108`):

  cell rule (exact to degree 3k+2):  WQ[q], PHI[q][i], DPHI[d][q][i]
  facet rule (NQF Gauss points):     WF[q], PHIF[e][q][i], LEG[m][q]  (Legendre of degree <= k+1)
  BDM:  interior-moment test functions NED[w][d][q] at the cell rule points
  VINV[i][j]      = modal <- nodal map of the equispaced P_{k+1} Lagrange nodes (CG projection, tracer path)
"""

from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200 import refelem as R  # noqa: E402

LD = np.longdouble


def snap(a, tol=1e-17):
    a = np.array(a, dtype=LD)
    scale = max(1.0, float(np.abs(a).max()))
    a[np.abs(a) < tol * scale * 100] = 0
    return a.astype(np.float64)


def fmt(a):
    a = np.asarray(a)
    if a.ndim == 0:
        v = float(a)
        return "0.0" if v == 0.0 else repr(v)
    return "{" + ",".join(fmt(x) for x in a) + "}"


def decl(name, a):
    a = np.asarray(a)
    dims = "".join(f"[{n}]" for n in a.shape)
    return f"__device__ constexpr double {name}{dims} = {fmt(a)};\n"


def accessor(ns, name, a):
    nd = np.asarray(a).ndim
    args = ", ".join(f"int i{j}" for j in range(nd))
    idx = "".join(f"[i{j}]" for j in range(nd))
    return f"  static __device__ __forceinline__ constexpr double {name}({args}) {{ return {ns}::{name}{idx}; }}\n"


def nedelec_ref(k, xq):
    """same space as oracle._nedelec1_basis but built here independently in long double"""
    nq = xq.shape[0]
    out = []
    if k == 0:
        return np.zeros((0, 2, nq), dtype=LD)
    ph = R.dubiner(k - 1, xq)
    for c in range(2):
        for i in range(ph.shape[0]):
            v = np.zeros((2, nq), dtype=LD)
            v[c] = ph[i]
            out.append(v)
    xi, eta = xq[:, 0], xq[:, 1]
    for a in range(k):
        q = xi ** a * eta ** (k - 1 - a)
        out.append(np.stack([-eta * q, xi * q]))
    return np.array(out)


def tables(k):
    nQ1, np_, nl1 = R.ncell(k + 1), R.ncell(k), k + 1
    xq, wq = R.triangle_quadrature(2 * k + 4, LD)
    phiQ = R.dubiner(k + 1, xq)
    dphiQ = R.dubiner_grad(k + 1, xq)
    phiP = R.dubiner(k, xq)
    D = np.einsum("q,aq,iqd->dai", wq, phiP, dphiQ)
    s, ws = R.gauss_legendre(k + 4, LD)
    ell = R.legendre01(k, s)
    E = np.array([np.einsum("q,mq,iq->mi", ws, ell, R.dubiner(k + 1, R.facet_points(e, s).astype(LD))) for e in range(3)])
    F = np.array([np.einsum("q,mq,aq->ma", ws, ell, R.dubiner(k, R.facet_points(e, s).astype(LD))) for e in range(3)])
    KK = np.array([D[0] @ D[0].T, D[0] @ D[1].T + D[1] @ D[0].T, D[1] @ D[1].T])
    TT = np.array([F[e].T @ F[e] for e in range(3)])
    LL = np.array([[E[e] @ D[d].T for d in range(2)] for e in range(3)])
    NN = np.array([[E[e] @ E[f].T for f in range(3)] for e in range(3)])
    out = dict(D=D, E=E, F=F, KK=KK, TT=TT, LL=LL, NN=NN)
    # ---- advection / BDM tabulations -------------------------------------------------------
    # cell rule exact for degree 3k+2 (w in P_{k+1}, Q* in P_{k+1}, grad Q in P_k)
    sym = R.triangle_quadrature_sym(3 * k + 2, LD)  # 7 / 16 points for k = 1 / 2 instead of 9 / 25
    xa, wa = sym if sym is not None else R.triangle_quadrature_gj(3 * k + 2, LD)
    out["WQ"] = wa
    out["PHI"] = R.dubiner(k + 1, xa).T  # [q][i]
    out["DPHI"] = np.moveaxis(R.dubiner_grad(k + 1, xa), [0, 1, 2], [2, 1, 0])  # [d][q][i]
    out["DPSI"] = np.moveaxis(R.dubiner_grad(k, xa), [0, 1, 2], [2, 1, 0])  # [d][q][a]
    nqf = (3 * k + 4 + 1) // 2
    sf, wf = R.gauss_legendre(nqf, LD)
    out["SF"] = sf
    out["WF"] = wf
    out["PHIF"] = np.array([R.dubiner(k + 1, R.facet_points(e, sf).astype(LD)).T for e in range(3)])  # [e][q][i]
    out["PSIF"] = np.array([R.dubiner(k, R.facet_points(e, sf).astype(LD)).T for e in range(3)])  # [e][q][a]
    out["LEG"] = R.legendre01(k + 1, sf)  # [m][q], m <= k+1
    out["DPHIF"] = np.array([np.moveaxis(R.dubiner_grad(k + 1, R.facet_points(e, sf).astype(LD)), [0, 1, 2], [2, 1, 0])
                             for e in range(3)])  # [e][d][q][i]
    # BDM facet functionals in modal form: BF[e][j][i] = int_0^1 phi_i(x_e(s)) l_j(s) ds, j <= k+1
    s2, w2 = R.gauss_legendre(k + 4, LD)
    leg2 = R.legendre01(k + 1, s2)
    out["BF"] = np.array(
        [np.einsum("q,jq,iq->ji", w2, leg2, R.dubiner(k + 1, R.facet_points(e, s2).astype(LD))) for e in range(3)])
    # Gram blocks of the facet normal-moment functionals (penalty Schur complement of the tentative
    # velocity solver): GG[e][f][j][l] = sum_i BF[e][j][i] BF[f][l][i]
    out["GG"] = np.array([[out["BF"][e] @ out["BF"][f].T for f in range(3)] for e in range(3)])
    # BDM interior functionals: BI[w][d][i] = int_T^ ned_w[d] phi_i
    xb, wb = R.triangle_quadrature(2 * k + 3, LD)
    ned = nedelec_ref(k, xb)
    out["BI"] = np.einsum("q,wdq,iq->wdi", wb, ned, R.dubiner(k + 1, xb)) if k > 0 else np.zeros((1, 2, nQ1))
    # LIFT = columns of N^-1 that belong to the facet functionals, N = all BDM functionals on the
    # Piola-pulled-back coefficients Qhat[c][i]:
    #   facet (e,j):  |e^| n^_e[c] BF[e][j][i]      with |e^| n^_e = (1,1), (-1,0), (0,-1)
    #   interior w :  BI[w][c][i]
    nref = np.array([[1, 1], [-1, 0], [0, -1]], dtype=LD)
    nfm = 3 * (k + 2)
    N = np.zeros((2 * nQ1, 2, nQ1), dtype=LD)
    for e in range(3):
        for j in range(k + 2):
            for c in range(2):
                N[e * (k + 2) + j, c, :] = nref[e, c] * out["BF"][e][j]
    if k > 0:
        N[nfm:, :, :] = out["BI"]
    Ninv = np.linalg.inv(N.reshape(2 * nQ1, 2 * nQ1).astype(np.float64))
    # one step of iterative refinement in long double
    Nl = N.reshape(2 * nQ1, 2 * nQ1)
    Ninv = Ninv.astype(LD)
    Ninv = Ninv + Ninv @ (np.eye(2 * nQ1, dtype=LD) - Nl @ Ninv)
    Ninv = Ninv + Ninv @ (np.eye(2 * nQ1, dtype=LD) - Nl @ Ninv)
    out["LIFT"] = Ninv[:, :nfm].reshape(2, nQ1, 3, k + 2)  # [c][i][e][j]
    # modal <- nodal map of the equispaced Lagrange nodes of P_{k+1} (CG velocity projection of the tracer
    # path, common.py:119-122): VINV = V^-1 with V[n][i] = phi_i(node_n); two refinement steps in long double
    nodes = R.lagrange_nodes_cell(k + 1).astype(LD)
    V = R.dubiner(k + 1, nodes).T
    Vi = np.linalg.inv(V.astype(np.float64)).astype(LD)
    Vi = Vi + Vi @ (np.eye(nQ1, dtype=LD) - V @ Vi)
    Vi = Vi + Vi @ (np.eye(nQ1, dtype=LD) - V @ Vi)
    out["VINV"] = Vi  # [i][j]
    return {n: snap(v) for n, v in out.items()}, dict(nQ1=nQ1, np_=np_, nl1=nl1, nq=len(wa), nqf=nqf,
                                                       nint=k * (k + 2))


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "..", "incompressibleeulerhdg_b200", "csrc", "hdg_tables.inc")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_tables.py -- do not edit.\n")
        f.write("// Reference-element constants (orthonormal Dubiner / Legendre bases), FP64 rounded from long double.\n")
        f.write("#pragma once\n\ntemplate <int K> struct RefTables;\n\n")
        for k in (1, 2, 3, 4):
            t, dims = tables(k)
            ns = f"hdg_tab_k{k}"
            f.write(f"namespace {ns} {{\n")
            for name, arr in t.items():
                f.write(decl(name, arr))
            f.write("}\n")
            f.write(f"template <> struct RefTables<{k}> {{\n")
            f.write(f"  static constexpr int NQ1 = {dims['nQ1']}, NP = {dims['np_']}, NL1 = {dims['nl1']};\n")
            f.write(f"  static constexpr int NQ = {dims['nq']}, NQF = {dims['nqf']}, NINT = {dims['nint']};\n")
            for name, arr in t.items():
                f.write(accessor(ns, name, arr))
            f.write("};\n\n")
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
