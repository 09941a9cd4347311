"""Reference-element machinery: orthonormal bases and quadrature (numpy, FP64 / long double).

The engine and the oracle both work in *modal* bases that are orthonormal on the
reference entities:

* cells:  Dubiner basis of P_m on the reference triangle T^ = {(xi,eta): xi,eta>=0, xi+eta<=1},
          normalised such that  int_T^ phi_i phi_j = delta_ij.  On an affine cell K with
          Jacobian J this makes the mass matrix  |det J| * I  (SURVEY.md H6).
* facets: Legendre basis of P_k on [0,1] with int_0^1 l_m l_n ds = delta_mn.

The reference (Firedrake) uses nodal bases for DG_{k+1}, DG_k and DGT_k
(`hdg_imex.py:65-68`, `hdg_implicit.py:46-49`); nodal<->modal conversion matrices for
equispaced Lagrange nodes are provided by :func:`nodal_to_modal_cell` /
:func:`nodal_to_modal_facet` so that an adapter can marshal `Function.dat.data` arrays.

Everything here is dtype generic: pass ``np.longdouble`` arrays to get ~19 digit tables (used by
``tools/gen_tables.py`` to emit correctly rounded FP64 constants for the CUDA kernels).
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "ncell",
    "cell_indices",
    "gauss_legendre",
    "jacobi",
    "grad_jacobi",
    "dubiner",
    "dubiner_grad",
    "legendre01",
    "legendre01_deriv",
    "triangle_quadrature",
    "triangle_quadrature_gj",
    "gauss_jacobi10",
    "facet_points",
    "FACET_VERTS",
    "REF_VERTS",
    "lagrange_nodes_cell",
    "lagrange_nodes_facet",
    "nodal_to_modal_cell",
    "nodal_to_modal_facet",
]

#: local facet e of the reference triangle runs from local vertex (e+1)%3 to (e+2)%3, i.e. it is
#: the facet opposite vertex e (the FIAT/Firedrake numbering) traversed counter-clockwise.
FACET_VERTS = ((1, 2), (2, 0), (0, 1))
REF_VERTS = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])


def ncell(m: int) -> int:
    """dimension of P_m on a triangle"""
    return (m + 1) * (m + 2) // 2


def cell_indices(m: int):
    """(i, j) index pairs of the Dubiner basis of P_m, graded by total degree.

    The grading makes the basis hierarchical: the first ncell(k) functions of the P_{k+1} basis are
    the P_k basis.
    """
    return [(i, n - i) for n in range(m + 1) for i in range(n + 1)]


def gauss_legendre(n: int, dtype=np.float64):
    """n-point Gauss-Legendre rule on [0, 1] (nodes, weights), Newton-refined in `dtype`."""
    x0, _ = np.polynomial.legendre.leggauss(n)
    x = np.asarray(x0, dtype=dtype)
    one = dtype(1)
    for _ in range(4):
        # Legendre recurrence for P_n and P_n'
        p0 = np.ones_like(x)
        p1 = x.copy()
        for j in range(2, n + 1):
            p0, p1 = p1, ((2 * j - 1) * x * p1 - (j - 1) * p0) / dtype(j)
        if n == 1:
            p0, p1 = np.ones_like(x), x.copy()
        dp = n * (x * p1 - p0) / (x * x - one)
        x = x - p1 / dp
    p0 = np.ones_like(x)
    p1 = x.copy()
    for j in range(2, n + 1):
        p0, p1 = p1, ((2 * j - 1) * x * p1 - (j - 1) * p0) / dtype(j)
    dp = n * (x * p1 - p0) / (x * x - one)
    w = 2 / ((one - x * x) * dp * dp)
    return (x + one) / 2, w / 2


def jacobi(x, alpha: int, n: int):
    """Jacobi polynomial P_n^{(alpha,0)} on [-1,1], orthonormal w.r.t. (1-x)^alpha.

    Three-term recurrence for the orthonormal polynomials (beta = 0).
    """
    x = np.asarray(x)
    dt = x.dtype.type
    a = dt(alpha)
    # gamma0 = 2^(a+1)/(a+1) * Gamma(a+1)Gamma(1)/Gamma(a+1) = 2^(a+1)/(a+1)
    gamma0 = dt(2) ** (a + 1) / (a + 1)
    p_prev = np.full_like(x, 1 / np.sqrt(gamma0))
    if n == 0:
        return p_prev
    gamma1 = (a + 1) / (a + 3) * gamma0
    p = ((a + 2) * x / 2 + a / 2) / np.sqrt(gamma1)
    if n == 1:
        return p
    aold = 2 / (2 + a) * np.sqrt((a + 1) / (a + 3))
    for i in range(1, n):
        i_ = dt(i)
        h1 = 2 * i_ + a
        anew = 2 / (h1 + 2) * np.sqrt((i_ + 1) * (i_ + 1 + a) * (i_ + 1 + a) * (i_ + 1) / (h1 + 1) / (h1 + 3))
        bnew = -(a * a) / h1 / (h1 + 2)
        p_prev, p = p, (-aold * p_prev + (x - bnew) * p) / anew
        aold = anew
    return p


def grad_jacobi(x, alpha: int, n: int):
    """derivative of :func:`jacobi`"""
    x = np.asarray(x)
    if n == 0:
        return np.zeros_like(x)
    dt = x.dtype.type
    return np.sqrt(dt(n) * (n + alpha + 1)) * _jacobi_ab(x, alpha + 1, 1, n - 1)


def _jacobi_ab(x, alpha: int, beta: int, n: int):
    """orthonormal Jacobi polynomial with general integer (alpha, beta) (needed for derivatives)"""
    x = np.asarray(x)
    dt = x.dtype.type
    a, b = dt(alpha), dt(beta)

    def gam(z):
        return _gamma_int(z, dt)

    gamma0 = dt(2) ** (a + b + 1) / (a + b + 1) * gam(alpha + 1) * gam(beta + 1) / gam(alpha + beta + 1)
    p_prev = np.full_like(x, 1 / np.sqrt(gamma0))
    if n == 0:
        return p_prev
    gamma1 = (a + 1) * (b + 1) / (a + b + 3) * gamma0
    p = ((a + b + 2) * x / 2 + (a - b) / 2) / np.sqrt(gamma1)
    if n == 1:
        return p
    aold = 2 / (2 + a + b) * np.sqrt((a + 1) * (b + 1) / (a + b + 3))
    for i in range(1, n):
        i_ = dt(i)
        h1 = 2 * i_ + a + b
        anew = 2 / (h1 + 2) * np.sqrt((i_ + 1) * (i_ + 1 + a + b) * (i_ + 1 + a) * (i_ + 1 + b) / (h1 + 1) / (h1 + 3))
        bnew = -(a * a - b * b) / h1 / (h1 + 2)
        p_prev, p = p, (-aold * p_prev + (x - bnew) * p) / anew
        aold = anew
    return p


def _gamma_int(z: int, dt):
    """Gamma(z) for positive integer z, exactly representable for the small z used here"""
    r = dt(1)
    for i in range(2, int(z)):
        r = r * i
    return r


def _collapse(pts):
    """reference triangle (xi, eta) -> collapsed coordinates (a, b) on [-1,1]^2"""
    pts = np.asarray(pts)
    r = 2 * pts[..., 0] - 1
    s = 2 * pts[..., 1] - 1
    one = pts.dtype.type(1)
    denom = one - s
    safe = np.where(denom == 0, one, denom)
    a = np.where(denom == 0, -one, 2 * (one + r) / safe - one)
    return a, s


def dubiner(m: int, pts):
    """orthonormal Dubiner basis of P_m at points `pts` [..., 2] of T^.  Returns [ncell(m), ...]."""
    pts = np.asarray(pts)
    dt = pts.dtype.type
    a, b = _collapse(pts)
    out = []
    for i, j in cell_indices(m):
        h1 = jacobi(a, 0, i)
        h2 = jacobi(b, 2 * i + 1, j)
        # sqrt(2)*h1*h2*(1-b)^i is orthonormal on the biunit triangle (area 2); T^ has area 1/2
        out.append(2 * np.sqrt(dt(2)) * h1 * h2 * (1 - b) ** i)
    return np.array(out)


def dubiner_grad(m: int, pts):
    """gradients (w.r.t. xi, eta) of the Dubiner basis.  Returns [ncell(m), ..., 2]."""
    pts = np.asarray(pts)
    dt = pts.dtype.type
    a, b = _collapse(pts)
    half = dt(1) / 2
    out = []
    for i, j in cell_indices(m):
        fa = jacobi(a, 0, i)
        dfa = grad_jacobi(a, 0, i)
        gb = jacobi(b, 2 * i + 1, j)
        dgb = grad_jacobi(b, 2 * i + 1, j)
        dr = dfa * gb
        ds = dfa * (gb * (half * (1 + a)))
        if i > 0:
            pw = (half * (1 - b)) ** (i - 1)
            dr = dr * pw
            ds = ds * pw
        tmp = dgb * (half * (1 - b)) ** i
        if i > 0:
            tmp = tmp - half * i * gb * (half * (1 - b)) ** (i - 1)
        ds = ds + fa * tmp
        scale = dt(2) ** (i + half)
        # d/dxi = 2 d/dr ; extra factor 2 for the area normalisation
        out.append(np.stack([4 * scale * dr, 4 * scale * ds], axis=-1))
    return np.array(out)


def legendre01(k: int, s):
    """orthonormal Legendre basis of P_k on [0,1] at points s.  Returns [k+1, ...]."""
    s = np.asarray(s)
    dt = s.dtype.type
    x = 2 * s - 1
    out = []
    p0 = np.ones_like(x)
    p1 = x.copy()
    for n in range(k + 1):
        if n == 0:
            p = p0
        elif n == 1:
            p = p1
        else:
            p0, p1 = p1, ((2 * n - 1) * x * p1 - (n - 1) * p0) / dt(n)
            p = p1
        out.append(np.sqrt(dt(2 * n + 1)) * p)
    return np.array(out)


def legendre01_deriv(k: int, s):
    """d/ds of :func:`legendre01`"""
    s = np.asarray(s)
    dt = s.dtype.type
    x = 2 * s - 1
    P = [np.ones_like(x), x.copy()]
    dP = [np.zeros_like(x), np.ones_like(x)]
    for n in range(2, k + 1):
        P.append(((2 * n - 1) * x * P[n - 1] - (n - 1) * P[n - 2]) / dt(n))
        dP.append(dP[n - 2] + (2 * n - 1) * P[n - 1])
    return np.array([2 * np.sqrt(dt(2 * n + 1)) * dP[n] for n in range(k + 1)])


def triangle_quadrature(degree: int, dtype=np.float64):
    """collapsed Gauss-Legendre rule on T^, exact for polynomials of total degree <= `degree`.

    Returns (points [nq,2], weights [nq]) with sum(weights) = 1/2.
    """
    n = degree // 2 + 2  # the (1-v) Jacobian raises the degree in v by one
    x, wx = gauss_legendre(n, dtype)
    u, v = np.meshgrid(x, x, indexing="ij")
    wu, wv = np.meshgrid(wx, wx, indexing="ij")
    # Duffy map: xi = u (1-v), eta = v
    pts = np.stack([(u * (1 - v)).ravel(), v.ravel()], axis=-1)
    w = (wu * wv * (1 - v)).ravel()
    return pts, w


def gauss_jacobi10(n: int, dtype=np.float64):
    """n-point Gauss-Jacobi rule for the weight (1-v) on [0,1] (nodes, weights), Newton-refined."""
    from scipy.special import roots_jacobi

    x0, _ = roots_jacobi(n, 1.0, 0.0)
    x = np.asarray(x0, dtype=dtype)
    for _ in range(4):
        # orthonormal Jacobi P_n^{(1,0)} and its derivative
        x = x - jacobi(x, 1, n) / grad_jacobi(x, 1, n)
    # weights from the Christoffel function: w_i = 1 / sum_{j<n} p_j(x_i)^2  (orthonormal p_j)
    den = sum(jacobi(x, 1, j) ** 2 for j in range(n))
    w = 1 / den
    # map [-1,1] with weight (1-x) to [0,1] with weight (1-v): x = 2v-1, (1-x) = 2(1-v), dx = 2 dv
    return (x + 1) / 2, w / 4


def triangle_quadrature_gj(degree: int, dtype=np.float64):
    """collapsed Gauss-Legendre x Gauss-Jacobi rule, n^2 points with n = ceil((degree+1)/2)"""
    n = degree // 2 + 1
    xu, wu = gauss_legendre(n, dtype)
    xv, wv = gauss_jacobi10(n, dtype)
    u, v = np.meshgrid(xu, xv, indexing="ij")
    a, b = np.meshgrid(wu, wv, indexing="ij")
    pts = np.stack([(u * (1 - v)).ravel(), v.ravel()], axis=-1)
    return pts, (a * b).ravel()


def facet_points(e: int, s):
    """points of local facet e at parameters s in [0,1] (running from vertex (e+1)%3 to (e+2)%3)"""
    s = np.asarray(s)
    v0 = REF_VERTS[FACET_VERTS[e][0]].astype(s.dtype)
    v1 = REF_VERTS[FACET_VERTS[e][1]].astype(s.dtype)
    return v0[None, :] * (1 - s)[:, None] + v1[None, :] * s[:, None]


def lagrange_nodes_cell(m: int):
    """equispaced Lagrange nodes of P_m on T^ (interior-shifted for m = 0)"""
    if m == 0:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]])
    return np.array([[i / m, j / m] for j in range(m + 1) for i in range(m + 1 - j)])


def lagrange_nodes_facet(k: int):
    """equispaced nodes of P_k on [0,1]"""
    if k == 0:
        return np.array([0.5])
    return np.linspace(0.0, 1.0, k + 1)


def nodal_to_modal_cell(m: int, nodes=None):
    """matrix V^{-1} with modal = V^{-1} nodal, where V[n,i] = phi_i(node_n)"""
    nodes = lagrange_nodes_cell(m) if nodes is None else nodes
    V = dubiner(m, nodes).T
    return np.linalg.inv(V)


def nodal_to_modal_facet(k: int, nodes=None):
    """matrix V^{-1} with modal = V^{-1} nodal on a facet"""
    nodes = lagrange_nodes_facet(k) if nodes is None else nodes
    V = legendre01(k, nodes).T
    return np.linalg.inv(V)


# ---- symmetric (Dunavant-type) rules: fewer points than the collapsed tensor rules ---------------------
_SYM_SEEDS = {
    # degree: (centroid weight or None, [(a, w) S21 orbits], [(a, b, w) S111 orbits]); weights sum to 1
    5: (0.225, [(0.470142064105115, 0.132394152788506), (0.101286507323456, 0.125939180544827)], []),
    8: (0.144315607677787,
        [(0.459292588292723, 0.095091634267285), (0.170569307751760, 0.103217370534718),
         (0.050547228317031, 0.032458497623198)],
        [(0.008394777409958, 0.263112829634638, 0.027230314174435)]),
}


def _sym_expand(params, n21, n111, has_c):
    pts, wts = [], []
    i = 0
    if has_c:
        pts.append((1 / 3, 1 / 3))
        wts.append(params[0])
        i = 1
    for _ in range(n21):
        a, w = params[i], params[i + 1]
        i += 2
        b = 1 - 2 * a
        pts += [(a, a), (a, b), (b, a)]
        wts += [w] * 3
    for _ in range(n111):
        a, b, w = params[i], params[i + 1], params[i + 2]
        i += 3
        c = 1 - a - b
        pts += [(a, b), (b, a), (a, c), (c, a), (b, c), (c, b)]
        wts += [w] * 6
    return np.array(pts), np.array(wts)


_SYM_CACHE = {}


def triangle_quadrature_sym(degree: int, dtype=np.float64):
    """fully symmetric rule on T^ exact for total degree <= `degree` (7 points for 5, 16 for 8), Newton-
    refined to round-off from the classical Dunavant parameters by enforcing orthogonality of the
    Dubiner modes against the constant.  Returns None when no rule is tabulated for `degree`.
    Weights sum to 1/2."""
    key = min((d for d in _SYM_SEEDS if d >= degree), default=None)
    if key is None:
        return None
    if key not in _SYM_CACHE:
        from scipy.optimize import least_squares

        cw, s21, s111 = _SYM_SEEDS[key]
        p0 = ([cw] if cw is not None else []) + [v for o in s21 for v in o] + [v for o in s111 for v in o]
        target = np.zeros(ncell(key))
        target[0] = 1.0 / dubiner(0, np.array([[1 / 3, 1 / 3]]))[0, 0]  # int phi_0 = 1/phi_0 (orthonormal)

        def resid(p):
            x, w = _sym_expand(p, len(s21), len(s111), cw is not None)
            return dubiner(key, x) @ (0.5 * w) - target

        sol = least_squares(resid, np.array(p0), xtol=3e-16, ftol=None, gtol=None, max_nfev=200)
        assert np.abs(resid(sol.x)).max() < 5e-16, np.abs(resid(sol.x)).max()
        x, w = _sym_expand(sol.x, len(s21), len(s111), cw is not None)
        assert np.all(w > 0) and np.all(x > 0) and np.all(x.sum(axis=1) < 1)
        _SYM_CACHE[key] = (x, 0.5 * w)
    x, w = _SYM_CACHE[key]
    return x.astype(dtype), w.astype(dtype)
