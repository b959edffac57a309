"""Numeric constants of the 2-D Fourier-Bessel basis.  Mirrors the reference's basis_func.py
(:7-29 Jn / Jn_zeros, :55-60 normalisers, :74-81 sph_harm_prefactor) but produces the float32
TABLES the CUDA kernels consume instead of sympy expression trees.  The symbolic helpers
(`bessel_basis`, `real_sph_harm`, ...) are kept for API compatibility and build their
expressions lazily (sympy is only imported if they are called).
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
from scipy import special as _sp
from scipy.optimize import brentq as _brentq


def Jn(r, n):
    """Spherical Bessel function j_n(r) via J_{n+1/2} (dtype of `r` is preserved)."""
    return np.sqrt(np.pi / (2 * r)) * _sp.jv(n + 0.5, r)


def Jn_zeros(n: int, k: int) -> np.ndarray:
    """First k positive zeros of j_l for l < n, float32 [n,k] (the reference stores float32).
    Zeros of j_l interlace those of j_{l-1}, so each row brackets the next."""
    zeros = np.zeros((n, k), dtype="float32")
    zeros[0] = np.arange(1, k + 1) * np.pi
    brackets = np.arange(1, k + n) * np.pi
    roots = np.zeros(k + n - 1, dtype="float32")
    for l in range(1, n):
        for j in range(k + n - 1 - l):
            roots[j] = _brentq(Jn, brackets[j], brackets[j + 1], (l,))
        brackets = roots
        zeros[l, :k] = roots[:k]
    return zeros


def bessel_normalizers(zeros: np.ndarray) -> np.ndarray:
    """N_ln = 1/sqrt(0.5 j_{l+1}(z_ln)^2), evaluated in the dtype of `zeros` (float32)."""
    n, k = zeros.shape
    rows = []
    for l in range(n):
        rows.append(1 / np.array([0.5 * Jn(zeros[l, i], l + 1) ** 2 for i in range(k)]) ** 0.5)
    return np.asarray(rows, dtype="float32")


@lru_cache(maxsize=None)
def bessel_tables(num_spherical: int, num_radial: int):
    """(zeros, normalisers) as contiguous float32 [L,R] arrays."""
    z = Jn_zeros(num_spherical, num_radial)
    return np.ascontiguousarray(z), np.ascontiguousarray(bessel_normalizers(z))


def sph_harm_prefactor(l: int, m: int) -> float:
    return ((2 * l + 1) * math.factorial(l - abs(m)) / (4 * np.pi * math.factorial(l + abs(m)))) ** 0.5


# ---------------------------------------------------------------- symbolic compatibility layer
def spherical_bessel_formulas(n: int):
    """sympy expressions of j_l(x), l < n, by the upward recurrence."""
    import sympy as sym
    x = sym.symbols("x")
    f = [sym.sin(x) / x]
    if n > 1:
        f.append(sym.sin(x) / x ** 2 - sym.cos(x) / x)
    for l in range(1, n - 1):
        f.append(sym.simplify((2 * l + 1) / x * f[l] - f[l - 1]))
    return f[:n]


def bessel_basis(n: int, k: int):
    """sympy expressions N_ln * j_l(z_ln * x), [n][k]."""
    import sympy as sym
    zeros, norm = bessel_tables(n, k)
    f = spherical_bessel_formulas(n)
    x = sym.symbols("x")
    return [[sym.simplify(float(norm[l, i]) * f[l].subs(x, float(zeros[l, i]) * x)) for i in range(k)]
            for l in range(n)]


def associated_legendre_polynomials(l: int, zero_m_only: bool = True):
    import sympy as sym
    if not zero_m_only:
        raise NotImplementedError("only m = 0 is used by X2-GNN")
    z = sym.symbols("z")
    P = [[sym.Integer(1)]]
    if l > 1:
        P.append([z])
    for j in range(2, l):
        P.append([sym.simplify(((2 * j - 1) * z * P[j - 1][0] - (j - 1) * P[j - 2][0]) / j)])
    return P[:l]


def real_sph_harm(l: int, zero_m_only: bool = True, spherical_coordinates: bool = True):
    """sympy expressions of Y_l0 (as [[expr], ...]), in cos(theta) or z."""
    import sympy as sym
    P = associated_legendre_polynomials(l, zero_m_only)
    theta, z = sym.symbols("theta z")
    out = []
    for i in range(l):
        e = sph_harm_prefactor(i, 0) * P[i][0]
        if spherical_coordinates:
            e = sym.sympify(e).subs(z, sym.cos(theta))
        out.append([sym.simplify(e)])
    return out
