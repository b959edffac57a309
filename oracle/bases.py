"""oracle.bases -- TEST INFRASTRUCTURE.  CPU restatement of the reference basis
expansions: envelop.py:5-32, radial_basis_layer.py:19-40, angular_basis_layer.py:12-93
and the constants/formulas of basis_func.py:7-155.  dtype follows the input tensors, so
the same code gives the fp32 behaviour and the fp64 ground truth (SURVEY.md App. B:
the reference's fp32 closed-form j_l cancels catastrophically for l >= 4, so parity of
the sbf basis is defined against the fp64 evaluation)."""
import math
import numpy as np
import torch
from scipy import special as sp
from scipy.optimize import brentq


# ---------------------------------------------------------------- envelope
def envelope_coeffs(exponent: int):
    """envelop.py:8-13."""
    p = exponent + 1
    return p, -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2


def poly_envelop(d: torch.Tensor, cutoff: float = 5.0, exponent: int = 5) -> torch.Tensor:
    """envelop.py:16-21: 1/x + a x^(p-1) + b x^p + c x^(p+1), x = d/cutoff, no clamp."""
    p, a, b, c = envelope_coeffs(exponent)
    x = d * (1 / cutoff)
    return 1 / x + a * x ** (p - 1) + b * x ** p + c * x ** (p + 1)


# ---------------------------------------------------------------- radial basis
def radial_basis(d: torch.Tensor, frequencies: torch.Tensor, cutoff: float = 5.0):
    """radial_basis_layer.py:36-40: sin(freq_n * d / cutoff) -> [E,R]."""
    return torch.sin(frequencies * (d * (1 / cutoff)).unsqueeze(-1))


def radial_frequencies(R: int) -> torch.Tensor:
    """radial_basis_layer.py:32: pi * (1..R), float32."""
    return np.pi * torch.arange(1, R + 1, dtype=torch.float32)


# ---------------------------------------------------------------- Bessel tables
def _jn(r, n):
    """basis_func.py:7-11 (dtype of r is preserved: float32 in -> float32 out)."""
    return np.sqrt(np.pi / (2 * r)) * sp.jv(n + 0.5, r)


def bessel_zeros(n: int, k: int) -> np.ndarray:
    """basis_func.py:14-29: first k zeros of j_l, l < n, stored float32."""
    zerosj = np.zeros((n, k), dtype="float32")
    zerosj[0] = np.arange(1, k + 1) * np.pi
    points = np.arange(1, k + n) * np.pi
    racines = np.zeros(k + n - 1, dtype="float32")
    for i in range(1, n):
        for j in range(k + n - 1 - i):
            racines[j] = brentq(_jn, points[j], points[j + 1], (i,))
        points = racines
        zerosj[i][:k] = racines[:k]
    return zerosj


def bessel_normalizers(zeros: np.ndarray) -> np.ndarray:
    """basis_func.py:55-60: N_ln = 1/sqrt(0.5 * j_{l+1}(z_ln)^2), float32 arithmetic
    because `zeros` is float32."""
    n, k = zeros.shape
    out = []
    for order in range(n):
        tmp = [0.5 * _jn(zeros[order, i], order + 1) ** 2 for i in range(k)]
        out.append(1 / np.array(tmp) ** 0.5)
    return np.asarray(out)


class _SphericalJl(torch.autograd.Function):
    """j_l(x) through scipy in float64 with its derivative (scipy's spherical_jn(derivative=True)), so that the
    oracle is differentiable w.r.t. the distances exactly where the reference's lambdified torch closed forms are."""

    @staticmethod
    def forward(ctx, x, l):
        ctx.save_for_backward(x)
        ctx.l = l
        return torch.from_numpy(sp.spherical_jn(l, x.detach().double().numpy())).to(x.dtype)

    @staticmethod
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        dj = torch.from_numpy(sp.spherical_jn(ctx.l, x.detach().double().numpy(), derivative=True)).to(x.dtype)
        return go * dj, None


def spherical_jl(l: int, x: torch.Tensor) -> torch.Tensor:
    """j_l(x) through scipy in float64 -- mathematically identical to the closed forms
    of basis_func.py:32-44 (differentiable, see _SphericalJl)."""
    if x.requires_grad:
        return _SphericalJl.apply(x, l)
    return torch.from_numpy(sp.spherical_jn(l, x.detach().double().numpy()))


def sph_harm_prefactor(l: int) -> float:
    """basis_func.py:74-81 with m = 0."""
    return ((2 * l + 1) / (4 * np.pi)) ** 0.5


def y_l0(L: int, theta: torch.Tensor) -> torch.Tensor:
    """basis_func.py:84-155 (zero_m_only): Y_l0(theta) = prefactor * P_l(cos theta),
    Legendre via the same three-term recurrence (basis_func.py:94-96) -> [T,L]."""
    c = torch.cos(theta)
    P = [torch.ones_like(c)]
    if L > 1:
        P.append(c)
    for j in range(2, L):
        P.append(((2 * j - 1) * c * P[j - 1] - (j - 1) * P[j - 2]) / j)
    return torch.stack([sph_harm_prefactor(l) * P[l] for l in range(L)], dim=1)


_TABLE_CACHE = {}


def bessel_tables(L: int, R: int):
    key = (L, R)
    if key not in _TABLE_CACHE:
        z = bessel_zeros(L, R)
        _TABLE_CACHE[key] = (z, bessel_normalizers(z))
    return _TABLE_CACHE[key]


def f_b_2d(d: torch.Tensor, angles: torch.Tensor, edge_index_1: torch.Tensor,
           num_spherical: int, num_radial: int, cutoff: float = 5.0,
           envelope_exponent: int = 5) -> torch.Tensor:
    """angular_basis_layer.py:80-93: sbf[t, l*R+n] = env(d_s) N_ln j_l(z_ln d_s/c) Y_l0(th_t)
    with s = edge_index_1[t]; the envelope cutoff is hard-coded 5.0 (:60)."""
    L, R = num_spherical, num_radial
    z, norm = bessel_tables(L, R)
    x = d / cutoff
    cols = []
    for l in range(L):
        for n in range(R):
            arg = (float(z[l, n]) * x)
            cols.append(float(norm[l, n]) * spherical_jl(l, arg).to(d.dtype))
    rbf = torch.stack(cols, dim=1)
    rbf_env = poly_envelop(d, 5.0, envelope_exponent)[:, None] * rbf
    rbf_env = rbf_env[edge_index_1.long()]
    cbf = y_l0(L, angles).repeat_interleave(R, dim=1)
    return rbf_env * cbf


def angular_basis(angles: torch.Tensor, num_sph: int) -> torch.Tensor:
    """angular_basis_layer.py:12-32 (AngularBasisLayer.forward): [T, num_sph]."""
    return y_l0(num_sph, angles)
