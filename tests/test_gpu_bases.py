"""GPU parity: envelope / radial / 2-D Fourier-Bessel basis kernels vs the oracle and the
reference-generated golden vectors (fp64 evaluation of the reference formulas)."""
import math

import pytest
import torch

from oracle import bases as obases
from util import relerr

pytestmark = pytest.mark.gpu


def test_envelope(golden):
    from x2gnn_b200 import envelop
    b = golden("bases")
    d = b["d"].cuda()
    env = envelop.poly_envelop(5.0, 5)
    got = env(d)
    scale = float(b["env_f64"].abs().max())
    assert float((got.cpu().double() - b["env_f64"]).abs().max()) < 2e-6 * scale
    assert torch.allclose(envelop.poly_envelop_func(d), got)
    assert float(env(torch.tensor([2.5], device="cuda"))) == 1.7109375        # App. B KAT
    assert not list(env.state_dict().keys())


def test_radial_fwd_bwd(golden):
    from x2gnn_b200 import radial_basis_layer as rbl
    b = golden("bases")
    d = b["d"].cuda()
    layer = rbl.RadialBasis(6, 5.0).cuda()
    assert list(layer.state_dict().keys()) == ["frequencies"]
    out = layer(d)
    assert float((out.cpu().double() - b["rbf_f64"]).abs().max()) < 2e-6
    assert torch.allclose(rbl.RadialBasis_func(d, 5.0, 6), out)
    # backward to the trainable frequencies (and to d) vs autograd of the oracle in fp64
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(0))
    d_req = d.clone().requires_grad_(True)
    layer(d_req).backward(g.cuda())
    f64 = obases.radial_frequencies(6).double().requires_grad_(True)
    d64 = b["d"].double().requires_grad_(True)
    obases.radial_basis(d64, f64).backward(g.double())
    assert relerr(layer.frequencies.grad, f64.grad) < 1e-5
    assert relerr(d_req.grad, d64.grad) < 1e-5
    # deterministic two-stage reduction: bitwise identical across runs, larger n
    dd = (0.9 + 4.1 * torch.rand(50000, generator=torch.Generator().manual_seed(1))).cuda()
    gg = torch.randn(50000, 6, generator=torch.Generator().manual_seed(2)).cuda()
    grads = []
    for _ in range(2):
        layer.zero_grad()
        layer(dd).backward(gg)
        grads.append(layer.frequencies.grad.clone())
    assert torch.equal(grads[0], grads[1])
    f64 = obases.radial_frequencies(6).double().requires_grad_(True)
    obases.radial_basis(dd.cpu().double(), f64).backward(gg.cpu().double())
    assert relerr(grads[0], f64.grad) < 1e-5


@pytest.mark.parametrize("LR", [(7, 6), (3, 4)])
def test_f_b_2d_golden(golden, LR):
    from x2gnn_b200 import angular_basis_layer as abl
    b = golden("bases")
    L, R = LR
    layer = abl.F_B_2D(L, R, 5.0, 5)
    assert not list(layer.state_dict().keys())
    got = layer(b["d"].cuda(), b["angles"].cuda(), b["src"].cuda())
    ref = b[f"sbf_{L}_{R}_f64"]
    err = float((got.cpu().double() - ref).abs().max())
    # atol = 2e-5 * max|.| against the fp64 evaluation of the reference formulas
    assert err < 2e-5 * float(ref.abs().max()), err
    # and in fact we are ~fp32-rounding accurate, far better than the reference's own fp32 path
    ref32_err = float((b[f"sbf_{L}_{R}_f32"].double() - ref).abs().max())
    assert err < 1e-6 * float(ref.abs().max()) or err < ref32_err


def test_f_b_2d_small_distances_and_sizes():
    """x = z_ln d/c < l exercises the power-series branch; T not a multiple of the tile."""
    from x2gnn_b200 import angular_basis_layer as abl
    gen = torch.Generator().manual_seed(3)
    d = torch.cat([torch.tensor([0.05, 0.2, 0.5, 0.94]), 0.9 + 4.1 * torch.rand(200, generator=gen)])
    T = 1000 + 37
    ang = math.pi * torch.rand(T, generator=gen)
    ang[:3] = torch.tensor([0.0, math.pi, math.pi / 2])
    src = torch.randint(0, d.numel(), (T,), generator=gen)
    for L, R in ((7, 6), (7, 16), (16, 3)):
        got = abl.F_B_2D(L, R, 5.0, 5)(d.cuda(), ang.cuda(), src.cuda())
        ref = obases.f_b_2d(d.double(), ang.double(), src, L, R)
        assert relerr(got, ref) < 1e-6, (L, R)
    empty = abl.F_B_2D(7, 6, 5.0, 5)(d.cuda(), ang[:0].cuda(), src[:0].cuda())
    assert empty.shape == (0, 42)



@pytest.mark.parametrize("LR", [(7, 6), (3, 4)])
def test_geometry_gradients_golden(golden, LR):
    """x2_sbf_bwd / x2_envelope_bwd / x2_angular_bwd against autograd through the reference's own expressions in
    fp64 (golden), 1e-5 relative; deterministic (the by-bond sum has a fixed order)."""
    from x2gnn_b200 import angular_basis_layer as abl, envelop
    b = golden("bases")
    L, R = LR
    layer = abl.F_B_2D(L, R, 5.0, 5)
    go = b[f"sbf_{L}_{R}_go"].float().cuda()
    res = []
    for _ in range(2):
        d = b["d"].cuda().requires_grad_(True)
        a = b["angles"].cuda().requires_grad_(True)
        out = layer(d, a, b["src"].cuda())
        assert getattr(out, "_x2_factors", None) is None        # the conv must read (and differentiate) the tensor
        res.append(torch.autograd.grad(out, (d, a), go))
    assert relerr(out, b[f"sbf_{L}_{R}_f64"]) < 1e-6
    assert relerr(res[0][0], b[f"sbf_{L}_{R}_gd_f64"]) < 1e-5
    assert relerr(res[0][1], b[f"sbf_{L}_{R}_gang_f64"]) < 1e-5
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    # only one of the two requested
    d = b["d"].cuda().requires_grad_(True)
    (gd,) = torch.autograd.grad(layer(d, b["angles"].cuda(), b["src"].cuda()), d, go)
    assert torch.equal(gd, res[0][0])
    a = b["angles"].cuda().requires_grad_(True)
    (ga,) = torch.autograd.grad(layer(b["d"].cuda(), a, b["src"].cuda()), a, go)
    assert torch.equal(ga, res[0][1])
    if LR == (7, 6):
        d = b["d"].cuda().requires_grad_(True)
        (ge,) = torch.autograd.grad(envelop.poly_envelop(5.0, 5)(d), d, b["env_go"].float().cuda())
        assert relerr(ge, b["env_gd_f64"]) < 1e-5
        a = b["angles"].cuda().requires_grad_(True)
        (gc,) = torch.autograd.grad(abl.AngularBasisLayer(7)(a), a, b["cbf_7_go"].float().cuda())
        assert relerr(gc, b["cbf_7_gang_f64"]) < 1e-5


def test_geometry_gradients_vs_oracle_sizes():
    """Small distances (power-series branch of j_l and j_l'), the poles of Y_l0, bonds with no triplet, class-default
    and wide bases, against the fp64 oracle."""
    from x2gnn_b200 import angular_basis_layer as abl
    gen = torch.Generator().manual_seed(5)
    d0 = torch.cat([torch.tensor([0.05, 0.2, 0.5, 0.94]), 0.9 + 4.1 * torch.rand(300, generator=gen)])
    T = 2000 + 13
    ang0 = math.pi * torch.rand(T, generator=gen)
    ang0[:3] = torch.tensor([0.0, math.pi, math.pi / 2])
    src = torch.randint(0, d0.numel() - 20, (T,), generator=gen)       # the last 20 bonds are no triplet's source
    for L, R in ((7, 6), (7, 16), (16, 3), (8, 32)):
        go = torch.randn(T, L * R, generator=gen)
        d = d0.cuda().requires_grad_(True)
        a = ang0.cuda().requires_grad_(True)
        gd, ga = torch.autograd.grad(abl.F_B_2D(L, R, 5.0, 5)(d, a, src.cuda()), (d, a), go.cuda())
        d64 = d0.double().requires_grad_(True)
        a64 = ang0.double().requires_grad_(True)
        rd, ra = torch.autograd.grad(obases.f_b_2d(d64, a64, src, L, R), (d64, a64), go.double())
        # the four small-d bonds have gradients 1e3..1e6 times the others: each on its own scale, the rest together
        gdc = gd.cpu().double()
        assert relerr(gdc[4:], rd[4:]) < 1e-5, (L, R)
        for i in range(4):
            assert abs(float(gdc[i] - rd[i])) < 1e-4 * abs(float(rd[i])), (L, R, i)
        assert relerr(ga, ra) < 1e-5, (L, R)
        assert float(gd[-20:].abs().max()) == 0.0
    with pytest.raises(Exception):
        d = d0.cuda().requires_grad_(True)
        torch.autograd.grad(abl.F_B_2D(16, 32, 5.0, 5)(d, ang0.cuda(), src.cuda()).sum(), d)   # L R = 512 > 256


def test_dimenet_radialbasis_twin():
    """radial_basis_layer.py:6-17 `radialbasis` (unused by the model, kept importable): sqrt(2/c) sin(n pi r / c) / r
    for one distance and for a column of distances, against the formula in fp64."""
    import numpy as np
    from x2gnn_b200 import radial_basis_layer as rbl
    for r in (torch.tensor([1.3]), torch.linspace(0.8, 4.9, 37).unsqueeze(1)):
        got = rbl.radialbasis(r.cuda(), 5.0, 6)
        n = torch.arange(1, 7, dtype=torch.float64).unsqueeze(0)
        rr = r.double().reshape(-1, 1)
        want = (2 / 5.0) ** 0.5 * torch.sin(rr * n * np.pi / 5.0) / rr
        assert got.shape == want.shape
        assert float((got.cpu().double() - want).abs().max()) < 2e-6


def test_angular_basis(golden):
    from x2gnn_b200 import angular_basis_layer as abl
    b = golden("bases")
    got = abl.AngularBasisLayer(7)(b["angles"].cuda())
    assert float((got.cpu().double() - b["cbf_7_f64"]).abs().max()) < 5e-7
    assert torch.allclose(abl.AngularBasisLayer_func(b["angles"].cuda(), 7), got)


def test_rotation_translation_invariance():
    """sbf / rbf depend on positions only through distances and angles."""
    from x2gnn_b200 import angular_basis_layer as abl, synth
    b = synth.qm9_batch(2, seed=9)
    pos = torch.from_numpy(b["atom_pos"]).double()
    ei = torch.from_numpy(b["edge_index"])
    tri, aj, ai, ak = (torch.from_numpy(a) for a in synth.triplets_host(b["edge_index"], len(b["x"])))
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)))
    outs = []
    for p in (pos, pos @ q.T + torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)):
        d = (p[ei[0]] - p[ei[1]]).norm(dim=1)
        ji, jk = p[ai] - p[aj], p[ak] - p[aj]
        ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
        outs.append(abl.F_B_2D(7, 6, 5.0, 5)(d.float().cuda(), ang.float().cuda(), tri[0].cuda()))
    assert relerr(outs[1], outs[0]) < 2e-5


def test_fused_geometry_kernels_match_the_torch_expressions():
    """x2_bond_lengths / x2_triplet_angles (xgnn.py:46,60-66 as two kernels) against the torch expressions in fp32
    (same operation order: equal to the last bits) and fp64, on a QM9-shaped batch and on degenerate triplets
    (collinear, zero-length)."""
    from x2gnn_b200 import geometry, synth
    b = synth.qm9_batch(6, seed=4)
    pos = torch.from_numpy(b["atom_pos"]).cuda()
    ei = torch.from_numpy(b["edge_index"]).cuda()
    tri, aj, ai, ak = (torch.from_numpy(a).cuda().long() for a in synth.triplets_host(b["edge_index"], len(b["x"])))
    d = geometry.bond_lengths(pos, ei[0].contiguous(), ei[1].contiguous())
    d32 = torch.norm(pos[ei[0]] - pos[ei[1]], dim=1)
    d64 = torch.norm(pos.double()[ei[0]] - pos.double()[ei[1]], dim=1)
    assert float((d - d32).abs().max()) <= 2.4e-7 * float(d32.max()) and relerr(d, d64) < 2e-7
    ang = geometry.triplet_angles(pos, ai, aj, ak)
    ji, jk = pos[ai] - pos[aj], pos[ak] - pos[aj]
    a32 = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    p64 = pos.double()
    ji, jk = p64[ai] - p64[aj], p64[ak] - p64[aj]
    a64 = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    assert ang.shape == a32.shape and float((ang - a32).abs().max()) < 1e-6
    assert float((ang.double() - a64).abs().max()) < 2e-6
    # degenerate: collinear (0 and pi) and a zero vector (atan2(0, 0) = 0)
    p = torch.tensor([[0., 0, 0], [1, 0, 0], [2, 0, 0], [-1, 0, 0]], device="cuda")
    i = torch.tensor([1, 1, 0], device="cuda"); j = torch.tensor([0, 0, 0], device="cuda"); k = torch.tensor([2, 3, 1], device="cuda")
    got = geometry.triplet_angles(p, i, j, k)
    assert torch.allclose(got, torch.tensor([0.0, math.pi, 0.0], device="cuda"), atol=1e-7)
    # positions that require grad keep the differentiable torch path
    pg = pos.clone().requires_grad_(True)
    geometry.triplet_angles(pg, ai, aj, ak).sum().backward()
    assert pg.grad is not None and geometry.bond_lengths(pos[:0], ei[0][:0].contiguous(), ei[1][:0].contiguous()).numel() == 0
