"""CPU check of the kernels' per-Gaussian arithmetic (mobgs_b200/csrc/gs_math.cuh compiled for
the host by tests/host_math/harness.cpp) against the oracle and its autograd."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gsplat_ref as G
from oracle import mobgs_ref as M


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _scene(n, seed, W=160, H=96):
    g = torch.Generator().manual_seed(seed)
    z = 0.5 + 9 * torch.rand(n, generator=g)
    means = torch.stack([(torch.rand(n, generator=g) * 2 - 1) * 1.4 * z, (torch.rand(n, generator=g) * 2 - 1) * 0.9 * z, z], -1)
    quats = torch.randn(n, 4, generator=g)
    scales = torch.exp(torch.log(0.01 * z)[:, None] + 1.5 * torch.rand(n, 3, generator=g))
    ang = 0.1
    view = torch.eye(4)
    view[0, 0], view[0, 2], view[2, 0], view[2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
    view[:3, 3] = torch.tensor([0.05, -0.02, 0.3])
    Kmat = torch.tensor([[0.9 * W, 0, W / 2 + 3], [0, 0.8 * W, H / 2 - 2], [0, 0, 1]])
    return means, quats, scales, view, Kmat, W, H


def _run_fwd(lib, means, quats, scales, view, Kmat, W, H):
    n = means.shape[0]
    out = np.zeros((n, 7), np.float32)
    lib.hm_project_fwd(n, _fp(means.numpy()), _fp(quats.numpy()), _fp(scales.numpy()), _fp(view.numpy().copy()),
                       _fp(Kmat.numpy().copy()), W, H, C.c_float(0.3), C.c_float(0.01), C.c_float(1e10),
                       C.c_float(0.0), _fp(out))
    return out


@pytest.mark.parametrize("seed", [0, 1])
def test_projection_forward_matches_oracle(host_math, seed):
    means, quats, scales, view, Kmat, W, H = _scene(4000, seed)
    out = _run_fwd(host_math, means, quats, scales, view, Kmat, W, H)
    radii, m2d, dep, con, _ = G.fully_fused_projection(means, None, quats, scales, view[None], Kmat[None], W, H)
    radii = radii[0].numpy()
    # radius = ceil(3 sqrt(..)) can flip by one at exact integers; culling decisions must agree otherwise
    same = (out[:, 6] > 0) == (radii > 0)
    assert same.mean() > 0.999
    vis = (out[:, 6] > 0) & (radii > 0)
    assert vis.sum() > 1000
    assert np.abs(out[vis, 6] - radii[vis]).max() <= 1
    np.testing.assert_allclose(out[vis, 0:2], m2d[0].numpy()[vis], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(out[vis, 2], dep[0].numpy()[vis], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out[vis, 3:6], con[0].numpy()[vis], rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("seed", [0, 3])
def test_projection_vjp_matches_autograd(host_math, seed):
    means, quats, scales, view, Kmat, W, H = _scene(3000, seed)
    n = means.shape[0]
    g = torch.Generator().manual_seed(100 + seed)
    v_in = torch.randn(n, 6, generator=g)
    md, qd, sd, vd = (t.double().requires_grad_(True) for t in (means, quats, scales, view))
    radii, m2d, dep, con, _ = G.fully_fused_projection(md, None, qd, sd, vd[None], Kmat.double()[None], W, H)
    vis = radii[0] > 0
    vin_d = v_in.double() * vis[:, None]
    loss = (m2d[0] * vin_d[:, 0:2]).sum() + (dep[0] * vin_d[:, 2]).sum() + (con[0] * vin_d[:, 3:6]).sum()
    loss.backward()

    v_out = np.zeros((n, 10), np.float32)
    v_view = np.zeros(12, np.float32)
    host_math.hm_project_bwd(n, _fp(means.numpy()), _fp(quats.numpy()), _fp(scales.numpy()),
                             _fp(view.numpy().copy()), _fp(Kmat.numpy().copy()), W, H, C.c_float(0.3),
                             C.c_float(0.01), C.c_float(1e10), C.c_float(0.0), _fp(v_in.numpy().copy()),
                             _fp(v_out), _fp(v_view))
    fwd = _run_fwd(host_math, means, quats, scales, view, Kmat, W, H)
    agree = torch.from_numpy(fwd[:, 6] > 0) == vis
    m = agree.numpy()

    def close(a, b, name):
        a, b = a[m], b[m]
        scale = np.abs(b).max() + 1e-12
        err = np.abs(a - b).max() / scale
        assert err < 2e-3, (name, err)

    close(v_out[:, 0:3], md.grad.numpy(), "v_means")
    close(v_out[:, 3:7], qd.grad.numpy(), "v_quats")
    close(v_out[:, 7:10], sd.grad.numpy(), "v_scales")
    if m.all():
        gv = vd.grad.numpy()
        ref = np.concatenate([gv[:3, :3].reshape(-1), gv[:3, 3]])
        assert np.abs(v_view - ref).max() / (np.abs(ref).max() + 1e-12) < 2e-3


def test_hermite_taps_match_reference_formula(host_math):
    g = torch.Generator().manual_seed(7)
    P = 12
    ctrl = torch.randn(64, P, 3, generator=g).double()
    out = np.zeros(8, np.float32)
    for n in range(2, P + 1):
        for t in [0.0, 1e-6, 0.1, 0.37, 0.5, 0.77, 0.999, 1.0, 1.08, -0.05]:
            nn_ = torch.full((64, 1), n, dtype=torch.int64)
            ref = M.hermite_spline(ctrl, torch.tensor(t, dtype=torch.float64), nn_)
            host_math.hm_hermite_taps(C.c_float(t), n, _fp(out))
            idx = out[:4].astype(int)
            w = torch.from_numpy(out[4:].astype(np.float64))
            got = sum(w[i] * ctrl[:, idx[i], :] for i in range(4))
            assert torch.allclose(got, ref, atol=2e-5), (n, t)


@pytest.mark.parametrize("lanes", [32, 16, 8])
def test_unit_mask_never_culls_a_contributing_pixel(host_math, lanes):
    """blend_units.cuh unit_mask: every pixel the blend loop would accept (alpha >= 1/255) lies in a
    unit whose bit is set; and the mask is tight (few units flagged that hold no such pixel)."""
    rng = np.random.default_rng(lanes)
    n = 40000
    # centres in and around the tile, 1-sigma sizes 0.3 .. 12 px, any orientation, opacities down to ~1/255
    mx, my = rng.uniform(-12, 28, n), rng.uniform(-12, 28, n)
    s1, s2 = np.exp(rng.uniform(np.log(0.3), np.log(12), n)), np.exp(rng.uniform(np.log(0.3), np.log(12), n))
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    cov = np.stack([c * c * s1**2 + s * s * s2**2, c * s * (s1**2 - s2**2), s * s * s1**2 + c * c * s2**2], -1)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    conic = np.stack([cov[:, 2] / det, -cov[:, 1] / det, cov[:, 0] / det], -1)
    opac = np.where(rng.random(n) < 0.2, rng.uniform(0.003, 0.01, n), rng.uniform(0.01, 1.0, n))
    geom = np.ascontiguousarray(np.concatenate([mx[:, None], my[:, None], opac[:, None], conic], 1), np.float32)
    mask, need = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    up = C.POINTER(C.c_uint32)
    host_math.hm_unit_mask(lanes, n, _fp(geom), mask.ctypes.data_as(up), need.ctypes.data_as(up))
    assert ((need & ~mask) == 0).all(), "a contributing pixel was culled"
    pop = lambda a: sum(int(bin(int(v)).count("1")) for v in a)
    extra = pop(mask & ~need) / max(1, pop(need))
    assert pop(need) > 10000 and extra < 0.08, extra
