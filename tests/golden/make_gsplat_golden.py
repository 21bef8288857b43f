#!/usr/bin/env python
"""Pin the gsplat half of the oracle — run this ON A MACHINE WHERE `import gsplat` WORKS (gsplat==1.4.0, the
version the reference pins in README.md:26; a CUDA device is required by gsplat itself).

gsplat cannot be installed in the build image (no network, not in the wheelhouse), so
oracle/gsplat_ref.py is a restatement whose parity is UNPINNED.  This script closes that gap the day gsplat is
available: it calls the real `gsplat.rendering.rasterization` / `fully_fused_projection` with exactly the
keyword sets of the reference's seven call sites in render() (gaussian_renderer/__init__.py:143, :163, :190,
:201, :236, :255, :274) on a seeded scene and writes inputs, outputs and the gradients of a seeded weighted
loss to tests/golden/gsplat_callsites.npz.  tests/test_gsplat_golden.py then holds BOTH the oracle (CPU,
-m "not gpu") and the CUDA path (-m gpu) to those vectors; while the file is absent those tests skip and say so.

    python tests/golden/make_gsplat_golden.py            # writes tests/golden/gsplat_callsites.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

W, H, NS, ND = 160, 96, 900, 600


def scene(device):
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    from oracle import mobgs_ref as M
    stat, dyn, intr = synthetic_scene(NS, ND, W, H, seed=77, device=device, footprint_px=(0.6, 4.0))
    cam = make_camera(intr, subframe_w2c(1, 3, device=device), time=0.4)
    with torch.no_grad():
        t = torch.tensor(0.4 + 0.3 / 23, device=device)
        d = [x.detach().clone() for x in M.dynamic_attributes(dyn, t, clamp_time=True)]
        s = [x.detach().clone() for x in M.static_attributes(stat)]
        o = [x.detach().clone() for x in M.dynamic_attributes(dyn, torch.tensor(0.4, device=device), clamp_time=False)]
    return s, d, o, cam


def callsites(s, d, o, cam, rasterization, fully_fused_projection, dev):
    """-> {name: (fn, kwargs (tensors are leaves), differentiable kwarg names)} mirroring render()."""
    s_means, s_quats, s_scales, s_opac, s_cols = s
    d_means, d_quats, d_scales, d_opac, d_cols = d
    viewmat = cam.world_view_transform.transpose(0, 1).contiguous()
    Kmat = cam.K
    bg9 = torch.tensor([0.1, 0.2, 0.3] * 3, device=dev)
    means, quats = torch.cat([s_means, d_means]), torch.cat([s_quats, d_quats])
    scales, opac, cols = torch.cat([s_scales, d_scales]), torch.cat([s_opac, d_opac]), torch.cat([s_cols, d_cols])
    common = dict(viewmats=viewmat[None], Ks=Kmat[None], width=W, height=H, packed=False)
    ones = lambda n: torch.ones(n, 1, device=dev)  # noqa: E731
    with torch.no_grad():
        _, ori_m2d, _, _, _ = fully_fused_projection(means=torch.cat([s_means, o[0]]), covars=None,
                                                     quats=torch.cat([s_quats, o[1]]), scales=scales,
                                                     viewmats=viewmat[None], Ks=Kmat[None], width=W, height=H)
        _, m2d, _, _, _ = fully_fused_projection(means=means, covars=None, quats=quats, scales=scales,
                                                 viewmats=viewmat[None], Ks=Kmat[None], width=W, height=H)
        flow_2d = (ori_m2d - m2d).squeeze(0)
    return {
        "l143_dyn_rgbed": (rasterization, dict(means=d_means, quats=d_quats, scales=d_scales, opacities=d_opac.squeeze(-1),
                                              colors=d_cols, backgrounds=bg9[None], render_mode="RGB+ED", **common)),
        "l163_dyn_alpha": (rasterization, dict(means=d_means, quats=d_quats, scales=d_scales, opacities=d_opac.squeeze(-1),
                                              colors=ones(ND), backgrounds=bg9[0:1][None], render_mode="RGB", **common)),
        "l190_project": (fully_fused_projection, dict(means=means, covars=None, quats=quats, scales=scales,
                                                      viewmats=viewmat[None], Ks=Kmat[None], width=W, height=H)),
        "l201_all_rgbed": (rasterization, dict(means=means, quats=quats, scales=scales, opacities=opac.squeeze(-1),
                                              colors=cols, backgrounds=bg9[None], render_mode="RGB+ED", **common)),
        "l236_stat_rgbed": (rasterization, dict(means=s_means, quats=s_quats, scales=s_scales, opacities=s_opac.squeeze(-1),
                                               colors=s_cols, backgrounds=bg9[None], render_mode="RGB+ED", **common)),
        "l255_stat_alpha": (rasterization, dict(means=s_means, quats=s_quats, scales=s_scales, opacities=s_opac.squeeze(-1),
                                               colors=ones(NS), backgrounds=bg9[0:1][None], render_mode="RGB", **common)),
        "l274_flow": (rasterization, dict(means=means, quats=quats, scales=scales, opacities=opac.squeeze(-1),
                                          colors=flow_2d, backgrounds=None, render_mode="RGB", **common)),
    }


DIFF = ("means", "quats", "scales", "opacities", "colors", "viewmats")


def run_callsite(fn, kwargs, seed):
    """Calls fn with leaf copies of the differentiable kwargs, backpropagates a seeded weighted sum of the float
    outputs.  -> (inputs, outputs, gradients) as dicts of numpy arrays."""
    kw = dict(kwargs)
    leaves = {}
    for k in DIFF:
        if torch.is_tensor(kw.get(k)):
            leaves[k] = kw[k].detach().clone().requires_grad_(True)
            kw[k] = leaves[k]
    out = fn(**kw)
    if isinstance(out[-1], dict):                       # rasterization -> (colors, alphas, meta)
        outs = {"render_colors": out[0], "render_alphas": out[1], "radii": out[2]["radii"], "means2d": out[2]["means2d"]}
        float_outs = ("render_colors", "render_alphas")
    else:                                               # fully_fused_projection -> (radii, means2d, depths, conics, comp)
        outs = {"radii": out[0], "means2d": out[1], "depths": out[2], "conics": out[3]}
        float_outs = ("means2d", "depths", "conics")
    g = torch.Generator().manual_seed(seed)
    loss = 0.0
    wts = {}
    for k in float_outs:
        wts[k] = (torch.rand(outs[k].shape, generator=g) * 2 - 1).to(outs[k].device)
        valid = outs[k] if k in ("render_colors", "render_alphas") else outs[k] * (outs["radii"] > 0).reshape(
            outs["radii"].shape + (1,) * (outs[k].dim() - outs["radii"].dim()))
        loss = loss + (valid * wts[k]).sum()
    loss.backward()
    npy = lambda t: t.detach().cpu().numpy()  # noqa: E731
    ins = {k: npy(v) for k, v in kwargs.items() if torch.is_tensor(v)}
    return ins, {k: npy(v) for k, v in outs.items()}, {k: npy(v.grad) for k, v in leaves.items() if v.grad is not None}, \
        {k: npy(v) for k, v in wts.items()}


def main():
    try:
        import gsplat
        from gsplat.rendering import fully_fused_projection, rasterization
    except ImportError as e:
        raise SystemExit(f"gsplat is not importable here ({e}); run this where gsplat==1.4.0 is installed")
    dev = "cuda"
    s, d, o, cam = scene(dev)
    blob = {"gsplat_version": np.array(getattr(gsplat, "__version__", "unknown"))}
    for i, (name, (fn, kw)) in enumerate(callsites(s, d, o, cam, rasterization, fully_fused_projection, dev).items()):
        ins, outs, grads, wts = run_callsite(fn, kw, seed=100 + i)
        for k, v in ins.items():
            blob[f"{name}/in/{k}"] = v
        for k, v in outs.items():
            blob[f"{name}/out/{k}"] = v
        for k, v in grads.items():
            blob[f"{name}/grad/{k}"] = v
        for k, v in wts.items():
            blob[f"{name}/w/{k}"] = v
        blob[f"{name}/render_mode"] = np.array(kw.get("render_mode", ""))
    path = os.path.join(ROOT, "tests", "golden", "gsplat_callsites.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, "from gsplat", blob["gsplat_version"])


if __name__ == "__main__":
    main()
