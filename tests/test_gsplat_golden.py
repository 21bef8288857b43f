"""Holds the oracle (CPU) and the CUDA operators (GPU) to golden vectors written by the REAL gsplat 1.4.0 at
the reference's seven render() call sites — tests/golden/gsplat_callsites.npz, produced by
tests/golden/make_gsplat_golden.py on a machine where gsplat is installed.  gsplat is not installable in the
build image, so the file is absent today and these tests SKIP: the gsplat half of the oracle stays
"parity unpinned" (oracle/gsplat_ref.py header) until someone runs that script.

A second test always runs: it drives the generator's own call-site table through the ORACLE (standing in for
gsplat) so that the script, its kwarg sets and this checker cannot rot."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "gsplat_callsites.npz")
sys.path.insert(0, os.path.join(HERE, "golden"))


def _close(got, want, name, atol=1e-4, rtol=1e-3, frac=1e-3):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(1.0, float(np.abs(want).max())) if "grad" in name else 1.0
    bad = np.abs(got - want) > atol * scale + rtol * np.abs(want)
    assert bad.mean() <= frac, (name, float(np.abs(got - want).max()), float(bad.mean()))


def _check_against(blob, rasterization, fully_fused_projection, device):
    import make_gsplat_golden as G
    names = sorted({k.split("/")[0] for k in blob.files if "/" in k})
    assert len(names) == 7
    for i, name in enumerate(names):
        ins = {k.split("/")[2]: torch.from_numpy(blob[k]).to(device) for k in blob.files if k.startswith(name + "/in/")}
        kw = dict(ins, width=G.W, height=G.H)
        if name == "l190_project":
            fn, kw["covars"] = fully_fused_projection, None
        else:
            fn = rasterization
            kw.update(packed=False, render_mode=str(blob[name + "/render_mode"]))
            kw.setdefault("backgrounds", None)
        leaves = {}
        for k in G.DIFF:
            if torch.is_tensor(kw.get(k)):
                leaves[k] = kw[k].clone().requires_grad_(True)
                kw[k] = leaves[k]
        out = fn(**kw)
        if name == "l190_project":
            outs = {"radii": out[0], "means2d": out[1], "depths": out[2], "conics": out[3]}
        else:
            outs = {"render_colors": out[0], "render_alphas": out[1], "radii": out[2]["radii"], "means2d": out[2]["means2d"]}
        vis = torch.from_numpy(blob[name + "/out/radii"]).to(device) > 0
        assert ((outs["radii"] > 0) != vis).float().mean() <= 1e-3, name
        loss = 0.0
        for k in [k.split("/")[2] for k in blob.files if k.startswith(name + "/w/")]:
            w = torch.from_numpy(blob[f"{name}/w/{k}"]).to(device)
            o = outs[k]
            if k not in ("render_colors", "render_alphas"):          # gsplat leaves culled entries uninitialised
                m = vis.reshape(vis.shape + (1,) * (o.dim() - vis.dim()))
                o = o * m
                _close((o).detach().cpu(), blob[f"{name}/out/{k}"] * m.cpu().numpy(), f"{name}/out/{k}")
            else:
                _close(o.detach().cpu(), blob[f"{name}/out/{k}"], f"{name}/out/{k}")
            loss = loss + (o * w).sum()
        loss.backward()
        for k, v in leaves.items():
            key = f"{name}/grad/{k}"
            if key in blob.files:
                assert v.grad is not None, key
                _close(v.grad.cpu(), blob[key], key)


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/gsplat_callsites.npz absent: gsplat parity UNPINNED "
                    "(run tests/golden/make_gsplat_golden.py where gsplat==1.4.0 is installed)")
def test_oracle_matches_real_gsplat_golden():
    from oracle import gsplat_ref as O
    _check_against(np.load(GOLD), O.rasterization, O.fully_fused_projection, "cpu")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/gsplat_callsites.npz absent: gsplat parity UNPINNED")
def test_cuda_operators_match_real_gsplat_golden():
    from mobgs_b200.rendering import fully_fused_projection, rasterization
    _check_against(np.load(GOLD), rasterization, fully_fused_projection, "cuda")


def test_generator_table_runs_on_the_oracle(tmp_path):
    """make_gsplat_golden's call-site table + run_callsite + this file's checker, with the oracle standing in for
    gsplat: writes a blob exactly as the real script would and checks the oracle against it (trivially equal —
    the point is that the seven kwarg sets are accepted and every gradient the checker expects exists)."""
    import make_gsplat_golden as G
    from oracle import gsplat_ref as O
    s, d, o, cam = G.scene("cpu")
    blob = {}
    table = G.callsites(s, d, o, cam, O.rasterization, O.fully_fused_projection, "cpu")
    assert list(table) == ["l143_dyn_rgbed", "l163_dyn_alpha", "l190_project", "l201_all_rgbed", "l236_stat_rgbed",
                           "l255_stat_alpha", "l274_flow"]
    for i, (name, (fn, kw)) in enumerate(table.items()):
        ins, outs, grads, wts = G.run_callsite(fn, kw, seed=100 + i)
        assert "means" in grads and "viewmats" in grads, name
        for grp, dct in (("in", ins), ("out", outs), ("grad", grads), ("w", wts)):
            for k, v in dct.items():
                blob[f"{name}/{grp}/{k}"] = v
        blob[f"{name}/render_mode"] = np.array(kw.get("render_mode", ""))
    path = tmp_path / "blob.npz"
    np.savez(path, **blob)
    _check_against(np.load(path), O.rasterization, O.fully_fused_projection, "cpu")
