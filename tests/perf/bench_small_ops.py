"""Timing of the two single-launch helpers of round 2's third session next to what they replace:
  * main_utils.get_normals (train.py:590): mobgs_depth_normals vs the reference's statements (numpy pixel grid on the host,
    [H,W,3] upload, torch ops) restated with torch on the device;
  * HexPlane-MLP operand packing: mobgs_pack_operands vs the torch formulation (mask / subtract / reshape / cat per tile).
    python tests/perf/bench_small_ops.py > gpurun_out/r2c_small_ops.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n        # device ms, wall ms per call


def reference_style_normals(z, ppx, ppy, sfx, sfy):
    """main_utils.py:95-141 as the reference runs it: numpy grid + upload + torch ops"""
    H, W = z.shape[-2:]
    xx, yy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    px = np.stack([xx, yy], axis=-1) + 0.5
    y = (px[..., 1] - ppy) / sfy
    x = (px[..., 0] - ppx - y * 0.0) / sfx
    viewdirs = torch.from_numpy(np.stack([x, y, np.ones_like(x)], axis=-1)).to(z.device)
    coords = (viewdirs[None] * z[..., None]).squeeze(0)
    hd, wd, _ = coords.shape
    l2r = coords[1:hd - 1, 2:wd] - coords[1:hd - 1, 0:wd - 2]
    b2t = coords[0:hd - 2, 1:wd - 1] - coords[2:hd, 1:wd - 1]
    n = torch.nn.functional.normalize(torch.cross(l2r, b2t, dim=-1), p=2, dim=-1)
    return torch.nn.functional.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant")[None]


def main():
    from mobgs_b200 import deformation as D
    from mobgs_b200.main_utils import depth_normals
    from test_oracle import _hexplane_args
    out = {"normals": {}, "operand_pack": {}}
    for W, H in ((512, 288), (1920, 1080)):
        z = (1 + 3 * torch.rand(1, H, W)).cuda()
        a = (W / 2, H / 2, 0.9 * W, 0.9 * W)
        ours = timeit(lambda: depth_normals(z, *a))
        ref = timeit(lambda: reference_style_normals(z, *a), n=10, warm=2)
        err = float((depth_normals(z, *a) - reference_style_normals(z, *a)).abs().max())
        px = W * H
        out["normals"][f"{W}x{H}"] = {"ours_device_ms": ours[0], "ours_wall_ms": ours[1], "reference_statements_device_ms": ref[0],
                                      "reference_statements_wall_ms": ref[1], "max_abs_diff": err,
                                      "hbm_GBps": 16 * px / (ours[0] * 1e-3) / 1e9}
    args = _hexplane_args(64)
    args.multires = [1, 2]
    args.kplanes_config = dict(args.kplanes_config, resolution=[64, 64, 64, 150])     # arguments/stereo/*.py
    net = D.HexPlaneMLP(args).cuda()
    d = net.deformation_net
    heads = [(s[1].weight, s[1].bias, s[3].weight, s[3].bias) for s in (d.pos_deform, d.scales_deform, d.rotations_deform)]
    planes = [[p for p in level] for level in d.grid.grids]
    w0, b0 = d.feature_out[0].weight, d.feature_out[0].bias
    pack = D._operand_pack(planes, w0, b0, heads)

    def device_pack():
        pack.versions = None
        pack.refresh()

    def tile(w):
        rows, K = w.shape
        hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)
        lo = w - hi
        t = lambda m: m.reshape(rows, K // 4, 4).permute(1, 0, 2).contiguous().reshape(-1)  # noqa: E731
        return torch.cat([t(hi), t(lo)])

    def torch_pack():
        with torch.no_grad():
            r = [torch.cat([tile(w0[h * 64:(h + 1) * 64]) for h in range(2)])]
            r.append(torch.cat([tile(Wa[h * 64:(h + 1) * 64]) for Wa, _, _, _ in heads for h in range(2)]))
            for _, _, Wb, _ in heads:
                pad = torch.zeros(16, 128, device="cuda"); pad[:Wb.shape[0]] = Wb
                r.append(tile(pad))
                r += [tile(pad.t().contiguous()[h * 64:(h + 1) * 64]) for h in range(2)]
            w0t = w0.t().contiguous()
            r += [tile(w0t[:64])] + ([tile(w0t[64:])] if w0t.shape[0] > 64 else [])
            r.append(torch.cat([tile(Wa.t().contiguous()[h * 64:(h + 1) * 64]) for Wa, _, _, _ in heads for h in range(2)]))
            r += [g.detach()[0].permute(1, 2, 0).contiguous() for level in planes for g in level]
            return r

    dp, tp = timeit(device_pack), timeit(torch_pack, n=10, warm=2)
    n_plane = sum(g.numel() for level in planes for g in level)
    out["operand_pack"] = {"config": "net_width 128, 2 levels of [64,64,64,150] planes", "plane_floats": n_plane,
                           "device_kernel_device_ms": dp[0], "device_kernel_wall_ms": dp[1],
                           "torch_ops_device_ms": tp[0], "torch_ops_wall_ms": tp[1]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
