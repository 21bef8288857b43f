"""Time the fused HexPlane+MLP forward (a11) against the PyTorch restatement of the reference
module on the same GPU (18 grid_sample + cuBLAS launches).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from argparse import Namespace  # noqa: E402
from mobgs_b200.deformation import HexPlaneMLP  # noqa: E402
from oracle.hexplane_ref import deform_forward_ref  # noqa: E402  (the timed torch baseline)


def main(n=150_000, iters=20):
    args = Namespace(net_width=128, defor_depth=1, bounds=1.6, grid_pe=0,
                     kplanes_config={"grid_dimensions": 2, "input_coordinate_dim": 4,
                                     "output_coordinate_dim": 32, "resolution": [64, 64, 64, 12]},
                     multires=[1, 2, 4], no_dx=False, no_grid=False, no_ds=False, no_dr=False,
                     empty_voxel=False, static_mlp=False, apply_rotation=False)
    torch.manual_seed(0)
    net = HexPlaneMLP(args).cuda()
    pts = (torch.rand(n, 3, device="cuda") * 3 - 1.5)
    scales = torch.randn(n, 3, device="cuda")
    rots = torch.randn(n, 4, device="cuda")
    t = torch.rand(n, 1, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = False

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    with torch.no_grad():
        ms_fused = timeit(lambda: net(pts, scales, rots, t))
        ms_torch = timeit(lambda: deform_forward_ref(net, pts, scales, rots, t))
        a = net(pts, scales, rots, t)
        b = deform_forward_ref(net, pts, scales, rots, t)
        err = max(float((x - y).abs().max()) for x, y in zip(a, b))
    # forward + backward (every gradient: inputs, planes, weights): tcgen05 path vs torch autograd (cuBLAS fp32)
    ins = [x.clone().requires_grad_(True) for x in (pts, scales, rots)]
    params = [p for p in net.parameters() if p.requires_grad]

    def train_fused():
        for p in params + ins:
            p.grad = None
        o = net(ins[0], ins[1], ins[2], t)
        (o[0].sum() + o[1].sum() + o[2].sum()).backward()

    def train_torch():
        for p in params + ins:
            p.grad = None
        o = deform_forward_ref(net, ins[0], ins[1], ins[2], t)
        (o[0].sum() + o[1].sum() + o[2].sum()).backward()

    ms_fb_fused, ms_fb_torch = timeit(train_fused), timeit(train_torch)
    from mobgs_b200 import _lib
    _lib.TIMING = {}
    train_fused()
    torch.cuda.synchronize()
    kern = {k: round(sum(a.elapsed_time(b) for a, b in v), 4) for k, v in _lib.TIMING.items()}
    _lib.TIMING = None
    flops = 2 * 63232 * n * 3          # 3xTF32: three tensor-core MACs per logical MAC
    print(json.dumps({"points": n, "fused_ms": ms_fused, "torch_fp32_ms": ms_torch, "speedup": ms_torch / ms_fused,
                      "max_abs_err_vs_torch": err, "issued_tf32_tflops": flops / ms_fused / 1e9,
                      "fwd_bwd_fused_ms": ms_fb_fused, "fwd_bwd_torch_fp32_ms": ms_fb_torch,
                      "fwd_bwd_speedup": ms_fb_torch / ms_fb_fused, "fwd_bwd_kernel_ms": kern,
                      "note": "fused_ms includes the per-call host-side weight tiling + channels-last plane copies"}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 150_000)
