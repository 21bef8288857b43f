"""f2 measurement: photometric loss forward + backward (train.py:621-628) on a [B,3,H,W] batch —
mobgs_b200.losses.photo_loss (2 launches) against the reference's own call pattern in PyTorch
(oracle/loss_ref.py == utils/loss_utils.py l1_loss + ssim: 5 grouped conv2d + elementwise, autograd
backward) on the same GPU.  One JSON line per shape, with the HBM roofline of the two kernels:
44 algorithmic bytes per element (fwd: read x, y, write 3 partial maps; bwd: read x, y, 3 maps, write v)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def timeit(fn, flush, steps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def main():
    from mobgs_b200 import _lib, losses
    from oracle import loss_ref as L           # the reference's call pattern (test/bench infrastructure)
    dev = torch.device("cuda")
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
    peak = 6462.1
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    for shape in ((2, 3, 288, 512), (2, 3, 1080, 1920)):
        g = torch.Generator().manual_seed(0)
        gt = torch.rand(shape, generator=g).to(dev)
        img = torch.rand(shape, generator=g).to(dev).requires_grad_(True)

        def ours():
            img.grad = None
            losses.photo_loss(img, gt, 0.2).backward()

        def ref():
            img.grad = None
            L.photo_loss(img, gt, 0.2).backward()

        t_ref = timeit(ref, flush)
        torch.backends.cudnn.allow_tf32 = False
        t_ref_fp32 = timeit(ref, flush)
        torch.backends.cudnn.allow_tf32 = True
        t_ours = timeit(ours, flush)
        _lib.TIMING = {}
        for _ in range(10):
            flush.zero_(); ours()
        torch.cuda.synchronize()
        k = {n: sum(a.elapsed_time(b) for a, b in ev) / 10 for n, ev in _lib.TIMING.items()}
        _lib.TIMING = None
        kern = k["mobgs_photo_loss_fwd"] + k["mobgs_photo_loss_bwd"]
        n = img.numel()
        ach = 44.0 * n / (kern * 1e-3) / 1e9
        print(json.dumps({"what": "f2 fused L1 + SSIM photometric loss, forward + backward", "shape": list(shape),
                          "ms": {"ours": t_ours, "ours_kernels_only": kern, "kernels": k, "torch_reference_pattern": t_ref,
                                 "torch_reference_pattern_no_tf32": t_ref_fp32},
                          "speedup_vs_torch": t_ref / t_ours,
                          "roofline": {"bound": "hbm", "kernel": "mobgs_photo_loss_fwd+bwd", "achieved": ach, "peak": peak,
                                       "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes": 44.0 * n}}))


if __name__ == "__main__":
    main()
