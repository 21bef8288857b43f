"""Writes tests/golden/normals.npz: inputs and outputs of the UNMODIFIED reference `main_utils.get_normals`
(/root/reference/main_utils.py:95-141) executed here on the CPU with the reference's own dycheck_geometry Camera as
`camera_metadata` — the fixture that pins oracle/normals_ref.py and the CUDA kernel (mobgs_depth_normals).

    python tests/golden/make_normals_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_env  # noqa: E402

CASES = [  # name, W, H, focal, principal point, skew, pixel aspect ratio, use_center
    ("centre", 40, 28, 36.0, (20.0, 14.0), 0.0, 1.0, True),
    ("skewed", 33, 21, 27.5, (15.25, 11.5), 0.15, 1.1, False),
    ("thin", 3, 3, 5.0, (1.5, 1.5), 0.0, 1.0, True),
]


def main():
    ref_env.setup_paths()
    from dycheck_geometry.camera import Camera as DyCamera
    from main_utils import get_normals
    out = {}
    g = torch.Generator().manual_seed(11)
    for name, W, H, f, pp, skew, par, uc in CASES:
        cam = DyCamera(orientation=np.eye(3, dtype=np.float32), position=np.zeros(3, np.float32), focal_length=np.float32(f),
                       principal_point=np.array(pp, np.float32), image_size=np.array([W, H], np.uint32), skew=skew,
                       pixel_aspect_ratio=par, use_center=uc)
        # a smooth depth ramp plus noise
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        z = 2.0 + 0.03 * xx + 0.05 * yy + 0.2 * torch.rand(H, W, generator=g)
        z = (z[None] + 1e-6)           # train.py:590 passes pred_depth + 1e-6
        n = get_normals(z, cam)
        assert n.shape == (1, 3, H, W) and n.dtype == torch.float32
        out[name + "_z"] = z.numpy()
        out[name + "_normals"] = n.numpy()
        out[name + "_intr"] = np.array([pp[0], pp[1], float(cam.scale_factor_x), float(cam.scale_factor_y), skew,
                                        0.5 if uc else 0.0], np.float32)
    np.savez_compressed(os.path.join(HERE, "normals.npz"), **out)
    print("wrote normals.npz:", sorted(out))


if __name__ == "__main__":
    main()
