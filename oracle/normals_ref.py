"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's main_utils.get_normals
(/root/reference/main_utils.py:95-141).  Pinned: tests/golden/normals.npz holds outputs of the unmodified reference
function (tests/golden/make_normals_golden.py); tests/test_oracle.py holds this file to them.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it."""
import numpy as np


def get_normals(z, ppx, ppy, sfx, sfy, skew=0.0, pixel_offset=0.5):
    """z [B,H,W] float32 -> normals [B,3,H,W] float32 (main_utils.py:95-141, evaluated per batch element)."""
    z = np.asarray(z, np.float32)
    B, H, W = z.shape
    f32 = np.float32
    xx, yy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))      # :96 get_pixels
    px, py = xx + f32(pixel_offset), yy + f32(pixel_offset)
    y = (py - f32(ppy)) / f32(sfy)                                                             # :97
    x = (px - f32(ppx) - y * f32(skew)) / f32(sfx)                                             # :98-100
    viewdirs = np.stack([x, y, np.ones_like(x)], axis=-1)                                      # :101
    out = np.zeros((B, 3, H, W), np.float32)
    if H < 3 or W < 3:
        return out
    for b in range(B):
        coords = viewdirs * z[b][..., None]                                                    # :104
        bottom, top = coords[2:H, 1:W - 1], coords[0:H - 2, 1:W - 1]                           # :131-132
        right, left = coords[1:H - 1, 2:W], coords[1:H - 1, 0:W - 2]                           # :133-134
        n = np.cross(right - left, top - bottom).astype(np.float32)                            # :135-137
        norm = np.sqrt((n * n).sum(-1, keepdims=True, dtype=np.float32))
        n = n / np.maximum(norm, f32(1e-12))                                                   # :138 F.normalize
        out[b, :, 1:H - 1, 1:W - 1] = n.transpose(2, 0, 1)                                     # :139 zero border
    return out
