"""simple_knn._C.distCUDA2 (scene/gaussian_model.py:10, used at init only): mean squared distance to the
3 nearest neighbours.  CUDA tensors go to the library (mobgs_b200.knn, tiled brute force); the torch
brute force below only serves CPU tensors, i.e. running the reference's Python on a GPU-less box."""
import torch


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if points.is_cuda:
        from mobgs_b200.knn import distCUDA2 as native
        return native(points)
    pts = points.float()
    out = torch.empty(pts.shape[0], device=pts.device)
    for s in range(0, pts.shape[0], 256):
        d = pts[s:s + 256, None, :] - pts[None, :, :]
        d2 = (d * d).sum(-1)
        out[s:s + 256] = d2.topk(4, dim=1, largest=False).values[:, 1:].mean(dim=1)
    return out
