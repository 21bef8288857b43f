"""simple_knn._C.distCUDA2 on the CUDA library: mean squared distance of every
point to its 3 nearest neighbours, used by GaussianModel.create_from_pcd* to initialise the scales
(scene/gaussian_model.py:420-421, :514).  CUDA only; compat/simple_knn routes here."""
import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("mobgs_b200.knn.distCUDA2 needs a CUDA tensor (there is no CPU fallback)")
    pts = points.detach().float().contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise RuntimeError(f"distCUDA2 expects [N,3] points, got {tuple(pts.shape)}")
    out = torch.empty(pts.shape[0], device=pts.device)
    _lib.knn3_mean_dist2(pts.data_ptr(), out.data_ptr(), pts.shape[0], _lib.current_stream())
    return out
