"""K9 / f4: mobgs_b200.knn.distCUDA2 (and its compat/simple_knn route) against the brute-force oracle.
Tolerance (fp32, written here): 1e-5 relative (+1e-12: duplicated points have distance 0)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [4, 257, 1024, 3001])
def test_dist_cuda2_matches_oracle(n):
    from mobgs_b200.knn import distCUDA2
    from oracle.knn_ref import dist2_mean3
    g = torch.Generator().manual_seed(n)
    pts = torch.randn(n, 3, generator=g) * torch.tensor([2.0, 1.0, 0.3])
    if n > 100:
        pts[10] = pts[11]                       # exact duplicates: a zero distance counts
        pts[50:60] *= 1e-3                      # a dense clump
    out = distCUDA2(pts.cuda()).cpu().numpy()
    ref = dist2_mean3(pts.numpy())
    assert np.all(np.abs(out - ref) <= 1e-5 * ref + 1e-12), float(np.abs(out - ref).max())


def test_compat_simple_knn_routes_to_the_library():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))
    from simple_knn._C import distCUDA2
    from mobgs_b200 import _lib
    pts = torch.rand(500, 3, generator=torch.Generator().manual_seed(0))
    before = _lib.LAUNCH_COUNT
    a = distCUDA2(pts.cuda())
    assert _lib.LAUNCH_COUNT == before + 1
    b = distCUDA2(pts)                          # CPU tensors: the torch stand-in
    assert torch.allclose(a.cpu(), b, rtol=1e-4, atol=1e-10)
    with pytest.raises(RuntimeError):
        from mobgs_b200.knn import distCUDA2 as native
        native(pts)
