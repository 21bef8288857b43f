"""ORACLE — test infrastructure only.  PARITY UNPINNED: simple-knn (gitlab.inria.fr/bkerbl/simple-knn, the
`submodules/simple-knn` entry of the reference's .gitmodules, no pinned revision) is not vendored in
/root/reference and cannot be installed here; this restates the quantity its `distCUDA2` returns — for each
point the mean of the squared Euclidean distances to its 3 nearest neighbours, the point itself excluded
(simple_knn.cu `boxMeanDist`: `dists[i] = (best[0] + best[1] + best[2]) / 3.0f`) — by brute force in float64.
Reference call sites: scene/gaussian_model.py:420-421, :514."""
import numpy as np


def dist2_mean3(points: np.ndarray) -> np.ndarray:
    p = points.astype(np.float64)
    n = p.shape[0]
    out = np.empty(n)
    for s in range(0, n, 1024):
        d2 = ((p[s:s + 1024, None, :] - p[None, :, :]) ** 2).sum(-1)
        d2[np.arange(d2.shape[0]), np.arange(s, s + d2.shape[0])] = np.inf     # the point itself
        out[s:s + 1024] = np.sort(d2, axis=1)[:, :3].mean(axis=1)
    return out
