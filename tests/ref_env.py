"""Test infrastructure: run the UNMODIFIED reference Python (/root/reference) in this GPU-less build
container, with `gsplat.rendering` served by the CPU oracle and the reference's hard-coded CUDA device
literals (`.cuda()`, device="cuda", torch.cuda.Event — SURVEY.md Appendix A) redirected to the CPU.

Nothing here ships and nothing here edits a reference file: the redirect patches torch's factory
functions / Tensor.cuda / Module.cuda in THIS process only, and only when no CUDA device exists.  Tests
that use it skip when /root/reference is absent (the GPU box).  What it is for:

  * constructing the reference's real GaussianModel / Camera / blceKernel through compat/ and checking
    that everything mobgs_b200's renderer reads from them exists with the right dtype / shape;
  * executing the reference's real `train.scene_reconstruction` loop (train.py:202-823) for a few
    iterations on a synthetic Stereo-Blur-shaped scene and recording what it computes, as the golden
    fixture the GPU harness (tests/harness_train_loop.py) replays on the CUDA path.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "gaussian_renderer"))


def setup_paths(oracle_gsplat: bool = True):
    """compat/ shims + repo + reference on sys.path; `gsplat.rendering` = the CPU oracle."""
    for p in (REF, ROOT, os.path.join(ROOT, "compat")):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT, REF]
    if oracle_gsplat:
        import oracle.gsplat_ref as og
        pkg = types.ModuleType("gsplat")
        pkg.rendering = og
        sys.modules["gsplat"] = pkg
        sys.modules["gsplat.rendering"] = og


class _FakeEvent:
    def __init__(self, *a, **k):
        import time
        self._t = time.perf_counter()

    def record(self, *a, **k):
        import time
        self._t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other._t - self._t) * 1e3


_FACTORIES = ("tensor", "zeros", "ones", "empty", "rand", "randn", "full", "arange", "eye", "linspace", "randint",
              "zeros_like", "ones_like", "empty_like", "rand_like", "randn_like", "as_tensor", "tril", "randperm")


def _is_cuda_dev(d):
    return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda") \
        or isinstance(d, int)


@contextlib.contextmanager
def cuda_to_cpu():
    """Within the context every request for a CUDA device lands on the CPU (no-op on a machine with CUDA)."""
    if torch.cuda.is_available():
        yield
        return
    saved = {}

    def wrap_factory(fn):
        def w(*a, **k):
            if _is_cuda_dev(k.get("device")) and not isinstance(k.get("device"), int):
                k["device"] = "cpu"
            return fn(*a, **k)
        return w

    for n in _FACTORIES:
        saved[("torch", n)] = getattr(torch, n)
        setattr(torch, n, wrap_factory(getattr(torch, n)))
    saved["Tensor.cuda"] = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    saved["Module.cuda"] = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, *a, **k: self
    t_to, m_to = torch.Tensor.to, torch.nn.Module.to

    def fix_args(a, k):
        a = tuple("cpu" if (_is_cuda_dev(x) and not isinstance(x, int)) else x for x in a)
        if _is_cuda_dev(k.get("device")) and not isinstance(k.get("device"), int):
            k = dict(k, device="cpu")
        return a, k

    def tensor_to(self, *a, **k):
        a, k = fix_args(a, k)
        return t_to(self, *a, **k)

    def module_to(self, *a, **k):
        a, k = fix_args(a, k)
        return m_to(self, *a, **k)

    saved["Tensor.to"], saved["Module.to"] = t_to, m_to
    torch.Tensor.to, torch.nn.Module.to = tensor_to, module_to
    saved["device"] = torch.device
    cuda_names = ("Event", "synchronize", "empty_cache", "manual_seed_all", "set_device")
    for n in cuda_names:
        saved[("cuda", n)] = getattr(torch.cuda, n)
    torch.cuda.Event = _FakeEvent
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda *a, **k: None
    torch.cuda.manual_seed_all = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    try:
        yield
    finally:
        for key, v in saved.items():
            if isinstance(key, tuple) and key[0] == "torch":
                setattr(torch, key[1], v)
            elif isinstance(key, tuple) and key[0] == "cuda":
                setattr(torch.cuda, key[1], v)
        torch.Tensor.cuda, torch.nn.Module.cuda = saved["Tensor.cuda"], saved["Module.cuda"]
        torch.Tensor.to, torch.nn.Module.to = saved["Tensor.to"], saved["Module.to"]


# ---------------------------------------------------------------------------------------------
# A synthetic Stereo-Blur-shaped scene made of the reference's own objects
# ---------------------------------------------------------------------------------------------
def hyper_args():
    """ModelHiddenParams as arguments/stereo/default.py + seesaw.py configure them (net_width 128, defor_depth 1,
    3 multires levels of a [16,16,16,12] HexPlane — the base resolution is reduced from 64 to keep this small)."""
    from argparse import ArgumentParser
    from arguments import ModelHiddenParams
    p = ArgumentParser()
    hp = ModelHiddenParams(p)
    args = p.parse_args([])
    args.kplanes_config = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32,
                           "resolution": [16, 16, 16, 12]}
    args.multires = [1, 2, 4]
    args.net_width = 128
    args.defor_depth = 1
    return hp.extract(args)


def group_args(cls, **over):
    from argparse import ArgumentParser
    p = ArgumentParser()
    g = cls(p)
    args = p.parse_args([])
    for k, v in over.items():
        setattr(args, k, v)
    return g.extract(args)


def make_metadata(w2c: np.ndarray, focal: float, W: int, H: int):
    """dycheck_geometry.camera.Camera as scene/dataset_readers.py:1533-1548 builds it for Stereo-Blur."""
    from dycheck_geometry.camera import Camera as DyCamera
    c2w = np.linalg.inv(w2c)
    return DyCamera(orientation=w2c[:3, :3].astype(np.float32), position=c2w[:3, 3].astype(np.float32),
                    focal_length=np.float32(focal), principal_point=np.array([W / 2.0, H / 2.0], np.float32),
                    image_size=np.array([W, H], np.uint32))


def make_camera(uid, w2c: np.ndarray, focal, W, H, time, max_time, image, depth, mask, device="cpu"):
    """The reference's scene.cameras.Camera, built the way FourDGSdataset does (scene/dataset.py)."""
    from scene.cameras import Camera
    from utils.graphics_utils import focal2fov
    R = np.transpose(w2c[:3, :3]).astype(np.float32)      # Camera.R is the transposed world-to-camera rotation
    T = w2c[:3, 3].astype(np.float32)
    Kmat = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]], np.float32)
    normal = np.zeros((H, W, 3), np.float32); normal[..., 2] = 1
    return Camera(colmap_id=uid, R=R, T=T, FoVx=focal2fov(focal, W), FoVy=focal2fov(focal, H), image=image,
                  gt_alpha_mask=None, image_name=f"{uid:05d}", uid=uid, max_time=max_time, data_device=device, time=time,
                  mask=mask, metadata=make_metadata(w2c, focal, W, H), normal=normal, depth=depth,
                  target_ts=np.zeros(3, np.float32), K=Kmat)


class FakeScene:
    """Duck-typed scene.Scene for train.scene_reconstruction (train.py:250-256, :743-773, :822): camera lists,
    the two models, a model path; saving is a no-op."""

    def __init__(self, train_cams, test_cams, stat, dyn, model_path):
        self.train_cams, self.test_cams = train_cams, test_cams
        self.stat_gaussians, self.dyn_gaussians = stat, dyn
        self.model_path = model_path
        self.dataset_type = "nvidia"
        self.cameras_extent = 1.0
        self.maxtime = train_cams[0].max_time

    def getTrainCameras(self): return self.train_cams
    def getTestCameras(self): return self.test_cams
    def getVideoCameras(self): return self.test_cams
    def save(self, *a, **k): pass
    def save_best_psnr(self, *a, **k): pass


def synthetic_reference_scene(n_static=260, n_dynamic=140, W=64, H=48, n_views=4, seed=3, model_path="/tmp/mobgs_ref_scene"):
    """Real GaussianModels initialised by the reference's own create_from_pcd / create_from_pcd_dynamic from a
    seeded point cloud (SURVEY §8d geometry), and real Cameras around it.  Call inside cuda_to_cpu()."""
    from scene.gaussian_model import GaussianModel
    from utils.graphics_utils import BasicPointCloud
    rng = np.random.default_rng(seed)
    focal = 0.9 * W
    hyper = hyper_args()

    def cloud(n):
        z = 2 + 6 * rng.random(n)
        x = (rng.random(n) * 2 - 1) * (W / 2 / focal) * z
        y = (rng.random(n) * 2 - 1) * (H / 2 / focal) * z
        return np.stack([x, y, z], -1).astype(np.float32)

    sp, dp = cloud(n_static), cloud(n_dynamic)
    stat, dyn = GaussianModel(3, hyper), GaussianModel(3, hyper)
    stat.create_from_pcd(BasicPointCloud(points=sp, colors=rng.random((n_static, 3)).astype(np.float32),
                                         normals=np.zeros_like(sp), times=rng.random((n_static, 1)).astype(np.float32)),
                         spatial_lr_scale=5, time_line=0)
    n_t = 6
    traj = dp[:, None, :] + np.cumsum(0.02 * rng.standard_normal((n_dynamic, n_t, 3)), axis=1).astype(np.float32)
    dyn.create_from_pcd_dynamic(BasicPointCloud(points=dp, colors=rng.random((n_dynamic, 3)).astype(np.float32),
                                                normals=np.zeros_like(dp), times=rng.random((n_dynamic, 1)).astype(np.float32)),
                                spatial_lr_scale=5, time_line=0, dyn_tracjectory=torch.from_numpy(traj))
    dyn._deformation.deformation_net.set_aabb(sp.max(0), sp.min(0), ref_type=dyn.get_xyz)
    with torch.no_grad():      # distCUDA2 scales of a 400-point cloud are huge: use the benchmark's pixel footprints
        for pc, pts in ((stat, sp), (dyn, dp)):
            z = torch.from_numpy(pts[:, 2:3])
            lo, hi = torch.log(0.8 / focal * z), torch.log(3.0 / focal * z)
            pc._scaling.copy_(lo + (hi - lo) * torch.from_numpy(rng.random((pts.shape[0], 3)).astype(np.float32)))
            pc._rotation.copy_(torch.from_numpy(rng.standard_normal((pts.shape[0], 4)).astype(np.float32)))
            pc._opacity.copy_(torch.from_numpy((1.5 * rng.standard_normal((pts.shape[0], 1))).astype(np.float32)))
        dyn._omega.copy_(torch.from_numpy((0.05 * rng.standard_normal((n_dynamic, 4))).astype(np.float32)))
        dyn._features_t.copy_(torch.from_numpy((0.1 * rng.standard_normal((n_dynamic, 3))).astype(np.float32)))
        for p in dyn.rgbdecoder.parameters():
            bound = float(np.sqrt(6.0 / (p.shape[0] + p.shape[1])))
            p.copy_(torch.from_numpy(((rng.random(tuple(p.shape)) * 2 - 1) * bound).astype(np.float32)))

    def cams(offset):
        out = []
        for v in range(n_views):
            a = np.radians(1.5 * (v - (n_views - 1) / 2) + offset)
            w2c = np.eye(4, dtype=np.float64)
            w2c[0, 0], w2c[0, 2], w2c[2, 0], w2c[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
            w2c[0, 3] = 0.03 * (v - (n_views - 1) / 2) + 0.01 * offset
            img = torch.from_numpy(rng.random((3, H, W)).astype(np.float32))
            depth = (2 + 6 * rng.random((H, W, 1))).astype(np.float32)
            mask = (rng.random((H, W, 1)) > 0.7).astype(np.float32)
            out.append(make_camera(v, w2c, focal, W, H, time=v / max(n_views - 1, 1), max_time=n_views - 1, image=img,
                                   depth=depth, mask=mask))
        return out

    train, test = cams(0.0), cams(0.4)
    os.makedirs(model_path, exist_ok=True)
    return stat, dyn, FakeScene(train, test, stat, dyn, model_path), hyper
