"""Light stand-ins for the reference's `GaussianModel` / `Camera` / `Sandwich` objects exposing
exactly the attribute API the renderer reads (scene/gaussian_model.py:209-257,
435-470; scene/cameras.py:109-146; helper_model.py:7-28), plus the seeded synthetic scene of
SURVEY.md §8d used by tests, smoke() and bench.py (no dataset exists in this environment).

The reference's own GaussianModel / Camera instances work with mobgs_b200.gaussian_renderer
unchanged — these classes exist so that tests and benchmarks do not need /root/reference.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Sandwich(nn.Module):
    """Same parameter names / shapes as helper_model.Sandwich (state_dict compatible)."""

    def __init__(self, dim=9, outdim=3, bias=False):
        super().__init__()
        self.mlp1 = nn.Conv2d(12, 6, kernel_size=1, bias=bias)
        self.mlp2 = nn.Conv2d(6, 3, kernel_size=1, bias=bias)

    def forward(self, input, rays, time=None):
        albedo, spec, timefeature = input.chunk(3, dim=1)
        h = torch.relu(self.mlp1(torch.cat([spec, timefeature, rays], dim=1)))
        return torch.sigmoid(albedo + self.mlp2(h))


class GaussianSet:
    """Parameter container with the GaussianModel getter API used by render()/get_flow()."""

    control_num = 12

    def __init__(self):
        self.scaling_activation = torch.exp
        self.opacity_activation = torch.sigmoid
        self.rotation_activation = F.normalize
        self.rgbdecoder = None
        self._deformation = None

    # getters (scene/gaussian_model.py:209-257)
    @property
    def get_xyz(self): return self._xyz
    @property
    def get_control_xyz(self): return self.control_xyz
    @property
    def get_scaling(self): return self.scaling_activation(self._scaling)
    @property
    def get_rotation_stat(self): return self.rotation_activation(self._rotation)
    @property
    def get_opacity(self): return self.opacity_activation(self._opacity)
    @property
    def get_trbfcenter(self): return self._trbf_center
    @property
    def get_features_static(self): return torch.cat((self._features_dc, 0.0 * self._features_t), dim=1)

    def get_features(self, deltat): return torch.cat((self._features_dc, deltat * self._features_t), dim=1)

    def get_rotation_dy(self, rotation, delta_t): return rotation + delta_t * self._omega

    PARAM_NAMES = ("_xyz", "control_xyz", "_features_dc", "_features_t", "_opacity", "_scaling",
                   "_rotation", "_omega", "_trbf_center")

    def parameters(self):
        ps = [getattr(self, n) for n in self.PARAM_NAMES]
        if self.rgbdecoder is not None:
            ps += list(self.rgbdecoder.parameters())
        return ps

    def requires_grad_(self, flag=True):
        for p in self.parameters():
            if p.is_floating_point():
                p.requires_grad_(flag)
        return self

    def to(self, device):
        for n in self.PARAM_NAMES + ("current_control_num",):
            t = getattr(self, n)
            rg = t.requires_grad
            setattr(self, n, t.detach().to(device).requires_grad_(rg))
        if self.rgbdecoder is not None:
            self.rgbdecoder.to(device)
        return self


class PinholeCamera:
    """Attribute API of scene/cameras.py Camera that the renderer touches."""

    def __init__(self, w2c: torch.Tensor, fx, fy, cx, cy, width, height, time=0.0, max_time=23, uid=0):
        self.image_width, self.image_height = int(width), int(height)
        self.time, self.max_time, self.uid = float(time), int(max_time), uid
        dev = w2c.device
        self.world_view_transform = w2c.transpose(0, 1)      # reference stores the transpose
        self.K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32, device=dev)
        self.FoVx = 2 * math.atan(width / (2 * fx))
        self.FoVy = 2 * math.atan(height / (2 * fy))
        # cam_ray [1,6,H,W] = camera centre | normalised world-space view direction (cameras.py:132-146)
        c2w = torch.inverse(w2c)
        self.camera_center = c2w[:3, 3]
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32, device=dev) + 0.5,
                                torch.arange(width, dtype=torch.float32, device=dev) + 0.5, indexing="ij")
        d = torch.stack([(xs - cx) / fx, (ys - cy) / fy, torch.ones_like(xs)], dim=-1)
        d = d / d.norm(dim=-1, keepdim=True)
        d = d @ c2w[:3, :3].T
        origin = self.camera_center.expand_as(d)
        self.cam_ray = torch.cat([origin, d], dim=-1).permute(2, 0, 1).unsqueeze(0).contiguous()

    def get_pixels(self, image_size_x, image_size_y, use_center=None):
        xx, yy = np.meshgrid(np.arange(image_size_x, dtype=np.float32),
                             np.arange(image_size_y, dtype=np.float32))
        return np.stack([xx, yy], axis=-1) + (0.5 if use_center else 0)


def synthetic_scene(n_static: int, n_dynamic: int, width: int, height: int, seed: int = 1234,
                    device="cpu", requires_grad: bool = True, footprint_px=(0.4608, 2.304)):
    """Seeded synthetic Gaussians of SURVEY.md §8d: depths U(2,10), ~2-10 px radius footprints.
    `footprint_px` = range of the projected 1-sigma size in pixels; the default equals the spec's
    log-scale ~ U(ln 0.004 z, ln 0.02 z) at 128 px width (fx = 115.2) and keeps the same pixel
    footprint at every resolution."""
    g = torch.Generator().manual_seed(seed)
    fx = fy = 0.9 * width
    cx, cy = width / 2, height / 2

    def rand(*s): return torch.rand(*s, generator=g)
    def randn(*s): return torch.randn(*s, generator=g)

    def base(n):
        z = 2 + 8 * rand(n)
        x = (rand(n) * 2 - 1) * 1.2 * (width / 2 / fx) * z
        y = (rand(n) * 2 - 1) * 1.2 * (height / 2 / fy) * z
        xyz = torch.stack([x, y, z], -1)
        lo, hi = torch.log(footprint_px[0] / fx * z)[:, None], torch.log(footprint_px[1] / fx * z)[:, None]
        scaling = lo + (hi - lo) * rand(n, 3)
        return xyz, scaling, randn(n, 4), 2 * randn(n, 1), rand(n, 6)

    stat, dyn = GaussianSet(), GaussianSet()
    xyz, sc, rot, op, fdc = base(n_static)
    stat._xyz, stat._scaling, stat._rotation, stat._opacity, stat._features_dc = xyz, sc, rot, op, fdc
    stat._features_t = torch.zeros(n_static, 3)
    stat._omega = torch.zeros(n_static, 4)
    stat._trbf_center = torch.zeros(n_static, 1)
    stat.control_xyz = torch.zeros(n_static, 12, 3)
    stat.current_control_num = torch.full((n_static, 1), 12, dtype=torch.int64)

    xyz, sc, rot, op, fdc = base(n_dynamic)
    dyn._xyz, dyn._scaling, dyn._rotation, dyn._opacity, dyn._features_dc = xyz, sc, rot, op, fdc
    dyn.control_xyz = 100 * (xyz[:, None, :] + torch.cumsum(0.01 * randn(n_dynamic, 12, 3), dim=1))
    dyn.current_control_num = torch.randint(4, 13, (n_dynamic, 1), generator=g, dtype=torch.int64)
    dyn._omega = 0.05 * randn(n_dynamic, 4)
    dyn._features_t = 0.1 * randn(n_dynamic, 3)
    dyn._trbf_center = rand(n_dynamic, 1)
    dec = Sandwich()
    with torch.no_grad():
        for p in dec.parameters():
            bound = math.sqrt(6.0 / (p.shape[0] + p.shape[1]))
            p.copy_((rand(*p.shape) * 2 - 1) * bound)
    dyn.rgbdecoder = dec
    stat.rgbdecoder = Sandwich()
    for s in (stat, dyn):
        s.to(device)
        if requires_grad:
            s.requires_grad_(True)
    intr = SimpleNamespace(fx=fx, fy=fy, cx=cx, cy=cy, width=width, height=height)
    return stat, dyn, intr


def subframe_w2c(k: int, K: int, device="cpu") -> torch.Tensor:
    """Sub-frame k of K: rotate about y by 0.2 deg * (k - K//2), translate x by 0.01 * (k - K//2)."""
    d = k - K // 2
    a = math.radians(0.2 * d)
    w2c = torch.eye(4)
    w2c[0, 0], w2c[0, 2], w2c[2, 0], w2c[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    w2c[0, 3] = 0.01 * d
    return w2c.to(device)


def make_camera(intr, w2c, time=0.5, max_time=23, uid=0) -> PinholeCamera:
    return PinholeCamera(w2c, intr.fx, intr.fy, intr.cx, intr.cy, intr.width, intr.height, time, max_time, uid)
