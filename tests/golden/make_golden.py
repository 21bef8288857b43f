"""Generate tests/golden/*.npz by executing the UNMODIFIED reference Python from /root/reference.

Runs only in the build container (it needs /root/reference); the fixtures it writes are what
travels to the GPU box.  Two families:

  render_*.npz   reference gaussian_renderer.render / get_flow / get_flow_static (source file
                 executed as-is) on the seeded stand-in scene of mobgs_b200.scene, with
                 `gsplat.rendering` replaced by the CPU oracle (injected into sys.modules by this script) and
                 the reference's own helper_model.Sandwich as the decoder.  This pins everything
                 *around* the two gsplat operators (spline, activations, concat order, decoder,
                 flow wiring, dict keys) to the reference's own code.
  hexplane_*.npz reference scene.deformation.deform_network (HexPlaneField + MLP heads, net_width 128,
                 3 levels x 6 planes x 32 features as in arguments/stereo/*.py but base resolution
                 16 instead of 64 to keep the fixture small), imported and run on CPU: inputs,
                 state_dict, outputs and gradients.

The reference hard-codes `.cuda()` / device="cuda" in the renderer; on this
GPU-less container those calls are redirected to the CPU through a proxy of the `torch` module
object *inside the imported reference module only* — the reference source is not edited.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT, REF]

# `gsplat.rendering` as seen by the reference renderer = the CPU oracle (test infrastructure; the
# compat/ shim itself only ever routes to the CUDA library)
import oracle.gsplat_ref as _oracle_gsplat  # noqa: E402
_pkg = types.ModuleType("gsplat")
_pkg.rendering = _oracle_gsplat
sys.modules["gsplat"] = _pkg
sys.modules["gsplat.rendering"] = _oracle_gsplat

import numpy as np  # noqa: E402
import torch  # noqa: E402


class _TorchCPU(types.ModuleType):
    """`torch` as seen by the reference renderer module: device='cuda' -> CPU, nothing else."""

    def __init__(self):
        super().__init__("torch")

    def __getattr__(self, name):
        attr = getattr(torch, name)
        if name in ("ones", "zeros", "tensor", "empty"):
            def wrapped(*a, **k):
                if k.get("device") == "cuda":
                    k.pop("device")
                return attr(*a, **k)
            return wrapped
        return attr


def load_reference_renderer():
    torch.Tensor.cuda = lambda self, *a, **k: self          # .cuda() is a no-op on this box
    import gaussian_renderer as ref
    assert ref.__file__.startswith(REF), ref.__file__
    ref.torch = _TorchCPU()
    return ref


def reference_decoder(weights_from):
    from helper_model import Sandwich
    dec = Sandwich(9, 3)
    dec.load_state_dict(weights_from.state_dict())
    return dec


def _np(v):
    return None if v is None else v.detach().cpu().numpy()


RENDER_CASES = {
    # name: (ns, nd, W, H, seed, time, kwargs)
    "render_plain": (220, 120, 64, 48, 11, 0.5, dict(get_static=True, get_dynamic=True)),
    "render_warped_flow": (220, 120, 64, 48, 12, 0.35, dict(get_static=True, get_dynamic=True,
                                                            delta_exposure=0.4, get_flow=True)),
    "render_time_clamped": (150, 100, 48, 48, 13, 0.97, dict(delta_exposure=1.0)),
}


def make_render_goldens(out_dir):
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    ref = load_reference_renderer()
    bg = torch.tensor([0.1, 0.4, 0.8, 1.0])
    for name, (ns, nd, W, H, seed, t, kw) in RENDER_CASES.items():
        stat, dyn, intr = synthetic_scene(ns, nd, W, H, seed=seed)
        dyn.rgbdecoder = reference_decoder(dyn.rgbdecoder)
        cam = make_camera(intr, subframe_w2c(1, 4), time=t)
        out = ref.render(cam, stat, dyn, None, bg, **kw)
        keys = ["render", "s_render", "s_depth", "d_render", "d_depth", "d_alpha", "s_alpha", "depth",
                "viewspace_points", "radii", "means_3d_final", "colors_precomp_final", "ori_flow",
                "ori_coord_map", "means_3d"]
        loss = out["render"].sum() + 0.1 * out["depth"].sum()
        if out["d_alpha"] is not None:
            loss = loss + out["d_alpha"].sum() + out["s_render"].mean()
        if out["ori_flow"] is not None:
            loss = loss + 0.01 * out["ori_flow"].sum()
        loss.backward()
        blob = {"out_" + k: _np(out[k]) for k in keys if out[k] is not None}
        blob.update({"grad_stat" + n: _np(getattr(stat, n).grad) for n in
                     ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc")})
        blob.update({"grad_dyn" + n: _np(getattr(dyn, n).grad) for n in
                     ("control_xyz", "_rotation", "_omega", "_scaling", "_opacity", "_features_dc", "_features_t")})
        blob["grad_viewspace"] = _np(out["viewspace_points"].grad)
        blob["grad_dec1"] = _np(dyn.rgbdecoder.mlp1.weight.grad)
        blob["grad_dec2"] = _np(dyn.rgbdecoder.mlp2.weight.grad)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
        print("wrote", name, {k: v.shape for k, v in blob.items() if k.startswith("out_")})

    # get_flow / get_flow_static
    stat, dyn, intr = synthetic_scene(200, 150, 64, 48, seed=21)
    dyn.rgbdecoder = reference_decoder(dyn.rgbdecoder)
    cam = make_camera(intr, subframe_w2c(2, 4), time=0.45)
    e2m, m2e, limg, lalpha = ref.get_flow(cam, stat, dyn, None, bg, delta_exposure=-0.7)
    cams = [make_camera(intr, subframe_w2c(k, 5)) for k in (0, 4, 2)]
    with torch.no_grad():
        f2d, rflow = ref.get_flow_static(*cams, stat, dyn, None, bg)
    np.savez_compressed(os.path.join(out_dir, "get_flow.npz"), exp2mid=_np(e2m), mid2exp=_np(m2e),
                        latent_img=_np(limg), latent_alpha=_np(lalpha), static_flow_2d=_np(f2d),
                        static_rendered_flow=_np(rflow))
    print("wrote get_flow")


def make_hexplane_golden(out_dir):
    from argparse import Namespace
    from scene.deformation import deform_network
    torch.manual_seed(5)
    args = Namespace(net_width=128, timebase_pe=4, defor_depth=1, posebase_pe=10, scale_rotation_pe=2,
                     opacity_pe=2, timenet_width=64, timenet_output=32, bounds=1.6, grid_pe=0,
                     kplanes_config={"grid_dimensions": 2, "input_coordinate_dim": 4,
                                     "output_coordinate_dim": 32, "resolution": [16, 16, 16, 6]},
                     multires=[1, 2, 4], no_dx=False, no_grid=False, no_ds=False, no_dr=False, no_do=True,
                     no_dshs=True, empty_voxel=False, static_mlp=False, apply_rotation=False)
    net = deform_network(args)
    # non-trivial planes / biases (time planes initialise to 1, biases to 0)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
        for g in net.deformation_net.grid.grids:
            for plane in g:
                plane.add_(0.05 * torch.randn_like(plane))
    net.deformation_net.set_aabb([1.2, 1.0, 1.4], [-1.3, -0.9, -1.1])
    n = 257
    pts = torch.rand(n, 3) * 3.0 - 1.5          # some outside the AABB -> clamped
    scales = torch.randn(n, 3) * 0.3 - 3.0
    rots = torch.randn(n, 4)
    t = torch.rand(n, 1)
    t[:8, 0] = torch.tensor([0.0, 1.0, 0.5, 0.25, 0.999, 1e-4, 0.75, 0.1])
    pts.requires_grad_(True)
    p2, s2, r2 = net(pts, scales, rots, t)
    w = torch.randn(n, 10, generator=torch.Generator().manual_seed(1))
    loss = (p2 * w[:, :3]).sum() + (s2 * w[:, 3:6]).sum() + (r2 * w[:, 6:]).sum()
    loss.backward()
    blob = {"in_pts": _np(pts), "in_scales": _np(scales), "in_rots": _np(rots), "in_t": _np(t),
            "out_pts": _np(p2), "out_scales": _np(s2), "out_rots": _np(r2), "loss_w": _np(w),
            "grad_pts": _np(pts.grad)}
    sd = net.state_dict()
    for k, v in sd.items():
        blob["sd::" + k] = _np(v)
    for k, p in net.named_parameters():
        if p.grad is not None:
            blob["pgrad::" + k] = _np(p.grad)
    np.savez_compressed(os.path.join(out_dir, "hexplane_w128.npz"), **blob)
    print("wrote hexplane", {k: v.shape for k, v in blob.items() if not k.startswith(("sd::", "pgrad::"))})


def make_photo_loss_golden(out_dir):
    """photo_loss.npz: the reference's own utils/loss_utils.py l1_loss + ssim (train.py:621-628) on CPU fp32:
    a [2,3,45,70] batch (not a multiple of the 16-px tile, like 288 = 18 x 16 but 70 != k x 16) with
    smooth + noisy content, plus a tiny [1,3,7,9] case smaller than the 11x11 window."""
    import utils.loss_utils as L          # the reference module, unmodified
    g = torch.Generator().manual_seed(11)
    blob = {}
    for tag, shape in (("a", (2, 3, 45, 70)), ("b", (1, 3, 7, 9))):
        B, C, H, W = shape
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        base = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + c) * torch.cos(4 * yy) for c in range(C)])[None].expand(B, -1, -1, -1)
        gt = (base + 0.05 * torch.randn(shape, generator=g)).clamp(0, 1).contiguous()
        img = (gt + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1).contiguous().requires_grad_(True)
        lam = 0.2
        l1 = L.l1_loss(img, gt)
        ss = L.ssim(img, gt)
        loss = l1 + lam * (1.0 - ss)
        loss.backward()
        blob.update({f"{tag}_img": _np(img), f"{tag}_gt": _np(gt), f"{tag}_l1": _np(l1), f"{tag}_ssim": _np(ss),
                     f"{tag}_loss": _np(loss), f"{tag}_grad": _np(img.grad), f"{tag}_lambda": np.float32(lam)})
    np.savez_compressed(os.path.join(out_dir, "photo_loss.npz"), **blob)
    print("wrote photo_loss", {k: getattr(v, "shape", None) for k, v in blob.items()})


def make_camera_rays_golden(out_dir):
    """camera_rays.npz: cam_ray of K = 3 cameras through the reference's own scene/cameras.py methods
    (get_pixels_torch :244-253, pixels_to_local_viewdirs_torch :255-266, pixels_to_viewdirs_torch :268-284,
    executed unmodified on a stub `self`) and the four lines of Camera.__init__ :140-146 that assemble the
    tensor, plus the gradients of a weighted sum w.r.t. R and the camera centre."""
    from types import SimpleNamespace as NS
    from scene.cameras import Camera          # the reference class, unmodified
    W, H, K = 37, 21, 3
    g = torch.Generator().manual_seed(3)
    meta = NS(principal_point_x=W / 2 + 0.7, principal_point_y=H / 2 - 0.4, scale_factor_x=31.0, scale_factor_y=29.5)
    rots, cens, rays, g_rot, g_cen = [], [], [], [], []
    wgt = torch.randn(K, 6, H, W, generator=g)
    for k in range(K):
        Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        Q = (Q * (1.0 + 0.05 * k)).requires_grad_(True)          # not exactly orthonormal: the re-normalisation matters
        c = torch.randn(3, generator=g).requires_grad_(True)
        stub = NS(metadata=meta, R=Q, data_device="cpu")
        stub.pixels_to_local_viewdirs_torch = lambda px, stub=stub: Camera.pixels_to_local_viewdirs_torch(stub, px)
        pixels = Camera.get_pixels_torch(stub, W, H, use_center=True)
        viewdirs = Camera.pixels_to_viewdirs_torch(stub, pixels)
        cam_origin, _ = torch.broadcast_tensors(c, viewdirs)                  # cameras.py:142
        cam_ray = torch.cat((cam_origin, viewdirs), dim=-1)                   # :143
        cam_ray = cam_ray.permute(2, 0, 1).unsqueeze(0)                       # :144
        (cam_ray[0] * wgt[k]).sum().backward()
        rots.append(_np(Q)); cens.append(_np(c)); rays.append(_np(cam_ray[0]))
        g_rot.append(_np(Q.grad)); g_cen.append(_np(c.grad))
    blob = {"rot": np.stack(rots), "centre": np.stack(cens), "rays": np.stack(rays), "weight": _np(wgt),
            "g_rot": np.stack(g_rot), "g_centre": np.stack(g_cen),
            "geom": np.array([meta.principal_point_x, meta.principal_point_y, meta.scale_factor_x, meta.scale_factor_y, W, H], np.float64)}
    np.savez_compressed(os.path.join(out_dir, "camera_rays.npz"), **blob)
    print("wrote camera_rays", {k: v.shape for k, v in blob.items()})


def make_flow_warp_golden(out_dir):
    """flow_warp_loss.npz: train.py:656-676 executed line by line (the statements are inline in the training
    loop, so they are reproduced here verbatim, in-place normalisation included) with the reference's own
    utils.loss_utils.l1_loss and torch's F.grid_sample, CPU fp32.  B = 2 views, K = 3 exposures, 23 x 31 px;
    coordinates = pixel grid + smooth offsets, some pushed outside the image (border clipping)."""
    import torch.nn.functional as F
    from utils.loss_utils import l1_loss          # the reference function, unmodified
    g = torch.Generator().manual_seed(21)
    B, K, H, W = 2, 3, 23, 31
    exposure_length = K
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    base = torch.stack([xs, ys], -1)[None, None].expand(B, K, -1, -1, -1)

    def coords():
        c = base + 3.0 * torch.randn(B, K, 1, 1, 2, generator=g) + 1.5 * torch.randn(B, K, H, W, 2, generator=g)
        c[:, :, :2, :, 1] -= 6.0          # rows sampled above the image
        c[:, :, :, -2:, 0] += 9.0         # columns sampled right of it
        return c.contiguous().requires_grad_(True)

    e2m_leaf, m2e_leaf = coords(), coords()
    ori_image_tensor = torch.rand(B, 3, H, W, generator=g).requires_grad_(True)   # live in train.py (:469, :607)
    latent_leaf = torch.rand(B, K, 3, H, W, generator=g).requires_grad_(True)
    la_leaf = torch.rand(B, K, 1, H, W, generator=g).requires_grad_(True)
    da_leaf = torch.rand(B, 1, H, W, generator=g).requires_grad_(True)
    exp2mid_coord_final_tensor, mid2exp_coord_final_tensor = e2m_leaf.clone(), m2e_leaf.clone()
    latent_img_final_tensor, latent_alpha_final_tensor, d_alpha_tensor = latent_leaf, la_leaf, da_leaf
    # ---- train.py:658-675 ----
    norm_exp2mid_coord_final_tensor = exp2mid_coord_final_tensor
    norm_exp2mid_coord_final_tensor[..., 0] = norm_exp2mid_coord_final_tensor[..., 0] / (W - 1)
    norm_exp2mid_coord_final_tensor[..., 1] = norm_exp2mid_coord_final_tensor[..., 1] / (H - 1)
    norm_exp2mid_coord_final_tensor = 2.0 * norm_exp2mid_coord_final_tensor - 1.0
    norm_exp2mid_coord_final_tensor = norm_exp2mid_coord_final_tensor.flatten(0, 1)
    warped_exp2mid_img_tensor = F.grid_sample(ori_image_tensor.unsqueeze(1).expand(-1, exposure_length, -1, -1, -1).flatten(0,1), norm_exp2mid_coord_final_tensor, mode='bilinear', padding_mode='border').reshape(-1, exposure_length, 3, H, W)
    norm_mid2exp_coord_final_tensor = mid2exp_coord_final_tensor
    norm_mid2exp_coord_final_tensor[..., 0] = norm_mid2exp_coord_final_tensor[..., 0] / (W - 1)
    norm_mid2exp_coord_final_tensor[..., 1] = norm_mid2exp_coord_final_tensor[..., 1] / (H - 1)
    norm_mid2exp_coord_final_tensor = 2.0 * norm_mid2exp_coord_final_tensor - 1.0
    norm_mid2exp_coord_final_tensor = norm_mid2exp_coord_final_tensor.flatten(0, 1)
    warped_mid2exp_img_tensor = F.grid_sample(latent_img_final_tensor.flatten(0,1), norm_mid2exp_coord_final_tensor, mode='bilinear', padding_mode='border').reshape(-1, exposure_length, 3, H, W)
    flow_loss = (l1_loss(warped_exp2mid_img_tensor.flatten(0,1), latent_img_final_tensor.flatten(0,1), mask=latent_alpha_final_tensor.flatten(0,1)) + l1_loss(warped_mid2exp_img_tensor.flatten(0,1), ori_image_tensor.unsqueeze(1).expand(-1, exposure_length, -1, -1, -1).flatten(0,1), mask=d_alpha_tensor.unsqueeze(1).expand(-1, exposure_length, -1, -1, -1).flatten(0,1)))
    # --------------------------
    flow_loss.backward()
    blob = {"ori": _np(ori_image_tensor), "latent": _np(latent_leaf), "exp2mid": _np(e2m_leaf), "mid2exp": _np(m2e_leaf),
            "latent_alpha": _np(la_leaf), "d_alpha": _np(da_leaf), "loss": _np(flow_loss),
            "g_latent": _np(latent_leaf.grad), "g_exp2mid": _np(e2m_leaf.grad), "g_mid2exp": _np(m2e_leaf.grad),
            "g_latent_alpha": _np(la_leaf.grad), "g_d_alpha": _np(da_leaf.grad), "g_ori": _np(ori_image_tensor.grad)}
    np.savez_compressed(os.path.join(out_dir, "flow_warp_loss.npz"), **blob)
    print("wrote flow_warp_loss", float(flow_loss))


def make_reg_loss_golden(out_dir):
    """reg_loss.npz: train.py:651-655 with the reference's own l1_loss / entropy_loss / sparsity_loss (CPU fp32)."""
    from utils.loss_utils import entropy_loss, l1_loss, sparsity_loss
    g = torch.Generator().manual_seed(31)
    B, H, W = 2, 19, 27
    depth_tensor = (1.0 + 4.0 * torch.rand(B, 1, H, W, generator=g)).requires_grad_(True)
    gt_depth_tensor = 1.0 + 4.0 * torch.rand(B, 1, H, W, generator=g)
    d_alpha_tensor = torch.rand(B, 1, H, W, generator=g)
    d_alpha_tensor[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1e-7, 1 - 1e-7])      # the epsilon matters at the ends
    d_alpha_tensor.requires_grad_(True)
    depth_loss = l1_loss(depth_tensor, gt_depth_tensor)
    reg = 0.2 * depth_loss
    mask_loss = 1e-7 * entropy_loss(d_alpha_tensor) + 1e-7 * sparsity_loss(d_alpha_tensor)
    reg = reg + mask_loss
    reg.backward()
    blob = {"depth": _np(depth_tensor), "gt_depth": _np(gt_depth_tensor), "d_alpha": _np(d_alpha_tensor), "reg": _np(reg),
            "depth_loss": _np(depth_loss), "entropy": _np(entropy_loss(d_alpha_tensor)), "sparsity": _np(sparsity_loss(d_alpha_tensor)),
            "g_depth": _np(depth_tensor.grad), "g_d_alpha": _np(d_alpha_tensor.grad)}
    np.savez_compressed(os.path.join(out_dir, "reg_loss.npz"), **blob)
    print("wrote reg_loss", float(reg))


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    if "--only-loss" in sys.argv:
        make_photo_loss_golden(out)
        sys.exit(0)
    if "--only-flow-warp" in sys.argv:
        make_flow_warp_golden(out)
        sys.exit(0)
    if "--only-reg" in sys.argv:
        make_reg_loss_golden(out)
        sys.exit(0)
    if "--only-rays" in sys.argv:
        make_camera_rays_golden(out)
        sys.exit(0)
    make_render_goldens(out)
    make_hexplane_golden(out)
    make_photo_loss_golden(out)
    make_camera_rays_golden(out)
    make_flow_warp_golden(out)
    make_reg_loss_golden(out)
