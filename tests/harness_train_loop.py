"""Replay of the reference's training-loop body on a chosen backend (test infrastructure).

tests/golden/train_loop.npz holds what the UNMODIFIED `train.scene_reconstruction` computed on a synthetic
scene (tests/golden/make_train_golden.py).  /root/reference does not travel to the GPU box, so this module
restates the loop body of train.py:430-680 — per view: centre render (:441), the K-1 warped sub-frame renders
with their exposure offsets (:502-516), the blur mean (:540-541), K get_flow calls (:563-579); then
photo_loss.backward(retain_graph=True) (:621-629), the regulariser sum (:651-676) and loss.backward() (:680) —
over stand-in model / camera objects loaded from the golden, importing nothing from the reference.  Backends:

  "oracle"  CPU: oracle.mobgs_ref.render_ref / get_flow_ref + oracle.loss_ref — must reproduce the golden to
            float rounding, which proves that this restatement IS the reference's loop body (run on CPU here);
  "dropin"  CUDA: mobgs_b200.gaussian_renderer.render / get_flow called exactly as train.py calls them
            (K render() calls, K get_flow() calls per view) + the fused losses;
  "fused"   CUDA: mobgs_b200.subframes.render_blurry_view + gaussian_renderer.get_flow_batched + fused losses —
            the single-launch-chain form INTEGRATION.md §1 gives a maintainer.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_loop.npz")
STAT_GROUPS = ("xyz", "f_dc", "opacity", "scaling", "rotation")
DYN_GROUPS = ("control_xyz", "f_dc", "f_t", "opacity", "scaling", "rotation", "omega")
ATTR_OF = {"xyz": "_xyz", "control_xyz": "control_xyz", "f_dc": "_features_dc", "f_t": "_features_t", "opacity": "_opacity",
           "scaling": "_scaling", "rotation": "_rotation", "omega": "_omega", "trbf_center": "_trbf_center"}


def load_models(z, it, device):
    """Stand-in GaussianSets holding the parameters both reference models had at the start of iteration `it`."""
    from mobgs_b200.scene import GaussianSet, Sandwich
    out = []
    for tag in ("stat", "dyn"):
        pc = GaussianSet()
        for g, attr in ATTR_OF.items():
            t = torch.from_numpy(z[f"it{it}/{tag}/{g}"]).to(device)
            setattr(pc, attr, t.requires_grad_(True))
        pc.current_control_num = torch.from_numpy(z[f"it{it}/{tag}/current_control_num"]).to(device)
        out.append(pc)
    stat, dyn = out
    dec = Sandwich().to(device)
    with torch.no_grad():
        dec.mlp1.weight.copy_(torch.from_numpy(z[f"it{it}/dec/mlp1"]))
        dec.mlp2.weight.copy_(torch.from_numpy(z[f"it{it}/dec/mlp2"]))
    dyn.rgbdecoder = dec
    stat.rgbdecoder = Sandwich().to(device)
    return stat, dyn


def make_cam(z, uid, w2c, time, device, pose_grad):
    """Camera stand-in with the attribute API the renderer reads; cam_ray as scene/cameras.py:140-146 builds it
    for a warped camera: origin = inverse(world_view_transform)[3,:3] (differentiable), directions rotated by the
    separately held camera-to-world rotation `R` (cameras.py:276-279), normalised twice (:266, :282)."""
    W, H = int(z["W"]), int(z["H"])
    Kmat = torch.from_numpy(z[f"cam{uid}/K"]).to(device)
    w2c = torch.as_tensor(w2c, dtype=torch.float32, device=device).clone().requires_grad_(pose_grad)
    c2w = torch.inverse(w2c)
    centre = c2w[:3, 3]
    rot = w2c[:3, :3].transpose(0, 1).detach()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device) + 0.5,
                            torch.arange(W, dtype=torch.float32, device=device) + 0.5, indexing="ij")
    local = torch.stack([(xs - Kmat[0, 2]) / Kmat[0, 0], (ys - Kmat[1, 2]) / Kmat[1, 1], torch.ones_like(xs)], dim=-1)
    local = local / torch.norm(local, dim=-1, keepdim=True)
    dirs = torch.matmul(rot, local.reshape(-1, 3)[..., None])[..., 0]
    dirs = (dirs / torch.norm(dirs, dim=-1, keepdim=True)).view(H, W, 3)
    origin, _ = torch.broadcast_tensors(centre, dirs)
    cam_ray = torch.cat((origin, dirs), dim=-1).permute(2, 0, 1).unsqueeze(0)
    if cam_ray.requires_grad:
        cam_ray.retain_grad()

    def get_pixels(w, h, use_center=None):
        xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
        return np.stack([xx, yy], axis=-1) + (0.5 if use_center else 0)

    return SimpleNamespace(uid=uid, w2c=w2c, world_view_transform=w2c.transpose(0, 1), K=Kmat, time=float(time),
                           max_time=int(z["max_time"]), image_width=W, image_height=H, cam_ray=cam_ray,
                           get_pixels=get_pixels, camera_center=centre)


def run_iteration(z, it, backend):
    """-> dict: photo, reg (the two backpropagated scalars), grads {stat/<g>, dyn/<g>, dec/mlp1, dec/mlp2},
    grad_w2c [calls,4,4], grad_cam_ray_sum [calls,6], grad_viewspace [calls,1,N,2], render_mean, depth_mean,
    flow_means [K calls,4] — in the reference's call order."""
    device = "cpu" if backend == "oracle" else "cuda"
    stat, dyn = load_models(z, it, device)
    K, W, H = int(z["num_warp"]), int(z["W"]), int(z["H"])
    half = K // 2
    lam_dssim, lam_flow = float(z["lambda_dssim"]), float(z["lambda_flow_loss"])
    bg = torch.tensor([0.0, 0.0, 0.0, -10.0], device=device)              # train.py:228 (fine stage, black background)
    uids, w2cs = z[f"it{it}/render/uid"], z[f"it{it}/render/w2c"]
    times, deltas = z[f"it{it}/render/time"], z[f"it{it}/render/delta"]
    n_views = len(uids) // K

    if backend == "oracle":
        from oracle import loss_ref as LR
        from oracle import mobgs_ref as M
        render, get_flow = M.render_ref, M.get_flow_ref

        def photo_fn(img, gt): return LR.photo_loss(img, gt, lam_dssim)
        def reg_fn(depth, gt_depth, d_alpha): return LR.reg_loss(depth, gt_depth, d_alpha)
        flow_fn = LR.flow_warp_loss
    else:
        from mobgs_b200 import losses as LS
        from mobgs_b200.gaussian_renderer import get_flow, get_flow_batched, render
        from mobgs_b200.subframes import render_blurry_view

        def photo_fn(img, gt): return LS.photo_loss(img, gt, lam_dssim)
        def reg_fn(depth, gt_depth, d_alpha): return LS.reg_loss(depth, gt_depth, d_alpha, 0.2, 1e-7)[0]
        flow_fn = LS.flow_warp_loss

    cams_all, vsps = [], []
    images, depth_list, d_alphas, oris, gts, gt_depths = [], [], [], [], [], []
    lat_img, lat_alpha, e2m, m2e = [], [], [], []
    render_mean, depth_mean, flow_means = [], [], []
    for v in range(n_views):
        calls = range(v * K, (v + 1) * K)                # call order per view: centre, then k = 0..K-1 without `half`
        uid = int(uids[v * K])
        centre = make_cam(z, uid, w2cs[v * K], times[v * K], device, pose_grad=False)
        warped, expo = {}, {}
        ks = [k for k in range(K) if k != half]
        for c, k in zip(list(calls)[1:], ks):
            warped[k] = make_cam(z, uid, w2cs[c], times[c], device, pose_grad=True)
            expo[k] = float(deltas[c])
        gts.append(torch.from_numpy(z[f"cam{uid}/image"]).to(device)[None])
        gt_depths.append(torch.from_numpy(z[f"cam{uid}/depth"]).to(device)[None])
        cams_all += [centre] + [warped[k] for k in ks]

        if backend == "fused":
            cam_list = [centre if k == half else warped[k] for k in range(K)]
            et = torch.tensor([0.0 if k == half else expo[k] for k in range(K)], device=device)
            pkg = render_blurry_view(centre, cam_list, et, stat, dyn, None, bg)
            pred_image, image_ori, pred_depth, d_alpha = pkg["render"], pkg["render_center"], pkg["depth"], pkg["d_alpha"]
            vsps += [pkg["viewspace_points"]] + [None] * (K - 1)
            render_mean += [float(pkg["subframes"][half].mean())] + [float(pkg["subframes"][k].mean()) for k in ks]
            depth_mean += [float(pkg["depths"][half].mean())] + [float(pkg["depths"][k].mean()) for k in ks]
        else:
            pkg = render(centre, stat, dyn, None, bg, stage="fine", cam_type="nvidia", get_static=True, get_dynamic=True,
                         iter_fact=it, ref_wc=None, flow=None, target_ts=None, target_w2cs=None)
            image_ori, pred_depth, d_alpha = pkg["render"], pkg["depth"], pkg["d_alpha"]
            if pkg["viewspace_points"].requires_grad and not pkg["viewspace_points"].is_leaf:
                pkg["viewspace_points"].retain_grad()
            vsps.append(pkg["viewspace_points"])
            render_mean.append(float(image_ori.mean())); depth_mean.append(float(pred_depth.mean()))
            rendered = []
            for k in range(K):
                if k == half:
                    rendered.append(image_ori)
                    continue
                p2 = render(warped[k], stat, dyn, None, bg, stage="fine", cam_type="nvidia", get_static=True,
                            get_dynamic=True, iter_fact=it, ref_wc=None, flow=None, target_ts=None, target_w2cs=None,
                            delta_exposure=expo[k])
                if p2["viewspace_points"].requires_grad and not p2["viewspace_points"].is_leaf:
                    p2["viewspace_points"].retain_grad()
                rendered.append(p2["render"])
                vsps.append(p2["viewspace_points"])
                render_mean.append(float(p2["render"].mean())); depth_mean.append(float(p2["depth"].mean()))
            pred_image = torch.mean(torch.stack(rendered, dim=0), dim=0) + 1e-10

        flow_deltas = [1.0 * (k - half) / half for k in range(K)]           # train.py:566-568
        if backend == "fused":
            a, b, li, la = get_flow_batched(centre, stat, dyn, None, bg, flow_deltas)
            la = la[:, None]
        else:
            outs = [get_flow(centre, stat, dyn, None, bg, delta_exposure=d) for d in flow_deltas]
            a = torch.cat([o[0] for o in outs]); b = torch.cat([o[1] for o in outs])
            li = torch.cat([o[2].unsqueeze(0) for o in outs]); la = torch.cat([o[3].unsqueeze(0) for o in outs])
        for k in range(K):
            flow_means.append([float(a[k].double().mean()), float(b[k].double().mean()), float(li[k].double().mean()),
                               float(la[k].double().mean())])
        e2m.append(a[None]); m2e.append(b[None]); lat_img.append(li[None]); lat_alpha.append(la[None])
        images.append(pred_image[None]); depth_list.append(pred_depth[None]); d_alphas.append(d_alpha[None])
        oris.append(image_ori[None])

    image_tensor, gt_image = torch.cat(images), torch.cat(gts)
    photo = photo_fn(image_tensor, gt_image[:, :3])
    photo.backward(retain_graph=True)                                              # train.py:629
    vsp_grads = [None if v is None or v.grad is None else v.grad.detach().clone() for v in vsps]   # :634-648
    depth_tensor, gt_depth, d_alpha_tensor = torch.cat(depth_list), torch.cat(gt_depths), torch.cat(d_alphas)
    reg = reg_fn(depth_tensor, gt_depth, d_alpha_tensor)
    ori_t, lat_t = torch.cat(oris), torch.cat(lat_img)
    flow = flow_fn(ori_t, lat_t, torch.cat(e2m).clone(), torch.cat(m2e).clone(), torch.cat(lat_alpha), d_alpha_tensor)
    reg = reg + lam_flow * flow
    reg.backward()                                                                 # train.py:680
    if device == "cuda":
        torch.cuda.synchronize()

    grads = {}
    for tag, pc, groups in (("stat", stat, STAT_GROUPS), ("dyn", dyn, DYN_GROUPS)):
        for g in groups:
            t = getattr(pc, ATTR_OF[g]).grad
            grads[f"{tag}/{g}"] = None if t is None else t.detach().cpu().numpy()
    grads["dec/mlp1"] = dyn.rgbdecoder.mlp1.weight.grad.detach().cpu().numpy()
    grads["dec/mlp2"] = dyn.rgbdecoder.mlp2.weight.grad.detach().cpu().numpy()
    g_w2c = np.stack([(c.w2c.grad.detach().cpu().numpy() if c.w2c.grad is not None else np.zeros((4, 4), np.float32))
                      for c in cams_all])
    g_ray = np.stack([(c.cam_ray.grad.detach().sum(dim=(0, 2, 3)).cpu().numpy() if c.cam_ray.grad is not None
                       else np.zeros(6, np.float32)) for c in cams_all])
    # the densification statistics read the viewspace gradient as it stands after BOTH backward passes (the list
    # holds the same tensors), so report the final .grad like the golden does
    vsp_final = [None if v is None or v.grad is None else v.grad.detach().cpu().numpy() for v in vsps]
    return dict(photo=float(photo.detach()), reg=float(reg.detach()), grads=grads, grad_w2c=g_w2c, grad_cam_ray_sum=g_ray,
                grad_viewspace=vsp_final, grad_viewspace_after_photo=vsp_grads, render_mean=np.array(render_mean),
                depth_mean=np.array(depth_mean), flow_means=np.array(flow_means))


EVAL_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_tto.npz")


def run_eval_tto(z, backend):
    """eval.py:96-150 (render_test_tto's per-view body) restated: Adam on (t, q) of the world-to-camera pose through
    render(cam, stat, dyn, pipe, bg, stage="fine", get_static=False, get_dynamic=False, w2c=curr_w2c), loss = -PSNR,
    CosineAnnealingLR after `decay_start`; all Gaussian tensors frozen (eval.py:246-255).
    -> dict(w2c [steps,4,4], loss [steps], g_t [steps,3], g_q [steps,4], solved_pose [4,4])."""
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    from pytorch3d.transforms import matrix_to_quaternion, quaternion_to_matrix   # compat/ stand-in (wxyz)
    from mobgs_b200.scene import GaussianSet, Sandwich
    device = "cpu" if backend == "oracle" else "cuda"
    models = []
    for tag in ("stat", "dyn"):
        pc = GaussianSet()
        for g, attr in ATTR_OF.items():
            setattr(pc, attr, torch.from_numpy(z[f"{tag}/{g}"]).to(device))          # frozen: requires_grad False
        pc.current_control_num = torch.from_numpy(z[f"{tag}/current_control_num"]).to(device)
        models.append(pc)
    stat, dyn = models
    dec = Sandwich().to(device)
    with torch.no_grad():
        dec.mlp1.weight.copy_(torch.from_numpy(z["dec/mlp1"])); dec.mlp2.weight.copy_(torch.from_numpy(z["dec/mlp2"]))
    for p in dec.parameters():
        p.requires_grad_(False)
    dyn.rgbdecoder, stat.rgbdecoder = dec, dec
    zz = {"W": z["W"], "H": z["H"], "max_time": z["max_time"], "cam0/K": z["K"]}
    cam = make_cam(zz, 0, z["w2c0"], float(z["time"]), device, pose_grad=False)
    if backend == "oracle":
        from oracle.mobgs_ref import render_ref as render
    else:
        from mobgs_b200.gaussian_renderer import render
    gt_rgb = torch.from_numpy(z["gt_rgb"]).to(device).float()
    bg = torch.zeros(10, device=device)
    steps, decay_start, lr, lr_final = int(z["steps"]), int(z["decay_start"]), float(z["lr"]), float(z["lr_final"])
    T_bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)
    w2c = cam.world_view_transform.transpose(0, 1).detach()
    t_init = torch.nn.Parameter(w2c[:3, 3].clone().detach(), requires_grad=True)
    q_init = torch.nn.Parameter(matrix_to_quaternion(w2c[:3, :3]).clone().detach(), requires_grad=True)
    optimizer = torch.optim.Adam([{"params": t_init, "lr": lr}, {"params": q_init, "lr": lr}])
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=steps - decay_start, eta_min=lr_final)
    out = {"w2c": [], "loss": [], "g_t": [], "g_q": []}
    for _step in range(steps):
        optimizer.zero_grad()
        curr = torch.cat([quaternion_to_matrix(q_init), t_init[:, None]], 1)
        curr = torch.cat([curr, T_bottom[None]], 0)
        pkg = render(cam, stat, dyn, None, bg, stage="fine", get_static=False, get_dynamic=False, cam_type="nvidia", w2c=curr)
        pred = pkg["render"].permute(1, 2, 0)
        mse = ((pred - gt_rgb) ** 2).mean()
        loss = -(20 * torch.log10(1.0 / torch.sqrt(mse)))
        loss.backward()
        out["w2c"].append(curr.detach().cpu().numpy()); out["loss"].append(float(loss.detach()))
        out["g_t"].append(t_init.grad.detach().cpu().numpy().copy()); out["g_q"].append(q_init.grad.detach().cpu().numpy().copy())
        optimizer.step()
        if _step >= decay_start:
            scheduler.step()
    solved = torch.cat([torch.cat([quaternion_to_matrix(q_init), t_init[:, None]], 1), T_bottom[None]], 0)
    return {k: np.stack(v) if k != "loss" else np.array(v) for k, v in out.items()} | {"solved_pose": solved.detach().cpu().numpy()}
