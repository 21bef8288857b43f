#!/usr/bin/env python
"""tests/golden/train_loop.npz — what the reference's REAL training loop computes.

Executes the UNMODIFIED `train.scene_reconstruction` (/root/reference/train.py:202-823) for a few iterations
on a synthetic Stereo-Blur-shaped scene made of the reference's own GaussianModel / Camera / blceKernel
objects (tests/ref_env.py), with `gsplat.rendering` served by the CPU oracle, and records — through wrappers
installed around `train.render`, `train.get_flow`, `torch.Tensor.backward` and `torch.optim.Adam.step`, no
reference file is touched — per iteration:

  * the parameters of both Gaussian models and the decoder at the start of the iteration,
  * every render() / get_flow() call the loop issued: camera uid, world-to-camera matrix, time, delta_exposure
    (i.e. the K warped sub-frame cameras blceKernel.get_warped_cams produced, train.py:472) and cheap checksums
    of what it returned,
  * the two scalars the loop backpropagates (photo_loss at :629, the regulariser sum at :680),
  * d loss / d parameter of every optimiser group at the moment the optimisers step (:796-800), the pose
    gradients of the warped cameras (d loss / d world_view_transform, d loss / d cam_ray) and the
    `viewspace_points` gradient the densification statistics read (:634-648).

tests/harness_train_loop.py replays those iterations on the CUDA path (drop-in render()/get_flow() call
pattern AND the fused render_blurry_view / get_flow_batched path) and tests/test_train_loop_gpu.py holds the
results to this file.  Runs only where /root/reference exists (the build container); ~20 s on CPU.

    python tests/golden/make_train_golden.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_env as E  # noqa: E402

ITERS, NUM_WARP, W, H = 3, 5, 64, 48
STAT_GROUPS = ("xyz", "f_dc", "opacity", "scaling", "rotation")
DYN_GROUPS = ("control_xyz", "f_dc", "f_t", "opacity", "scaling", "rotation", "omega")
ATTR_OF = {"xyz": "_xyz", "control_xyz": "control_xyz", "f_dc": "_features_dc", "f_t": "_features_t", "opacity": "_opacity",
           "scaling": "_scaling", "rotation": "_rotation", "omega": "_omega", "trbf_center": "_trbf_center"}


def _np(t):
    return t.detach().cpu().numpy().copy()


def main(out_path=None):
    E.setup_paths()
    blob = {}
    with E.cuda_to_cpu():
        import train as T
        from arguments import ModelParams, OptimizationParams, PipelineParams, blceParams
        from utils.timer import Timer
        torch.manual_seed(0); random.seed(0); np.random.seed(0)
        stat, dyn, scene, hyper = E.synthetic_reference_scene(W=W, H=H)
        dataset = E.group_args(ModelParams, model_path=scene.model_path, debug_process=False)
        opt = E.group_args(OptimizationParams, batch_size=2, lambda_flow_loss=1e-2)
        pipe = E.group_args(PipelineParams)
        blceopt = E.group_args(blceParams, num_warp=NUM_WARP)

        # static scene description the harness needs to rebuild the cameras
        for i, cam in enumerate(scene.train_cams):
            blob[f"cam{i}/image"] = _np(cam.original_image)
            blob[f"cam{i}/depth"] = _np(cam.depth)
            blob[f"cam{i}/mask"] = _np(cam.mask)
            blob[f"cam{i}/w2c"] = _np(cam.world_view_transform.transpose(0, 1))
            blob[f"cam{i}/K"] = _np(cam.K)
            blob[f"cam{i}/time"] = np.array(cam.time, np.float64)
            blob[f"cam{i}/cam_ray"] = _np(cam.cam_ray)
        blob["n_cams"] = np.array(len(scene.train_cams))
        blob["max_time"] = np.array(scene.train_cams[0].max_time)
        blob["num_warp"], blob["W"], blob["H"] = np.array(NUM_WARP), np.array(W), np.array(H)
        blob["lambda_dssim"], blob["lambda_flow_loss"] = np.array(opt.lambda_dssim), np.array(opt.lambda_flow_loss)

        state = {"it": 0, "calls": [], "backward": [], "live": []}

        def snapshot_params(it):
            for tag, pc, groups in (("stat", stat, STAT_GROUPS + ("control_xyz", "f_t", "omega", "trbf_center")),
                                    ("dyn", dyn, DYN_GROUPS + ("xyz", "trbf_center"))):
                for g in groups:
                    blob[f"it{it}/{tag}/{g}"] = _np(getattr(pc, ATTR_OF[g]))
                blob[f"it{it}/{tag}/current_control_num"] = _np(pc.current_control_num)
            blob[f"it{it}/dec/mlp1"] = _np(dyn.rgbdecoder.mlp1.weight)
            blob[f"it{it}/dec/mlp2"] = _np(dyn.rgbdecoder.mlp2.weight)

        real_render, real_get_flow = T.render, T.get_flow

        def render_rec(cam, *a, **k):
            if state["it"] == 0:          # first call of iteration 1 happens before any step
                state["it"] = 1
                snapshot_params(1)
            wvt = cam.world_view_transform
            if torch.is_tensor(wvt) and wvt.requires_grad:
                wvt.retain_grad()
            if torch.is_tensor(cam.cam_ray) and cam.cam_ray.requires_grad:
                cam.cam_ray.retain_grad()
            out = real_render(cam, *a, **k)
            de = k.get("delta_exposure")
            state["calls"].append(dict(kind="render", uid=cam.uid, w2c=_np(wvt.transpose(0, 1)), time=float(cam.time),
                                       delta=(None if de is None else float(de)), cam=cam, out=out,
                                       mean=float(out["render"].mean()), depth_mean=float(out["depth"].mean())))
            return out

        def get_flow_rec(cam, *a, **k):
            out = real_get_flow(cam, *a, **k)
            state["calls"].append(dict(kind="get_flow", uid=cam.uid, delta=float(k["delta_exposure"]),
                                       sums=[float(o.double().mean()) for o in out]))
            return out

        real_backward = torch.Tensor.backward

        def backward_rec(self, *a, **k):
            if self.dim() == 0:
                state["backward"].append(float(self.detach()))
            return real_backward(self, *a, **k)

        real_step = torch.optim.Adam.step
        stepped = {"n": 0}

        def step_rec(self, *a, **k):
            it = state["it"]
            names = [g.get("name") for g in self.param_groups]
            if "control_xyz" in names:                         # one of the two Gaussian models (stat steps first, :796-797)
                tag = "stat" if stepped["n"] % 2 == 0 else "dyn"
                stepped["n"] += 1
                for g in self.param_groups:
                    if len(g["params"]) == 1 and g["params"][0].grad is not None and g["name"] in ATTR_OF:
                        blob[f"it{it}/grad/{tag}/{g['name']}"] = _np(g["params"][0].grad)
                    if g["name"] == "decoder" and tag == "dyn":
                        blob[f"it{it}/grad/dec/mlp1"] = _np(g["params"][0].grad)
                        blob[f"it{it}/grad/dec/mlp2"] = _np(g["params"][1].grad)
                if tag == "dyn":
                    flush_iteration(it)
            return real_step(self, *a, **k)

        def flush_iteration(it):
            calls = state["calls"]
            r = [c for c in calls if c["kind"] == "render"]
            f = [c for c in calls if c["kind"] == "get_flow"]
            blob[f"it{it}/render/uid"] = np.array([c["uid"] for c in r])
            blob[f"it{it}/render/w2c"] = np.stack([c["w2c"] for c in r])
            blob[f"it{it}/render/time"] = np.array([c["time"] for c in r])
            blob[f"it{it}/render/delta"] = np.array([np.nan if c["delta"] is None else c["delta"] for c in r])
            blob[f"it{it}/render/mean"] = np.array([c["mean"] for c in r])
            blob[f"it{it}/render/depth_mean"] = np.array([c["depth_mean"] for c in r])
            g_w2c, g_ray, vsp = [], [], []
            for c in r:
                wvt = c["cam"].world_view_transform
                g = wvt.grad if (torch.is_tensor(wvt) and wvt.grad is not None) else torch.zeros(4, 4)
                g_w2c.append(_np(g.transpose(0, 1)))
                ray = c["cam"].cam_ray
                gr = ray.grad if (torch.is_tensor(ray) and ray.grad is not None) else torch.zeros_like(ray)
                g_ray.append(_np(gr.sum(dim=(0, 2, 3))))             # 6 sums: enough to pin the pose path
                v = c["out"]["viewspace_points"]
                vsp.append(_np(v.grad) if v.grad is not None else np.zeros(tuple(v.shape), np.float32))
            blob[f"it{it}/render/grad_w2c"] = np.stack(g_w2c)
            blob[f"it{it}/render/grad_cam_ray_sum"] = np.stack(g_ray)
            blob[f"it{it}/render/grad_viewspace"] = np.stack(vsp)
            blob[f"it{it}/get_flow/uid"] = np.array([c["uid"] for c in f])
            blob[f"it{it}/get_flow/delta"] = np.array([c["delta"] for c in f])
            blob[f"it{it}/get_flow/means"] = np.array([c["sums"] for c in f])
            blob[f"it{it}/backward"] = np.array(state["backward"])
            state["calls"], state["backward"] = [], []
            state["it"] = it + 1
            if it + 1 <= ITERS:
                # parameters at the start of the next iteration = after both steps; snapshot lazily at its first render
                state["pending_snapshot"] = it + 1

        def render_rec2(cam, *a, **k):
            if state.get("pending_snapshot"):
                snapshot_params(state.pop("pending_snapshot"))
            return render_rec(cam, *a, **k)

        T.render, T.get_flow = render_rec2, get_flow_rec
        torch.Tensor.backward = backward_rec
        torch.optim.Adam.step = step_rec
        try:
            timer = Timer(); timer.start()
            T.scene_reconstruction(dataset, opt, hyper, pipe, blceopt, [], [], [], None, -1, dyn, stat, scene, "fine", None,
                                   ITERS, timer)
        finally:
            T.render, T.get_flow = real_render, real_get_flow
            torch.Tensor.backward = real_backward
            torch.optim.Adam.step = real_step
    blob["iters"] = np.array(ITERS)
    out_path = out_path or os.path.join(HERE, "train_loop.npz")
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, f"{os.path.getsize(out_path) / 1e6:.2f} MB;",
          {k: blob[k] for k in blob if k.endswith("/backward")})
    return blob


EVAL_STEPS, EVAL_DECAY_START = 6, 2
EVAL_LR = 3e-4          # eval.py:258-260 (lr_p = lr_q = 0.0003, lr_final = 1e-6)


def make_eval_golden(out_path=None):
    """tests/golden/eval_tto.npz: the UNMODIFIED `eval.render_test_tto` (eval.py:43-166, test-time pose optimisation
    through render(w2c=...)) for one test view, EVAL_STEPS Adam steps, all Gaussians frozen as eval.py:246-255 does.
    Recorded: frozen parameters, the view (pose, K, time, ground-truth image as eval.py loads it from PNG), and per
    step the w2c handed to render(), the loss (-PSNR), d loss / d (t, q); plus the solved pose."""
    import inspect
    from PIL import Image
    E.setup_paths()
    blob = {}
    with E.cuda_to_cpu():
        import eval as EV
        from arguments import PipelineParams
        EV.models.PerceptualLoss = lambda *a, **k: None          # LPIPS (needs downloaded AlexNet weights) is built but unused in the loop
        torch.manual_seed(0); random.seed(0); np.random.seed(0)
        stat, dyn, scene, hyper = E.synthetic_reference_scene(W=W, H=H)
        for pc in (dyn, stat):                                   # eval.py:246-255
            for attr in inspect.getmembers(pc):
                try:
                    attr[1].requires_grad = False
                except Exception:  # noqa: BLE001
                    pass
        for tag, pc in (("stat", stat), ("dyn", dyn)):
            for g, attr in ATTR_OF.items():
                blob[f"{tag}/{g}"] = _np(getattr(pc, attr))
            blob[f"{tag}/current_control_num"] = _np(pc.current_control_num)
        blob["dec/mlp1"], blob["dec/mlp2"] = _np(dyn.rgbdecoder.mlp1.weight), _np(dyn.rgbdecoder.mlp2.weight)
        cam = scene.test_cams[2]
        gt_dir = os.path.join(scene.model_path, "inference_images")
        os.makedirs(gt_dir, exist_ok=True)
        img8 = (cam.original_image.permute(1, 2, 0).numpy() * 255).astype("uint8")
        Image.fromarray(img8).save(os.path.join(gt_dir, f"{cam.image_name}.png"))
        blob["gt_rgb"] = (img8 / 255.0).astype(np.float32)       # what eval.py:89-95 ends up with ([H,W,3])
        blob["w2c0"] = _np(cam.world_view_transform.transpose(0, 1))
        blob["K"], blob["time"], blob["max_time"] = _np(cam.K), np.array(cam.time), np.array(cam.max_time)
        blob["cam_ray"] = _np(cam.cam_ray)
        blob["W"], blob["H"] = np.array(W), np.array(H)
        blob["steps"], blob["decay_start"], blob["lr"], blob["lr_final"] = (np.array(EVAL_STEPS), np.array(EVAL_DECAY_START),
                                                                             np.array(EVAL_LR), np.array(1e-6))
        rec = {"w2c": [], "loss": [], "g_t": [], "g_q": []}
        real_render, real_backward, real_step = EV.render, torch.Tensor.backward, torch.optim.Adam.step

        def render_rec(camera, *a, **k):
            if k.get("w2c") is not None and k["w2c"].requires_grad:
                rec["w2c"].append(_np(k["w2c"]))
            return real_render(camera, *a, **k)

        def backward_rec(self, *a, **k):
            if self.dim() == 0:
                rec["loss"].append(float(self.detach()))
            return real_backward(self, *a, **k)

        def step_rec(self, *a, **k):
            ps = [g["params"][0] for g in self.param_groups]
            if len(ps) == 2 and tuple(ps[0].shape) == (3,) and tuple(ps[1].shape) == (4,):
                rec["g_t"].append(_np(ps[0].grad)); rec["g_q"].append(_np(ps[1].grad))
            return real_step(self, *a, **k)

        EV.render, torch.Tensor.backward, torch.optim.Adam.step = render_rec, backward_rec, step_rec
        try:
            pipe = E.group_args(PipelineParams)
            bg = torch.tensor([0] * 9 + [0], dtype=torch.float32)          # eval.py:222-223
            save_dir = os.path.join(scene.model_path, "eval_out")
            EV.render_test_tto(H=H, W=W, scene=scene, test_cams=[cam], save_dir=save_dir, gt_rgb_dir=gt_dir,
                               tto_steps=EVAL_STEPS, decay_start=EVAL_DECAY_START, lr_p=EVAL_LR, lr_q=EVAL_LR, lr_final=1e-6,
                               use_sgd=False, initialize_from_previous_camera=False, initialize_from_previous_step_factor=1,
                               initialize_from_previous_lr_factor=1.0, fg_mask_th=0.1, local_viewdirs=None, batch_shape=None,
                               renderArgs=[pipe, bg])
        finally:
            EV.render, torch.Tensor.backward, torch.optim.Adam.step = real_render, real_backward, real_step
        blob["solved_pose"] = np.load(os.path.join(save_dir, "solved_poses.npy"))[0]
        for k, v in rec.items():
            blob["step/" + k] = np.stack(v) if k != "loss" else np.array(v)
    out_path = out_path or os.path.join(HERE, "eval_tto.npz")
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, "losses", blob["step/loss"])
    return blob


if __name__ == "__main__":
    if "--only-eval" not in sys.argv:
        main()
    make_eval_golden()
