"""Debug aid: per-output-key gradient error of the drop-in render() vs the fp32 and fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import mobgs_ref as M
from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
from mobgs_b200.gaussian_renderer import render

def to64(pc):
    for n in pc.PARAM_NAMES:
        t = getattr(pc, n); setattr(pc, n, t.detach().double().requires_grad_(True))
    if pc.rgbdecoder is not None: pc.rgbdecoder.double()
    return pc

ns, nd, W, H = 700, 300, 128, 128
so, do, intr = synthetic_scene(ns, nd, W, H)
s64, d64, _ = synthetic_scene(ns, nd, W, H); to64(s64); to64(d64)
sc, dc, _ = synthetic_scene(ns, nd, W, H, device="cuda")
cam_o = make_camera(intr, subframe_w2c(0, 1))
cam_c = make_camera(intr, subframe_w2c(0, 1, device="cuda"))
cam_64 = make_camera(intr, subframe_w2c(0, 1)); cam_64.world_view_transform = cam_64.world_view_transform.double(); cam_64.K = cam_64.K.double(); cam_64.cam_ray = cam_64.cam_ray.double()
bg = torch.tensor([0.2, 0.5, 0.7, 1.0])
kw = dict(get_static=True, get_dynamic=True)
oc = render(cam_c, sc, dc, None, bg.cuda(), **kw)
oo = M.render_ref(cam_o, so, do, None, bg, **kw)
o64 = M.render_ref(cam_64, s64, d64, None, bg.double(), **kw)
g = torch.Generator().manual_seed(0)
for key in ["render", "s_render", "d_render", "depth", "d_depth", "d_alpha", "s_alpha"]:
    w = torch.rand(oo[key].shape, generator=g)
    for pcs in ((sc, dc), (so, do), (s64, d64)):
        for pc in pcs:
            for n in pc.PARAM_NAMES:
                getattr(pc, n).grad = None
    (oc[key] * w.cuda()).sum().backward(retain_graph=True)
    (oo[key] * w).sum().backward(retain_graph=True)
    (o64[key] * w.double()).sum().backward(retain_graph=True)
    fwd = (oc[key].detach().cpu() - o64[key].detach()).abs().max().item()
    fwd32 = (oo[key].detach() - o64[key].detach()).abs().max().item()
    line = f"{key:9s} fwd err cuda={fwd:.2e} cpu32={fwd32:.2e} |"
    for name, pc, po, p64 in (("s_xyz", sc, so, s64), ("d_ctrl", dc, do, d64)):
        attr = "_xyz" if name == "s_xyz" else "control_xyz"
        gc, go, g6 = getattr(pc, attr).grad, getattr(po, attr).grad, getattr(p64, attr).grad
        if g6 is None or gc is None:
            line += f" {name}: none |"; continue
        sc_ = g6.abs().max().item() + 1e-30
        ec = (gc.cpu().double() - g6).abs(); eo = (go.double() - g6).abs()
        line += f" {name}: scale={sc_:.2e} cuda max={ec.max().item()/sc_:.2e} n>1e-4={(ec > 1e-4*sc_).sum().item()} cpu32 max={eo.max().item()/sc_:.2e} n>1e-4={(eo > 1e-4*sc_).sum().item()} |"
    print(line)
