"""(GPU box) per-tensor error of the tcgen05 HexPlane+MLP backward against the oracle's autograd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_oracle import _hexplane_args
from mobgs_b200.deformation import HexPlaneMLP
from oracle.hexplane_ref import deform_forward_ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.manual_seed(3)
net = HexPlaneMLP(_hexplane_args(16))
with torch.no_grad():
    for name, p in net.named_parameters():
        if p.dim() == 1: p.uniform_(-0.2, 0.2)
        elif "grids" in name: p.add_(0.1 * torch.randn_like(p))
        elif p.requires_grad: torch.nn.init.xavier_uniform_(p)
net.set_aabb([1.3, 1.1, 1.2], [-1.2, -1.0, -1.4])
pts = (torch.rand(n, 3) * 2 - 1.0).requires_grad_(True)
scales = (torch.randn(n, 3) * 0.3 - 3).requires_grad_(True)
rots = torch.randn(n, 4).requires_grad_(True)
t = torch.rand(n, 1)
which = sys.argv[2] if len(sys.argv) > 2 else "all"
w = [torch.randn(n, 3), torch.randn(n, 3), torch.randn(n, 4)]
if which == "pts": w[1].zero_(); w[2].zero_()
if which == "scales": w[0].zero_(); w[2].zero_()
if which == "rots": w[0].zero_(); w[1].zero_()
outs = deform_forward_ref(net, pts, scales, rots, t)
sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
ref_in = [pts.grad.clone(), scales.grad.clone(), rots.grad.clone()]
net.zero_grad(set_to_none=True); net.cuda()
ins = [x.detach().cuda().requires_grad_(True) for x in (pts, scales, rots)]
got = net(ins[0], ins[1], ins[2], t.cuda())
for o, r, nm in zip(got, outs, ("pts", "scales", "rots")):
    print("fwd", nm, float((o.detach().cpu() - r.detach()).abs().max()))
sum((o * wi.cuda()).sum() for o, wi in zip(got, w)).backward()
torch.cuda.synchronize()
for g, r, nm in zip(ins, ref_in, ("pts", "scales", "rots")):
    print(f"in  {nm:8s} err {float((g.grad.cpu() - r).abs().max()):.3e}  scale {float(r.abs().max()):.3e}")
for k, p in net.named_parameters():
    if k in ref:
        e = float((p.grad.cpu() - ref[k]).abs().max()); s = float(ref[k].abs().max())
        print(f"par {k:50s} err {e:.3e} scale {s:.3e} {'BAD' if e > 1e-3 * max(s, 1e-6) else ''}")
