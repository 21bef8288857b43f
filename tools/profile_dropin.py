"""Host-side profile of the drop-in render() at the reference's training shape (cProfile)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mobgs_b200.gaussian_renderer import render  # noqa: E402
from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene  # noqa: E402

dev = torch.device("cuda")
stat, dyn, intr = synthetic_scene(100_000, 50_000, 512, 288, device=dev)
cam = make_camera(intr, subframe_w2c(0, 1, device=dev))
bg = torch.zeros(3, device=dev)
tgt = torch.rand(3, 288, 512, device=dev)
params = [p for pc in (stat, dyn) for p in pc.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    out = render(cam, stat, dyn, None, bg, get_static=True, get_dynamic=True, delta_exposure=0.2)
    (out["render"] - tgt).abs().mean().backward()


for _ in range(10):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
print("ms per render fwd+bwd:", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
