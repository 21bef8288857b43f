"""Host-side profile (cProfile) of the K-batched train step at the reference's training shape (sb_150k_512x288_K9),
where the step is host-bound: which Python lines cost what per step."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mobgs_b200 import _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sb_150k_512x288_K9"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_lib.load()
job = bench.GpuJob(wl, dev, 0, 1, False)
for _ in range(10):
    job.step(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    job.step(True)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {(t1 - t0) * 10:.3f} ms/step, with final sync {(t2 - t0) * 10:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    job.step(True)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(35)
