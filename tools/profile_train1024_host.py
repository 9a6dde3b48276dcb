"""Host-side profile of the batch-1k training step through the public API (config 2 is launch / host bound):
cProfile over 3000 eager steps, top functions by own time, plus the wall time per step with and without sync."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200.loss import TripletLoss  # noqa: E402

dev = torch.device("cuda", 0)
a, v = synth_embeddings(1024, 666, dev)
mod = TripletLoss(0.2)
vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)


def step():
    vv.grad = None
    aa.grad = None
    loss = mod(vv, aa)
    loss.backward()
    return loss


for _ in range(50):
    step()
torch.cuda.synchronize()
n = 3000
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step (launch only) {1e6 * (t1 - t0) / n:.1f} us; with final sync {1e6 * (t2 - t0) / n:.1f} us")
# forward only / no autograd
with torch.no_grad():
    t0 = time.perf_counter()
    for _ in range(n):
        mod(vv, aa)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
print(f"forward without grad (no fused step: separate kernels) host {1e6 * (t1 - t0) / n:.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
