"""Host profile of the reference's evaluation call on a 1467-clip gallery (bench.py --workload eval1467):
resampled_recall_at_1_to_n(size=100, n_samples=500, N=10) + score_triplets(n_samples=500), cProfile over 5 calls."""
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200 import metrics, triplet  # noqa: E402

dev = torch.device("cuda", 0)
n = 1467
a, v = synth_embeddings(n, 666, dev)
g = torch.Generator().manual_seed(5)
dur = torch.randint(20, 60, (n,), generator=g).float() / 10.0


def step():
    torch.manual_seed(666)
    random.seed(666)
    rec = metrics.resampled_recall_at_1_to_n(v, a, size=100, n_samples=500, N=10)
    acc = triplet.score_triplets(v, a, dur, n_samples=500)["accuracy"]
    return rec, acc


for _ in range(3):
    step()
torch.cuda.synchronize()
t = time.time()
for _ in range(10):
    step()
torch.cuda.synchronize()
print(f"eval1467: {(time.time() - t) / 10 * 1e3:.1f} ms per evaluation (native sampler: {triplet._native_sampler()})")
for name, fn in (("resampled_recall_at_1_to_n", lambda: metrics.resampled_recall_at_1_to_n(v, a, size=100, n_samples=500, N=10)),
                 ("score_triplets", lambda: triplet.score_triplets(v, a, dur, n_samples=500))):
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(f"  {name}: {(time.time() - t) / 10 * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
