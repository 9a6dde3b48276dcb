"""ncu target for the kernels outside the gallery step: rank pass (16384^2, BASELINE config 3), row log-sum-exp pass
and MIL-NCE gradient pass (north-star kernel a), encoder tail (2^20 rows), triplet kernel (2^20 triplets).
Profile with ``ncu --profile-from-start off`` (one launch of each sits in a profiler range)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
a16, v16 = synth_embeddings(16384, 666, dev)
a32, v32 = synth_embeddings(32768, 667, dev)
ra, _ = ops.row_norms(a16)
rv, _ = ops.row_norms(v16)
_, thr = ops.sim_diag(a16, v16, ra, rv)
idx = torch.arange(16384, device=dev)
g, ld = ops.gmat_alloc(32768, 32768, dev)
den = torch.full((32768,), 12.0, device=dev)
x = torch.randn(1 << 20, 512, device=dev).bfloat16()
w = (torch.randn(512, 512, device=dev) / 512 ** 0.5).bfloat16()
b = torch.randn(512, device=dev) * 0.1
t = 1 << 20
ta, tp, tn = (torch.randn(t, 512, device=dev).bfloat16() for _ in range(3))


def run():
    ops.sim_rank(a16, v16, ra, rv, thr, idx)
    ops.sim_lse_rows(a32, v32, scale=1.0 / 0.07)
    ops.sim_lse_grad(a32, v32, den, den, g, ld, scale=1.0 / 0.07)
    ops.project_normalize(x, w, b)
    ops.triplet_score(ta, tp, tn)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
