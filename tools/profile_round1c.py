"""ncu target for the kernels changed in the third session of round 1: the one-pass row + column log-sum-exp
(pb2_sim_lse_both) and the hinge + rank pass with the select-tree column totals, on a 32768 x 32768 block.
Profile with ``ncu --profile-from-start off`` (one launch of each sits in a profiler range)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n = 32768
a, v = synth_embeddings(n, 667, dev)
ra, _ = ops.row_norms(a)
rv, _ = ops.row_norms(v)
diag, thr = ops.sim_diag(a, v, ra, rv)
g, ld = ops.gmat_alloc(n, n, dev)
rc = torch.zeros(n, dtype=torch.int32, device=dev)
cc = torch.zeros(n, dtype=torch.int32, device=dev)
rk = torch.zeros(n, dtype=torch.int32, device=dev)
bound = ops.logit_bound(a, v, 1.0 / 0.07)


def run():
    ops.sim_lse_both(a, v, bound, scale=1.0 / 0.07)
    ops.sim_hinge(a, v, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
