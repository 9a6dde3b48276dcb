"""ncu target: the one-pass MIL-NCE statistics (pb2_sim_lse_both) and the same pass with the recall ranks fused in
(pb2_sim_lse_both_rank) on a 32768 x 32768 block, temperature 0.07, one profiled launch each
(``ncu --profile-from-start off``)."""
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
_, thr = ops.sim_diag(a, v, ra, rv)
bound = ops.logit_bound(a, v, 1.0 / 0.07)
cnt = torch.zeros(n, dtype=torch.int32, device=dev)


def run():
    ops.sim_lse_both(a, v, bound, scale=1.0 / 0.07)
    cnt.zero_()
    ops.sim_lse_both(a, v, bound, scale=1.0 / 0.07, rank=(ra, rv, thr, 0, 0, cnt))


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
