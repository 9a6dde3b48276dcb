"""A/B of the similarity pass's cluster variants inside one process: the 65536-clip MIL-NCE step (public API, fwd + bwd)
and the 16384^2 retrieval call with pb2_debug_sim_pair(0) (independent CTAs), (1) (CTA pairs, cta_group::2), (2)
(multicast clusters) and (-1) (the product's default), alternately."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from gpu_probe import _t  # noqa: E402
from peppa_b200 import _cabi, loss, metrics  # noqa: E402
_cabi.use_measurement_library()      # the pb2_debug_* selectors live in the measurement build only

lib = _cabi.lib()
dev = torch.device("cuda", 0)
a, v = synth_embeddings(65536, 666, dev)
a16, v16 = a[:16384].contiguous(), v[:16384].contiguous()
mod = loss.MILNCELoss(temperature=0.07)


def step():
    vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)
    mod(vv, aa).backward()


for rep in range(3):
    for mode in (0, 1, 2, -1):
        lib.pb2_debug_sim_pair(mode)
        ms = _t(step, iters=5, warm=2)
        ms_r = _t(lambda: metrics.recall_at_1_to_n(v16, a16, None, N=10), iters=20, warm=3)
        print(f"sim_pair={mode:2d}: MIL-NCE 65536 step {ms:.3f} ms   recall_at_1_to_n 16384^2 {ms_r:.4f} ms", flush=True)
lib.pb2_debug_sim_pair(-1)
