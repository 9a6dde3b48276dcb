"""Small fixed workload for ncu: one gallery step (loss fwd+bwd + recall@1..10) on a 32768 x 32768
gallery = 1 sim_hinge+rank launch and 2 grad_gemm launches per step; 1 warm-up step + 1 step.
``python tools/profile_target.py [n] [steps]``"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200 import _cabi  # noqa: E402

if os.environ.get("PB2_SIM_PAIR"):       # measurement build: pick the similarity pass's cluster variant (0 / 1 / 2)
    _cabi.use_measurement_library().pb2_debug_sim_pair(int(os.environ["PB2_SIM_PAIR"]))
from peppa_b200.gallery import GalleryStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
a, v = synth_embeddings(n, 666, dev)
step = GalleryStep(n, 512, device=dev)
for _ in range(1 + steps):
    out = step.run(a, v)
torch.cuda.synchronize()
print("loss", out["loss"].item(), "R@10", out["recall"][10].item())
