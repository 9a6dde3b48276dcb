"""A/B of the similarity pass's cluster variants inside the HEADLINE step (2^20 x 2^20 gallery, hinge fwd + bwd +
recall), alternating in one process on one board (measurement build): pb2_debug_sim_pair 0 = independent CTAs,
1 = CTA pairs (cta_group::2), -1 = the product's default.
    python tools/ab_gallery_1m.py [steps per figure] [rounds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import rank_hash_terms, synth_embeddings  # noqa: E402
from gpu_probe import _t  # noqa: E402
from peppa_b200 import _cabi  # noqa: E402

lib = _cabi.use_measurement_library()
from peppa_b200.gallery import GalleryStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << 20
dev = torch.device("cuda", 0)
a, v = synth_embeddings(n, 666, dev)
step = GalleryStep(n, 512, device=dev)
step.run(a, v)
for rep in range(rounds):
    for mode in (0, 1, -1):
        lib.pb2_debug_sim_pair(mode)
        ms = _t(lambda: step.run(a, v), iters=steps, warm=1)
        out = step.run(a, v)
        h = int(rank_hash_terms(out["ranks"], 0).sum()) & 0xFFFFFFFFFFFFFFFF
        print(f"pair={mode:2d} gallery {n}: {ms:.1f} ms/step  loss {out['loss'].item():.9f} rank_hash {h:016x}", flush=True)
lib.pb2_debug_sim_pair(-1)
