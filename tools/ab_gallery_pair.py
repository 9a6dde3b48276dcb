"""A/B of the similarity pass's cluster variants INSIDE the real gallery step (131072 clips = 4 x 4 blocks, one-byte
gradient matrix, kind::i8 products), alternating in one process on one board (measurement build):
  pair=0 independent CTAs, pair=1 CTA pairs on one M = 256 MMA (the product), pair=2 clusters of 2 with a multicast Y tile,
  pair=3 CTA pairs with a resident X strip
    python tools/ab_gallery_pair.py [modes, e.g. 1,3] [clips]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from gpu_probe import _t  # noqa: E402
from peppa_b200 import _cabi  # noqa: E402

lib = _cabi.use_measurement_library()
from peppa_b200.gallery import GalleryStep  # noqa: E402

modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
dev = torch.device("cuda", 0)
a, v = synth_embeddings(n, 666, dev)
step = GalleryStep(n, 512, device=dev)
ref = None
for rep in range(3):
    for mode in modes:
        lib.pb2_debug_sim_pair(mode)
        out = step.run(a, v)
        chk = (out["loss"].item(), int(out["ranks"].sum()), float(out["dA"].abs().sum()))
        ref = ref or chk
        ms = _t(lambda: step.run(a, v), iters=8, warm=2)
        print(f"pair={mode} gallery {n}: {ms:.2f} ms/step  same results as pair={modes[0]}: {chk[1:] == ref[1:] and abs(chk[0] - ref[0]) <= 1e-6 * abs(ref[0])}", flush=True)
lib.pb2_debug_sim_pair(-1)
