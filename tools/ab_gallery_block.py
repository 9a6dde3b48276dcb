"""Block edge of the gallery step: does a gradient-matrix block small enough to stay in L2 pay?  131072 clips, hinge
step (one-byte gradient matrix, kind::i8 products), block edges alternating in one process on one board.
    python tools/ab_gallery_block.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from gpu_probe import _t  # noqa: E402
from peppa_b200.gallery import GalleryStep  # noqa: E402

n = 131072
dev = torch.device("cuda", 0)
a, v = synth_embeddings(n, 666, dev)
steps = {b: GalleryStep(n, 512, device=dev, block=b) for b in (32768, 16384, 8192, 4096)}
for rep in range(2):
    for b, step in steps.items():
        ms = _t(lambda: step.run(a, v), iters=6, warm=2)
        out = step.run(a, v)
        print(f"block {b:5d} ({(n // b) ** 2:4d} blocks, G block {b * b / 2**20:6.0f} MiB): {ms:7.2f} ms/step   loss {out['loss'].item():.9f}", flush=True)
