"""Timing of the encoder tail (Linear 512 -> 512 + normalise + bf16 + rinv) on 2^20 rows, 50 launches back to back."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gpu_probe import _t  # noqa: E402
from peppa_b200 import ops  # noqa: E402

n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, 512, generator=g, device="cuda").bfloat16()
w = (torch.randn(512, 512, generator=g, device="cuda") / 512 ** 0.5).bfloat16()
b = torch.randn(512, generator=g, device="cuda") * 0.1
for rep in range(3):
    ms = _t(lambda: ops.project_normalize(x, w, b), iters=50, warm=5)
    print(f"encoder tail, {n} rows: {ms:.4f} ms = {n / ms / 1e6:.3f} G rows/s", flush=True)
