"""A/B of the encoder tail's variants (pb2_debug_proj_variant of the measurement build: 0 the product's choice = resident x
with 256-column phases for this shape, 1 interleaved kernel, 2 column split, 3 phased halves, 4 sixteen epilogue warps, 5
resident x with 128-column phases) on 2^20 rows 512 -> 512: variants alternate in one process on one board, 30 launches
back to back per figure, minimum (~alone) and median (~sustained) over the rounds.
    python tools/ab_tail.py [variants, default 1,0] [rounds, default 8]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gpu_probe import _t  # noqa: E402
from peppa_b200 import _cabi, ops  # noqa: E402

lib = _cabi.use_measurement_library()
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "0"])]
n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, 512, generator=g, device="cuda").bfloat16()
w = (torch.randn(512, 512, generator=g, device="cuda") / 512 ** 0.5).bfloat16()
b = torch.randn(512, generator=g, device="cuda") * 0.1
ref = None
times = {v: [] for v in variants}
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for rep in range(reps):
    for v in variants:
        lib.pb2_debug_proj_variant(v)
        out, rinv, nrm = ops.project_normalize(x, w, b)
        if ref is None:
            ref = out.float()
        dmax = (out.float() - ref).abs().max().item() if rep == 0 else float("nan")
        ms = _t(lambda: ops.project_normalize(x, w, b), iters=30, warm=3)
        times[v].append(ms)
        if rep == 0:
            print(f"variant {v}: max |out - variant {variants[0]}| {dmax:.2e}", flush=True)
lib.pb2_debug_proj_variant(0)
for v in variants:
    ts = sorted(times[v])
    med = ts[len(ts) // 2]
    print(f"variant {v}: median {med:.4f} ms (min {ts[0]:.4f}, max {ts[-1]:.4f}) = {2.0 * n * 512 * 512 / med / 1e9:.0f} TF/s, "
          f"{n * (1024 + 1024 + 8) / med / 1e9:.2f} TB/s", flush=True)
