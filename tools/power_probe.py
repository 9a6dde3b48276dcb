"""Clocks and power while one kernel runs back to back for ~3 s: is the fused hinge pass power-cap bound?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import ClockSampler, synth_embeddings  # noqa: E402
from peppa_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n = 32768
a, v = synth_embeddings(n, 666, dev)
ra, _ = ops.row_norms(a)
rv, _ = ops.row_norms(v)
diag, thr = ops.sim_diag(a, v, ra, rv)
g, ld = ops.gmat_alloc(n, n, dev)
gh = torch.randn(n, n, device=dev).half()
rc = torch.zeros(n, dtype=torch.int32, device=dev)
cc = torch.zeros(n, dtype=torch.int32, device=dev)
rk = torch.zeros(n, dtype=torch.int32, device=dev)
idx = torch.arange(n, device=dev)
vh = ops.rows_scale_f16(v, rv)
cases = {
    "sim_rank": lambda: ops.sim_rank(a, v, ra, rv, thr, idx, rank=rk),
    "sim_hinge+rank+G": lambda: ops.sim_hinge(a, v, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk),
    "grad_gemm": lambda: ops.grad_gemm(g, n, n, ld, vh, transpose=False),
    "torch.matmul bf16 32768x32768x512": lambda: torch.matmul(a, v.T),
}
for name, fn in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with ClockSampler(0) as clk:
        t0 = time.time()
        it = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 3.0:
            for _ in range(50):
                fn()
            it += 50
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    pw = [float(r[2]) for r in clk.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
    s = clk.summary()
    print(f"{name:36s} {ms:7.3f} ms/launch  {2 * n * n * 512 / ms / 1e9:7.1f} TFLOP/s  sm clock median {s['sm_mhz']} MHz  "
          f"power mean {sum(pw) / max(1, len(pw)):.0f} W max {max(pw) if pw else 0:.0f} W  reasons {s['reasons']}", flush=True)
