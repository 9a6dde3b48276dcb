"""A/B of the gradient product of one 32768 x 32768 gradient-matrix block x [32768, 512] embeddings, timed
ALTERNATELY in one process on one board (each figure: `reps` launches back to back, i.e. at the board's power cap like
the gallery step; rounds interleave the variants so that clock drift hits them equally):

  i8     pb2_grad_gemm_ws, one-byte G x two 8-bit planes, tcgen05 kind::i8 (this round's product path)
  f16    pb2_grad_gemm_ws, fp16 G x fp16 embeddings, tcgen05 kind::f16 (round 1's path; still MIL-NCE's)
  cublas torch.matmul(G_fp16, Z_fp16) -> cuBLAS (nvjet), and the G^T product as torch.matmul(G.T, Z)

    python tools/ab_gradgemm.py [reps] [rounds]  > gpurun_out/ab_gradgemm.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from peppa_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, d = 32768, 512
dev = "cuda"
torch.manual_seed(0)
g8, ld8 = ops.gmat_alloc(n, n, dev, torch.uint8)
g8.copy_(torch.randint(0, 3, (n, ld8), device=dev, dtype=torch.uint8))
g16, ld16 = ops.gmat_alloc(n, n, dev)
g16.copy_(g8[:, :ld16].half())
z = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1).bfloat16()
rinv, _ = ops.row_norms(z)
zh = ops.rows_scale_f16(z, rinv)
zq = ops.rows_quant_i8(z, rinv)
out = torch.zeros(n, d, device=dev)
gt = g16.t()        # a view: cuBLAS reads it transposed


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


variants = [
    ("i8     G   Z", lambda: ops.grad_gemm(g8, n, n, ld8, zq, transpose=False, out=out, accumulate=True)),
    ("f16    G   Z", lambda: ops.grad_gemm(g16, n, n, ld16, zh, transpose=False, out=out, accumulate=True)),
    ("cublas G   Z", lambda: torch.matmul(g16, zh)),
    ("i8     G^T Z", lambda: ops.grad_gemm(g8, n, n, ld8, zq, transpose=True, out=out, accumulate=True)),
    ("f16    G^T Z", lambda: ops.grad_gemm(g16, n, n, ld16, zh, transpose=True, out=out, accumulate=True)),
    ("cublas G^T Z", lambda: torch.matmul(gt, zh)),
]
print(f"gradient product of a {n} x {n} block x [{n}, {d}], {reps} launches back to back per figure, {rounds} interleaved rounds")
print(f"(algorithmic work 2 * {n}^2 * {d} = {2.0 * n * n * d / 1e12:.2f} TFLOP per launch; ms per launch, equivalent TFLOP/s)")
res = {k: [] for k, _ in variants}
for r in range(rounds):
    for k, fn in variants:
        res[k].append(timed(fn))
for k, _ in variants:
    ms = sorted(res[k])[len(res[k]) // 2]
    print(f"{k}:  " + "  ".join(f"{x:.4f}" for x in res[k]) + f"   median {ms:.4f} ms = {2.0 * n * n * d / ms / 1e9:7.1f} TFLOP/s")
# the same numbers say what the i8 path moves: 1 GiB of G per launch instead of 2 GiB
ref = torch.matmul(g16.float()[:2048], zh.float())
got = torch.zeros(n, d, device=dev)
ops.grad_gemm(g8, n, n, ld8, zq, transpose=False, out=got)
print(f"i8 result vs fp32 matmul of the fp16 operands, rows 0..2047: max rel err {((got[:2048] - ref).abs().max() / ref.abs().max()).item():.2e}")
