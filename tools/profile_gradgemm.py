"""ncu target: the gradient GEMM of one 32768 x 32768 gradient-matrix block (G fp16 x Z fp16 [32768, 512]) as
(1) whole 128 x 256 tiles, (2) CTA pairs + stream-K, (3) the library (torch.matmul -> cuBLAS) on the same
operands.  Profile with ``ncu --profile-from-start off`` (the three launches sit in a profiler range)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from peppa_b200 import _cabi, ops  # noqa: E402
_cabi.use_measurement_library()      # the pb2_debug_* selectors live in the measurement build only

lib = _cabi.lib()
n = 32768
gm, ld = ops.gmat_alloc(n, n, "cuda")
gm.copy_(torch.randint(0, 3, (n, ld), device="cuda").half())
z = (torch.randn(n, 512, device="cuda") * 0.05).half()
out = torch.zeros(n, 512, device="cuda")


def run():
    lib.pb2_debug_gg_pair(0)
    ops.grad_gemm(gm, n, n, ld, z, transpose=False, out=out, accumulate=True, stream_k=False)
    lib.pb2_debug_gg_pair(1)
    ops.grad_gemm(gm, n, n, ld, z, transpose=False, out=out, accumulate=True, stream_k=True)
    return torch.matmul(gm[:, :n], z)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
