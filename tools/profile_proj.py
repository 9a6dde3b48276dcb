"""ncu target: the encoder tail (Linear 512 -> 512 + normalise + bf16 + rinv) on 2^20 rows, one profiled launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from peppa_b200 import _cabi, ops  # noqa: E402

variant = int(os.environ.get("PB2_PROJ_VARIANT", "0"))   # measurement build's kernel variants (proj.cu: g_variant)
if variant:
    _cabi.use_measurement_library().pb2_debug_proj_variant(variant)
n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, 512, generator=g, device="cuda").bfloat16()
w = (torch.randn(512, 512, generator=g, device="cuda") / 512 ** 0.5).bfloat16()
b = torch.randn(512, generator=g, device="cuda") * 0.1
ops.project_normalize(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.project_normalize(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
