"""ncu target: triplet_accuracy on 2^20 triplets x 512 bf16 (BASELINE config 4), 3 launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from peppa_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(666)
a, p, n = (torch.randn(1 << 20, 512, generator=g, device="cuda").bfloat16() for _ in range(3))
for _ in range(3):
    out = ops.triplet_score(a, p, n)
torch.cuda.synchronize()
print("mean accuracy", out.mean().item())
