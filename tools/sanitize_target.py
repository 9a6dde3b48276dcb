"""Small shapes through every kernel family for compute-sanitizer --tool memcheck (one tool per gpurun call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from peppa_b200 import encoder, loss, metrics, ops  # noqa: E402
from peppa_b200.gallery import GalleryStep  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
V = torch.nn.functional.normalize(torch.randn(1100, 512), dim=1).to(dev)
A = torch.nn.functional.normalize(2 * V.cpu() + torch.randn(1100, 512), dim=1).to(dev)
v, a = V.clone().requires_grad_(True), A.clone().requires_grad_(True)
loss.TripletLoss(0.2)(v, a).backward()
v.grad = a.grad = None
loss.MILNCELoss(0.5)(v, a).backward()
loss.MILNCELoss()(v[:300], a[:900]).sum().backward()          # K = 3 candidates per clip
metrics.recall_at_1_to_n(V, A, None, N=10)
metrics.triplet_accuracy(A, V, V.roll(1, 0))
GalleryStep(1100, 512, block=384).run(A.bfloat16(), V.bfloat16())
GalleryStep(1100, 512, block=384, loss="milnce").run(A.bfloat16(), V.bfloat16())
# stream-K / CTA-pair gradient GEMM (workspace path) on a ragged shape, both orientations
gm, ld = ops.gmat_alloc(19000, 1000, dev)
gm.zero_()
gm[:, :1000] = torch.randint(0, 3, (19000, 1000), device=dev).half()
ops.grad_gemm(gm, 19000, 1000, ld, torch.randn(1000, 512, device=dev).half(), transpose=False)
ops.grad_gemm(gm, 19000, 1000, ld, torch.randn(19000, 512, device=dev).half(), transpose=True)
gm2, ld2 = ops.gmat_alloc(1000, 19000, dev)
gm2.zero_()
ops.grad_gemm(gm2, 1000, 19000, ld2, torch.randn(1000, 512, device=dev).half(), transpose=True)
# encoder tail, ragged rows / narrow outputs
for rows, n_in, n_out in ((300, 28, 512), (129, 512, 64), (1000, 768, 384)):
    encoder.ProjectNormalize(n_in, n_out).to(dev)(torch.randn(rows, n_in, device=dev))
# 256-wide log-sum-exp tiles with a ragged last 128-column unit
ops.sim_lse_rows(torch.randn(20000, 512, device=dev).bfloat16(), torch.randn(1300, 512, device=dev).bfloat16())
ops.sim_lse_rows(torch.randn(40000, 512, device=dev).bfloat16(), torch.randn(300, 512, device=dev).bfloat16())
torch.cuda.synchronize()
print("sanitize target ok")
