"""ncu target: TripletLoss fwd+bwd at batch 1024 x 512 bf16 (BASELINE config 2), 3 eager steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200.loss import TripletLoss  # noqa: E402

dev = torch.device("cuda", 0)
a, v = synth_embeddings(1024, 666, dev)
vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)
mod = TripletLoss(0.2)
for _ in range(3):
    vv.grad = aa.grad = None
    loss = mod(vv, aa)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
