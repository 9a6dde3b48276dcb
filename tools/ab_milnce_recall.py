"""What recall@k costs on top of the MIL-NCE step (65536 clips, temperature 0.07, fwd + bwd through GalleryStep): no
recall, recall from the SAME pass as the loss statistics (pb2_sim_lse_both_rank, the product), recall as a separate
pb2_sim_rank pass (round 2's first form).  Modes alternate in one process on one board; ranks are compared."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import DIM, synth_embeddings  # noqa: E402
from gpu_probe import _t  # noqa: E402
from peppa_b200.gallery import GalleryStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
a, v = synth_embeddings(n, 666, dev)
plain = GalleryStep(n, DIM, device=dev, loss="milnce", temperature=0.07)
fused = GalleryStep(n, DIM, device=dev, loss="milnce", temperature=0.07, with_recall=True)
split = GalleryStep(n, DIM, device=dev, loss="milnce", temperature=0.07, with_recall=True)
split.fuse_rank = False
of, os_ = fused.run(a, v), split.run(a, v)
print("ranks identical:", bool(torch.equal(of["ranks"], os_["ranks"])), " loss identical:", bool(torch.equal(of["loss"], os_["loss"])),
      " recall@1/5/10:", [round(of["recall"][k].item(), 4) for k in (1, 5, 10)], flush=True)
for rep in range(3):
    t0 = _t(lambda: plain.run(a, v), iters=5, warm=2)
    t1 = _t(lambda: fused.run(a, v), iters=5, warm=2)
    t2 = _t(lambda: split.run(a, v), iters=5, warm=2)
    print(f"MIL-NCE {n} step: no recall {t0:.3f} ms | recall from the statistics pass {t1:.3f} ms (+{t1 - t0:.3f}) | "
          f"separate rank pass {t2:.3f} ms (+{t2 - t0:.3f})", flush=True)
