"""Multi-GPU check of peppa_b200.gallery.GalleryStep on real GPUs (run under torchrun, one rank per GPU):
every rank runs its shard of a seeded gallery over NCCL; rank 0 also runs the whole gallery on its own GPU
(world = 1) and the two must agree (loss, recall, ranks bit-identical; gradient rows to 1e-5 -- the fp32
reduce-scatter sums partials in a different order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py [n_total]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from peppa_b200.gallery import GalleryStep  # noqa: E402


def emb(n, seed=666, alpha=4.0, d=512):
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    return A.bfloat16(), V.bfloat16()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    A, V = emb(n)
    nl = n // world
    sl = slice(rank * nl, (rank + 1) * nl)
    for block in (32768, 1536):
        out = GalleryStep(nl, 512, rank=rank, world=world, device=dev, block=block).run(A[sl].to(dev), V[sl].to(dev))
        torch.cuda.synchronize()
        if rank == 0:
            ref = GalleryStep(n, 512, device=dev, block=block).run(A.to(dev), V.to(dev))
            rel = lambda x, r: ((x - r).abs().max() / r.abs().max()).item()  # noqa: E731
            e_loss = abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item())
            e_da, e_dv = rel(out["dA"], ref["dA"][sl]), rel(out["dV"], ref["dV"][sl])
            same_ranks = torch.equal(out["ranks"], ref["ranks"][sl])
            same_recall = torch.equal(out["recall"], ref["recall"])
            ok = e_loss < 1e-6 and e_da < 1e-5 and e_dv < 1e-5 and same_ranks and same_recall
            print(f"world={world} n={n} block={block}: loss rel {e_loss:.2e} dA rel {e_da:.2e} dV rel {e_dv:.2e} "
                  f"ranks identical {same_ranks} recall identical {same_recall} -> {'OK' if ok else 'MISMATCH'}", flush=True)
            assert ok
        dist.barrier()
    # MIL-NCE over the same sharding: cross-rank merge of the column log-sum-exp partials
    out = GalleryStep(nl, 512, rank=rank, world=world, device=dev, loss="milnce", temperature=0.5).run(A[sl].to(dev), V[sl].to(dev))
    torch.cuda.synchronize()
    if rank == 0:
        ref = GalleryStep(n, 512, device=dev, loss="milnce", temperature=0.5).run(A.to(dev), V.to(dev))
        rel = lambda x, r: ((x - r).abs().max() / r.abs().max()).item()  # noqa: E731
        e_loss = abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item())
        e_da, e_dv = rel(out["dA"], ref["dA"][sl]), rel(out["dV"], ref["dV"][sl])
        ok = e_loss < 1e-5 and e_da < 1e-3 and e_dv < 1e-3
        print(f"world={world} n={n} milnce: loss rel {e_loss:.2e} dA rel {e_da:.2e} dV rel {e_dv:.2e} -> {'OK' if ok else 'MISMATCH'}",
              flush=True)
        assert ok
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
