"""Timeline of ONE 2^20-gallery step per rank (run under torchrun): where the collectives sit against the tensor work.
CUDA events on the compute stream at the phase boundaries of GalleryStep.run (GalleryStep.timeline), after warm-up;
every rank's marks are gathered and printed by rank 0, plus the per-kernel event totals of that step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/timeline_multi_gpu.py [n_total] > gpurun_out/timeline_8gpu.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import synth_embeddings  # noqa: E402
from peppa_b200 import ops  # noqa: E402
from peppa_b200.gallery import GalleryStep  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nl = n // world
    A, V = synth_embeddings(n, 666, dev)
    a_loc, v_loc = A[rank * nl:(rank + 1) * nl].clone(), V[rank * nl:(rank + 1) * nl].clone()
    del A, V
    for mode in ("p2p", "nccl"):
        step = GalleryStep(nl, 512, rank=rank, world=world, device=dev, dv_reduce=mode)
        for _ in range(3):
            step.run(a_loc, v_loc)
        dist.barrier()
        torch.cuda.synchronize()
        step.timeline = []
        ops.EVENT_LOG = []
        step.run(a_loc, v_loc)
        torch.cuda.synchronize()
        marks, step.timeline = step.timeline, None
        log, ops.EVENT_LOG = ops.EVENT_LOG, None
        t0 = marks[0][1]
        mine = [(lab, t0.elapsed_time(e)) for lab, e in marks]
        kern = {}
        for name, _, s, e in log:
            k = kern.setdefault(name, [0, 0.0])
            k[0] += 1
            k[1] += s.elapsed_time(e)
        allm = [None] * world
        dist.all_gather_object(allm, (mine, kern))
        if rank == 0:
            used = "peer memory (pb2_peer_reduce)" if step.peers is not None else "ncclReduce per column block"
            print(f"=== {world} GPUs, gallery {n}, dV reduction: {used}; one step after 3 warm-up steps; ms since the step's start")
            labels = [lab for lab, _ in mine]
            for i, lab in enumerate(labels):
                ts = [m[0][i][1] for m in allm]
                print(f"{min(ts):9.2f} .. {max(ts):9.2f}  {lab}")
            print("per-kernel device time in that step (rank 0; launches, total ms):",
                  {k: (v[0], round(v[1], 2)) for k, v in allm[0][1].items()})
            ends = [m[0][-1][1] for m in allm]
            print(f"step: {max(ends):.2f} ms (max over ranks; fastest rank {min(ends):.2f} ms)\n", flush=True)
        del step
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
