"""Multi-GPU check of the row-sharded gallery on real GPUs (run under torchrun, one rank per GPU).  The record it
prints (and writes to gpurun_out/check_multi_gpu_<world>gpu.json) is committed under profiles/.

  1. GalleryStep over NCCL + peer memory (dv_reduce = "p2p": our pb2_peer_reduce kernel over CUDA-IPC mapped buffers)
     and over NCCL alone (dv_reduce = "nccl") against the SAME gallery run by rank 0 on its own GPU (world = 1):
     ranks / recall bit-identical, rank hash equal, loss to 1e-6, gradient rows to 1e-5 (the cross-rank sum orders its
     fp32 terms differently); hinge (one-byte and fp16 gradient matrix, blocked) and MIL-NCE.
  2. The sharded result against the ORACLE: exact ranks of 64 seeded rows and their dA / dV rows against the blockwise
     restatement of the reference's formulas (oracle/blockwise.py), like bench.py's check.verified.
  3. The C-ABI NCCL entry points (pb2_nccl_gallery_allgather / _colstat_merge / _dv_reduce_scatter) driven with a raw
     ncclComm_t created through ctypes, against torch.distributed's results -- the path of a non-torch host.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py [n_total]
"""
import ctypes
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import rank_hash_terms, synth_embeddings, verify_gallery  # noqa: E402
from peppa_b200 import _cabi, ops  # noqa: E402
from peppa_b200.gallery import GalleryStep  # noqa: E402


def rel(x, r):
    return ((x - r).abs().max() / r.abs().max()).item()


def raw_nccl_comm(rank, world):
    """ncclComm_t built with ctypes on the NCCL torch already loaded (what a C / Go / Rust host would do natively)."""
    path = next((line.split()[-1] for line in open("/proc/self/maps") if "libnccl.so" in line), None)
    lib = ctypes.CDLL(path or "libnccl.so.2")

    class UID(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]

    uid = UID()
    if rank == 0:
        assert lib.ncclGetUniqueId(ctypes.byref(uid)) == 0
    box = [bytes(uid) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctypes.memmove(ctypes.byref(uid), box[0], 128)
    comm = ctypes.c_void_p()
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UID, ctypes.c_int]
    assert lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
    return lib, comm


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    A, V = synth_embeddings(n, 666, dev)            # one global seed: every rank holds the whole gallery, uses its rows
    nl = n // world
    sl = slice(rank * nl, (rank + 1) * nl)
    a_loc, v_loc = A[sl].contiguous(), V[sl].contiguous()
    record = {"world": world, "n": n, "checks": []}
    ok_all = True

    def note(name, ok, **kw):
        nonlocal ok_all
        ok_all &= bool(ok)
        if rank == 0:
            record["checks"].append({"check": name, "ok": bool(ok), **kw})
            print(f"world={world} n={n} {name}: " + " ".join(f"{k}={v}" for k, v in kw.items()) + f" -> {'OK' if ok else 'MISMATCH'}", flush=True)

    refs = {}
    for mode in ("p2p", "nccl"):
        for block, byte_g in ((32768, None), (1536, None), (2048, False)):
            step = GalleryStep(nl, 512, rank=rank, world=world, device=dev, block=block, byte_gmat=byte_g, dv_reduce=mode)
            used = "p2p" if step.peers is not None else "nccl"
            out = step.run(a_loc, v_loc)
            out2 = step.run(a_loc, v_loc)             # second step on the same buffers (the cross-step hazards of the peer pull)
            torch.cuda.synchronize()
            same_again = torch.equal(out["dV"], out2["dV"]) and torch.equal(out["ranks"], out2["ranks"])
            h = rank_hash_terms(out["ranks"], rank * nl).sum().reshape(1)
            dist.all_reduce(h)
            key = (block, byte_g)
            if rank == 0 and key not in refs:
                r1 = GalleryStep(n, 512, device=dev, block=block, byte_gmat=byte_g).run(A, V)
                refs[key] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in r1.items()}
                refs[key]["hash"] = int(rank_hash_terms(r1["ranks"], 0).sum())
            if rank == 0:
                ref = refs[key]
                e_loss = abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item())
                e_da, e_dv = rel(out["dA"], ref["dA"][sl]), rel(out["dV"], ref["dV"][sl])
                same_ranks = torch.equal(out["ranks"], ref["ranks"][sl])
                same_recall = torch.equal(out["recall"], ref["recall"])
                same_hash = int(h.item()) == ref["hash"]
                ok = e_loss < 1e-6 and e_da < 1e-5 and e_dv < 1e-5 and same_ranks and same_recall and same_hash and same_again
                note(f"hinge dv_reduce={mode}(ran {used}) block={block} gmat={'u8' if step.byte_gmat else 'f16'} vs world=1", ok,
                     loss_rel=f"{e_loss:.2e}", dA_rel=f"{e_da:.2e}", dV_rel=f"{e_dv:.2e}", ranks_identical=same_ranks,
                     recall_identical=same_recall, rank_hash_equal=same_hash, second_step_identical=same_again)
            if block == 32768 and byte_g is None:     # against the oracle, like bench.py check.verified
                ver = verify_gallery(out, A, V, rank, world)
                vt = torch.tensor([ver["rows"], ver["rank_mismatch"]], dtype=torch.int64, device=dev)
                ve = torch.tensor([ver["dA_rel_err"], ver["dV_rel_err"]], dtype=torch.float64, device=dev)
                dist.all_reduce(vt)
                dist.all_reduce(ve, op=dist.ReduceOp.MAX)
                note(f"hinge dv_reduce={mode} vs oracle/blockwise.py (64 seeded rows)", int(vt[0]) == 64 and int(vt[1]) == 0 and
                     float(ve.max()) < 1e-3, rows=int(vt[0]), rank_mismatches=int(vt[1]), dA_rel=f"{float(ve[0]):.2e}",
                     dV_rel=f"{float(ve[1]):.2e}")
            del step, out, out2
            gc.collect()
            dist.barrier()
        # MIL-NCE over the same sharding: cross-rank merge of the column log-sum-exp partials, recall from the same gallery
        step = GalleryStep(nl, 512, rank=rank, world=world, device=dev, loss="milnce", temperature=0.5, dv_reduce=mode,
                           with_recall=True)
        out = step.run(a_loc, v_loc)
        torch.cuda.synchronize()
        if rank == 0:
            if "milnce" not in refs:
                refs["milnce"] = GalleryStep(n, 512, device=dev, loss="milnce", temperature=0.5, with_recall=True).run(A, V)
            ref = refs["milnce"]
            e_loss = abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item())
            e_da, e_dv = rel(out["dA"], ref["dA"][sl]), rel(out["dV"], ref["dV"][sl])
            same_ranks = torch.equal(out["ranks"], ref["ranks"][sl])
            if "rank_pass" not in refs:     # the ranks of the fused statistics + rank pass against the separate rank pass
                ra_, _ = ops.row_norms(A)
                rv_, _ = ops.row_norms(V)
                _, thr_ = ops.sim_diag(A, V, ra_, rv_)
                refs["rank_pass"] = ops.sim_rank(A, V, ra_, rv_, thr_, torch.arange(n, device=dev))
            same_pass = torch.equal(ref["ranks"], refs["rank_pass"])
            note(f"milnce dv_reduce={mode} vs world=1 (recall ranks from the statistics pass)",
                 e_loss < 1e-5 and e_da < 1e-3 and e_dv < 1e-3 and same_ranks and same_pass,
                 loss_rel=f"{e_loss:.2e}", dA_rel=f"{e_da:.2e}", dV_rel=f"{e_dv:.2e}", ranks_identical=same_ranks,
                 ranks_equal_separate_rank_pass=same_pass)
        del step, out
        gc.collect()
        dist.barrier()

    # ---- the C ABI with a raw ncclComm_t
    lib = _cabi.lib()
    assert lib.pb2_nccl_available() == 1
    nccl, comm = raw_nccl_comm(rank, world)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    full = torch.empty(n, 512, dtype=torch.bfloat16, device=dev)
    _cabi.check(lib.pb2_nccl_gallery_allgather(comm, ctypes.c_void_p(v_loc.data_ptr()), nl, 512 * 2, ctypes.c_void_p(full.data_ptr()), st))
    torch.cuda.synchronize()
    note("C ABI pb2_nccl_gallery_allgather (raw ncclComm_t)", torch.equal(full, V))
    cnt = torch.full((n,), rank + 1, dtype=torch.int32, device=dev)
    loss = torch.tensor([float(rank + 1)], device=dev)
    hits = torch.arange(11, dtype=torch.float32, device=dev) * (rank + 1)
    _cabi.check(lib.pb2_nccl_colstat_merge(comm, ctypes.c_void_p(cnt.data_ptr()), n, ctypes.c_void_p(loss.data_ptr()),
                                           ctypes.c_void_p(hits.data_ptr()), 11, st))
    torch.cuda.synchronize()
    tot = world * (world + 1) // 2
    note("C ABI pb2_nccl_colstat_merge", bool((cnt == tot).all()) and loss.item() == tot and hits[10].item() == 10 * tot)
    part = torch.randn(n, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    mine = torch.empty(nl, 512, device=dev)
    _cabi.check(lib.pb2_nccl_dv_reduce_scatter(comm, ctypes.c_void_p(part.data_ptr()), nl, 512, ctypes.c_void_p(mine.data_ptr()), st))
    want = torch.empty(nl, 512, device=dev)
    dist.reduce_scatter_tensor(want, part.clone())
    torch.cuda.synchronize()
    note("C ABI pb2_nccl_dv_reduce_scatter vs torch.distributed", rel(mine, want) < 1e-6)
    # peer memory through the C ABI alone: every rank pulls its rows of `part` from all ranks
    peers = __import__("peppa_b200.gallery", fromlist=["_PeerRows"])._PeerRows(part, rank, world, None, dev)
    dist.barrier()
    pulled = torch.empty(nl, 512, device=dev)
    ops.peer_reduce([p + rank * nl * 512 * 4 for p in peers.ptrs], pulled)
    torch.cuda.synchronize()
    note("C ABI pb2_ipc_export / pb2_ipc_open / pb2_peer_reduce vs ncclReduceScatter", rel(pulled, want) < 1e-6)
    dist.barrier()
    peers.close()
    nccl.ncclCommDestroy(comm)
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        record["ok"] = bool(flag.item())
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"check_multi_gpu_{world}gpu.json"), "w") as f:
            json.dump(record, f, indent=1)
        print("ALL OK" if record["ok"] else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
