"""CPU, world_size 2 over gloo: the row-sharded gallery step (peppa_b200/gallery.py) -- all-gather of
the video block, merge of the partial column counts, reduce-scatter of the dV partials -- reproduces
the single-process closed forms of the oracle.  Kernel entry points are emulated (tests/emu_ops.py);
the CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _emb(n, alpha, d=128, seed=666):
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    return V.bfloat16(), A.bfloat16()


def _worker_milnce(rank, world, port, n_local, block, outdir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu_ops
        from peppa_b200.gallery import GalleryStep
        V, A = _emb(n_local * world, 4.0)
        sl = slice(rank * n_local, (rank + 1) * n_local)
        step = GalleryStep(n_local, V.shape[1], rank=rank, world=world, device="cpu", block=block, backend=emu_ops,
                           loss="milnce", temperature=0.5, with_recall=True)
        out = step.run(A[sl].contiguous(), V[sl].contiguous())
        torch.save((rank, out["loss"].item(), out["dA"].clone(), out["dV"].clone(), out["ranks"].clone(), out["recall"].clone()),
                   os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, n_local, block, outdir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu_ops
        from peppa_b200.gallery import GalleryStep
        V, A = _emb(n_local * world, 4.0)
        sl = slice(rank * n_local, (rank + 1) * n_local)
        step = GalleryStep(n_local, V.shape[1], margin=0.2, top_n=10, rank=rank, world=world, device="cpu", block=block,
                           backend=emu_ops)
        out = step.run(A[sl].contiguous(), V[sl].contiguous())
        out2 = step.run(A[sl].contiguous(), V[sl].contiguous())           # buffers are reusable
        assert torch.equal(out["dA"], out2["dA"]) and torch.equal(out["ranks"], out2["ranks"])
        torch.save((rank, out["loss"].item(), out["recall"].clone(), out["dA"].clone(), out["dV"].clone(),
                    out["ranks"].clone()), os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_local,block", [(96, 32768), (80, 48)])
def test_two_rank_gallery_matches_single_process_oracle(n_local, block, tmp_path):
    from oracle import pig_oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, block, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    V, A = _emb(n_local * world, 4.0)
    # rows = audio, cols = video: contrastive(cosine_matrix(A, V)); symmetric, so it is TripletLoss(V, A)
    loss, dA, dV = O.hinge_loss_and_grads(A.float(), V.float(), 0.2)
    ref_loss_va = O.triplet_loss(V.float(), A.float(), 0.2)
    ranks, near = O.ranks_identity(V.float(), A.float())
    gdA = torch.cat([r[3] for r in res])
    gdV = torch.cat([r[4] for r in res])
    granks = torch.cat([r[5] for r in res]).long()
    for r in res:
        assert abs(r[1] - loss.item()) < 1e-5 * abs(loss.item()) and abs(r[1] - ref_loss_va.item()) < 1e-5
    assert (gdA.double() - dA).abs().max() / dA.abs().max() < 1e-3      # fp16 operand rounding of the emulation
    assert (gdV.double() - dV).abs().max() / dV.abs().max() < 1e-3
    assert bool(((granks == ranks) | near).all())
    recall = res[0][2]
    assert recall[0] == 0 and torch.equal(res[0][2], res[1][2])
    for k in (1, 5, 10):
        assert abs(recall[k].item() - (granks < k).float().mean().item()) < 1e-6


@pytest.mark.parametrize("n_local,block", [(72, 32768), (64, 40)])
def test_two_rank_milnce_gallery_matches_closed_form(n_local, block, tmp_path):
    """Cross-rank merge of the column log-sum-exp partials (all-gather + pb2_lse_combine contract)."""
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker_milnce, args=(r, world, port, n_local, block, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    V, A = _emb(n_local * world, 4.0)
    a = A.double().requires_grad_(True)
    v = V.double().requires_grad_(True)
    x = (a @ v.T) / 0.5                                    # rows = audio, columns = video, temperature 0.5
    den = torch.logaddexp(torch.logsumexp(x, 1), torch.logsumexp(x, 0))
    ref = (den - torch.diagonal(x)).mean()
    ref.backward()
    for r in res:
        assert abs(r[1] - ref.item()) < 1e-5 * abs(ref.item())
    gdA = torch.cat([r[2] for r in res]).double()
    gdV = torch.cat([r[3] for r in res]).double()
    assert (gdA - a.grad).abs().max() / a.grad.abs().max() < 2e-3      # fp16 gradient matrix of the emulation
    assert (gdV - v.grad).abs().max() / v.grad.abs().max() < 2e-3
    # the ranking fused into the statistics pass (with_recall=True): cosine ranks of the whole gallery, row-sharded
    cos = torch.nn.functional.normalize(A.float(), dim=1) @ torch.nn.functional.normalize(V.float(), dim=1).T
    d = 1.0 - cos
    closer = d < torch.diagonal(d)[:, None]
    granks = closer.sum(dim=1).to(torch.int32)
    near = ((d - torch.diagonal(d)[:, None]).abs() < 1e-6).sum(dim=1) > 1
    got = torch.cat([r[4] for r in res])
    assert bool(((got == granks) | near).all())
    assert torch.equal(res[0][5], res[1][5]) and abs(res[0][5][10].item() - (got < 10).float().mean().item()) < 1e-6


def test_column_block_walk_and_reduction_order():
    """Host logic of the sharded step without any process group: every rank walks ALL columns exactly once, its own
    blocks first, blocks never straddle two owners; and the NCCL-path reductions are started in ONE canonical order
    on every rank (owner by owner) however the ranks interleave their own blocks."""
    sys.path.insert(0, HERE)
    import emu_ops
    from peppa_b200.gallery import GalleryStep
    world, nl, block = 4, 100, 48
    orders = []
    for rank in range(world):
        step = GalleryStep(nl, 128, rank=rank, world=world, device="cpu", block=block, backend=emu_ops, with_grad=False)
        blocks = step._column_blocks()
        cols = sorted((c0, c1) for c0, c1, _ in blocks)
        assert cols[0][0] == 0 and cols[-1][1] == world * nl and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
        assert all(c0 // nl == (c1 - 1) // nl for c0, c1, _ in blocks)                      # one owner per block
        n_own = len([b for b in blocks if b[2]])
        assert all(b[2] for b in blocks[:n_own]) and not any(b[2] for b in blocks[n_own:])  # own blocks first
        assert all(rank * nl <= c0 < (rank + 1) * nl for c0, _, loc in blocks if loc)
        issued = []
        step._reduce_to_owner = lambda full, c0, c1: issued.append((c0, c1)) or len(issued)
        step.p_v = None
        red = step._reduction_state()
        for c0, c1, _ in blocks:
            step._start_ready_reductions(c1, red)
        assert red["next"] == len(red["order"]) and len(red["pending"]) == len(blocks)
        orders.append(issued)
    assert all(o == orders[0] for o in orders) and orders[0] == sorted(orders[0])          # same sequence on every rank
