"""Row-sharded audio<->video gallery: symmetric hinge loss (forward + gradients) and recall@1..N
from ONE pass over the similarity matrix, on 1..P GPUs (BASELINE config 5).

Rank r owns rows [r*N/P, (r+1)*N/P) of both the audio matrix A and the video matrix V.
Rows of S are the local audio clips (queries, as in pig/metrics.py:8 where ``references`` = audio),
columns are ALL video clips:

  1. all-gather the video block (bf16), its row norms and the diagonal scores (NCCL over NVLink), ASYNCHRONOUSLY:
     the strip is walked column block by column block starting with the blocks this rank owns -- they need no
     remote data -- and the gathered blocks are awaited only when the walk reaches the first remote one
  2. fused tcgen05 pass over each [rows x column block]: hinge loss partials, indicator counts, rank counts of the
     diagonal, and the gradient-matrix block (one byte per entry)            (pb2_sim_hinge + rank)
  3. gradient GEMMs (kind::i8): dA rows are complete locally; dV partials are [N, D] per rank
  4. all-reduce the column counts (int32), the scalar loss and the recall hits -- one small NCCL step that is also
     the barrier after which every rank's dV partials are final
  5. dV reduce-scatter over PEER MEMORY: every rank pulls the partial rows it owns from all ranks with P2P loads over
     NVLink and sums them in a fixed order (pb2_peer_reduce on buffers exchanged as CUDA IPC handles).  No NCCL
     kernel runs beside the persistent tensor-core grids during the step (round 1 reduced per column block with
     ncclReduce, whose CTAs took SMs from the 148-CTA GEMM grids: +4 % on every overlapped launch at 8 GPUs)
  6. normalisation Jacobian (pb2_hinge_finish) on the local rows

The loss is symmetric in (V, A) (pig/loss.py:41-48 adds the row and the column hinge), so
``contrastive(cosine_matrix(A, V))`` equals the reference's ``TripletLoss(V, A)``; the gradients are
returned under their own names.  With world_size == 1 no collective is issued.

``loss="milnce"`` runs pig/loss.py:13-26 (MILNCELoss, K = 1, optional temperature) over the same sharding:
row log-sum-exp of the local strip is complete locally; the column log-sum-exp is a partial per rank
(over its own rows) and is merged across ranks (all-gather of the [N] partials + pb2_lse_combine); the
backward recomputes the strip, writes the fp16 gradient-matrix blocks and reduces dV like the hinge path.
``with_recall=True`` adds recall@1..N of the same gallery from the SAME pass as the loss statistics
(pb2_sim_lse_both_rank; a separate pb2_sim_rank pass only when the logits have no usable bound).
"""
from __future__ import annotations

import torch

from . import ops

_BLOCK = 32768      # gradient-matrix block edge (rows x cols kept in HBM at once: 1 GiB of bytes, 2 GiB of fp16)


def _blocks(n, step, base=0):
    return [(base + s, base + min(n, s + step)) for s in range(0, n, step)]


class _PeerRows:
    """The [N, D] fp32 dV partial of every rank, mapped into this process over CUDA IPC (one process per GPU, one
    node): ``ptrs[q]`` is rank q's buffer as a device pointer usable by this GPU's kernels."""

    def __init__(self, buf, rank, world, group, device):
        import torch.distributed as dist
        self.device, self.bases, self.ptrs = device, [], []
        try:
            mine = ops.ipc_export(buf)
        except Exception:  # noqa: BLE001 -- every rank still takes part in the exchange below
            mine = None
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        if any(h is None for h in handles):
            raise RuntimeError("a rank could not export its buffer as a CUDA IPC handle")
        try:
            for q in range(world):
                if q == rank:
                    self.ptrs.append(buf.data_ptr())
                else:
                    base, ptr = ops.ipc_open(handles[q][0], handles[q][1], device)
                    self.bases.append(base)
                    self.ptrs.append(ptr)
        except Exception:
            self.close()
            raise
        self._keep = buf

    def close(self):
        for b in self.bases:
            try:
                ops.ipc_close(b, self.device)
            except Exception:  # noqa: BLE001 -- interpreter shutdown: the driver unmaps with the context
                pass
        self.bases = []


class GalleryStep:
    """Reusable buffers + one ``run`` per step.  ``group`` is a torch.distributed process group (or
    None for the default group); ``world == 1`` needs no initialised process group at all.

    ``dv_reduce``: how the dV partials reach their owners when world > 1 -- "p2p" (our own kernel over peer memory, see
    the module docstring), "nccl" (per column block ncclReduce overlapped with the tensor work; also what the gloo
    orchestration tests run) or "auto" (default: p2p if the IPC exchange succeeds on every rank, else nccl).

    ``with_recall``: recall@1..top_n (and the int32 ranks) of the same gallery; default on for the hinge loss (the rank
    counts come out of the hinge pass) and off for MIL-NCE, where switching it on takes the counts out of the one-pass
    log-sum-exp statistics (``fuse_rank``, an attribute for A/B tools: False runs a separate rank pass instead)."""

    def __init__(self, n_local: int, dim: int, margin: float = 0.2, top_n: int = 10, rank: int = 0, world: int = 1,
                 group=None, device=None, block: int = _BLOCK, with_grad: bool = True, backend=None,
                 loss: str = "hinge", temperature: float = 1.0, logit_bound=None, byte_gmat=None, dv_reduce="auto",
                 with_recall=None):
        if loss not in ("hinge", "milnce"):
            raise ValueError("loss must be 'hinge' or 'milnce'")
        self.loss, self.inv_tau = loss, 1.0 / float(temperature)
        # MIL-NCE: a known bound on |logit| (e.g. 1 / temperature for unit-norm embeddings) skips the per-step device
        # round trip that otherwise derives it from the row norms; float("inf") forces the two-pass log-sum-exp
        self.logit_bound = None if logit_bound is None else float(logit_bound)
        # ``backend`` exists for the world_size-2 gloo tests of the orchestration on CPU boxes: they
        # inject an emulation of the kernel entry points built from the oracle.  The product never
        # passes it; the default is the CUDA ops module and there is no automatic selection.
        self.ops = backend if backend is not None else ops
        self.n_local, self.dim, self.margin, self.top_n = n_local, dim, float(margin), top_n
        self.rank, self.world, self.group = rank, world, group
        self.n_total = n_local * world
        self.device = ops.require_cuda(device) if backend is None else torch.device(device or "cpu")
        self.block = block
        self.with_grad = with_grad
        self.with_recall = (loss == "hinge") if with_recall is None else bool(with_recall)
        self.fuse_rank = True       # MIL-NCE: rank counts from the statistics pass (False: a separate pb2_sim_rank pass; tools/ A/B)
        dev, n, nl = self.device, self.n_total, n_local
        f32, i32 = torch.float32, torch.int32
        self.v_full = torch.empty(n, dim, dtype=torch.bfloat16, device=dev) if world > 1 else None
        self.rv_full = torch.empty(n, dtype=f32, device=dev) if world > 1 else None
        self.diag_full = torch.empty(n, dtype=f32, device=dev) if world > 1 else None
        self.row_cnt = torch.empty(nl, dtype=i32, device=dev)
        self.col_cnt = torch.empty(n, dtype=i32, device=dev)
        self.ranks = torch.empty(nl, dtype=i32, device=dev)
        # loss partials of every (column block, row block) pass of a step: [passes, CTAs], cleared once per step and
        # folded once at its end (instead of a memset + a fold launch per pass)
        n_pass = len(_blocks(nl, block)) * len(_blocks(nl, block)) * world
        self.loss_parts = torch.empty(n_pass, ops.sim_grid(dev), dtype=f32, device=dev) if (backend is None and loss == "hinge") else None
        self.peers = None
        self.timeline = None        # set to [] to collect (label, CUDA event) marks of the next run (tools/timeline_multi_gpu.py)
        if dv_reduce not in ("auto", "p2p", "nccl"):
            raise ValueError("dv_reduce must be 'auto', 'p2p' or 'nccl'")
        if with_grad:
            self.p_a = torch.empty(nl, dim, dtype=f32, device=dev)
            self.p_v = torch.empty(n, dim, dtype=f32, device=dev)
            self.p_v_loc = torch.empty(nl, dim, dtype=f32, device=dev) if world > 1 else None
            br = bc = min(block, nl)             # column blocks never straddle two owners
            # hinge: the gradient matrix is exactly {0, 1, 2} -> one byte per entry and kind::i8 gradient GEMMs
            # (byte_gmat=False keeps the fp16 matrix; MIL-NCE's entries are softmax weights and stay fp16)
            self.byte_gmat = (loss == "hinge" and dim % 256 == 0 and block <= 32768) if byte_gmat is None else bool(byte_gmat)
            if self.byte_gmat and (loss != "hinge" or dim % 256 != 0 or block > 32768):
                raise ValueError("byte_gmat needs the hinge loss, dim % 256 == 0 and block <= 32768")
            if self.byte_gmat and backend is None:
                self.gmat, self.ld_g = self.ops.gmat_alloc(br, bc, dev, torch.uint8)
            else:
                self.byte_gmat = False
                self.gmat, self.ld_g = self.ops.gmat_alloc(br, bc, dev)
            if world > 1 and backend is None and dv_reduce != "nccl":
                try:
                    self.peers = _PeerRows(self.p_v, rank, world, group, dev)
                except Exception:  # noqa: BLE001 -- e.g. a virtual-memory allocator segment has no IPC handle
                    if dv_reduce == "p2p":
                        raise
                    self.peers = None
                self._agree_on_peers()

    def _agree_on_peers(self):
        """Every rank uses the peer-memory path or none does (a collective decision, made once)."""
        import torch.distributed as dist
        ok = torch.tensor([1 if self.peers is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0 and self.peers is not None:
            self.peers.close()
            self.peers = None

    def __del__(self):
        if getattr(self, "peers", None) is not None:
            self.peers.close()

    # -- collectives (no-ops for world == 1) ------------------------------------------------------
    def _all_gather(self, out, loc, async_op=False):
        import torch.distributed as dist
        return dist.all_gather_into_tensor(out, loc.contiguous(), group=self.group, async_op=async_op)

    def _reduce_to_owner(self, full, c0, c1):
        """NCCL / gloo path: start the reduction of rows [c0, c1) of a [N, D] partial to the rank owning them (summed in
        place there).  Returns the pending work handle."""
        import torch.distributed as dist
        owner = c0 // self.n_local
        dst = dist.get_global_rank(self.group, owner) if self.group is not None else owner
        return dist.reduce(full[c0:c1], dst=dst, group=self.group, async_op=True)

    def _mark(self, label):
        """Timeline mark on the compute stream (only when ``self.timeline`` is a list: measurement runs)."""
        if self.timeline is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))
            self.timeline.append((label, e))

    def _column_blocks(self):
        """[(c0, c1, local)] over all N columns, never straddling two owners: this rank's own blocks first (they need no
        gathered data), then the other owners' in rank order."""
        nl = self.n_local
        owners = [self.rank] + [q for q in range(self.world) if q != self.rank]
        return [(c0, c1, q == self.rank) for q in owners for (c0, c1) in _blocks(nl, self.block, q * nl)]

    def _start_ready_reductions(self, done_c1, state):
        """NCCL / gloo path: collectives must be issued in ONE order on every rank, while each rank walks its own
        blocks first.  Reductions are therefore started owner by owner in rank order, as soon as this rank has
        finished every block up to that point of the canonical order (its own blocks wait for their turn)."""
        state["done"].add(done_c1)
        while state["next"] < len(state["order"]):
            c0, c1 = state["order"][state["next"]]
            if c1 not in state["done"]:
                break
            state["pending"].append(self._reduce_to_owner(self.p_v, c0, c1))
            state["next"] += 1

    def _reduction_state(self):
        order = [(c0, c1) for q in range(self.world) for (c0, c1) in _blocks(self.n_local, self.block, q * self.n_local)]
        return {"order": order, "next": 0, "done": set(), "pending": []}

    def _own_dv_rows(self):
        """This rank's rows of the dV partials summed over the ranks: peer-memory pull (our kernel) or what NCCL left."""
        nl, r0g = self.n_local, self.rank * self.n_local
        if self.world == 1:
            return self.p_v
        if self.peers is not None:
            off = r0g * self.dim * 4
            ops.peer_reduce([p + off for p in self.peers.ptrs], self.p_v_loc)
            return self.p_v_loc
        return self.p_v[r0g:r0g + nl]        # reduced in place by _reduce_to_owner

    def run(self, a_loc: torch.Tensor, v_loc: torch.Tensor, rinv_a=None, rinv_v=None):
        """a_loc, v_loc: [n_local, dim] bf16 on this rank's GPU; rinv_a / rinv_v: their fp32 1/||row|| if the caller
        already holds them (the encoder tail emits them; rows tagged by it are recognised too).  Returns a dict with
        the global loss (0-d fp32), the local gradient rows ``dA``/``dV`` (fp32, None without grad), ``recall``
        ([top_n + 1] fp32: global recall@n, row 0 == 0) and the local int32 ``ranks``."""
        import torch.distributed as dist
        if self.loss == "milnce":
            return self._run_milnce(a_loc, v_loc)
        ops = self.ops
        nl, n, dev = self.n_local, self.n_total, self.device
        r0g = self.rank * nl                                  # global id of the first local row
        known = getattr(ops, "known_rinv", lambda *_: None)
        ra = rinv_a if rinv_a is not None else known(a_loc, a_loc)
        rv = rinv_v if rinv_v is not None else known(v_loc, v_loc)
        ra = ra if ra is not None else ops.row_norms(a_loc)[0]
        rv = rv if rv is not None else ops.row_norms(v_loc)[0]
        self._mark("step start")
        diag, pos_thr = ops.sim_diag(a_loc, v_loc, ra, rv)        # S_ii and its rank threshold
        gathers = []
        if self.world > 1:      # in flight while the local column blocks compute
            gathers = [self._all_gather(self.v_full, v_loc, True), self._all_gather(self.rv_full, rv, True),
                       self._all_gather(self.diag_full, diag, True)]
            self._mark("all-gathers issued (async); local column blocks start")
        self.row_cnt.zero_()
        self.col_cnt.zero_()
        self.ranks.zero_()
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        rblocks, cblocks = _blocks(nl, self.block), self._column_blocks()
        if self.loss_parts is not None:
            self.loss_parts.zero_()
        if self.with_grad:
            quant = ops.rows_quant_i8 if self.byte_gmat else ops.rows_scale_f16
            ah = quant(a_loc, ra)               # embedding operands of the gradient GEMMs (two 8-bit planes / fp16)
            vh_loc, vh_full = quant(v_loc, rv), None
        red = self._reduction_state()
        for ci, (c0, c1, local) in enumerate(cblocks):      # column blocks outermost: a block's dV partial completes early
            if local:       # this rank's own video rows: no gathered data needed
                vc, rvc, dc = v_loc[c0 - r0g:c1 - r0g], rv[c0 - r0g:c1 - r0g], diag[c0 - r0g:c1 - r0g]
                vhc = vh_loc[c0 - r0g:c1 - r0g] if self.with_grad else None
            else:
                if gathers:
                    self._mark("local column blocks done; waiting for the all-gathers")
                    for w in gathers:
                        w.wait()
                    gathers = []
                    self._mark("all-gathers complete on the compute stream; remote column blocks start")
                if self.with_grad and vh_full is None:
                    vh_full = quant(self.v_full, self.rv_full)
                vc, rvc, dc = self.v_full[c0:c1], self.rv_full[c0:c1], self.diag_full[c0:c1]
                vhc = vh_full[c0:c1] if self.with_grad else None
            for ri, (r0, r1) in enumerate(rblocks):
                extra = {} if self.loss_parts is None else {"part": self.loss_parts[ci * len(rblocks) + ri]}
                part = ops.sim_hinge(a_loc[r0:r1], vc, ra[r0:r1], rvc, diag[r0:r1], dc, self.margin, self.row_cnt[r0:r1],
                                     self.col_cnt[c0:c1], self.gmat if self.with_grad else None,
                                     self.ld_g if self.with_grad else 0, row_offset=r0g + r0, col_offset=c0,
                                     pos_thr=pos_thr[r0:r1], rank=self.ranks[r0:r1], **extra)
                if self.loss_parts is None:
                    ops.hinge_loss_terms(loss, partials=part)
                if self.with_grad:      # first contribution overwrites, later ones accumulate: no zero-fill pass, and
                    # nothing touches a remote owner's rows of p_v before the gathers above have completed (peers may
                    # still be pulling last step's partials until they enter this step's all-gather)
                    ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, vhc, transpose=False, out=self.p_a[r0:r1],
                                  accumulate=ci > 0)
                    ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, ah[r0:r1], transpose=True, out=self.p_v[c0:c1],
                                  accumulate=ri > 0)
            if self.with_grad and self.world > 1 and self.peers is None:
                self._start_ready_reductions(c1, red)
        for w in gathers:       # (a rank whose walk never left its own blocks: only possible with world == 1)
            w.wait()
        if self.loss_parts is not None:       # every pass's CTA partials in one fixed-order fold (fp64 inside)
            ops.hinge_loss_terms(loss, partials=self.loss_parts.view(-1))
        ops.hinge_loss_terms(loss, diag=diag, cnt=self.row_cnt, margin=self.margin)      # local rows' term
        hits = (self.ranks.unsqueeze(0) < torch.arange(self.top_n + 1, device=dev, dtype=torch.int32).unsqueeze(1))
        hits = hits.sum(dim=1).to(torch.float32)
        self._mark("last gradient GEMM enqueued; all-reduce of counts / loss / hits")
        if self.world > 1:
            # also the barrier of the peer-memory reduction: a rank enters it after its last gradient GEMM
            dist.all_reduce(self.col_cnt, group=self.group)
            dist.all_reduce(loss, group=self.group)
            dist.all_reduce(hits, group=self.group)
            for w in red["pending"]:
                w.wait()
            diag_full = self.diag_full
        else:
            diag_full = diag
        # column term from the merged counts (identical on every rank, added once after the all-reduce)
        ops.hinge_loss_terms(loss, diag=diag_full, cnt=self.col_cnt, margin=self.margin)
        inv_n2 = 1.0 / float(n) ** 2
        out = {"loss": loss * inv_n2, "recall": hits / float(n), "ranks": self.ranks, "dA": None, "dV": None}
        if self.with_grad:
            cc = self.col_cnt[r0g:r0g + nl]
            self._mark("all-reduces complete; dV pull over peer memory starts")
            p_v_own = self._own_dv_rows()
            self._mark("dV rows summed; Jacobians")
            out["dA"] = ops.hinge_finish(self.p_a, a_loc, v_loc, ra, rv, self.row_cnt, cc, inv_n2)
            out["dV"] = ops.hinge_finish(p_v_own, v_loc, a_loc, rv, ra, self.row_cnt, cc, inv_n2)
        self._mark("step end")
        return out

    def _run_milnce(self, a_loc, v_loc):
        import torch.distributed as dist
        ops = self.ops
        nl, n, dev, inv_tau = self.n_local, self.n_total, self.device, self.inv_tau
        r0g = self.rank * nl
        if self.world > 1:
            self._all_gather(self.v_full, v_loc)
            v_full = self.v_full
        else:
            v_full = v_loc
        rblocks = _blocks(nl, self.block)
        cblocks = [(c0, c1) for (c0, c1, _) in self._column_blocks()]
        lse_row = lse_col = None
        bound = self.logit_bound if self.logit_bound is not None else ops.logit_bound(a_loc, v_full, inv_tau)
        ranked = False
        if self.with_recall:        # recall@1..N of the same gallery (cosine ranking, pig/metrics.py:23-40) on the local strip
            ra, _ = ops.row_norms(a_loc)
            rv_full, _ = ops.row_norms(v_full)
            _, pos_thr = ops.sim_diag(a_loc, v_loc, ra, rv_full[r0g:r0g + nl].contiguous())
            self.ranks.zero_()
        if bound <= ops.LSE_BOTH_MAX_BOUND:
            # bounded logits: rows (complete locally) and columns (over THIS rank's rows) from one pass per block --
            # and the rank counts of the recall from the same pass (pb2_sim_lse_both_rank)
            lse_row = torch.full((nl,), float("-inf"), dtype=torch.float32, device=dev)
            lse_col = torch.full((n,), float("-inf"), dtype=torch.float32, device=dev)
            for (r0, r1) in rblocks:
                for (c0, c1) in cblocks:
                    fused = ((ra[r0:r1], rv_full[c0:c1], pos_thr[r0:r1], r0g + r0, c0, self.ranks[r0:r1])
                             if self.with_recall and self.fuse_rank else None)
                    ops.sim_lse_both(a_loc[r0:r1], v_full[c0:c1], bound, scale=inv_tau, lse_row=lse_row[r0:r1],
                                     lse_col=lse_col[c0:c1], rank=fused)
            ranked = self.with_recall and self.fuse_rank
        else:
            for (c0, c1) in cblocks:      # x = A_loc V^T / tau: rows complete locally
                lse_row = ops.sim_lse_rows(a_loc, v_full[c0:c1], scale=inv_tau, lse=lse_row)
            for (r0, r1) in rblocks:      # columns: log-sum-exp over THIS rank's rows only
                lse_col = ops.sim_lse_rows(v_full, a_loc[r0:r1], scale=inv_tau, lse=lse_col)
        if self.world > 1:
            parts = torch.empty(self.world, n, dtype=torch.float32, device=dev)
            self._all_gather(parts.view(-1), lse_col)
            lse_col = ops.lse_combine(parts)                  # cross-rank merge of the column statistics
            lse_row_full = torch.empty(n, dtype=torch.float32, device=dev)
            self._all_gather(lse_row_full, lse_row)
        else:
            lse_row_full = lse_row
        diag = ops.pair_dot(a_loc, v_loc)
        if inv_tau != 1.0:
            diag = diag * inv_tau
        mean_loc, den_loc = ops.milnce_loss(lse_row, lse_col[r0g:r0g + nl].contiguous(), diag)
        loss = mean_loc * float(nl)
        hits = None
        if self.with_recall:
            if not ranked:          # unbounded logits took the two-pass statistics: the ranking is its own pass
                cols = torch.arange(r0g, r0g + nl, device=dev, dtype=torch.int64)
                ops.sim_rank(a_loc, v_full, ra, rv_full, pos_thr, cols, rank=self.ranks)
            hits = (self.ranks.unsqueeze(0) < torch.arange(self.top_n + 1, device=dev, dtype=torch.int32).unsqueeze(1))
            hits = hits.sum(dim=1).to(torch.float32)
        if self.world > 1:
            dist.all_reduce(loss, group=self.group)
            if hits is not None:
                dist.all_reduce(hits, group=self.group)
        out = {"loss": loss / float(n), "recall": None if hits is None else hits / float(n),
               "ranks": self.ranks if hits is not None else None, "dA": None, "dV": None}
        if not self.with_grad:
            return out
        if self.world > 1:
            _, den_full = ops.milnce_loss(lse_row_full, lse_col, torch.zeros(n, dtype=torch.float32, device=dev))
        else:
            den_full = den_loc
        ah, vh = ops.rows_scale_f16(a_loc), ops.rows_scale_f16(v_full)
        red = self._reduction_state()
        for ci, (c0, c1) in enumerate(cblocks):
            for ri, (r0, r1) in enumerate(rblocks):
                ops.sim_lse_grad(a_loc[r0:r1], v_full[c0:c1], den_loc[r0:r1], den_full[c0:c1], self.gmat, self.ld_g,
                                 scale=inv_tau)
                ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, vh[c0:c1], transpose=False, out=self.p_a[r0:r1],
                              accumulate=ci > 0)
                ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, ah[r0:r1], transpose=True, out=self.p_v[c0:c1],
                              accumulate=ri > 0)
            if self.world > 1 and self.peers is None:
                self._start_ready_reductions(c1, red)
        if self.world > 1:
            if self.peers is not None:      # stream-ordered barrier: every rank's partials are final once it has passed
                dist.all_reduce(torch.zeros(1, dtype=torch.float32, device=dev), group=self.group)
            for w in red["pending"]:
                w.wait()
        p_v_own = self._own_dv_rows()
        out["dA"] = ops.milnce_finish(self.p_a, v_loc, inv_tau / float(n))
        out["dV"] = ops.milnce_finish(p_v_own, a_loc, inv_tau / float(n))
        return out
