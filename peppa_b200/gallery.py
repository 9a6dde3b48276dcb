"""Row-sharded audio<->video gallery: symmetric hinge loss (forward + gradients) and recall@1..N
from ONE pass over the similarity matrix, on 1..P GPUs (BASELINE config 5).

Rank r owns rows [r*N/P, (r+1)*N/P) of both the audio matrix A and the video matrix V.
Rows of S are the local audio clips (queries, as in pig/metrics.py:8 where ``references`` = audio),
columns are ALL video clips:

  1. all-gather the video block (bf16), its row norms and the diagonal scores  (NCCL over NVLink)
  2. fused tcgen05 pass over the [N/P x N] strip: hinge loss partials, indicator counts, rank
     counts of the diagonal, and the fp16 gradient-matrix block           (pb2_sim_hinge + rank)
  3. gradient GEMMs: dA rows are complete locally; dV partials are [N, D] per rank.  The strip is walked
     column block by column block, and as soon as a column block's dV partial is complete it is reduced
     to the rank that owns those video rows (asynchronously, on NCCL's stream, while the next column
     block computes): the 2 GiB reduce-scatter of a 2^20 gallery hides behind the tensor work
  4. all-reduce the column counts (int32) and the scalar loss
  5. normalisation Jacobian (pb2_hinge_finish) on the local rows

The loss is symmetric in (V, A) (pig/loss.py:41-48 adds the row and the column hinge), so
``contrastive(cosine_matrix(A, V))`` equals the reference's ``TripletLoss(V, A)``; the gradients are
returned under their own names.  With world_size == 1 no collective is issued.

``loss="milnce"`` runs pig/loss.py:13-26 (MILNCELoss, K = 1, optional temperature) over the same sharding:
row log-sum-exp of the local strip is complete locally; the column log-sum-exp is a partial per rank
(over its own rows) and is merged across ranks (all-gather of the [N] partials + pb2_lse_combine); the
backward recomputes the strip, writes the fp16 gradient-matrix blocks and reduce-scatters dV like the
hinge path.  (No recall in this mode.)
"""
from __future__ import annotations

import torch

from . import ops

_BLOCK = 32768      # gradient-matrix block edge (rows x cols fp16 kept in HBM at once: 2 GiB)


def _blocks(n, step):
    return [(s, min(n, s + step)) for s in range(0, n, step)]


class GalleryStep:
    """Reusable buffers + one ``run`` per step.  ``group`` is a torch.distributed process group (or
    None for the default group); ``world == 1`` needs no initialised process group at all."""

    def __init__(self, n_local: int, dim: int, margin: float = 0.2, top_n: int = 10, rank: int = 0, world: int = 1,
                 group=None, device=None, block: int = _BLOCK, with_grad: bool = True, backend=None,
                 loss: str = "hinge", temperature: float = 1.0, logit_bound=None, byte_gmat=None):
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
        dev, n, nl = self.device, self.n_total, n_local
        f32, i32 = torch.float32, torch.int32
        self.v_full = torch.empty(n, dim, dtype=torch.bfloat16, device=dev) if world > 1 else None
        self.rv_full = torch.empty(n, dtype=f32, device=dev) if world > 1 else None
        self.diag_full = torch.empty(n, dtype=f32, device=dev) if world > 1 else None
        self.row_cnt = torch.empty(nl, dtype=i32, device=dev)
        self.col_cnt = torch.empty(n, dtype=i32, device=dev)
        self.ranks = torch.empty(nl, dtype=i32, device=dev)
        if with_grad:
            self.p_a = torch.empty(nl, dim, dtype=f32, device=dev)
            self.p_v = torch.empty(n, dim, dtype=f32, device=dev)
            self.p_v_loc = torch.empty(nl, dim, dtype=f32, device=dev) if world > 1 else None
            br, bc = min(block, nl), min(block, n)
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

    # -- collectives (no-ops for world == 1) ------------------------------------------------------
    def _all_gather(self, out, loc):
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, loc.contiguous(), group=self.group)

    def _reduce_scatter(self, out, full):
        import torch.distributed as dist
        if dist.get_backend(self.group) == "gloo":      # gloo has no reduce_scatter: all-reduce + slice
            dist.all_reduce(full, group=self.group)
            out.copy_(full[self.rank * self.n_local:(self.rank + 1) * self.n_local])
        else:
            dist.reduce_scatter_tensor(out, full, group=self.group)

    def _reduce_to_owners(self, full, c0, c1):
        """Start the reduction of rows [c0, c1) of a [N, D] partial to the rank(s) owning them; the owner's
        rows are summed in place.  Returns the pending work handles."""
        import torch.distributed as dist
        works, a = [], c0
        while a < c1:
            owner = a // self.n_local
            b = min(c1, (owner + 1) * self.n_local)
            dst = dist.get_global_rank(self.group, owner) if self.group is not None else owner
            works.append(dist.reduce(full[a:b], dst=dst, group=self.group, async_op=True))
            a = b
        return works

    def run(self, a_loc: torch.Tensor, v_loc: torch.Tensor, rinv_a=None, rinv_v=None):
        """a_loc, v_loc: [n_local, dim] bf16 on this rank's GPU; rinv_a / rinv_v: their fp32 1/||row|| if the caller
        already holds them (the encoder tail emits them; rows tagged by it are recognised too).  Returns a dict with the global loss
        (0-d fp32), the local gradient rows ``dA``/``dV`` (fp32, None without grad), ``recall``
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
        diag, pos_thr = ops.sim_diag(a_loc, v_loc, ra, rv)        # S_ii and its rank threshold
        if self.world > 1:
            self._all_gather(self.v_full, v_loc)
            self._all_gather(self.rv_full, rv)
            self._all_gather(self.diag_full, diag)
            v_full, rv_full, diag_full = self.v_full, self.rv_full, self.diag_full
        else:
            v_full, rv_full, diag_full = v_loc, rv, diag
        self.row_cnt.zero_()
        self.col_cnt.zero_()
        self.ranks.zero_()
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        rblocks, cblocks = _blocks(nl, self.block), _blocks(n, self.block)
        if self.with_grad:
            if self.byte_gmat:      # two 8-bit planes per row: the operand of the kind::i8 gradient GEMMs
                ah = ops.rows_quant_i8(a_loc, ra)
                vh = ops.rows_quant_i8(v_full, rv_full)
            else:
                ah = ops.rows_scale_f16(a_loc, ra)
                vh = ops.rows_scale_f16(v_full, rv_full)
            acc_a, acc_v = len(cblocks) > 1, len(rblocks) > 1
            if acc_a:
                self.p_a.zero_()
            if acc_v:
                self.p_v.zero_()
        pending = []
        for (c0, c1) in cblocks:            # column blocks outermost: a block's dV partial completes early
            for (r0, r1) in rblocks:
                part = ops.sim_hinge(a_loc[r0:r1], v_full[c0:c1], ra[r0:r1], rv_full[c0:c1], diag[r0:r1],
                                     diag_full[c0:c1], self.margin, self.row_cnt[r0:r1], self.col_cnt[c0:c1],
                                     self.gmat if self.with_grad else None, self.ld_g if self.with_grad else 0,
                                     row_offset=r0g + r0, col_offset=c0, pos_thr=pos_thr[r0:r1],
                                     rank=self.ranks[r0:r1])
                ops.hinge_loss_terms(loss, partials=part)
                if self.with_grad:
                    ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, vh[c0:c1], transpose=False,
                                  out=self.p_a[r0:r1], accumulate=acc_a)
                    ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, ah[r0:r1], transpose=True,
                                  out=self.p_v[c0:c1], accumulate=acc_v)
            if self.with_grad and self.world > 1:
                pending += self._reduce_to_owners(self.p_v, c0, c1)
        ops.hinge_loss_terms(loss, diag=diag, cnt=self.row_cnt, margin=self.margin)      # local rows' term
        hits = (self.ranks.unsqueeze(0) < torch.arange(self.top_n + 1, device=dev, dtype=torch.int32).unsqueeze(1))
        hits = hits.sum(dim=1).to(torch.float32)
        if self.world > 1:
            dist.all_reduce(self.col_cnt, group=self.group)
            dist.all_reduce(loss, group=self.group)
            dist.all_reduce(hits, group=self.group)
            for w in pending:
                w.wait()
        # column term from the merged counts (identical on every rank, added once after the all-reduce)
        ops.hinge_loss_terms(loss, diag=diag_full, cnt=self.col_cnt, margin=self.margin)
        inv_n2 = 1.0 / float(n) ** 2
        out = {"loss": loss * inv_n2, "recall": hits / float(n), "ranks": self.ranks, "dA": None, "dV": None}
        if self.with_grad:
            cc = self.col_cnt[r0g:r0g + nl]
            p_v_loc = self.p_v[r0g:r0g + nl]      # this rank's rows: complete locally (world 1) or reduced in place
            out["dA"] = ops.hinge_finish(self.p_a, a_loc, v_loc, ra, rv, self.row_cnt, cc, inv_n2)
            out["dV"] = ops.hinge_finish(p_v_loc, v_loc, a_loc, rv, ra, self.row_cnt, cc, inv_n2)
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
        rblocks, cblocks = _blocks(nl, self.block), _blocks(n, self.block)
        lse_row = lse_col = None
        bound = self.logit_bound if self.logit_bound is not None else ops.logit_bound(a_loc, v_full, inv_tau)
        if bound <= ops.LSE_BOTH_MAX_BOUND:
            # bounded logits: rows (complete locally) and columns (over THIS rank's rows) from one pass per block
            lse_row = torch.full((nl,), float("-inf"), dtype=torch.float32, device=dev)
            lse_col = torch.full((n,), float("-inf"), dtype=torch.float32, device=dev)
            for (r0, r1) in rblocks:
                for (c0, c1) in cblocks:
                    ops.sim_lse_both(a_loc[r0:r1], v_full[c0:c1], bound, scale=inv_tau, lse_row=lse_row[r0:r1],
                                     lse_col=lse_col[c0:c1])
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
        if self.world > 1:
            dist.all_reduce(loss, group=self.group)
        out = {"loss": loss / float(n), "recall": None, "ranks": None, "dA": None, "dV": None}
        if not self.with_grad:
            return out
        if self.world > 1:
            _, den_full = ops.milnce_loss(lse_row_full, lse_col, torch.zeros(n, dtype=torch.float32, device=dev))
        else:
            den_full = den_loc
        ah, vh = ops.rows_scale_f16(a_loc), ops.rows_scale_f16(v_full)
        acc_a, acc_v = len(cblocks) > 1, len(rblocks) > 1
        if acc_a:
            self.p_a.zero_()
        if acc_v:
            self.p_v.zero_()
        for (r0, r1) in rblocks:
            for (c0, c1) in cblocks:
                ops.sim_lse_grad(a_loc[r0:r1], v_full[c0:c1], den_loc[r0:r1], den_full[c0:c1], self.gmat, self.ld_g,
                                 scale=inv_tau)
                ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, vh[c0:c1], transpose=False, out=self.p_a[r0:r1],
                              accumulate=acc_a)
                ops.grad_gemm(self.gmat, r1 - r0, c1 - c0, self.ld_g, ah[r0:r1], transpose=True, out=self.p_v[c0:c1],
                              accumulate=acc_v)
        if self.world > 1:
            self._reduce_scatter(self.p_v_loc, self.p_v)
        p_v_loc = self.p_v_loc if self.world > 1 else self.p_v
        out["dA"] = ops.milnce_finish(self.p_a, v_loc, inv_tau / float(n))
        out["dV"] = ops.milnce_finish(p_v_loc, a_loc, inv_tau / float(n))
        return out
