"""Drop-in for ``pig/loss.py``: same classes / functions / signatures, fused sm_100a kernels inside.

``TripletLoss(margin)(V, A)`` (pig/loss.py:28-39) and ``MILNCELoss()(V, A)`` (pig/loss.py:5-26)
never materialise the N x N similarity matrix: the tcgen05 GEMM's epilogue reduces it to the
loss, the indicator counts / log-sum-exp statistics and an fp16 gradient matrix, and two more
tensor-core GEMMs turn that into dV and dA.  The gradient products run during ``forward`` (the loss is a
scalar); ``backward`` is one launch that applies the normalisation Jacobians and ``grad_output``.
"""
from __future__ import annotations

import torch
import torch.nn

from . import _cabi, ops
from .util import cosine_matrix  # noqa: F401  (pig/loss.py re-exports its own copy, :51-55)

# Largest gradient-matrix block kept in HBM at once (rows x cols fp16).  Bigger problems are
# walked block by block with accumulating gradient GEMMs.
_MAX_BLOCK = 32768
_FAST_DTYPES = (torch.bfloat16, torch.float16, torch.float32)
# MIL-NCE problems from this many logits on take the one-pass row + column log-sum-exp (pb2_sim_lse_both)
_LSE_BOTH_MIN_PAIRS = 1 << 26


def _blocks(n, step):
    return [(s, min(n, s + step)) for s in range(0, n, step)]


class _HingeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V, A, margin):
        if V.dim() != 2 or A.dim() != 2 or V.shape[0] != A.shape[0]:
            raise RuntimeError(f"TripletLoss expects V [N, D] and A [N, D]; got {tuple(V.shape)} and {tuple(A.shape)}")
        # rows in their own dtype (fp32 / fp16 / bf16: nothing is rounded; mixed dtypes meet in fp32)
        vb, ab = ops.as_row_pair(V, A)
        dev = vb.device
        n = vb.shape[0]
        need_grad = any(ctx.needs_input_grad[:2])
        if need_grad and n <= _MAX_BLOCK:       # one gradient-matrix block: the fused step, 3 + 1 launches
            # forward: prep, similarity / hinge pass, both gradient products (+ the loss fold beside them); the Jacobians
            # wait for grad_output (backward), which is applied in fp32 before the rounding: an AMP GradScaler's 65536
            # must reach an fp16 gradient of ~1e-7 before the rounding does
            # (embeddings that come from the encoder tail carry their 1/||row||: nothing re-derives the norms)
            loss, state = ops.hinge_forward(vb, ab, margin, ops.known_rinv(V, vb), ops.known_rinv(A, ab))
            ctx.save_for_backward(state, vb, ab)
            ctx.fused = True
            ctx.meta = (V.dtype, V.device, A.dtype, A.device, V.shape[1], A.shape[1])
            return loss if loss.device == V.device else loss.to(V.device)
        ctx.fused = False
        if 0 < n <= _MAX_BLOCK:                 # the loss alone (validation): prep, the pass without a gradient matrix, the fold
            loss, _ = ops.hinge_forward(vb, ab, margin, ops.known_rinv(V, vb), ops.known_rinv(A, ab), want_state=False)
            return loss if loss.device == V.device else loss.to(V.device)
        rv, ra = ops.rinv_of(V, vb), ops.rinv_of(A, ab)
        diag = ops.pair_dot(vb, ab, rinv_x=rv, rinv_y=ra)       # M_ii, pig/loss.py:43
        vx, ax, fv, fa = ops.mma_pair(vb, ab, rv, ra)           # tensor-core operands + epilogue factors (split-fp16 for fp32 rows)
        row_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        col_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        inv_n2 = 1.0 / float(n) ** 2
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        pv = pa = None
        byte_g = need_grad and ops.byte_gmat_ok(vb.shape[1])       # entries are exactly {0, 1, 2}: one byte each
        if need_grad:       # normalised embeddings as operands of the gradient GEMMs: two 8-bit planes, or fp16 copies
            quant = ops.rows_quant_i8 if byte_g else ops.rows_scale_f16
            vh, ah = quant(vb, rv), quant(ab, ra)
        blocks = _blocks(n, _MAX_BLOCK)
        for (r0, r1) in blocks:
            for (c0, c1) in blocks:
                g = ld = None
                if need_grad:
                    g, ld = ops.gmat_alloc(r1 - r0, c1 - c0, dev, torch.uint8 if byte_g else torch.float16)
                part = ops.sim_hinge(vx[r0:r1], ax[c0:c1], fv[r0:r1], fa[c0:c1], diag[r0:r1], diag[c0:c1], margin,
                                     row_cnt[r0:r1], col_cnt[c0:c1], g, ld or 0, row_offset=r0, col_offset=c0)
                ops.hinge_loss_terms(loss, partials=part, alpha=inv_n2)
                if need_grad:
                    if pv is None:
                        pv = torch.zeros(n, vb.shape[1], dtype=torch.float32, device=dev) if len(blocks) > 1 else \
                            torch.empty(n, vb.shape[1], dtype=torch.float32, device=dev)
                        pa = torch.zeros_like(pv) if len(blocks) > 1 else torch.empty_like(pv)
                    acc = len(blocks) > 1
                    ops.grad_gemm(g, r1 - r0, c1 - c0, ld, ah[c0:c1], transpose=False, out=pv[r0:r1], accumulate=acc)
                    ops.grad_gemm(g, r1 - r0, c1 - c0, ld, vh[r0:r1], transpose=True, out=pa[c0:c1], accumulate=acc)
        ops.hinge_loss_terms(loss, diag=diag, cnt=row_cnt, margin=margin, alpha=inv_n2)
        ops.hinge_loss_terms(loss, diag=diag, cnt=col_cnt, margin=margin, alpha=inv_n2)
        # zero-norm rows: the reference yields NaN (division by a zero norm, pig/util.py:11-12)
        bad = ~(torch.isfinite(rv).all() & torch.isfinite(ra).all())
        loss = torch.where(bad, torch.full_like(loss, float("nan")), loss)
        if need_grad:
            dV = ops.hinge_finish(pv, vb, ab, rv, ra, row_cnt, col_cnt, inv_n2)
            dA = ops.hinge_finish(pa, ab, vb, ra, rv, row_cnt, col_cnt, inv_n2)
            ctx.save_for_backward(dV[:, :V.shape[1]], dA[:, :A.shape[1]])
            ctx.meta = (V.dtype, V.device, A.dtype, A.device)
        return loss.to(V.device)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.fused:       # one launch: both Jacobians times grad_output, already in the inputs' dtype
            state, vb, ab = ctx.saved_tensors
            vd, vdev, ad, adev, dv_, da_ = ctx.meta
            go = grad_out.detach()
            if go.device != state.device or go.dtype != torch.float32:
                go = go.to(device=state.device, dtype=torch.float32)
            # two fresh contiguous tensors (not views of one buffer): AccumulateGrad takes them without a copy;
            # scaled in fp32, rounded to the inputs' dtype last
            same = vd == ad and vd in (torch.float32, torch.bfloat16, torch.float16)
            g0, g1 = ops.hinge_backward(state, vb, ab, go, vd if same else torch.float32)
            # (this is the batch-1k training step, launch bound: no view, cast or copy that is not needed)
            d = vb.shape[1]
            gV = gA = None
            if ctx.needs_input_grad[0]:
                gV = g0 if dv_ == d else g0[:, :dv_]
                gV = gV if (gV.device == vdev and gV.dtype == vd) else gV.to(device=vdev, dtype=vd)
            if ctx.needs_input_grad[1]:
                gA = g1 if da_ == d else g1[:, :da_]
                gA = gA if (gA.device == adev and gA.dtype == ad) else gA.to(device=adev, dtype=ad)
            return gV, gA, None
        dV, dA = ctx.saved_tensors
        vd, vdev, ad, adev = ctx.meta
        go = grad_out.to(device=dV.device, dtype=torch.float32)
        gV = (dV * go).to(device=vdev, dtype=vd) if ctx.needs_input_grad[0] else None
        gA = (dA * go).to(device=adev, dtype=ad) if ctx.needs_input_grad[1] else None
        return gV, gA, None


class _MilNceFn(torch.autograd.Function):
    """pig/loss.py:13-26.  x = V A^T / tau is viewed as [N, N, K] with K = len(A) / len(V) candidates per
    clip (audio rows i*K .. i*K+K-1 belong to video i):
        num_i = LSE_k x[i, i, k]
        den_i = LSE( x[i, :, :]  U  x[:, i, :] ) = logaddexp(row LSE of video i, LSE_k column LSE of audio i*K+k)
        loss  = mean_i (den_i - num_i)
    Neither x nor its [N, 2NK] concatenation is materialised: two fused row-LSE passes give the statistics,
    the backward recomputes x tile by tile into the fp16 gradient matrix
        G[i, c] = exp(x_ic - den_i) + exp(x_ic - den_{c // K})      (the diagonal softmax term is rank-K)."""

    @staticmethod
    def forward(ctx, V, A, inv_tau):
        if V.dim() != 2 or A.dim() != 2:
            raise RuntimeError("MILNCELoss expects 2-D V and A")
        n = V.shape[0]
        if n == 0 or A.shape[0] % n != 0 or A.shape[0] == 0:   # the reference's x.view(N, N, -1) fails the same way
            raise RuntimeError(f"shape '[{n}, {n}, -1]' is invalid for input of size {n * A.shape[0]}")
        k = A.shape[0] // n
        vb, ab = ops.as_row_pair(V, A)
        vx, ax, fv, fa = ops.mma_pair(vb, ab)      # tensor-core operands (+ power-of-two row scales of the split-fp16 pair for fp32 rows)
        sl = (lambda t, a0, a1: None if t is None else t[a0:a1])
        dev = vb.device
        nk = n * k
        need_grad = any(ctx.needs_input_grad)
        vblocks, ablocks = _blocks(n, _MAX_BLOCK), _blocks(nk, _MAX_BLOCK)
        lse_row = lse_col = None
        # large problems: both directions of the log-sum-exp from ONE pass over the logits when they are bounded
        # (one device round trip for the bound; below _LSE_BOTH_MIN_PAIRS the two passes cost less than that)
        bound = ops.logit_bound(vb, ab, inv_tau) if n * nk >= _LSE_BOTH_MIN_PAIRS else None
        if bound is not None and bound <= ops.LSE_BOTH_MAX_BOUND:
            lse_row = torch.full((n,), float("-inf"), dtype=torch.float32, device=dev)
            lse_col = torch.full((nk,), float("-inf"), dtype=torch.float32, device=dev)
            for (r0, r1) in vblocks:
                for (c0, c1) in ablocks:
                    ops.sim_lse_both(vx[r0:r1], ax[c0:c1], bound, rinv_x=sl(fv, r0, r1), rinv_y=sl(fa, c0, c1), scale=inv_tau,
                                     lse_row=lse_row[r0:r1], lse_col=lse_col[c0:c1])
        else:
            for (c0, c1) in ablocks:   # rows = videos, columns = audio candidates
                lse_row = ops.sim_lse_rows(vx, ax[c0:c1], rinv_x=fv, rinv_y=sl(fa, c0, c1), scale=inv_tau, lse=lse_row)
            for (c0, c1) in vblocks:   # LSE over videos for every audio candidate = row LSE of A V^T
                lse_col = ops.sim_lse_rows(ax, vx[c0:c1], rinv_x=fa, rinv_y=sl(fv, c0, c1), scale=inv_tau, lse=lse_col)
        if k == 1:
            diag = ops.pair_dot(vb, ab)
            if inv_tau != 1.0:
                diag = diag * inv_tau
            loss, den = ops.milnce_loss(lse_row, lse_col, diag)
            den_a, w = den, None
        else:
            iv = torch.arange(n, device=dev, dtype=torch.int64).repeat_interleave(k)
            pos = ops.pair_dot(vb, ab, ix=iv) * inv_tau                      # x[i, i, k], [N*K]
            num = ops.lse_combine(pos.view(n, k).t().contiguous())           # LSE over the K candidates
            col_grp = ops.lse_combine(lse_col.view(n, k).t().contiguous())   # x[:, i, :] merged over k
            loss, den = ops.milnce_loss(lse_row, col_grp, num)
            den_a = den.repeat_interleave(k)                                 # den of the clip an audio row belongs to
            w = torch.exp(pos - num.repeat_interleave(k)).contiguous()       # softmax_k of the paired logits
        if need_grad:
            acc_v, acc_a = len(ablocks) > 1, len(vblocks) > 1
            vh, ah = ops.rows_scale_f16(vb), ops.rows_scale_f16(ab)
            d = vb.shape[1]
            pv = (torch.zeros if acc_v else torch.empty)(n, d, dtype=torch.float32, device=dev)
            pa = (torch.zeros if acc_a else torch.empty)(nk, d, dtype=torch.float32, device=dev)
            for (r0, r1) in vblocks:
                for (c0, c1) in ablocks:
                    g, ld = ops.gmat_alloc(r1 - r0, c1 - c0, dev)
                    ops.sim_lse_grad(vx[r0:r1], ax[c0:c1], den[r0:r1], den_a[c0:c1], g, ld, rinv_x=sl(fv, r0, r1),
                                     rinv_y=sl(fa, c0, c1), scale=inv_tau)
                    ops.grad_gemm(g, r1 - r0, c1 - c0, ld, ah[c0:c1], transpose=False, out=pv[r0:r1], accumulate=acc_v)
                    ops.grad_gemm(g, r1 - r0, c1 - c0, ld, vh[r0:r1], transpose=True, out=pa[c0:c1], accumulate=acc_a)
            if k == 1:
                dV = ops.milnce_finish(pv, ab, inv_tau / n)
                dA = ops.milnce_finish(pa, vb, inv_tau / n)
            else:
                dV = ops.milnce_finish_k(pv, ab, w, n, k, 1, inv_tau / n)
                dA = ops.milnce_finish_k(pa, vb, w, nk, 1, k, inv_tau / n)
            ctx.save_for_backward(dV[:, :V.shape[1]], dA[:, :A.shape[1]])
            ctx.meta = (V.dtype, V.device, A.dtype, A.device)
        return loss.to(V.device)

    @staticmethod
    def backward(ctx, grad_out):
        dV, dA = ctx.saved_tensors
        vd, vdev, ad, adev = ctx.meta
        go = grad_out.to(device=dV.device, dtype=torch.float32)
        gV = (dV * go).to(device=vdev, dtype=vd) if ctx.needs_input_grad[0] else None
        gA = (dA * go).to(device=adev, dtype=ad) if ctx.needs_input_grad[1] else None
        return gV, gA, None


class MILNCELoss(torch.nn.Module):
    """The loss implemented is: log(pos/(2 * pos + neg)) = log(pos/(pos + neg/2)) - log(2)
    (pig/loss.py:5-26; MIL-NCE of Miech et al.; K = len(A) / len(V) candidates per clip, K = 1 in the reference repo).

    ``temperature`` is an extension (default 1.0 = the reference, which has none): logits are divided by it."""

    def __init__(self, temperature=1.0):
        super(MILNCELoss, self).__init__()
        self.temperature = float(temperature)

    def forward(self, V, A):
        """Returns MIL-NCE loss.
        Args:
           V: Tensor of embeddings (e.g. video)
           A: Tensor of embeddings (e.g. audio)
        """
        return _MilNceFn.apply(V, A, 1.0 / self.temperature)


class TripletLoss(torch.nn.Module):
    def __init__(self, margin):
        super(TripletLoss, self).__init__()
        self.margin = margin

    def forward(self, V, A):
        """Returns Triplet loss with margin.
        Args:
           V: Tensor of embeddings (e.g. video)
           A: Tensor of embeddings (e.g. audio)
        """
        # the batch-~1k training step is host bound: inputs that need no conversion go through the C++ autograd node
        # (csrc/torch_fast.cpp: the same pb2_hinge_forward / pb2_hinge_backward calls without the Python round trips)
        if (V.is_cuda and A.is_cuda and V.dim() == 2 and V.shape == A.shape and V.dtype == A.dtype and V.device == A.device
                and V.dtype in _FAST_DTYPES and V.shape[1] % 64 == 0 and 0 < V.shape[0] <= _MAX_BLOCK
                and V.stride(1) == 1 and A.stride(1) == 1 and V.stride(0) % 8 == 0 and A.stride(0) % 8 == 0
                and (V.requires_grad or A.requires_grad) and torch.is_grad_enabled() and ops.EVENT_LOG is None
                and V.data_ptr() % 16 == 0 and A.data_ptr() % 16 == 0):
            fast = _cabi.fast()
            if fast is not None:
                return fast.triplet_loss(V, A, float(self.margin), ops.known_rinv(V, V), ops.known_rinv(A, A))
        return _HingeFn.apply(V, A, float(self.margin))


class _ContrastiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M, margin):
        if M.dim() != 2 or M.shape[0] != M.shape[1]:
            raise RuntimeError(f"contrastive expects a square similarity matrix, got {tuple(M.shape)}")
        dev = ops.require_cuda(M.device)
        m = M.detach().to(device=dev, dtype=torch.float32).contiguous()
        loss, grad = ops.contrastive_matrix(m, margin, ctx.needs_input_grad[0])
        if grad is not None:
            ctx.save_for_backward(grad)
            ctx.meta = (M.dtype, M.device)
        return loss.to(M.device)

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        dt, dev = ctx.meta
        return (grad * grad_out.to(device=grad.device, dtype=torch.float32)).to(device=dev, dtype=dt), None


def contrastive(M, margin=0.2):
    "Returns contrastive margin loss over similarity matrix M."
    return _ContrastiveFn.apply(M, float(margin))
