"""Drop-in for the hot-path part of ``pig/util.py``: ``cosine_matrix`` (pig/util.py:9-13) plus the
two helpers the triplet sampler consumes (``shuffled``/``grouped``, pig/util.py:31-35)."""
from __future__ import annotations

import random
from itertools import groupby

import torch

from . import ops


def _cosine_forward(U, V):
    ub, vb = ops.as_row_pair(U, V)          # own dtype (fp32 / fp16 / bf16), sizes checked like the reference's matmul
    ru, nu = ops.row_norms(ub)
    rv, nv = ops.row_norms(vb)
    return ops.sim_matrix(*ops.mma_pair(ub, vb, ru, rv)), (ub, vb, ru, nu, rv, nv)


class _CosineMatrix(torch.autograd.Function):
    """Forward on the tcgen05 kernel.  The backward (only reached by analysis code that
    differentiates through a materialised matrix, e.g. contrastive(cosine_matrix(..))) folds the
    incoming fp32 gradient into the same fp16 gradient-matrix format the fused losses use."""

    @staticmethod
    def forward(ctx, U, V):
        out, saved = _cosine_forward(U, V)
        ctx.save_for_backward(out, *saved)
        ctx.in_dtypes = (U.dtype, V.dtype)
        ctx.in_devices = (U.device, V.device)
        ctx.in_dims = (U.shape[1], V.shape[1])
        return out

    @staticmethod
    def backward(ctx, dS):
        S, ub, vb, ru, nu, rv, nv = ctx.saved_tensors
        dS = dS.to(device=S.device, dtype=torch.float32)
        # d/dÛ = dS V̂, d/dV̂ = dSᵀ Û; dS goes to fp16 with a power-of-two range scale
        amax = dS.abs().amax().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(torch.log2(16384.0 / amax)))
        r, c = S.shape
        g, ld = ops.gmat_alloc(r, c, S.device)
        g[:, :c] = (dS * scale).to(torch.float16)
        inv = (1.0 / scale).reshape(1)
        zeros_r = torch.zeros(r, dtype=torch.int32, device=S.device)
        zeros_c = torch.zeros(c, dtype=torch.int32, device=S.device)
        pu = ops.grad_gemm(g, r, c, ld, ops.rows_scale_f16(vb, rv), transpose=False)
        pv = ops.grad_gemm(g, r, c, ld, ops.rows_scale_f16(ub, ru), transpose=True)
        # hinge_finish with zero counts is exactly the normalisation Jacobian
        dU = ops.hinge_finish(pu, ub, ub, ru, ru, zeros_r, zeros_r, 1.0, inv)
        dV = ops.hinge_finish(pv, vb, vb, rv, rv, zeros_c, zeros_c, 1.0, inv)
        dU = dU[:, :ctx.in_dims[0]].to(device=ctx.in_devices[0], dtype=ctx.in_dtypes[0])
        dV = dV[:, :ctx.in_dims[1]].to(device=ctx.in_devices[1], dtype=ctx.in_dtypes[1])
        return dU, dV


def cosine_matrix(U, V):
    "Returns the matrix of cosine similarity between each row of U and each row of V."
    if torch.is_grad_enabled() and (U.requires_grad or V.requires_grad):
        out = _CosineMatrix.apply(U, V)
    else:
        out, _ = _cosine_forward(U, V)
    if U.dtype in (torch.float16, torch.bfloat16):
        out = out.to(U.dtype)
    return out.to(U.device)


def shuffled(xs):
    return sorted(xs, key=lambda _: random.random())


def grouped(xs, key=lambda x: x):
    return groupby(sorted(xs, key=key), key=key)
