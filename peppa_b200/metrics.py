"""Drop-in for ``pig/metrics.py``: recall@n ranking and triplet accuracy on the B200 kernels.

Ranking never sorts and never writes the [R, C] distance matrix: for every (query, target) pair
the fused similarity-and-rank kernel counts the candidates strictly closer than the target
(``fl32(1 - s) < fl32(1 - s_target)``, the comparison ``argsort`` resolves in
pig/metrics.py:8-12), and ``recall`` is ``count < n``.  Results equal the reference's except on
rows where another candidate lies within 1e-6 of the target (argsort's tie order there is
unspecified in the reference itself).

Argument order is the reference's: ``recall_at_n(candidates, references, correct)`` ranks the
*candidates* for every *reference* row (pig/metrics.py:8).
"""
from __future__ import annotations

import torch

from . import ops


def _targets(correct, n_rows, n_cols, device):
    """(row_idx, col_idx, per-row target count) from the reference's ``correct`` argument.

    Extensions (default behaviour unchanged): ``None`` = identity, a 1-D integer tensor = the
    target column of every row.
    """
    if correct is None:
        idx = torch.arange(n_rows, device=device)
        return idx, idx, None
    correct = torch.as_tensor(correct)
    if correct.dim() == 1 and not correct.dtype.is_floating_point and correct.dtype != torch.bool:
        return torch.arange(n_rows, device=device), correct.to(device=device, dtype=torch.int64), None
    if correct.dim() != 2 or correct.shape[0] != n_rows:
        raise IndexError(f"correct must be [{n_rows}, {n_cols}], got {tuple(correct.shape)}")
    nz = torch.nonzero(correct.to(device))
    rows, cols = nz[:, 0].contiguous(), nz[:, 1].contiguous()
    counts = torch.bincount(rows, minlength=n_rows)
    if bool((counts == 0).any()):
        # pig/metrics.py:20 / :39 divide by len(target)
        raise ZeroDivisionError("division by zero")
    if rows.numel() == n_rows:      # exactly one target per row: rows == arange
        return rows, cols, None
    return rows, cols, counts


def _pair_ranks(candidates, references, correct):
    """int32 rank of every (row, target) pair plus the bookkeeping to fold pairs back into rows."""
    qb, gb = ops.as_row_pair(references, candidates)      # own dtype, sizes checked like the reference's matmul
    dev = qb.device
    rq, rg = ops.rinv_of(references, qb), ops.rinv_of(candidates, gb)      # the encoder tail's 1/||row|| when tagged
    qb, gb, rq, rg = ops.mma_pair(qb, gb, rq, rg)         # tensor-core operands + epilogue factors (split-fp16 for fp32 rows)
    rows, cols, counts = _targets(correct, qb.shape[0], gb.shape[0], dev)
    if counts is None:
        q, rq_p = qb, rq
    else:                           # several targets per row: one query copy per (row, target) pair
        q, rq_p = qb[rows].contiguous(), rq[rows].contiguous()
    # the target's own score comes from the same tensor-core arithmetic as every candidate's
    # (``correct=None`` hands back one index tensor for rows and columns: no device round trip to find that out)
    identity = counts is None and cols.numel() == gb.shape[0] and (cols is rows or bool((cols == rows).all()))
    tgt, rt = (gb, rg) if identity else (gb[cols].contiguous(), rg[cols].contiguous())
    _, pos_thr = ops.sim_diag(q, tgt, rq_p, rt)
    rank = ops.sim_rank(q, gb, rq_p, rg, pos_thr, cols)
    return rank, rows, counts, qb.shape[0]


def _fold(hits, rows, counts, n_rows):
    """mean over the targets of each row of a [pairs] (or [k, pairs]) 0/1 tensor."""
    if counts is None:
        return hits
    out = torch.zeros(*hits.shape[:-1], n_rows, dtype=torch.float32, device=hits.device)
    out.index_add_(-1, rows, hits)
    return out / counts.to(torch.float32)


def recall_at_n(candidates, references, correct, n=1):
    rank, rows, counts, n_rows = _pair_ranks(candidates, references, correct)
    hits = (rank < n).to(torch.float32)
    return _fold(hits, rows, counts, n_rows).cpu()     # the reference returns torch.tensor(list): CPU float32


def recall_at_1_to_n(candidates, references, correct, N=1):
    rank, rows, counts, n_rows = _pair_ranks(candidates, references, correct)
    ns = torch.arange(0, N + 1, device=rank.device, dtype=torch.int32).unsqueeze(1)
    hits = (rank.unsqueeze(0) < ns).to(torch.float32)   # row 0 (n = 0) is identically zero
    return _fold(hits, rows, counts, n_rows).cpu()


def batch_triplet_accuracy(batch):
    return triplet_accuracy(batch.anchor, batch.positive, batch.negative)


def _rows_last(x, dim):
    x = x.movedim(dim, -1)
    return x.reshape(-1, x.shape[-1]), x.shape[:-1]


def triplet_accuracy(anchor, positive, negative, dim=1, discrete=True):
    out_device = anchor.device
    out_dtype = torch.promote_types(torch.promote_types(anchor.dtype, positive.dtype), negative.dtype)
    dev = ops.require_cuda(anchor.device)
    shape = torch.broadcast_shapes(anchor.shape, positive.shape, negative.shape)
    work_dtype = out_dtype if out_dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.float32
    mats = []
    for t in (anchor, positive, negative):
        t = t.detach().to(device=dev, dtype=work_dtype).expand(shape)
        m, lead = _rows_last(t, dim)
        mats.append(m.contiguous())
    out = ops.triplet_score(mats[0], mats[1], mats[2], discrete=discrete)
    out = out.reshape(lead)
    if out_dtype.is_floating_point:
        out = out.to(out_dtype)
    return out.to(out_device)


# One G x G score matrix serves every subset while it fits comfortably in HBM (fp32, 2^28 entries = 1 GiB).
_RESAMPLE_MATRIX_LIMIT = 1 << 28


def _resampled_ranks(candidates, references, size, n_samples):
    """int32 [n_samples, size] ranks of each subset's positives; the subset draws consume the CPU
    global generator exactly like the reference's loop (one randperm per sample, nothing else)."""
    ix = torch.stack([sample_indices(candidates, size) for _ in range(n_samples)]) if n_samples > 0 else \
        torch.empty(0, size, dtype=torch.int64)
    g = len(candidates)
    if g * g <= _RESAMPLE_MATRIX_LIMIT and n_samples > 0:
        qb, gb = ops.as_row_pair(references, candidates)
        rq, _ = ops.row_norms(qb)
        rg, _ = ops.row_norms(gb)
        scores = ops.sim_matrix(*ops.mma_pair(qb, gb, rq, rg))          # rows = references, as in pig/metrics.py:8
        return ops.subset_rank(scores, ix.to(qb.device))
    ranks = [_pair_ranks(candidates[i], references[i], None)[0] for i in ix]      # huge galleries: per subset
    return torch.stack(ranks) if ranks else torch.empty(0, size, dtype=torch.int32)


def resampled_recall(candidates, references, size=100, n_samples=100, n=1):
    assert len(candidates) == len(references)
    assert len(candidates) >= size
    rank = _resampled_ranks(candidates, references, size, n_samples)
    return (rank < n).to(torch.float32).cpu()


def resampled_recall_at_1_to_n(candidates, references, size=100, n_samples=100, N=1):
    assert len(candidates) == len(references)
    assert len(candidates) >= size
    rank = _resampled_ranks(candidates, references, size, n_samples)
    ns = torch.arange(0, N + 1, device=rank.device, dtype=torch.int32).view(1, N + 1, 1)
    return (rank.unsqueeze(1) < ns).to(torch.float32).cpu()      # [n_samples, N+1, size], row 0 == 0


def sample_indices(x, size):
    # stays on the CPU global generator so seeded runs draw the same subsets as pig/metrics.py:79-81
    ix = torch.randperm(x.size(0))[:size]
    return ix
