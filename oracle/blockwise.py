"""Blockwise restatement of the reference's formulas for galleries the full-matrix oracle cannot hold.
TEST INFRASTRUCTURE ONLY (same rule as ``oracle/pig_oracle.py``: imported by ``tests/``, ``tools/`` checks and the
verification / CPU legs of ``bench.py`` -- never by anything under ``peppa_b200/``).

The reference's own code needs the whole N x N matrix (``pig/util.py:9-13`` -> ``pig/metrics.py:8``,
``pig/loss.py:41-48``): 4 TiB of fp32 at N = 2^20.  The functions here evaluate THE SAME expressions for a
sample of query rows against the whole gallery, or block by block with fp64 accumulation, in plain torch on whatever
device the inputs live on (the CPU for small cases; a GPU's fp32 / fp64 library matmul for the 2^20 checks -- the
checker is torch, never this repo's kernels).  ``tests/test_oracle_golden.py::test_blockwise_matches_full_oracle``
pins them to the full-matrix oracle (itself pinned to the reference's outputs) at sizes where both run.
"""
from __future__ import annotations

import torch


def _unit(x, dtype):
    x = x.to(dtype)
    return x / torch.linalg.vector_norm(x, ord=2, dim=1, keepdim=True)      # pig/util.py:11-12, no epsilon


def sampled_ranks(candidates, references, rows, tol=1e-6, block=256):
    """pig/metrics.py:8-12 with ``correct = eye`` for the query rows ``rows`` only:
    ``dist = 1 - cosine_matrix(references[rows], candidates)`` in fp32, ``rank = #{c : dist[c] < dist[pos]}``.
    Returns (ranks int64 [k], near bool [k]): ``near`` marks rows where another candidate lies within ``tol`` of
    the positive -- argsort's order is unspecified there, so those rows are exempt from the identity check."""
    rows = torch.as_tensor(rows, device=references.device, dtype=torch.int64)
    C = _unit(candidates, torch.float32)
    ranks, near = [], []
    for s in range(0, rows.numel(), block):
        r = rows[s:s + block]
        d = 1 - _unit(references[r], torch.float32) @ C.T
        pos = d[torch.arange(r.numel(), device=d.device), r].unsqueeze(1)
        ranks.append((d < pos).sum(1))
        near.append(((d - pos).abs() <= tol).sum(1) > 1)
    return torch.cat(ranks), torch.cat(near)


def sampled_rank_bounds(candidates, references, rows, tol=1e-6, block=256):
    """For galleries of ~10^6 candidates the 1e-6 tie window around the positive is no longer empty for a sizeable
    share of the rows (fp32 scores are ~1e-7 apart there), so beside (ranks, near) this returns the interval every
    admissible ordering of the near-ties must respect: lo = #{c closer and outside the tie window},
    hi = #{c != pos closer or inside the window}.  Rows that are not near-ties have lo == hi == rank."""
    rows = torch.as_tensor(rows, device=references.device, dtype=torch.int64)
    C = _unit(candidates, torch.float32)
    out = [[], [], [], []]
    for s in range(0, rows.numel(), block):
        r = rows[s:s + block]
        d = 1 - _unit(references[r], torch.float32) @ C.T
        pos = d[torch.arange(r.numel(), device=d.device), r].unsqueeze(1)
        closer, tie = d < pos, (d - pos).abs() <= tol               # one tie predicate for near, lo and hi
        out[0].append(closer.sum(1))
        out[1].append(tie.sum(1) > 1)
        out[2].append((closer & ~tie).sum(1))
        out[3].append((closer | tie).sum(1) - 1)
    return tuple(torch.cat(o) for o in out)


def all_ranks(candidates, references, tol=1e-6, block=2048):
    """sampled_ranks for every row (blockwise; the GPU-side twin of pig_oracle.ranks_identity)."""
    n = references.shape[0]
    return sampled_ranks(candidates, references, torch.arange(n, device=references.device), tol, block)


def hinge_loss_blockwise(X, Y, margin, block=2048):
    """pig/loss.py:41-48 on M = cosine_matrix(X, Y) without holding M: scores per block in fp32 (the reference's
    arithmetic), the sum in fp64.  loss = sum_{i != j} [relu(m + M_ij - M_jj) + relu(m + M_ij - M_ii)] / N^2."""
    n = X.shape[0]
    Xn, Yn = _unit(X, torch.float32), _unit(Y, torch.float32)
    diag = (Xn * Yn).sum(1)
    total = torch.zeros((), dtype=torch.float64, device=X.device)
    for s in range(0, n, block):
        M = Xn[s:s + block] @ Yn.T
        b = M.shape[0]
        idx = torch.arange(b, device=X.device)
        cost = torch.clamp(margin + M - diag.unsqueeze(0), min=0) + torch.clamp(margin + M - diag[s:s + b].unsqueeze(1), min=0)
        cost[idx, s + idx] = 0                                                  # pig/loss.py:48 removes the diagonal
        total += cost.double().sum()
    return total / float(n) ** 2


def hinge_grad_rows(X, Y, rows, margin):
    """Rows ``rows`` of d loss / d X for loss = contrastive(cosine_matrix(X, Y), margin) (SURVEY 8a' closed form,
    verified there against the reference's autograd), everything in fp64:
        G_ij = ([m + S_ij - d_j >= 0] + [m + S_ij - d_i >= 0]) (i != j),  G_ii = -(sum_k [..]_ki + sum_k [..]_ik)
        g_i = sum_j G_ij Yhat_j,   dX_i = (g_i - Xhat_i <g_i, Xhat_i>) / ||X_i|| / N^2."""
    rows = torch.as_tensor(rows, device=X.device, dtype=torch.int64)
    n, k = X.shape[0], rows.numel()
    Xh, Yh = _unit(X, torch.float64), _unit(Y, torch.float64)
    diag = (Xh * Yh).sum(1)
    ar = torch.arange(k, device=X.device)
    S = Xh[rows] @ Yh.T                                                    # [k, n] rows of the score matrix
    ir = (margin + S - diag[rows].unsqueeze(1)) >= 0                       # row hinge active
    ic = (margin + S - diag.unsqueeze(0)) >= 0                             # column hinge active
    G = ir.double() + ic.double()
    G[ar, rows] = 0
    # the diagonal term needs the whole COLUMN of each sampled clip: [m + S_ki - d_i >= 0] over all rows k
    col_cnt = ((margin + (Xh @ Yh[rows].T) - diag[rows].unsqueeze(0)) >= 0).sum(0) - 1
    row_cnt = ir.sum(1) - 1
    g = G @ Yh - (row_cnt + col_cnt).double().unsqueeze(1) * Yh[rows]
    xr = Xh[rows]
    norm = torch.linalg.vector_norm(X[rows].double(), ord=2, dim=1, keepdim=True)
    return (g - xr * (g * xr).sum(1, keepdim=True)) / norm / float(n) ** 2


def hinge_kink_counts(X, Y, margin, tol=1e-6, block=2048):
    """The gradient of pig/loss.py:41-48 is DISCONTINUOUS where a hinge argument crosses zero: an entry with
    |m + M_ij - M_jj| or |m + M_ij - M_ii| below the arithmetic's resolution may legitimately fall on either side
    (fp32 against fp64, one summation order against another -- the reference against itself on another BLAS), and
    each such entry moves row i of dX and row j of dY by one unit vector (+ one unit of the diagonal count) / N^2.
    Returns k [N]: per clip, the number of entries of ITS row and ITS column within ``tol`` of a kink -- the
    gradient-side twin of the 1e-6 near-tie exemption of the ranks."""
    n = X.shape[0]
    Xn, Yn = _unit(X, torch.float64), _unit(Y, torch.float64)
    diag = (Xn * Yn).sum(1)
    k = torch.zeros(n, dtype=torch.int64, device=X.device)
    for s in range(0, n, block):
        M = Xn[s:s + block] @ Yn.T
        b = M.shape[0]
        idx = torch.arange(b, device=X.device)
        near = ((margin + M - diag.unsqueeze(0)).abs() <= tol).to(torch.int64) + \
               ((margin + M - diag[s:s + b].unsqueeze(1)).abs() <= tol).to(torch.int64)
        near[idx, s + idx] = 0
        k[s:s + b] += near.sum(1)
        k += near.sum(0)
    return k


def hinge_rows_within(got, ref, x, k, tol_rel=1e-3):
    """Row-wise gradient check with the kink allowance: ||got_i - ref_i|| <= tol_rel * max(||ref_i||, 1% of the median
    row norm) + 2 k_i / (||x_i|| N^2).  Returns (ok, worst ratio of error to allowance)."""
    got, ref = got.double(), ref.double().to(got.device)
    n = ref.shape[0]
    rn = ref.norm(dim=1)
    floor = (0.01 * rn.median()).clamp_min(1e-300)
    allow = tol_rel * torch.maximum(rn, floor) + 2.0 * k.to(got.device).double() / (x.double().to(got.device).norm(dim=1) * float(n) ** 2)
    ratio = ((got - ref).norm(dim=1) / allow).max().item()
    return ratio <= 1.0, ratio
