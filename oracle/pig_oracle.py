"""CPU oracle for the peppa contrastive-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (torch-CPU, fp32 unless told otherwise) of the
arithmetic in the reference's ``pig/util.py``, ``pig/loss.py``, ``pig/metrics.py``
and the scoring half of ``pig/triplet.py``.  It is the checker for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``peppa_b200/`` imports it, and the product path has no CPU fallback.

Pinning: the reference ships no golden vectors or automated tests for this path
(SURVEY.md section 4), so parity is pinned against *outputs of the reference itself*:
``oracle/make_golden.py`` imports ``/root/reference/pig`` in the build container
and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function here against those fixtures bit-for-bit (fp32) on CPU.

Every function cites the reference file:line it follows.  The restatement keeps
the reference's *algorithm* (per-row argsort loop, two cosine passes, the
Python sampler consuming ``random`` in the same order) because it also serves as
the "port" CPU baseline that bench.py times next to the GPU path.
"""
from __future__ import annotations

import random as _random
from itertools import groupby as _groupby

import torch
import torch.nn.functional as _F

# --------------------------------------------------------------------------- util


def cosine_matrix(U, V):
    """pig/util.py:9-13 (duplicate at pig/loss.py:51-55).

    Rows are divided by their L2 norm with NO epsilon (a zero row gives NaN),
    then the normalised matrices are multiplied.
    """
    Un = U / torch.linalg.vector_norm(U, ord=2, dim=1, keepdim=True)
    Vn = V / torch.linalg.vector_norm(V, ord=2, dim=1, keepdim=True)
    return Un @ Vn.T


# --------------------------------------------------------------------------- loss


def contrastive(M, margin=0.2):
    """pig/loss.py:41-48.  Symmetric hinge over a square similarity matrix.

    ``cost_col[i,j] = relu(margin + M[i,j] - M[j,j])`` and
    ``cost_row[i,j] = relu(margin + M[i,j] - M[i,i])``; the diagonal (2*margin per
    row) is removed again and the total divided by N**2 (not N*(N-1)).
    """
    neg = -M
    d = torch.diagonal(neg)
    cost_col = torch.clamp(margin - neg + d, min=0)
    cost_row = torch.clamp(margin - neg + d.reshape(-1, 1), min=0)
    cost = cost_col + cost_row
    return (cost.sum() - torch.diagonal(cost).sum()) / cost.shape[0] ** 2


def triplet_loss(V, A, margin):
    """pig/loss.py:28-39 ``TripletLoss(margin).forward(V, A)``."""
    return contrastive(cosine_matrix(V, A), margin=margin)


def milnce_loss(V, A):
    """pig/loss.py:13-26 ``MILNCELoss().forward(V, A)``.

    ``x = V @ A.T`` (no normalisation, no temperature) viewed as [N, N, K] with
    K = len(A) // len(V) candidates per video.  numerator_i = logsumexp_k x[i,i,k];
    denominator_i = logsumexp over row i of x and column i of x together (for
    K == 1 the diagonal entry is therefore counted twice).  Mean of den - num.
    """
    n = V.shape[0]
    x = (V @ A.T).reshape(n, n, -1)
    eye = torch.eye(n, dtype=x.dtype)[:, :, None]
    num = torch.logsumexp((x * eye).sum(dim=1), dim=1)
    both = torch.cat((x, x.permute(1, 0, 2)), dim=1).reshape(n, -1)
    den = torch.logsumexp(both, dim=1)
    return torch.mean(den - num)


# ------------------------------------------------------------------------ metrics


def recall_at_n(candidates, references, correct, n=1):
    """pig/metrics.py:7-21.  rows = references (queries), cols = candidates."""
    dist = 1 - cosine_matrix(references, candidates)
    out = []
    for j in range(dist.shape[0]):
        order = torch.argsort(dist[j])
        top = order[:n]
        target = torch.nonzero(correct[j])[:, 0]
        hits = (top.unsqueeze(0) == target.unsqueeze(1)).sum().item()
        out.append(hits / len(target))          # ZeroDivisionError when no target
    return torch.tensor(out)


def recall_at_1_to_n(candidates, references, correct, N=1):
    """pig/metrics.py:23-40.  Result is [N+1, R]; row 0 is identically zero."""
    dist = 1 - cosine_matrix(references, candidates)
    out = [[] for _ in range(N + 1)]
    out[0] = [0 for _ in range(dist.shape[0])]
    for j in range(dist.shape[0]):
        order = torch.argsort(dist[j])
        target = torch.nonzero(correct[j])[:, 0]
        for n in range(1, N + 1):
            top = order[:n]
            hits = (top.unsqueeze(0) == target.unsqueeze(1)).sum().item()
            out[n].append(hits / len(target))
    return torch.tensor(out)


def triplet_accuracy(anchor, positive, negative, dim=1, discrete=True):
    """pig/metrics.py:45-52.  {0, 0.5, 1} when discrete, raw cosine gap otherwise."""
    gap = _F.cosine_similarity(anchor, positive, dim=dim) - _F.cosine_similarity(anchor, negative, dim=dim)
    return (torch.sign(gap) + 1) / 2 if discrete else gap


def batch_triplet_accuracy(batch):
    """pig/metrics.py:42-43."""
    return triplet_accuracy(batch.anchor, batch.positive, batch.negative)


def sample_indices(x, size):
    """pig/metrics.py:79-81.  Consumes the global torch CPU generator."""
    return torch.randperm(x.size(0))[:size]


def resampled_recall(candidates, references, size=100, n_samples=100, n=1):
    """pig/metrics.py:54-64."""
    assert len(candidates) == len(references)
    assert len(candidates) >= size
    rows = []
    for _ in range(n_samples):
        ix = sample_indices(candidates, size)
        rows.append(recall_at_n(candidates[ix], references[ix], torch.eye(size), n=n))
    return torch.stack(rows)


def resampled_recall_at_1_to_n(candidates, references, size=100, n_samples=100, N=1):
    """pig/metrics.py:67-77."""
    assert len(candidates) == len(references)
    assert len(candidates) >= size
    rows = []
    for _ in range(n_samples):
        ix = sample_indices(candidates, size)
        rows.append(recall_at_1_to_n(candidates[ix], references[ix], torch.eye(size), N=N))
    return torch.stack(rows)


# ------------------------------------------------------------------------ triplets


def shuffled(xs):
    """pig/util.py:31-32.  One ``random.random()`` draw per element, stable sort."""
    return sorted(xs, key=lambda _: _random.random())


def grouped(xs, key=lambda x: x):
    """pig/util.py:34-35."""
    return _groupby(sorted(xs, key=key), key=key)


def pairs(xs):
    """pig/triplet.py:115-121.  Adjacent pairs; a trailing odd element is dropped."""
    return [xs[i:i + 2] for i in range(0, len(xs) - 1, 2)]


def _triplets(clips, criterion):
    """pig/triplet.py:99-104.  Within each equal-``criterion`` group: shuffle,
    pair up, and ``random.sample`` decides which of the pair is the target."""
    for _, items in grouped(clips, key=criterion):
        for pair in pairs(shuffled(items)):
            target, distractor = _random.sample(pair, 2)
            yield target, distractor


def sample_triplet_indices(duration):
    """The index draw shared by pig/triplet.py:66-68 and :86-88."""
    pos, neg = zip(*_triplets(range(len(duration)), lambda idx: duration[idx]))
    return torch.tensor(pos), torch.tensor(neg)


def score_triplets(video, audio, duration, n_samples=100):
    """pig/triplet.py:82-96 with the stray line :93 removed.

    At reference HEAD line 93 (``success.append(success)``) raises NameError on the
    first iteration; the intended behaviour (what the shipped checkpoints'
    ``valnarr_triplet`` scores were computed with) is the function without it.
    """
    accuracy, length = [], []
    for _ in range(n_samples):
        pos, neg = sample_triplet_indices(duration)
        acc = triplet_accuracy(anchor=audio[pos], positive=video[pos], negative=video[neg])
        accuracy.append(acc.mean().item())
        length.append(duration[pos])
    return {"accuracy": torch.tensor(accuracy), "duration": torch.cat(length)}


def comparative_score_triplets(video_set, audio_set, duration, n_samples=100):
    """pig/triplet.py:63-79.  One index draw per sample shared by all models."""
    success = [[] for _ in video_set]
    length = []
    for _ in range(n_samples):
        pos, neg = sample_triplet_indices(duration)
        for m in range(len(video_set)):
            success[m].append(triplet_accuracy(anchor=audio_set[m][pos], positive=video_set[m][pos],
                                               negative=video_set[m][neg], discrete=False))
        length.append(duration[pos])
    return {"success": [torch.cat(s) for s in success], "duration": torch.cat(length)}


# ----------------------------------------------------------- closed forms (SURVEY 8a')
# Used by tests at sizes where the reference's N x N temporaries or per-row Python
# loop are too slow; each is itself checked against the functions above at small N.


def ranks_identity(candidates, references, block=2048):
    """rank_j = #{c : fl32(1 - S[j,c]) < fl32(1 - S[j,j])} with S = cosine_matrix(references,
    candidates) -- the count form of pig/metrics.py:8-20 when ``correct`` is the identity.
    Also returns, per row, whether another candidate lies within 1e-6 of the positive
    (the rows whose reference rank depends on argsort's unspecified tie order)."""
    R = references / torch.linalg.vector_norm(references, dim=1, keepdim=True)
    C = candidates / torch.linalg.vector_norm(candidates, dim=1, keepdim=True)
    n = R.shape[0]
    ranks = torch.empty(n, dtype=torch.int64)
    near = torch.empty(n, dtype=torch.bool)
    for s in range(0, n, block):
        e = min(n, s + block)
        d = 1 - R[s:e] @ C.T
        pos = d[torch.arange(e - s), torch.arange(s, e)].unsqueeze(1)
        ranks[s:e] = (d < pos).sum(dim=1)
        near[s:e] = ((d - pos).abs() <= 1e-6).sum(dim=1) > 1
    return ranks, near


def hinge_loss_and_grads(V, A, margin, dtype=torch.float64):
    """Closed form of TripletLoss fwd+bwd (SURVEY 8a'), evaluated in ``dtype``."""
    V = V.to(dtype)
    A = A.to(dtype)
    n = V.shape[0]
    rv = 1 / torch.linalg.vector_norm(V, dim=1, keepdim=True)
    ra = 1 / torch.linalg.vector_norm(A, dim=1, keepdim=True)
    Vh, Ah = V * rv, A * ra
    S = Vh @ Ah.T
    d = torch.diagonal(S)
    zc = margin + S - d.unsqueeze(0)
    zr = margin + S - d.unsqueeze(1)
    off = ~torch.eye(n, dtype=torch.bool)
    loss = (torch.clamp(zc, min=0)[off].sum() + torch.clamp(zr, min=0)[off].sum()) / n ** 2
    Ic = ((zc >= 0) & off).to(dtype)
    Ir = ((zr >= 0) & off).to(dtype)
    G = (Ic + Ir) / n ** 2
    G = G - torch.diag(Ic.sum(dim=0) + Ir.sum(dim=1)) / n ** 2
    gV, gA = G @ Ah, G.T @ Vh
    dV = (gV - Vh * (gV * Vh).sum(dim=1, keepdim=True)) * rv
    dA = (gA - Ah * (gA * Ah).sum(dim=1, keepdim=True)) * ra
    return loss, dV, dA


def milnce_loss_and_grads(V, A, dtype=torch.float64):
    """Closed form of MILNCELoss fwd+bwd for K == 1 (SURVEY 8a')."""
    V = V.to(dtype)
    A = A.to(dtype)
    n = V.shape[0]
    x = V @ A.T
    den = torch.logaddexp(torch.logsumexp(x, dim=1), torch.logsumexp(x, dim=0))
    loss = (den - torch.diagonal(x)).mean()
    G = (torch.exp(x - den.unsqueeze(1)) + torch.exp(x - den.unsqueeze(0))) / n - torch.eye(n, dtype=dtype) / n
    return loss, G @ A, G.T @ V
