"""GPU parity: the CUDA path (through the C ABI) against the reference-generated golden fixtures and
the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): loss and gradients within 1e-3 relative error (norm-wise:
max|x - ref| / max|ref|; bf16 inputs, fp32 accumulate); ranks / recall identical except rows where
another candidate lies within 1e-6 of the positive; triplet outputs identical except |gap| < 1e-6.
"""
import random

import pytest
import torch

from conftest import golden_files, load_golden, rel_err, row_rel_err
from oracle import pig_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def emb(n, alpha, d=512, seed=666):
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    return V.bfloat16().float(), A.bfloat16().float()


@pytest.fixture(scope="module")
def pb():
    import peppa_b200.loss as loss
    import peppa_b200.metrics as metrics
    import peppa_b200.triplet as triplet
    import peppa_b200.util as util
    return type("PB", (), dict(loss=loss, metrics=metrics, triplet=triplet, util=util))


def near_rows_multi(candidates, references, correct, tol=1e-6):
    """Rows of a multi-target ``correct`` where some target has another candidate within ``tol`` of its distance
    (argsort's order -- and with it the reference's top-n overlap -- is unspecified there)."""
    d = 1 - O.cosine_matrix(references, candidates)
    near = torch.zeros(d.shape[0], dtype=torch.bool)
    for j, t in torch.nonzero(correct).tolist():
        near[j] |= int(((d[j] - d[j, t]).abs() <= tol).sum()) > 1
    return near


def near_rows_gpu(candidates, references, tol=1e-6, block=4096):
    """The oracle's near-tie mask (identity targets) for galleries too large for the CPU oracle: fp32 torch on the GPU."""
    from oracle import blockwise as B
    return B.all_ranks(candidates.cuda(), references.cuda(), tol=tol, block=block)[1].cpu()


def hinge_rows_ok(got, ref, x, k):
    """Row-wise 1e-3 bar for hinge gradients with the kink allowance of oracle/blockwise.py: an entry within 1e-6 of
    a hinge kink may fall on either side (the gradient is discontinuous there) and moves its row by ~2 / N^2."""
    from oracle import blockwise as B
    ok, ratio = B.hinge_rows_within(got.cpu(), ref, x, k, TOL)
    assert ok, ratio


def _grads(mod, V, A):
    v = V.cuda().requires_grad_(True)
    a = A.cuda().requires_grad_(True)
    out = mod(v, a)
    out.backward()
    return out.detach().cpu(), v.grad.cpu(), a.grad.cpu()


# ------------------------------------------------------------------------------------ golden
@pytest.mark.parametrize("name", golden_files("sim_n"))
def test_losses_against_reference_golden(pb, name):
    g = load_golden(name)
    for kind, mod in (("hinge", pb.loss.TripletLoss(0.2)), ("milnce", pb.loss.MILNCELoss())):
        loss, dV, dA = _grads(mod, g["V"], g["A"])
        assert rel_err(loss, g[f"{kind}_loss"]) < TOL, kind
        assert rel_err(dV, g[f"{kind}_dV"]) < TOL and rel_err(dA, g[f"{kind}_dA"]) < TOL, kind
        assert loss.dtype == torch.float32 and loss.dim() == 0


@pytest.mark.parametrize("name", golden_files("sim_n"))
def test_recall_against_reference_golden(pb, name):
    g = load_golden(name)
    V, A = g["V"], g["A"]
    n = V.shape[0]
    _, near = O.ranks_identity(V, A)
    eye = torch.eye(n, device="cuda")
    for k in (1, 5, 10):
        got = pb.metrics.recall_at_n(V.cuda(), A.cuda(), eye, n=k)
        assert got.dtype == torch.float32 and got.device.type == "cpu" and got.shape == (n,)
        assert bool(((got == g[f"recall_at_{k}"]) | near).all())
    got = pb.metrics.recall_at_1_to_n(V.cuda(), A.cuda(), eye, N=10)
    assert got.shape == (11, n) and bool((got[0] == 0).all())
    assert bool(((got == g["recall_at_1_to_10"]) | near.unsqueeze(0)).all())
    # general multi-target `correct`
    multi = g["correct_multi"]
    near_m = near_rows_multi(V, A, multi)               # rows where a TARGET of the row sits in a 1e-6 tie
    got = pb.metrics.recall_at_n(V.cuda(), A.cuda(), multi.cuda(), n=5)
    assert bool((((got - g["recall_multi_at_5"]).abs() < 1e-6) | near_m).all())
    got = pb.metrics.recall_at_1_to_n(V.cuda(), A.cuda(), multi.cuda(), N=10)
    assert bool((((got - g["recall_multi_1_to_10"]).abs() < 1e-6) | near_m.unsqueeze(0)).all())
    # cosine_matrix and contrastive(M)
    M = pb.util.cosine_matrix(V.cuda(), A.cuda())
    assert (M.cpu() - g["cosine_VA"]).abs().max() < 2e-6
    c = pb.loss.contrastive(g["cosine_VA"].cuda(), margin=0.2)
    assert rel_err(c.cpu(), g["contrastive_M"]) < 1e-5


def test_rectangular_retrieval_golden(pb):
    g = load_golden("sim_rect_40x96.npz")
    got = pb.metrics.recall_at_n(g["V"].cuda(), g["A"].cuda(), g["correct"].cuda(), n=3)
    assert torch.equal(got, g["recall_at_3"])
    got = pb.metrics.recall_at_1_to_n(g["V"].cuda(), g["A"].cuda(), g["correct"].cuda(), N=10)
    assert torch.equal(got, g["recall_at_1_to_10"])
    assert (pb.util.cosine_matrix(g["A"].cuda(), g["V"].cuda()).cpu() - g["cosine"]).abs().max() < 2e-6


@pytest.mark.parametrize("name", golden_files("triplet_t"))
def test_triplet_accuracy_golden(pb, name):
    g = load_golden(name)
    a, p, n = g["anchor"].cuda(), g["positive"].cuda(), g["negative"].cuda()
    for dt in (torch.float32, torch.bfloat16):      # fixtures are bf16-representable: both paths see the same values
        gap = pb.metrics.triplet_accuracy(a.to(dt), p.to(dt), n.to(dt), discrete=False).float().cpu()
        assert (gap - g["gap"]).abs().max() < (2e-6 if dt == torch.float32 else 1e-2)
        disc = pb.metrics.triplet_accuracy(a.to(dt), p.to(dt), n.to(dt)).float().cpu()
        assert bool(((disc == g["discrete"]) | (g["gap"].abs() < 1e-6)).all())
    if a.shape[0] >= 600:      # exact ties and zero vectors are exactly 0.5
        disc = pb.metrics.triplet_accuracy(a, p, n).cpu()
        assert disc[3] == 0.5 and disc[5] == 0.5 and disc[11] == 0.5
    batch = pb.triplet.TripletBatch(anchor=a, positive=p, negative=n)
    assert torch.equal(pb.metrics.batch_triplet_accuracy(batch), pb.metrics.triplet_accuracy(a, p, n))


def test_resampled_recall_golden(pb):
    g = load_golden("resampled_g300.npz")
    V, A = g["V"].cuda(), g["A"].cuda()
    _, near = O.ranks_identity(g["V"], g["A"])
    torch.manual_seed(666)
    got = pb.metrics.resampled_recall(V, A, size=100, n_samples=6, n=10)
    assert got.shape == (6, 100)
    ix = g["sample_indices"]
    assert bool(((got == g["resampled_recall_n10"]) | near[ix]).all())
    torch.manual_seed(666)
    got = pb.metrics.resampled_recall_at_1_to_n(V, A, size=100, n_samples=4, N=10)
    assert got.shape == (4, 11, 100)
    assert bool(((got == g["resampled_1_to_10"]) | near[ix[:4]].unsqueeze(1)).all())
    with pytest.raises(AssertionError):
        pb.metrics.resampled_recall(V[:50], A[:50], size=100)
    with pytest.raises(AssertionError):
        pb.metrics.resampled_recall(V, A[:-1], size=100)


def test_triplet_scoring_golden(pb):
    g = load_golden("triplet_sampler_g240.npz")
    dur = g["duration"]
    random.seed(666)
    comp = pb.triplet.comparative_score_triplets([g["V"].cuda(), g["V2"].cuda()], [g["A"].cuda(), g["A2"].cuda()], dur,
                                                 n_samples=5)
    assert (comp["success"][0].cpu() - g["comp_success0"]).abs().max() < 2e-6
    assert (comp["success"][1].cpu() - g["comp_success1"]).abs().max() < 2e-6
    assert torch.equal(comp["duration"], g["comp_duration"])
    random.seed(666)
    sc = pb.triplet.score_triplets(g["V"].cuda(), g["A"].cuda(), dur, n_samples=5)
    assert (sc["accuracy"] - g["score_accuracy"]).abs().max() < 1e-6
    assert torch.equal(sc["duration"], g["score_duration"])


def test_error_conventions(pb):
    g = load_golden("sim_n8_a4.0.npz")
    with pytest.raises(ZeroDivisionError):
        pb.metrics.recall_at_n(g["V"].cuda(), g["A"].cuda(), torch.zeros(8, 8), n=1)
    z = g["V"].clone()
    z[2] = 0                                    # zero row -> NaN like pig/util.py:11-12
    assert torch.isnan(pb.loss.TripletLoss(0.2)(z.cuda(), g["A"].cuda()))
    assert torch.isnan(pb.util.cosine_matrix(z.cuda(), g["A"].cuda())[2]).all()
    out = pb.loss.TripletLoss(0.2)(g["V"], g["A"])      # CPU tensors in -> CPU result, computed on the GPU
    assert out.device.type == "cpu"


# -------------------------------------------------------------------------- seeded, larger sizes
@pytest.mark.parametrize("n,alpha", [(1024, 4.0), (1024, 0.5), (1000, 4.0), (2304, 1.0)])
def test_losses_against_oracle(pb, n, alpha):
    V, A = emb(n, alpha)
    for kind, mod, ref in (("hinge", pb.loss.TripletLoss(0.2), lambda: O.hinge_loss_and_grads(V, A, 0.2)),
                           ("milnce", pb.loss.MILNCELoss(), lambda: O.milnce_loss_and_grads(V, A))):
        loss, dV, dA = _grads(mod, V, A)
        rl, rdv, rda = ref()
        assert rel_err(loss, rl) < TOL, kind
        assert rel_err(dV, rdv) < TOL and rel_err(dA, rda) < TOL, kind
        if kind == "hinge":           # row by row: no wrong row hides under a large max (kinks: see hinge_rows_ok)
            from oracle import blockwise as B
            k = B.hinge_kink_counts(V, A, 0.2)
            hinge_rows_ok(dV, rdv, V, k)
            hinge_rows_ok(dA, rda, A, k)
        else:
            assert row_rel_err(dV, rdv) < TOL and row_rel_err(dA, rda) < TOL, kind


def test_non_unit_norm_inputs(pb):
    """cosine_matrix is scale invariant per row; the fused path must be too (H2 of SURVEY 7)."""
    V, A = emb(512, 4.0)
    g = torch.Generator().manual_seed(3)
    V = (V * torch.empty(512, 1).uniform_(0.05, 20.0, generator=g)).bfloat16().float()
    A = (A * torch.empty(512, 1).uniform_(0.05, 20.0, generator=g)).bfloat16().float()
    loss, dV, dA = _grads(pb.loss.TripletLoss(0.2), V, A)
    rl, rdv, rda = O.hinge_loss_and_grads(V, A, 0.2)
    assert rel_err(loss, rl) < TOL and rel_err(dV, rdv) < TOL and rel_err(dA, rda) < TOL
    ranks, near = O.ranks_identity(V, A)
    got = pb.metrics.recall_at_n(V.cuda(), A.cuda(), None, n=5)
    assert bool((((ranks < 5).float() == got) | near).all())


@pytest.mark.parametrize("n,alpha", [(4096, 4.0), (4096, 0.5), (5000, 4.0)])
def test_ranks_against_oracle(pb, n, alpha):
    V, A = emb(n, alpha)
    ranks, near = O.ranks_identity(V, A)
    got = pb.metrics.recall_at_1_to_n(V.cuda(), A.cuda(), None, N=10)
    for k in range(1, 11):
        assert bool((((ranks < k).float() == got[k]) | near).all()), k
    if alpha >= 4.0:        # realistic regime: near-ties are rare (the adversarial one has ~5 % of rows)
        assert int(near.sum()) < n // 100


def test_retrieval_16k_properties(pb):
    """C3 size (16384 x 16384): full oracle takes ~40 s on CPU, so check (i) exact ranks on a random
    512-row subset against the blockwise oracle and (ii) size-independent properties."""
    n = 16384
    V, A = emb(n, 4.0)
    got = pb.metrics.recall_at_1_to_n(V.cuda(), A.cuda(), None, N=10)
    assert got.shape == (11, n)
    assert bool((got[1:] >= got[:-1]).all())                    # recall@n is monotone in n
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:512]
    R = A[rows] / A[rows].norm(dim=1, keepdim=True)             # queries = references = audio
    C = V / V.norm(dim=1, keepdim=True)
    d = 1 - R @ C.T
    pos = d[torch.arange(512), rows].unsqueeze(1)
    ranks = (d < pos).sum(1)
    near = ((d - pos).abs() <= 1e-6).sum(1) > 1
    for k in (1, 5, 10):
        assert bool((((ranks < k).float() == got[k][rows]) | near).all())
    # permutation invariance: shuffling the gallery (and the targets with it) leaves recall unchanged
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    got_p = pb.metrics.recall_at_n(V[perm].cuda(), A.cuda(), inv.cuda(), n=10)
    assert bool(((got_p == got[10]) | near_rows_gpu(V, A)).all())      # only rows in a 1e-6 near-tie may move


def test_triplets_1m_properties(pb):
    """C4 size: 1M triplets.  Antisymmetry (swap positive and negative -> 1 - acc) and agreement with
    the oracle on a 4096-row slice."""
    t = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(5)
    a, p, n = (torch.randn(t, 512, device="cuda", generator=g).bfloat16() for _ in range(3))
    acc = pb.metrics.triplet_accuracy(a, p, n)
    swapped = pb.metrics.triplet_accuracy(a, n, p)
    assert torch.equal(acc.float() + swapped.float(), torch.ones(t, device="cuda"))
    sl = slice(12345, 12345 + 4096)
    ref = O.triplet_accuracy(a[sl].float().cpu(), p[sl].float().cpu(), n[sl].float().cpu())
    gap = O.triplet_accuracy(a[sl].float().cpu(), p[sl].float().cpu(), n[sl].float().cpu(), discrete=False)
    assert bool(((acc[sl].float().cpu() == ref) | (gap.abs() < 1e-6)).all())


def test_empty_and_tiny_inputs(pb):
    e = torch.empty(0, 512, device="cuda")
    assert pb.metrics.triplet_accuracy(e, e, e).shape == (0,)
    V, A = emb(1, 4.0)
    assert pb.metrics.recall_at_n(V.cuda(), A.cuda(), torch.eye(1), n=1).tolist() == [1.0]
    loss, dV, dA = _grads(pb.loss.TripletLoss(0.2), *emb(2, 4.0))
    rl, rdv, rda = O.hinge_loss_and_grads(*emb(2, 4.0), 0.2)
    assert rel_err(loss, rl) < TOL


@pytest.mark.parametrize("n,block", [(1536, 32768), (1280, 512), (1000, 384)])
def test_gallery_step_single_gpu(pb, n, block):
    """peppa_b200.gallery.GalleryStep (loss fwd+bwd + recall from one S pass, blocked gradient matrix)
    against the oracle's closed forms; rows = audio, columns = video."""
    from peppa_b200.gallery import GalleryStep
    V, A = emb(n, 4.0)
    step = GalleryStep(n, 512, margin=0.2, top_n=10, block=block)
    out = step.run(A.cuda().bfloat16(), V.cuda().bfloat16())
    loss, dA, dV = O.hinge_loss_and_grads(A, V, 0.2)
    assert rel_err(out["loss"].cpu(), loss) < TOL
    assert rel_err(out["dA"].cpu(), dA) < TOL and rel_err(out["dV"].cpu(), dV) < TOL
    from oracle import blockwise as B
    k = B.hinge_kink_counts(A, V, 0.2)
    hinge_rows_ok(out["dA"], dA, A, k)
    hinge_rows_ok(out["dV"], dV, V, k)
    assert rel_err(out["loss"].cpu(), O.triplet_loss(V, A, 0.2)) < TOL        # symmetric in (V, A)
    ranks, near = O.ranks_identity(V, A)
    assert bool(((out["ranks"].cpu().long() == ranks) | near).all())
    for k in (1, 5, 10):
        assert abs(out["recall"][k].item() - (out["ranks"] < k).float().mean().item()) < 1e-6
    again = step.run(A.cuda().bfloat16(), V.cuda().bfloat16())                 # deterministic, reusable buffers
    assert torch.equal(again["dA"], out["dA"]) and again["loss"].item() == out["loss"].item()


@pytest.mark.parametrize("n", [3, 70, 130, 257])
def test_gallery_step_tiny_sizes(pb, n):
    """The one-byte gradient matrix / kind::i8 path on galleries smaller than one tile (TMA clips every box)."""
    from oracle import blockwise as B
    from peppa_b200.gallery import GalleryStep
    V, A = emb(n, 4.0)
    out = GalleryStep(n, 512).run(A.cuda().bfloat16(), V.cuda().bfloat16())
    loss, dA, dV = O.hinge_loss_and_grads(A, V, 0.2)
    assert rel_err(out["loss"].cpu(), loss) < TOL
    k = B.hinge_kink_counts(A, V, 0.2)
    if float(dA.abs().max()) > 0:
        assert rel_err(out["dA"].cpu(), dA) < 2 * TOL and rel_err(out["dV"].cpu(), dV) < 2 * TOL   # few terms: 16-bit planes average less
        assert B.hinge_rows_within(out["dA"].cpu(), dA, A, k, 3 * TOL)[0]
    ranks, near = O.ranks_identity(V, A)
    assert bool(((out["ranks"].cpu().long() == ranks) | near).all())


def test_triplet_loss_public_api_beyond_one_block(pb):
    """TripletLoss through the public API at N = 40000 > 32768: the block-walking path with the one-byte gradient
    matrix (blocks of 32768 and 7232), against the blockwise oracle (whole loss; 128 sampled gradient rows in fp64)."""
    from oracle import blockwise as B
    n = 40000
    g = torch.Generator(device="cuda").manual_seed(11)
    V = torch.nn.functional.normalize(torch.randn(n, 512, generator=g, device="cuda"), dim=1)
    A = torch.nn.functional.normalize(4.0 * V + torch.randn(n, 512, generator=g, device="cuda"), dim=1).bfloat16()
    V = V.bfloat16()
    v, a = V.clone().requires_grad_(True), A.clone().requires_grad_(True)
    loss = pb.loss.TripletLoss(0.2)(v, a)
    loss.backward()
    ref = B.hinge_loss_blockwise(V, A, 0.2)
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(2))[:128].cuda()
    gV = B.hinge_grad_rows(V, A, rows, 0.2)
    gA = B.hinge_grad_rows(A, V, rows, 0.2)
    # bf16 gradient tensors (the inputs' dtype): 2^-9 relative rounding per element on top of the 1e-3 bar
    assert v.grad.dtype == torch.bfloat16 and rel_err(v.grad[rows].float(), gV) < 2.0 ** -8 and rel_err(a.grad[rows].float(), gA) < 2.0 ** -8
    vf, af = V.float().requires_grad_(True), A.float().requires_grad_(True)          # fp32 tensors in: fp32 gradients out
    pb.loss.TripletLoss(0.2)(vf, af).backward()
    assert rel_err(vf.grad[rows], gV) < TOL and rel_err(af.grad[rows], gA) < TOL
    assert row_rel_err(vf.grad[rows], gV) < TOL and row_rel_err(af.grad[rows], gA) < TOL


def test_rank_ties_are_strict(pb):
    """Duplicate gallery rows give candidates whose score equals the positive's exactly: they must not
    count (dist < dist_pos is strict, pig/metrics.py:8-12), in both the rank kernel and the fused
    hinge+rank kernel."""
    from peppa_b200.gallery import GalleryStep
    n = 768
    V, A = emb(n, 4.0)
    V[1::2] = V[0::2]                      # every odd video row duplicates the even one before it
    d = 1 - O.cosine_matrix(A, V)
    pos = d[torch.arange(n), torch.arange(n)].unsqueeze(1)
    ranks = (d < pos).sum(1)
    near = (((d - pos).abs() <= 1e-6).sum(1) > 2)      # the positive and its duplicate are expected
    got = pb.metrics._pair_ranks(V.cuda(), A.cuda(), None)[0].cpu().long()
    assert bool(((got == ranks) | near).all())
    fused = GalleryStep(n, 512).run(A.cuda().bfloat16(), V.cuda().bfloat16())["ranks"].cpu().long()
    assert torch.equal(fused, got)          # same thresholds, same two roundings: bit-identical counts


def test_resampled_recall_paths_agree(pb):
    """The one-matrix fast path (pb2_sim_matrix + pb2_subset_rank) and the per-subset path draw the same
    subsets and give the same recall (up to near-ties, which the two arithmetic routes may order differently)."""
    V, A = emb(700, 4.0, seed=21)
    torch.manual_seed(7)
    fast = pb.metrics.resampled_recall_at_1_to_n(V.cuda(), A.cuda(), size=100, n_samples=20, N=10)
    old = pb.metrics._RESAMPLE_MATRIX_LIMIT
    pb.metrics._RESAMPLE_MATRIX_LIMIT = 0
    try:
        torch.manual_seed(7)
        slow = pb.metrics.resampled_recall_at_1_to_n(V.cuda(), A.cuda(), size=100, n_samples=20, N=10)
    finally:
        pb.metrics._RESAMPLE_MATRIX_LIMIT = old
    assert fast.shape == slow.shape == (20, 11, 100)
    torch.manual_seed(7)
    ix = torch.stack([torch.randperm(700)[:100] for _ in range(20)])           # the draws of pig/metrics.py:79-81
    # a query is exempt only if, INSIDE ITS SUBSET, another candidate lies within 1e-6 of the positive
    d = 1 - O.cosine_matrix(A, V)
    near = torch.stack([(((d[i][:, i] - d[i, i].unsqueeze(1)).abs() <= 1e-6).sum(1) > 1) for i in ix])
    assert bool(((fast == slow).all(dim=1) | near).all())
    torch.manual_seed(7)
    ref = O.resampled_recall_at_1_to_n(V, A, size=100, n_samples=20, N=10)
    assert bool(((fast == ref).all(dim=1) | near).all())


def test_cosine_matrix_odd_shapes_and_backward(pb):
    g = torch.Generator().manual_seed(9)
    U = torch.randn(130, 512, generator=g).bfloat16().float()
    W = torch.randn(257, 512, generator=g).bfloat16().float()
    M = pb.util.cosine_matrix(U.cuda(), W.cuda())
    assert M.shape == (130, 257)
    assert (M.cpu() - O.cosine_matrix(U, W)).abs().max() < 2e-6
    u = U.cuda().requires_grad_(True)
    w = W.cuda().requires_grad_(True)
    T = torch.randn(130, 257, generator=g)
    (pb.util.cosine_matrix(u, w) * T.cuda()).sum().backward()
    ur = U.clone().requires_grad_(True)
    wr = W.clone().requires_grad_(True)
    (O.cosine_matrix(ur, wr) * T).sum().backward()
    assert rel_err(u.grad.cpu(), ur.grad) < 2e-3 and rel_err(w.grad.cpu(), wr.grad) < 2e-3   # fp16 cast of dS


def test_api_conformance_dtypes_layouts(pb):
    """Inputs the reference accepts must keep working: fp16 / fp32 / bf16, non-contiguous views, CPU tensors,
    an embedding size that is not a multiple of 64 (zero-padded internally), and the result types."""
    g = torch.Generator().manual_seed(12)
    V = torch.nn.functional.normalize(torch.randn(200, 300, generator=g), dim=1)       # full fp32 significands:
    A = torch.nn.functional.normalize(2.0 * V + torch.randn(200, 300, generator=g), dim=1)   # not bf16-representable
    for make in (lambda x: x.cuda(), lambda x: x.cuda().half(), lambda x: x.cuda().bfloat16(),
                 lambda x: x.cuda().t().contiguous().t(), lambda x: x):
        v = make(V).requires_grad_(True)
        a = make(A).requires_grad_(True)
        # the oracle sees exactly the values handed in (fp32 math on them), whatever their dtype
        Vo, Ao = v.detach().float().cpu(), a.detach().float().cpu()
        rl, rdv, rda = O.hinge_loss_and_grads(Vo, Ao, 0.2)
        ranks, near = O.ranks_identity(Vo, Ao)
        loss = pb.loss.TripletLoss(0.2)(v, a)
        loss.backward()
        assert loss.device == v.device and v.grad.shape == (200, 300) and v.grad.dtype == v.dtype
        tol = TOL if v.dtype == torch.float32 else 6e-3          # half-precision gradients are rounded to 2^-9 / 2^-11
        assert rel_err(loss.float().cpu(), rl) < TOL
        assert rel_err(v.grad.float().cpu(), rdv) < tol and rel_err(a.grad.float().cpu(), rda) < tol
        rec = pb.metrics.recall_at_n(make(V), make(A), None, n=3)
        assert bool((((ranks < 3).float() == rec) | near).all())
    # triplet_accuracy: dim argument and broadcasting, like F.cosine_similarity
    a3 = torch.randn(5, 7, 64, generator=g)
    p3 = torch.randn(5, 7, 64, generator=g)
    n3 = torch.randn(1, 7, 64, generator=g)
    got = pb.metrics.triplet_accuracy(a3.cuda(), p3.cuda(), n3.cuda(), dim=2)
    assert got.shape == (5, 7) and torch.equal(got.cpu(), O.triplet_accuracy(a3, p3, n3, dim=2))
    got = pb.metrics.triplet_accuracy(a3.cuda(), p3.cuda(), n3.cuda(), dim=1, discrete=False)
    assert got.shape == (5, 64)
    assert (got.cpu() - O.triplet_accuracy(a3, p3, n3, dim=1, discrete=False)).abs().max() < 2e-6


@pytest.mark.parametrize("name", golden_files("milnce_n"))
def test_milnce_k_candidates_golden(pb, name):
    """MILNCELoss with K = len(A) / len(V) > 1 candidates per video against the reference's own outputs."""
    g = load_golden(name)
    v = g["V"].cuda().requires_grad_(True)
    a = g["A"].cuda().requires_grad_(True)
    loss = pb.loss.MILNCELoss()(v, a)
    loss.backward()
    assert rel_err(loss, g["loss"]) < 1e-3
    assert rel_err(v.grad, g["dV"]) < 1e-3 and rel_err(a.grad, g["dA"]) < 1e-3


def test_milnce_k_candidates_oracle_and_errors(pb):
    from oracle import pig_oracle as O
    gen = torch.Generator().manual_seed(5)
    n, k = 600, 4
    V = (2.0 * torch.nn.functional.normalize(torch.randn(n, 512, generator=gen), dim=1)).bfloat16().float()
    A = (2.0 * V.repeat_interleave(k, dim=0) + 0.5 * torch.randn(n * k, 512, generator=gen)).bfloat16().float()
    v0, a0 = V.clone().requires_grad_(True), A.clone().requires_grad_(True)
    ref = O.milnce_loss(v0, a0)
    ref.backward()
    v, a = V.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    loss = pb.loss.MILNCELoss()(v, a)
    loss.backward()
    assert rel_err(loss, ref.detach()) < 1e-3
    assert rel_err(v.grad, v0.grad) < 1e-3 and rel_err(a.grad, a0.grad) < 1e-3
    with pytest.raises(RuntimeError, match="invalid for input of size"):     # the reference's view() error
        pb.loss.MILNCELoss()(V.cuda(), A[:n * k - 1].cuda())


def test_install_patches_a_pig_package(pb):
    """peppa_b200.install() over a stand-in `pig` package: the hot-path names resolve to the B200 modules."""
    import sys
    import types

    import peppa_b200
    pig = types.ModuleType("pig_standin")
    util = types.ModuleType("pig_standin.util")
    util.cosine_matrix = lambda U, V: None
    sys.modules["pig_standin"], sys.modules["pig_standin.util"] = pig, util
    pig.util = util
    try:
        peppa_b200.install(pig)
        assert pig.loss is pb.loss and pig.metrics is pb.metrics
        assert sys.modules["pig_standin.util"].cosine_matrix is pb.util.cosine_matrix
        V, A = emb(32, 4.0)
        assert pig.loss.TripletLoss(0.2)(V.cuda(), A.cuda()).item() > 0
    finally:
        for k in [k for k in sys.modules if k.startswith("pig_standin")]:
            del sys.modules[k]


def test_milnce_temperature_extension(pb):
    """temperature defaults to the reference (none); with tau the logits are V A^T / tau."""
    V, A = emb(512, 4.0)
    tau = 0.25
    loss, dV, dA = _grads(pb.loss.MILNCELoss(temperature=tau), V, A)
    v = V.double().requires_grad_(True)
    a = A.double().requires_grad_(True)
    x = (v @ a.T) / tau
    den = torch.logaddexp(torch.logsumexp(x, 1), torch.logsumexp(x, 0))
    ref = (den - torch.diagonal(x)).mean()
    ref.backward()
    assert rel_err(loss, ref.detach()) < TOL and rel_err(dV, v.grad) < TOL and rel_err(dA, a.grad) < TOL


@pytest.mark.parametrize("n,block", [(1024, 32768), (1100, 384)])
def test_gallery_step_milnce_single_gpu(pb, n, block):
    """GalleryStep(loss='milnce'): rows = audio, columns = video; equals MILNCELoss up to the (V, A) swap."""
    from peppa_b200.gallery import GalleryStep
    V, A = emb(n, 4.0)
    out = GalleryStep(n, 512, block=block, loss="milnce").run(A.cuda().bfloat16(), V.cuda().bfloat16())
    loss, dA, dV = O.milnce_loss_and_grads(A, V)           # closed form with rows = first argument
    assert rel_err(out["loss"].cpu(), loss) < TOL
    assert rel_err(out["dA"].cpu(), dA) < TOL and rel_err(out["dV"].cpu(), dV) < TOL
    assert rel_err(out["loss"].cpu(), O.milnce_loss(V, A)) < TOL     # the loss itself is symmetric


@pytest.mark.parametrize("n,block,tau", [(1100, 384, 0.5), (5000, 2048, 0.07), (4096, 32768, 0.07)])
def test_gallery_step_milnce_with_recall(pb, n, block, tau):
    """GalleryStep(loss='milnce', with_recall=True): the recall ranks come out of the statistics pass
    (pb2_sim_lse_both_rank, block by block with row / column offsets).  Ranks equal the oracle's outside near-ties and the
    separate rank pass bit for bit; loss and gradients are those of the step without recall, bit for bit."""
    from peppa_b200.gallery import GalleryStep
    V, A = emb(n, 4.0)
    ab, vb = A.cuda().bfloat16(), V.cuda().bfloat16()
    fused = GalleryStep(n, 512, block=block, loss="milnce", temperature=tau, with_recall=True).run(ab, vb)
    split = GalleryStep(n, 512, block=block, loss="milnce", temperature=tau, with_recall=True)
    split.fuse_rank = False
    split = split.run(ab, vb)
    plain = GalleryStep(n, 512, block=block, loss="milnce", temperature=tau).run(ab, vb)
    assert torch.equal(fused["ranks"], split["ranks"]) and torch.equal(fused["recall"], split["recall"])
    assert torch.equal(fused["loss"], plain["loss"]) and torch.equal(fused["dA"], plain["dA"]) and torch.equal(fused["dV"], plain["dV"])
    ranks, near = O.ranks_identity(V, A)          # candidates = video (columns), references = audio (rows)
    got = fused["ranks"].cpu()
    assert bool(((got == ranks) | near).all())
    for k in (1, 5, 10):
        exact = (ranks < k).float().mean().item()
        assert abs(fused["recall"][k].item() - exact) <= near.float().mean().item() + 1e-6
    assert plain["recall"] is None and fused["recall"][0].item() == 0.0


@pytest.mark.parametrize("r,c,d,scale", [(1000, 1500, 512, 1.0), (300, 70, 64, 14.0), (4096, 4100, 256, 5.0), (129, 33, 512, 1.0 / 0.07)])
def test_sim_lse_both_against_fp64(pb, r, c, d, scale):
    """One pass over the logits gives both directions of the log-sum-exp (pig/loss.py:23-25: x and x.permute(1,0,2)):
    against fp64, against the two-pass kernels, block-wise accumulation, bit-reproducible, bound refusal."""
    from peppa_b200 import ops
    g = torch.Generator().manual_seed(r * 7 + c)
    X = torch.nn.functional.normalize(torch.randn(r, d, generator=g), dim=1).bfloat16()
    Y = torch.nn.functional.normalize(0.5 * X[torch.arange(c) % r].float() + torch.randn(c, d, generator=g), dim=1).bfloat16()
    xb, yb = X.cuda(), Y.cuda()
    bound = ops.logit_bound(xb, yb, scale)
    assert scale * 0.99 < bound < scale * 1.02                       # unit-norm rows up to bf16 rounding
    lr, lc = ops.sim_lse_both(xb, yb, bound, scale=scale)
    S = (X.double() @ Y.double().T) * scale
    assert (lr.cpu().double() - torch.logsumexp(S, 1)).abs().max() < 1e-4
    assert (lc.cpu().double() - torch.logsumexp(S, 0)).abs().max() < 1e-4
    assert (lr - ops.sim_lse_rows(xb, yb, scale=scale)).abs().max() < 1e-4
    assert (lc - ops.sim_lse_rows(yb, xb, scale=scale)).abs().max() < 1e-4
    lr2, lc2 = ops.sim_lse_both(xb, yb, bound, scale=scale)
    assert torch.equal(lr, lr2) and torch.equal(lc, lc2)             # fixed summation order, no atomics
    ar = torch.full((r,), float("-inf"), device="cuda")
    ac = torch.full((c,), float("-inf"), device="cuda")
    rh, chf = r // 2 + 3, c // 3 + 1
    for (r0, r1) in ((0, rh), (rh, r)):
        for (c0, c1) in ((0, chf), (chf, c)):
            ops.sim_lse_both(xb[r0:r1], yb[c0:c1], bound, scale=scale, lse_row=ar[r0:r1], lse_col=ac[c0:c1])
    assert (ar - lr).abs().max() < 1e-5 and (ac - lc).abs().max() < 1e-5
    with pytest.raises(RuntimeError):
        ops.sim_lse_both(xb, yb, 100.0, scale=scale)                 # 2^(-2 * 144) would underflow: refused


@pytest.mark.parametrize("n,scale", [(1000, 1.0), (4229, 1.0 / 0.07), (9000, 4.0)])
def test_sim_lse_both_rank_fused(pb, n, scale):
    """North-star kernels (a) + (b) from ONE pass over the scores: pb2_sim_lse_both_rank gives the log-sum-exp
    statistics of pb2_sim_lse_both (bit for bit) and the rank counts of pb2_sim_rank (bit for bit), whole and in
    blocks with row / column offsets; the ranks equal the oracle's outside near-ties."""
    from oracle import pig_oracle as O
    from peppa_b200 import ops
    V, A = emb(n, 4.0)
    vb, ab = V.cuda().bfloat16(), A.cuda().bfloat16()
    rv, _ = ops.row_norms(vb)
    ra, _ = ops.row_norms(ab)
    _, thr = ops.sim_diag(ab, vb, ra, rv)
    idx = torch.arange(n, device="cuda")
    bound = ops.logit_bound(ab, vb, scale)
    ref_rank = ops.sim_rank(ab, vb, ra, rv, thr, idx)
    ref_lr, ref_lc = ops.sim_lse_both(ab, vb, bound, scale=scale)
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    lr, lc = ops.sim_lse_both(ab, vb, bound, scale=scale, rank=(ra, rv, thr, 0, 0, cnt))
    assert torch.equal(lr, ref_lr) and torch.equal(lc, ref_lc)
    assert torch.equal(cnt, ref_rank)
    ranks, near = O.ranks_identity(V.bfloat16().float(), A.bfloat16().float())
    assert bool(((cnt.cpu() == ranks) | near).all())
    # blocks: the positive of local row i of a row block starting at r0 is global column r0 + i
    cnt2 = torch.zeros(n, dtype=torch.int32, device="cuda")
    ar = torch.full((n,), float("-inf"), device="cuda")
    ac = torch.full((n,), float("-inf"), device="cuda")
    rh, chf = n // 2 + 3, n // 3 + 1
    for (r0, r1) in ((0, rh), (rh, n)):
        for (c0, c1) in ((0, chf), (chf, n)):
            ops.sim_lse_both(ab[r0:r1], vb[c0:c1], bound, scale=scale, lse_row=ar[r0:r1], lse_col=ac[c0:c1],
                             rank=(ra[r0:r1], rv[c0:c1], thr[r0:r1], r0, c0, cnt2[r0:r1]))
    assert torch.equal(cnt2, ref_rank)
    assert (ar - ref_lr).abs().max() < 1e-5 and (ac - ref_lc).abs().max() < 1e-5


def test_cluster_variants_match_independent_ctas(pb):
    """The similarity pass runs on CTA pairs (cta_group::2, one M = 256 MMA per two SMs) once every SM has a tile;
    independent CTAs, the multicast clusters and pairs with a resident X strip (mode 3: one-byte hinge pass and rank
    pass) are the measurement build's options.  All four are bit-identical for
    every policy -- rank, both log-sum-exp passes, the MIL-NCE gradient matrix, the stored score matrix, the hinge pass
    with the fp16 and the one-byte gradient matrix -- also with an odd trailing row block."""
    from peppa_b200 import _cabi, ops
    n = 33 * 128 + 5
    V, A = emb(n, 4.0)
    vb, ab = V.cuda().bfloat16(), A.cuda().bfloat16()
    rv, _ = ops.row_norms(vb)
    ra, _ = ops.row_norms(ab)
    diag, thr = ops.sim_diag(ab, vb, ra, rv)
    idx = torch.arange(n, device="cuda")
    bound = ops.logit_bound(ab, vb, 4.0)

    def everything():
        out = [ops.sim_rank(ab, vb, ra, rv, thr, idx), *ops.sim_lse_both(ab, vb, bound, scale=4.0),
               ops.sim_lse_rows(ab, vb, scale=4.0), ops.sim_matrix(ab, vb, ra, rv).clone()]
        lr, lc = out[1], out[2]
        g, ld = ops.gmat_alloc(n, n, "cuda")
        g.zero_()
        ops.sim_lse_grad(ab, vb, lr, lc, g, ld, scale=4.0)
        out.append(g)
        for dt in (torch.float16, torch.uint8):
            g, ld = ops.gmat_alloc(n, n, "cuda", dt)
            g.zero_()
            rc, cc, rk = (torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(3))
            part = ops.sim_hinge(ab, vb, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk)
            out += [g, rc, cc, rk, part.sum()]
        return out

    res = {}
    with _cabi.measurement_library() as lib:      # the selectors exist in the measurement build only
        try:
            for mode in (0, 2, 3, 1):              # 3: pairs with a resident X strip (one-byte hinge pass, rank pass)
                lib.pb2_debug_sim_pair(mode)
                res[mode] = everything()
        finally:
            lib.pb2_debug_sim_pair(-1)
    res["product"] = everything()
    for x, y in zip(res[1], res["product"]):       # the product library runs the CTA pairs at this size
        assert torch.equal(x, y)
    for mode in (2, 3, 1):
        for k, (x, y) in enumerate(zip(res[0], res[mode])):
            if x.dim() == 0:                       # the loss: per-CTA partials, another tile-to-CTA deal
                assert abs(x.item() - y.item()) <= 1e-6 * abs(x.item()), (mode, k)
            else:
                assert torch.equal(x, y), (mode, k)
    ranks, near = O.ranks_identity(V, A)
    assert bool(((res[1][0].cpu() == ranks) | near).all())


def test_milnce_one_pass_statistics_path(pb):
    """MILNCELoss takes the one-pass row + column statistics from 2^26 logits on when the logits are bounded;
    force it at a small size (ragged blocks), and check that unbounded logits keep the two-pass kernels."""
    old = pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS
    pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS = 512, 0
    try:
        V, A = emb(1200, 4.0)
        for tau in (1.0, 0.07):
            loss, dV, dA = _grads(pb.loss.MILNCELoss(temperature=tau), V, A)
            v = V.double().requires_grad_(True)
            a = A.double().requires_grad_(True)
            x = (v @ a.T) / tau
            ref = (torch.logaddexp(torch.logsumexp(x, 1), torch.logsumexp(x, 0)) - torch.diagonal(x)).mean()
            ref.backward()
            assert rel_err(loss, ref.detach()) < TOL and rel_err(dV, v.grad) < TOL and rel_err(dA, a.grad) < TOL
        from peppa_b200 import ops
        calls = []
        orig = ops.sim_lse_both
        ops.sim_lse_both = lambda *a_, **k: calls.append(1) or orig(*a_, **k)
        try:
            Vs, As = (V * 8.0).bfloat16().float(), (A * 8.0).bfloat16().float()      # |logit| up to 64 > 41.6
            loss, dV, dA = _grads(pb.loss.MILNCELoss(), Vs, As)
            assert not calls
            rl, rdv, rda = O.milnce_loss_and_grads(Vs, As)
            assert rel_err(loss, rl) < TOL and rel_err(dV, rdv) < TOL and rel_err(dA, rda) < TOL
            _grads(pb.loss.MILNCELoss(), V, A)
            assert calls
        finally:
            ops.sim_lse_both = orig
    finally:
        pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS = old


def test_blocked_loss_paths(pb):
    """Batches larger than one gradient-matrix block walk blocks with accumulating gradient GEMMs; exercise
    that code with a tiny block edge."""
    old = pb.loss._MAX_BLOCK
    pb.loss._MAX_BLOCK = 512
    try:
        V, A = emb(1200, 4.0)
        for mod, ref in ((pb.loss.TripletLoss(0.2), lambda: O.hinge_loss_and_grads(V, A, 0.2)),
                         (pb.loss.MILNCELoss(), lambda: O.milnce_loss_and_grads(V, A))):
            loss, dV, dA = _grads(mod, V, A)
            rl, rdv, rda = ref()
            assert rel_err(loss, rl) < TOL and rel_err(dV, rdv) < TOL and rel_err(dA, rda) < TOL
    finally:
        pb.loss._MAX_BLOCK = old


def test_contrastive_matrix_backward(pb):
    g = torch.Generator().manual_seed(4)
    M = (torch.randn(193, 193, generator=g) * 0.2)
    m = M.cuda().requires_grad_(True)
    pb.loss.contrastive(m, margin=0.3).backward()
    mr = M.clone().requires_grad_(True)
    O.contrastive(mr, margin=0.3).backward()
    assert rel_err(m.grad.cpu(), mr.grad) < 1e-5


@pytest.mark.parametrize("d,dtype", [(64, torch.bfloat16), (768, torch.bfloat16), (100, torch.float32), (512, torch.float16)])
def test_triplet_kernel_shapes(pb, d, dtype):
    """Vector paths (D a multiple of 32 lanes x 16 bytes), the generic path (odd D) and sliced, unaligned views."""
    g = torch.Generator().manual_seed(d)
    a, p, n = (torch.randn(777, d, generator=g).to(dtype) for _ in range(3))
    ref = O.triplet_accuracy(a.float(), p.float(), n.float(), discrete=False)
    got = pb.metrics.triplet_accuracy(a.cuda(), p.cuda(), n.cuda(), discrete=False).float().cpu()
    assert (got - ref).abs().max() < (2e-6 if dtype == torch.float32 else 4e-3)       # result is cast to the input dtype
    wide = torch.randn(777, d + 3, generator=g).to(dtype).cuda()
    view = wide[:, 1:d + 1]                                                           # unaligned, strided rows
    got = pb.metrics.triplet_accuracy(view, p.cuda(), n.cuda(), discrete=False).float().cpu()
    ref = O.triplet_accuracy(view.float().cpu(), p.float(), n.float(), discrete=False)
    assert (got - ref).abs().max() < (2e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("r,c,d", [(20480, 2048, 512), (20000, 1000, 512), (9000, 4100, 256), (19000, 19000, 768),
                                   (19200, 19200, 1024), (300, 40000, 512)])
def test_grad_gemm_stream_k(pb, tr, r, c, d):
    """The stream-K decomposition of pb2_grad_gemm_ws (row blocks cut between CTAs, completed through the
    workspace) agrees with the whole-tile kernel and with fp64, and is bit-reproducible."""
    from peppa_b200 import ops
    torch.manual_seed(3)
    gm, ld = ops.gmat_alloc(r, c, "cuda")
    gm.zero_()
    gm[:, :c] = torch.randint(0, 3, (r, c), device="cuda").half()
    z = (torch.randn(r if tr else c, d, device="cuda") * 0.05).half()
    tiles = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=False)
    sk = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=True)
    sk2 = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=True)
    acc = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, out=torch.ones_like(sk), accumulate=True, alpha=0.5)
    G = gm[:, :c].double()
    ref = (G.T if tr else G) @ z.double()
    scale = ref.abs().max()
    tol = 1e-5 * max(1.0, ((r if tr else c) / 4096.0) ** 0.5)      # fp32 accumulation over the contraction length
    assert ((tiles - ref).abs().max() / scale).item() < tol
    assert ((sk - ref).abs().max() / scale).item() < tol
    assert ((acc - (1 + 0.5 * ref)).abs().max() / scale).item() < tol
    assert torch.equal(sk, sk2)


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("r,c,d", [(20480, 2048, 512), (20000, 1000, 512), (9000, 4100, 256), (300, 30000, 512),
                                   (32768, 32768, 512), (1200, 700, 768), (130, 70, 256), (5, 3, 512)])
def test_grad_gemm_i8_planes(pb, tr, r, c, d):
    """The kind::i8 gradient GEMM (one-byte G in {0, 1, 2} x two 8-bit planes of the normalised embeddings): EXACT
    against the integer product of the planes (s32 accumulation, exact fp32 join), within the quantisation step of
    the true product, stream-K == whole tiles bit for bit, accumulate, ragged edges."""
    from peppa_b200 import ops
    torch.manual_seed(5)
    gm, ld = ops.gmat_alloc(r, c, "cuda", torch.uint8)
    gm.fill_(7)                                                   # padding columns hold garbage: they must never be read
    gm[:, :c] = torch.randint(0, 3, (r, c), device="cuda", dtype=torch.uint8)
    k = r if tr else c
    z = torch.nn.functional.normalize(torch.randn(k, d, device="cuda"), dim=1).bfloat16()
    rinv, _ = ops.row_norms(z)
    planes = ops.rows_quant_i8(z, rinv)
    assert planes.shape == (k, 2 * d) and planes.dtype == torch.uint8
    hi = planes[:, :d].view(torch.int8).double()
    lo = planes[:, d:].double()
    q = 256.0 * hi + lo
    zh = z.double() * rinv.double().unsqueeze(1)
    assert (q / 32512.0 - zh).abs().max().item() <= 0.5 / 32512.0 + 1e-7          # round-to-nearest, 16 bits
    tiles = ops.grad_gemm(gm, r, c, ld, planes, transpose=tr, stream_k=False)
    sk = ops.grad_gemm(gm, r, c, ld, planes, transpose=tr, stream_k=True)
    acc = ops.grad_gemm(gm, r, c, ld, planes, transpose=tr, out=torch.ones_like(sk), accumulate=True, alpha=0.5)
    G = gm[:, :c].double()
    exact = ((G.T if tr else G) @ q) / 32512.0
    true = (G.T if tr else G) @ zh
    scale = true.abs().max()
    assert ((tiles.double() - exact).abs().max() / scale).item() < 2e-7          # integers are exact; one fp32 rounding
    assert torch.equal(sk, tiles)                                                # the stream-K fold adds integers
    assert ((acc.double() - (1 + 0.5 * exact)).abs().max() / scale).item() < 1e-6
    # 16-bit planes, one scale per tensor: each operand entry is off by at most 1.5e-5 (sigma 8.9e-6); against a RANDOM
    # {0, 1, 2} matrix that is a relative rms error of 2e-4 whatever K, and ~2e-4 of the largest entry at 5 sigma
    # (the hinge gradient matrix is far from random: whole-gradient errors are 4e-6 at N = 2^20, bench check.verified)
    err = tiles.double() - true
    assert (err.abs().max() / scale).item() < 6e-4
    assert (err.pow(2).mean().sqrt() / true.pow(2).mean().sqrt()).item() < 3e-4


def test_byte_gradient_matrix_equals_fp16_path(pb):
    """GalleryStep with the one-byte gradient matrix (default for the hinge loss) against the fp16 matrix: identical
    loss, counts and ranks (the similarity pass is the same arithmetic), gradients within the planes' 16-bit
    quantisation; and the fused pass writes exactly the bytes {0, 1, 2} the fp16 pass writes as halves."""
    from peppa_b200 import ops
    from peppa_b200.gallery import GalleryStep
    n = 3000                                                       # ragged: 3000 = 23 * 128 + 56 columns, blocks of 1024
    V, A = emb(n, 4.0)
    a, v = A.cuda().bfloat16(), V.cuda().bfloat16()
    o8 = GalleryStep(n, 512, block=1024, byte_gmat=True).run(a, v)
    o16 = GalleryStep(n, 512, block=1024, byte_gmat=False).run(a, v)
    assert o8["loss"].item() == o16["loss"].item() and torch.equal(o8["ranks"], o16["ranks"])
    assert rel_err(o8["dA"], o16["dA"]) < 3e-4 and rel_err(o8["dV"], o16["dV"]) < 3e-4
    loss, dA, dV = O.hinge_loss_and_grads(A, V, 0.2)
    assert rel_err(o8["dA"].cpu(), dA) < TOL and rel_err(o8["dV"].cpu(), dV) < TOL
    from oracle import blockwise as B
    k = B.hinge_kink_counts(A, V, 0.2)
    hinge_rows_ok(o8["dA"], dA, A, k)
    hinge_rows_ok(o8["dV"], dV, V, k)
    # the matrices themselves
    ra, _ = ops.row_norms(a)
    rv, _ = ops.row_norms(v)
    diag = ops.pair_dot(a, v, rinv_x=ra, rinv_y=rv)
    mats = {}
    for dt in (torch.uint8, torch.float16):
        g, ld = ops.gmat_alloc(n, n, "cuda", dt)
        rc, cc = torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda")
        ops.sim_hinge(a, v, ra, rv, diag, diag, 0.2, rc, cc, g, ld)
        mats[dt] = (g[:, :n].to(torch.int32), rc, cc)
    assert torch.equal(mats[torch.uint8][0], mats[torch.float16][0]) and int(mats[torch.uint8][0].max()) == 2
    assert torch.equal(mats[torch.uint8][1], mats[torch.float16][1]) and torch.equal(mats[torch.uint8][2], mats[torch.float16][2])


@pytest.mark.parametrize("rows,n_in,n_out,bias", [(1000, 512, 512, True), (5, 512, 512, False), (300, 28, 512, True),
                                                  (4096, 768, 256, True), (129, 512, 64, True), (2000, 1024, 384, True),
                                                  # several 256-row tiles per CTA pair (the resident-x kernel's barrier
                                                  # phases across tiles; the interleaved kernel's for n_in > 512)
                                                  (70001, 512, 512, True), (40000, 256, 256, False), (50000, 768, 512, True)])
def test_encoder_tail_project_normalize(pb, rows, n_in, n_out, bias):
    """nn.Linear + F.normalize of the reference encoders (pig/models.py:96-109, :130-150) as one
    tcgen05 kernel; bf16 operands, fp32 accumulate, bf16 output.  Tolerances: the output is the bf16 rounding of
    the fp32 result (2^-8 of the largest component); rinv / norm are fp32 statistics."""
    from peppa_b200 import encoder
    g = torch.Generator().manual_seed(rows + n_in)
    x = torch.randn(rows, n_in, generator=g).bfloat16().float()
    lin = torch.nn.Linear(n_in, n_out, bias=bias)
    with torch.no_grad():
        lin.weight.copy_(lin.weight.bfloat16().float())
    ref_in = x.clone().requires_grad_(True)
    ref = torch.nn.functional.normalize(lin(ref_in), p=2, dim=1)
    mod = encoder.ProjectNormalize.from_linear(lin).cuda()
    xin = x.cuda().requires_grad_(True)
    out, rinv = mod(xin, return_rinv=True)
    assert out.dtype == torch.bfloat16 and out.shape == (rows, n_out) and rinv.shape == (rows,)
    assert (out.float().cpu() - ref.detach()).abs().max().item() <= 2.0 ** -8 * ref.detach().abs().max().item() + 1e-6
    assert rel_err(rinv, 1.0 / out.float().norm(dim=1)) < 1e-5
    # backward through the normalisation Jacobian and the projection
    w = torch.randn(rows, n_out, generator=g)
    (ref * w).sum().backward()
    (out.float() * w.cuda()).sum().backward()
    assert rel_err(xin.grad, ref_in.grad) < 2e-2
    assert rel_err(mod.weight.grad, lin.weight.grad) < 2e-2
    if bias:
        assert rel_err(mod.bias.grad, lin.bias.grad) < 2e-2


def test_encoder_tail_column_split_variant(pb):
    """The product picks the resident-x phased kernel (n_in <= 512, n_out % 256 == 0) or the interleaved kernel; the
    measurement build's other kernels -- interleaved always, column split (row statistics exchanged through distributed
    shared memory with st.async), phased halves, sixteen epilogue warps, resident x with 128-column phases -- give the
    same embeddings."""
    from peppa_b200 import _cabi, ops
    g = torch.Generator().manual_seed(11)
    for rows, n_in, n_out in ((3000, 512, 512), (129, 256, 128), (2000, 1024, 384), (700, 512, 64), (1300, 192, 256)):
        x = torch.randn(rows, n_in, generator=g).bfloat16().cuda()
        w = (torch.randn(n_out, n_in, generator=g) / n_in ** 0.5).bfloat16().cuda()
        b = torch.randn(n_out, generator=g).cuda()
        o1, r1, m1 = ops.project_normalize(x, w, b)
        for variant in (1, 2, 3, 4, 5):  # 1: interleaved kernel; 2: column split over a CTA pair; 3: phased halves; 4: sixteen epilogue warps; 5: resident x with 128-column phases
            if variant == 2 and n_out % 128 != 0:
                continue
            with _cabi.measurement_library() as lib:
                lib.pb2_debug_proj_variant(variant)
                try:
                    o2, r2, m2 = ops.project_normalize(x, w, b)
                    torch.cuda.synchronize()
                finally:
                    lib.pb2_debug_proj_variant(0)
            assert (o1.float() - o2.float()).abs().max().item() <= 2.0 ** -8 * o1.float().abs().max().item()
            assert rel_err(r2, 1.0 / o2.float().norm(dim=1)) < 1e-5       # rinv belongs to ITS rounded rows
            assert rel_err(m2, m1) < 1e-5 and rel_err(r2, r1) < 1e-3


def test_encoder_tail_feeds_the_loss(pb):
    """The tail's (bf16 rows, rinv) pair is what the scoring kernels consume: TripletLoss on the fused
    embeddings equals the reference pipeline project -> normalize -> TripletLoss on the same bf16 weights."""
    from oracle import pig_oracle as O
    from peppa_b200 import encoder
    g = torch.Generator().manual_seed(3)
    n = 384
    fv, fa = torch.randn(n, 512, generator=g).bfloat16().float(), torch.randn(n, 512, generator=g).bfloat16().float()
    fa = (fa + 0.2 * fv).bfloat16().float()        # weakly related pairs: a non-trivial hinge loss (~0.07)
    pv, pa = torch.nn.Linear(512, 512), torch.nn.Linear(512, 512)
    with torch.no_grad():
        pa.weight.copy_(pv.weight)
        pa.bias.copy_(pv.bias)
        for m in (pv, pa):
            m.weight.copy_(m.weight.bfloat16().float())
    V = torch.nn.functional.normalize(pv(fv), dim=1)
    A = torch.nn.functional.normalize(pa(fa), dim=1)
    ref = O.triplet_loss(V.detach(), A.detach(), 0.2)
    ev = encoder.ProjectNormalize.from_linear(pv).cuda()(fv.cuda())
    ea = encoder.ProjectNormalize.from_linear(pa).cuda()(fa.cuda())
    got = pb.loss.TripletLoss(0.2)(ev, ea)
    assert ref.item() > 1e-3 and rel_err(got, ref) < 5e-3     # embeddings are bf16-rounded before the loss


def test_encoder_tail_rinv_is_consumed_not_recomputed(pb):
    """SURVEY 8f row 3, second half: the tail's fp32 1/||row|| travels with its bf16 rows (a tag on the tensor) and
    TripletLoss / recall_at_n / GalleryStep take it instead of running pb2_row_norms (or the norm part of
    pb2_hinge_prep) again.  The tail sums the squares in its own order, so the two 1/||row|| agree to fp32 rounding,
    not bit for bit; results agree to ~1e-6.  An in-place edit of the rows voids the tag."""
    from peppa_b200 import encoder, ops
    from peppa_b200.gallery import GalleryStep
    g = torch.Generator().manual_seed(3)
    n = 1024
    fv = torch.randn(n, 512, generator=g).bfloat16().cuda()
    fa = (fv.float().cpu() * 0.3 + torch.randn(n, 512, generator=g)).bfloat16().cuda()
    tail = encoder.ProjectNormalize(512, 512).cuda()
    ev, rv = tail(fv, return_rinv=True)
    ea, ra = tail(fa, return_rinv=True)
    assert ops.known_rinv(ev, ops.as_rows(ev)) is rv and ops.known_rinv(ea, ops.as_rows(ea)) is ra
    assert (rv - ops.row_norms(ev.detach())[0]).abs().max().item() < 1e-6
    calls = []
    orig = ops.row_norms
    ops.row_norms = lambda x: calls.append(tuple(x.shape)) or orig(x)
    try:
        lt = pb.loss.TripletLoss(0.2)(ev, ea)                                  # fused step: norms skipped in hinge_prep
        rt = pb.metrics.recall_at_n(ev, ea, None, n=10)
        gt = GalleryStep(n, 512).run(ea.detach(), ev.detach(), rinv_a=ra, rinv_v=rv)
        assert calls == []                                                     # nothing recomputed a norm
        lu = pb.loss.TripletLoss(0.2)(ev.detach().clone(), ea.detach().clone())   # untagged copies: norms recomputed
        ru = pb.metrics.recall_at_n(ev.detach().clone(), ea.detach().clone(), None, n=10)
        gu = GalleryStep(n, 512).run(ea.detach().clone(), ev.detach().clone())
        assert len(calls) >= 4
    finally:
        ops.row_norms = orig
    assert abs(lt.item() - lu.item()) < 1e-6 * abs(lu.item()) and abs(gt["loss"].item() - gu["loss"].item()) < 1e-6 * abs(lu.item())
    assert int((rt != ru).sum()) <= 2 and rel_err(gt["dA"], gu["dA"]) < 1e-5       # a 1-ulp rinv may move a near-tie
    lt.backward()                                                              # gradients flow back into the tail
    assert tail.weight.grad is not None and bool(torch.isfinite(tail.weight.grad).all())
    ev.detach().mul_(1.0)                                                      # any in-place edit bumps the version
    assert ops.known_rinv(ev, ops.as_rows(ev)) is None


def test_embedding_store_scoring(pb, tmp_path):
    """a gallery scored from the sharded store equals the same embeddings scored directly,
    and an evaluation row built from stores matches the oracle metrics with the reference's seeds."""
    from peppa_b200 import store
    from peppa_b200.gallery import GalleryStep
    V, A = emb(1536, 4.0)
    g = torch.Generator().manual_seed(5)
    dur = torch.randint(20, 60, (1536,), generator=g).float() / 10.0
    p = str(tmp_path / "fixed")
    with store.EmbeddingStoreWriter(p, 512, rows_per_shard=500) as w:
        w.append(V, A, dur)
    out = store.score_store(p, with_grad=False)
    ref = GalleryStep(1536, 512, with_grad=False).run(A.bfloat16().cuda(), V.bfloat16().cuda())
    assert torch.equal(out["ranks"], ref["ranks"]) and torch.equal(out["loss"], ref["loss"])
    assert rel_err(out["loss"], O.triplet_loss(V, A, 0.2)) < TOL
    # evaluation row (pig/evaluation.py:78-110) from stores, small sample counts
    Vj, Aj = emb(1536, 2.0, seed=7)
    pj = str(tmp_path / "jitter")
    with store.EmbeddingStoreWriter(pj, 512, rows_per_shard=4096) as w:
        w.append(Vj, Aj)
    random.seed(666)
    torch.manual_seed(666)
    row = store.evaluation_row("narration", False, p, pj, n_samples=3, size=100, N=10)
    random.seed(666)
    torch.manual_seed(666)
    ref_fixed = O.resampled_recall_at_1_to_n(V, A, size=100, n_samples=3, N=10)
    ref_jit = O.resampled_recall_at_1_to_n(Vj, Aj, size=100, n_samples=3, N=10)
    ref_acc = O.score_triplets(V, A, dur, n_samples=3)["accuracy"]
    assert row["recall_fixed"].shape == (3, 11, 100)
    # the subsets the reference's RNG draws (3 randperms for the fixed set, then 3 for the jittered one); a query is
    # exempt only when another candidate of ITS subset lies within 1e-6 of the positive
    torch.manual_seed(666)
    ix_f = torch.stack([torch.randperm(1536)[:100] for _ in range(3)])
    ix_j = torch.stack([torch.randperm(1536)[:100] for _ in range(3)])

    def near_of(Vx, Ax, ix):
        d = 1 - O.cosine_matrix(Ax, Vx)
        return torch.stack([(((d[i][:, i] - d[i, i].unsqueeze(1)).abs() <= 1e-6).sum(1) > 1) for i in ix])

    assert bool(((row["recall_fixed"].cpu() == ref_fixed).all(dim=1) | near_of(V, A, ix_f)).all())
    assert bool(((row["recall_jitter"].cpu() == ref_jit).all(dim=1) | near_of(Vj, Aj, ix_j)).all())
    # triplet accuracy per resample: identical unless one of its pairs has |gap| < 1e-6 (sign undefined there)
    random.seed(666)                # the sampler consumes `random` only (the recall draws above use torch's generator)
    gaps = O.comparative_score_triplets([V], [A], dur, n_samples=3)["success"][0]
    # the store holds bf16 embeddings, so -- like the reference on bf16 inputs -- the accuracies come back in bf16:
    # the fp32 oracle's mean, rounded to bf16
    acc_got = torch.as_tensor(row["triplet_acc"]).float().cpu()
    acc_ref = torch.as_tensor(ref_acc).float().bfloat16().float()
    if bool((gaps.abs() < 1e-6).any()):
        assert torch.allclose(acc_got, acc_ref, atol=float((gaps.abs() < 1e-6).sum()) / (gaps.numel() / 3) + 2.0 ** -8)
    else:
        assert torch.equal(acc_got, acc_ref)
    assert torch.equal(row["recall_at_10_fixed"], row["recall_fixed"][:, 10, :])


def test_gallery_131k_properties(pb):
    """C5-shaped gallery at 2^17 clips (4 x 4 gradient-matrix blocks: the stream-K / CTA-pair gradient GEMMs and the
    blocked hinge pass exactly as in the 2^20 bench), checked through size-independent properties -- the reference
    cannot run it (N^2 fp32 = 64 GiB) -- plus exact ranks and gradient rows on random subsets against the oracle."""
    from peppa_b200.gallery import GalleryStep
    n = 1 << 17
    g = torch.Generator(device="cuda").manual_seed(666)
    V = torch.nn.functional.normalize(torch.randn(n, 512, generator=g, device="cuda"), dim=1)
    A = torch.nn.functional.normalize(2.0 * V + torch.randn(n, 512, generator=g, device="cuda"), dim=1).bfloat16()
    V = V.bfloat16()
    step = GalleryStep(n, 512, margin=0.2, top_n=10)
    out = step.run(A, V)
    loss, dA, dV, ranks = out["loss"].item(), out["dA"].clone(), out["dV"].clone(), out["ranks"].clone()
    # (1) the normalisation Jacobian makes every gradient row orthogonal to its input row
    for grad, x in ((dA, A), (dV, V)):
        dots = (grad.double() * x.double()).sum(1).abs().max().item()
        assert dots < 1e-4 * grad.double().norm(dim=1).max().item()
    # (2) recall is the histogram of the ranks; monotone in n
    rec = out["recall"].cpu()
    assert rec[0] == 0 and bool((rec[1:] >= rec[:-1]).all())
    for k in (1, 5, 10):
        assert abs(rec[k].item() - (ranks < k).float().mean().item()) < 1e-6
    # (3) the loss is symmetric in the roles of the two modalities (pig/loss.py:41-48 adds both directions)
    swapped = step.run(V, A)
    assert abs(swapped["loss"].item() - loss) < 1e-6 * abs(loss)
    assert rel_err(swapped["dA"], dV) < 1e-5 and rel_err(swapped["dV"], dA) < 1e-5
    # (4) permuting the clips permutes ranks and gradient rows (near-ties may move a rank)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(3)).cuda()
    p = step.run(A[perm].contiguous(), V[perm].contiguous())
    assert abs(p["loss"].item() - loss) < 1e-6 * abs(loss)
    near_all = near_rows_gpu(V, A, block=2048).cuda()                       # rows = audio queries, as in step.run(A, V)
    assert bool(((p["ranks"] == ranks[perm]) | near_all[perm]).all())       # only rows in a 1e-6 near-tie may move
    assert rel_err(p["dA"], dA[perm]) < 1e-5
    # (5) deterministic
    again = step.run(A, V)
    assert torch.equal(again["dA"], dA) and torch.equal(again["dV"], dV) and again["loss"].item() == loss
    # (6) exact ranks of 256 random queries against the oracle formulation, and their gradient rows in fp64
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:256].cuda()
    Af, Vf = A.double(), V.double()
    Ah, Vh = Af / Af.norm(dim=1, keepdim=True), Vf / Vf.norm(dim=1, keepdim=True)
    S = (Ah[rows].float() @ Vh.float().T)
    d = 1 - S
    pos = d[torch.arange(256, device="cuda"), rows].unsqueeze(1)
    want = (d < pos).sum(1)
    near = ((d - pos).abs() <= 1e-6).sum(1) > 1
    assert bool(((ranks[rows].long() == want) | near).all())
    diag = (Ah * Vh).sum(1)
    S64 = Ah[rows] @ Vh.T                                                    # [256, n] rows of the score matrix
    ir = (0.2 + S64 - diag[rows].unsqueeze(1)) >= 0                          # row hinge active
    ic = (0.2 + S64 - diag.unsqueeze(0)) >= 0                                # column hinge active
    G = ir.double() + ic.double()
    G[torch.arange(256, device="cuda"), rows] = 0
    # diagonal term: -(row count + column count); the column count of a sampled clip needs its whole column
    col_cnt = ((0.2 + (Ah @ Vh[rows].T) - diag[rows].unsqueeze(0)) >= 0).sum(0) - 1
    row_cnt = ir.sum(1) - 1
    gA = G @ Vh - (row_cnt + col_cnt).double().unsqueeze(1) * Vh[rows]
    gA = (gA - Ah[rows] * (gA * Ah[rows]).sum(1, keepdim=True)) / Af[rows].norm(dim=1, keepdim=True) / float(n) ** 2
    assert rel_err(dA[rows], gA) < TOL


def test_graft_entry_smoke(pb):
    """The driver's smoke(): one small invocation of the hot path on cuda:0 checked against the oracle."""
    import __graft_entry__ as entry
    entry.smoke()


def test_bench_line_contract(pb):
    """bench.py prints exactly one JSON line with the keys the driver reads (a single-GPU workload, seconds)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "retrieval16k", "--steps", "5", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "clocks", "e2e", "gpu_launches"):
        assert key in rec, key
    assert rec["value"] > 0 and rec["gpu_launches"] > 0 and rec["vs_baseline"] is None and "workload" in rec["config"]
    roof = rec["roofline"]
    assert roof["bound"] == "tensor" and 0.3 < roof["frac"] < 1.3 and roof["peak"] > 0 and "frac_of_burst" in roof
    assert rec["e2e"]["h2d_bytes_per_step"] > 0 and rec["e2e"]["d2h_bytes_per_step"] > 0 and rec["e2e"]["value"] < rec["value"]
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert set(rec["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_triplet_loss_cxx_autograd_node_matches_ctypes_path(pb, dtype):
    """The launch-bound training step (BASELINE config 2) goes through csrc/torch_fast.cpp when the inputs need no
    conversion: a C++ autograd node around the SAME pb2_hinge_step / pb2_scale_pair calls.  Loss and gradients are bit
    for bit those of the Python autograd.Function; grad_output is applied in fp32 before the rounding; the node works
    under CUDA graph capture; inputs that need a conversion still take the Python path."""
    from peppa_b200 import _cabi
    from peppa_b200.loss import _HingeFn
    assert _cabi.fast() is not None, "csrc/_pb2_fast.so not built / not loadable"
    V, A = emb(1024, 4.0)
    for scale in (1.0, 65536.0):
        v1, a1 = V.cuda().to(dtype).requires_grad_(True), A.cuda().to(dtype).requires_grad_(True)
        v2, a2 = V.cuda().to(dtype).requires_grad_(True), A.cuda().to(dtype).requires_grad_(True)
        l1 = pb.loss.TripletLoss(0.2)(v1, a1)
        assert "HingeStepFn" in l1.grad_fn.name()                            # the C++ node
        (l1 * scale).backward()
        l2 = _HingeFn.apply(v2, a2, 0.2)
        (l2 * scale).backward()
        assert torch.equal(l1, l2) and torch.equal(v1.grad, v2.grad) and torch.equal(a1.grad, a2.grad)
        assert v1.grad.dtype == dtype and a1.grad.dtype == dtype
    ref_loss, ref_dv, ref_da = O.hinge_loss_and_grads(v1.detach().float().cpu(), a1.detach().float().cpu(), 0.2)
    assert rel_err(l1.detach().cpu(), ref_loss) < TOL
    # conversions needed -> the Python path (zero-padded feature dimension; non-contiguous rows)
    vp, ap = V[:, :500].cuda().to(dtype).requires_grad_(True), A[:, :500].cuda().to(dtype).requires_grad_(True)
    assert "HingeStepFn" not in pb.loss.TripletLoss(0.2)(vp, ap).grad_fn.name()
    # CUDA graph capture of forward + backward through the C++ node
    vs, as_ = V.cuda().to(dtype).requires_grad_(True), A.cuda().to(dtype).requires_grad_(True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            vs.grad = as_.grad = None
            pb.loss.TripletLoss(0.2)(vs, as_).backward()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    vs.grad = as_.grad = None
    with torch.cuda.graph(g):
        lg = pb.loss.TripletLoss(0.2)(vs, as_)
        lg.backward()
    g.replay()
    torch.cuda.synchronize()
    gv, ga = _grad_of(pb, V, A, dtype)
    assert torch.equal(lg, l1) and torch.equal(vs.grad, gv) and torch.equal(as_.grad, ga)


def _grad_of(pb, V, A, dtype):
    v, a = V.cuda().to(dtype).requires_grad_(True), A.cuda().to(dtype).requires_grad_(True)
    pb.loss.TripletLoss(0.2)(v, a).backward()
    return v.grad, a.grad


@pytest.mark.parametrize("dtype,n", [(torch.bfloat16, 1024), (torch.float16, 1000), (torch.bfloat16, 2048), (torch.float32, 1024)])
def test_training_step_operands_written_right_before(pb, dtype, n):
    """Inside pb2_hinge_step the similarity pass fetches its operand tiles BEFORE griddepcontrol.wait (they are the
    caller's rows, complete before hinge_prep got past its own wait; csrc/step.cu, host_util.h: OperandsReadyScope),
    and the backward's scale launch triggers its dependents early.  Chain of 24 steps without a host sync in which
    every step's inputs are written by the kernels right ahead of it -- the previous step's gradients as they leave
    pb2_scale_pair, then rows a torch kernel derives from them -- against the same chain with a device sync around
    every step: the kernels are deterministic, so any operand fetched too early shows as a different bit."""
    V0, A0 = emb(n, 4.0)

    def chain(sync):
        v, a = V0.cuda().to(dtype), A0.cuda().to(dtype)
        out = []
        for it in range(24):
            v, a = v.detach().requires_grad_(True), a.detach().requires_grad_(True)
            if sync:
                torch.cuda.synchronize()
            loss = pb.loss.TripletLoss(0.2)(v, a)
            (loss * float(n * n)).backward()      # gradient entries of order 1 in every dtype (fp16 would flush 1 / N^2)
            if sync:
                torch.cuda.synchronize()
            out.append((loss.detach(), v.grad, a.grad))
            if it % 2 == 0:          # the gradients themselves (directions matter, cosine ignores their scale) ...
                v, a = v.grad, a.grad
            else:                    # ... or rows a torch kernel writes from them right before the next step
                v, a = (V0.cuda() + 0.1 * v.grad.float()).to(dtype), (A0.cuda() + 0.1 * a.grad.float()).to(dtype)
        torch.cuda.synchronize()
        return out

    free, synced = chain(False), chain(True)
    for it, ((l1, dv1, da1), (l2, dv2, da2)) in enumerate(zip(free, synced)):
        assert torch.isfinite(l2), it
        assert torch.equal(l1, l2) and torch.equal(dv1, dv2) and torch.equal(da1, da2), it


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n", [1024, 1000, 2304, 4100])
def test_hinge_forward_backward_pair_equals_one_call_step(pb, dtype, n):
    """pb2_hinge_forward + pb2_hinge_backward (3 + 1 launches: the loss folded beside the gradient products -- in a CTA
    of their grid at n = 1024 / 1000, as a launch of its own when they do not fit one grid --, grad_output applied in
    the Jacobian kernel before the rounding) against pb2_hinge_step + pb2_scale_pair (4 + 1 launches, fp32 gradients in
    between): loss and gradients bit for bit, for grad_output 1 and an AMP loss scale; the backward may run twice."""
    from peppa_b200 import ops
    V, A = emb(n, 4.0)
    vb, ab = ops.as_row_pair(V.cuda().to(dtype), A.cuda().to(dtype))
    loss1, grads = ops.hinge_step(vb, ab, 0.2)
    loss2, state = ops.hinge_forward(vb, ab, 0.2)
    assert torch.equal(loss1, loss2) and torch.isfinite(loss1)
    ref_loss, _, _ = O.hinge_loss_and_grads(vb.float().cpu(), ab.float().cpu(), 0.2)
    assert rel_err(loss2.cpu(), ref_loss) < TOL
    for scale in (1.0, 65536.0, 1.0):
        go = torch.tensor(scale, dtype=torch.float32, device="cuda")
        g0, g1 = ops.scale_pair(grads[0], grads[1], go, dtype)
        h0, h1 = ops.hinge_backward(state, vb, ab, go, dtype)
        assert torch.equal(g0, h0) and torch.equal(g1, h1)
        assert h0.dtype == dtype and float(h0.float().abs().sum()) > 0


@pytest.mark.parametrize("dtype,n", [(torch.bfloat16, 1024), (torch.float16, 1000), (torch.float32, 300), (torch.bfloat16, 4100)])
def test_loss_alone_equals_the_training_loss(pb, dtype, n):
    """A forward no backward will follow (validation: torch.no_grad, or inputs without requires_grad) takes
    pb2_hinge_forward without a state buffer -- prep, the hinge pass without a gradient matrix, the fold: three
    launches, and the same bits as the loss of the training step (and the oracle's value to 1e-3)."""
    V, A = emb(n, 4.0)
    v, a = V.cuda().to(dtype), A.cuda().to(dtype)
    mod = pb.loss.TripletLoss(0.2)
    with torch.no_grad():
        l0 = mod(v, a)
    l1 = mod(v, a)                                                    # no requires_grad
    l2 = mod(v.clone().requires_grad_(True), a.clone().requires_grad_(True))
    assert not l0.requires_grad and l2.requires_grad
    assert torch.equal(l0, l1) and torch.equal(l0, l2.detach())
    ref_loss, _, _ = O.hinge_loss_and_grads(v.float().cpu(), a.float().cpu(), 0.2)
    assert rel_err(l0.cpu(), ref_loss) < TOL
    z = v.clone()
    z[3] = 0                                                          # a zero row: NaN like the reference's 0 / 0
    with torch.no_grad():
        assert torch.isnan(mod(z, a))
