"""CPU: the oracle restatement reproduces the reference's own outputs (tests/golden, written by
oracle/make_golden.py from /root/reference) -- this is what pins parity for the whole repo."""
import random

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden, rel_err
from oracle import pig_oracle as O


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)       # fixtures were written with a fixed summation order
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", golden_files("sim_n"))
def test_loss_and_recall_match_reference(name):
    g = load_golden(name)
    V, A = g["V"], g["A"]
    n = V.shape[0]
    assert torch.equal(O.cosine_matrix(V, A), g["cosine_VA"])
    assert torch.equal(O.contrastive(g["cosine_VA"], 0.2), torch.as_tensor(g["contrastive_M"]))
    for kind, fn in (("hinge", lambda v, a: O.triplet_loss(v, a, 0.2)), ("milnce", O.milnce_loss)):
        v = V.clone().requires_grad_(True)
        a = A.clone().requires_grad_(True)
        loss = fn(v, a)
        loss.backward()
        assert torch.equal(loss.detach(), torch.as_tensor(g[f"{kind}_loss"])), kind
        assert torch.equal(v.grad, g[f"{kind}_dV"]) and torch.equal(a.grad, g[f"{kind}_dA"]), kind
    eye = torch.eye(n)
    for k in (1, 5, 10):
        assert torch.equal(O.recall_at_n(V, A, eye, n=k), g[f"recall_at_{k}"])
    assert torch.equal(O.recall_at_1_to_n(V, A, eye, N=10), g["recall_at_1_to_10"])
    assert torch.equal(O.recall_at_n(V, A, g["correct_multi"], n=5), g["recall_multi_at_5"])
    assert torch.equal(O.recall_at_1_to_n(V, A, g["correct_multi"], N=10), g["recall_multi_1_to_10"])


@pytest.mark.parametrize("name", golden_files("sim_n"))
def test_closed_forms_match_reference(name):
    """SURVEY 8(a') closed forms (what the fused kernels implement) against reference autograd."""
    g = load_golden(name)
    V, A = g["V"], g["A"]
    loss, dV, dA = O.hinge_loss_and_grads(V, A, 0.2)
    assert rel_err(loss, g["hinge_loss"]) < 2e-6
    assert rel_err(dV, g["hinge_dV"]) < 2e-5 and rel_err(dA, g["hinge_dA"]) < 2e-5
    loss, dV, dA = O.milnce_loss_and_grads(V, A)
    assert rel_err(loss, g["milnce_loss"]) < 2e-6
    assert rel_err(dV, g["milnce_dV"]) < 2e-5 and rel_err(dA, g["milnce_dA"]) < 2e-5
    ranks, near = O.ranks_identity(V, A)
    for k in (1, 5, 10):
        got = (ranks < k).float()
        ok = (got == g[f"recall_at_{k}"]) | near
        assert bool(ok.all())
    assert int(near.sum()) <= max(1, V.shape[0] // 50)


def test_rectangular_retrieval():
    g = load_golden("sim_rect_40x96.npz")
    assert torch.equal(O.cosine_matrix(g["A"], g["V"]), g["cosine"])
    assert torch.equal(O.recall_at_n(g["V"], g["A"], g["correct"], n=3), g["recall_at_3"])
    assert torch.equal(O.recall_at_1_to_n(g["V"], g["A"], g["correct"], N=10), g["recall_at_1_to_10"])


@pytest.mark.parametrize("name", golden_files("triplet_t"))
def test_triplet_accuracy_matches_reference(name):
    g = load_golden(name)
    a, p, n = g["anchor"], g["positive"], g["negative"]
    assert torch.equal(O.triplet_accuracy(a, p, n), g["discrete"])
    assert torch.equal(O.triplet_accuracy(a, p, n, discrete=False), g["gap"])
    if a.shape[0] >= 600:   # ties and zero vectors resolve to exactly 0.5
        assert g["discrete"][3] == 0.5 and g["discrete"][5] == 0.5 and g["discrete"][11] == 0.5


def test_resampled_recall_matches_reference():
    g = load_golden("resampled_g300.npz")
    torch.manual_seed(666)
    assert torch.equal(O.resampled_recall(g["V"], g["A"], size=100, n_samples=6, n=10), g["resampled_recall_n10"])
    torch.manual_seed(666)
    assert torch.equal(O.resampled_recall_at_1_to_n(g["V"], g["A"], size=100, n_samples=4, N=10), g["resampled_1_to_10"])
    torch.manual_seed(666)
    ix = torch.stack([O.sample_indices(g["V"], 100) for _ in range(6)])
    assert torch.equal(ix, g["sample_indices"])
    with pytest.raises(AssertionError):
        O.resampled_recall(g["V"][:50], g["A"][:50], size=100)
    with pytest.raises(AssertionError):
        O.resampled_recall(g["V"], g["A"][:-1], size=100)


def test_recall_without_target_raises_like_reference():
    g = load_golden("sim_n8_a4.0.npz")
    with pytest.raises(ZeroDivisionError):
        O.recall_at_n(g["V"], g["A"], torch.zeros(8, 8), n=1)


def test_triplet_sampler_and_scores_match_reference():
    g = load_golden("triplet_sampler_g240.npz")
    dur = g["duration"]
    random.seed(666)
    for k in range(5):
        pos, neg = O.sample_triplet_indices(dur)
        assert np.array_equal(pos.numpy(), g["draws"][k, 0].numpy()) and np.array_equal(neg.numpy(), g["draws"][k, 1].numpy())
    random.seed(666)
    comp = O.comparative_score_triplets([g["V"], g["V2"]], [g["A"], g["A2"]], dur, n_samples=5)
    assert torch.equal(comp["success"][0], g["comp_success0"]) and torch.equal(comp["success"][1], g["comp_success1"])
    assert torch.equal(comp["duration"], g["comp_duration"])
    random.seed(666)
    sc = O.score_triplets(g["V"], g["A"], dur, n_samples=5)
    assert torch.equal(sc["accuracy"], g["score_accuracy"]) and torch.equal(sc["duration"], g["score_duration"])
    assert "NameError" in str(g["head_error"])      # the reference's own score_triplets is broken at HEAD


@pytest.mark.parametrize("name", golden_files("milnce_n"))
def test_milnce_k_candidates_bit_exact(name):
    """MILNCELoss with K > 1 audio candidates per video (pig/loss.py:19-25): oracle == reference, bit for bit."""
    g = load_golden(name)
    v = g["V"].clone().requires_grad_(True)
    a = g["A"].clone().requires_grad_(True)
    torch.set_num_threads(1)
    loss = O.milnce_loss(v, a)
    loss.backward()
    assert a.shape[0] == int(g["k"]) * v.shape[0]
    assert np.array_equal(loss.detach().numpy(), g["loss"])
    assert torch.equal(v.grad, g["dV"]) and torch.equal(a.grad, g["dA"])


@pytest.mark.parametrize("name", ["sim_n257_a4.0.npz", "sim_n100_a0.5.npz"])
def test_blockwise_matches_full_oracle(name):
    """oracle/blockwise.py (the checker at 2^20, where the N x N matrix does not exist) gives the reference's own
    recall vectors / loss / gradient rows at sizes where the reference ran: recall against the golden fixtures
    (reference outputs), loss and gradient rows against the reference's autograd gradients stored there."""
    from oracle import blockwise as B
    if name not in golden_files("sim_n"):
        pytest.skip(f"{name} not among the fixtures")
    g = load_golden(name)
    V, A = g["V"], g["A"]
    n = V.shape[0]
    ranks, near = B.all_ranks(V, A, block=96)                       # ragged blocks on purpose
    r0, n0 = O.ranks_identity(V, A)
    assert torch.equal(ranks, r0) and torch.equal(near, n0)
    for k in (1, 5, 10):
        assert bool((((ranks < k).float() == g[f"recall_at_{k}"]) | near).all())
    assert abs(B.hinge_loss_blockwise(V, A, 0.2, block=96).item() - float(g["hinge_loss"])) < 1e-6 * abs(float(g["hinge_loss"]))
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:64]
    dv = B.hinge_grad_rows(V, A, rows, 0.2)
    da = B.hinge_grad_rows(A, V, rows, 0.2)                         # the loss is symmetric in its two arguments
    scale_v, scale_a = g["hinge_dV"].abs().max(), g["hinge_dA"].abs().max()
    assert (dv - g["hinge_dV"][rows].double()).abs().max() < 1e-5 * scale_v
    assert (da - g["hinge_dA"][rows].double()).abs().max() < 1e-5 * scale_a
