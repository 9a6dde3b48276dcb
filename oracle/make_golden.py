"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (/root/reference/pig).

Run in the build container only (``python oracle/make_golden.py``); the GPU box has
no /root/reference, so the resulting fixtures are committed.  Inputs are seeded,
bf16-rounded and stored as their 16-bit patterns so every consumer (oracle, CUDA path) sees the
exact same values.  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PEPPA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    # pig.triplet imports moviepy / pytorch_lightning / pig.data at module scope; none of
    # them is used by the scoring functions, so empty stand-ins are enough to import it.
    for name in ("moviepy", "moviepy.editor", "pytorch_lightning", "pig.data"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    import pig.util, pig.loss, pig.metrics, pig.triplet  # noqa
    return pig


def embeddings(n, alpha, d=512, seed=666):
    """V = normalize(randn), A = normalize(alpha*V + randn); bf16-rounded."""
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    return V.bfloat16().float(), A.bfloat16().float()


def triplet_inputs(t, d=512, seed=666, related=False):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(t, d, generator=g)
    p = a + 2.0 * torch.randn(t, d, generator=g) if related else torch.randn(t, d, generator=g)
    n = torch.randn(t, d, generator=g)
    return a.bfloat16().float(), p.bfloat16().float(), n.bfloat16().float()


def bits(x):
    """bf16-representable float32 tensor -> uint16 bit patterns (halves the fixture size)."""
    assert torch.equal(x.bfloat16().float(), x)
    return x.bfloat16().view(torch.int16).numpy().view(np.uint16)


def milnce_k_fixtures(pig):
    """MILNCELoss with K = len(A) / len(V) > 1 audio candidates per video (pig/loss.py:19-25 views the
    logits as [N, N, K]); forward value and autograd gradients of the reference itself."""
    for n, k, alpha in [(16, 3, 4.0), (40, 2, 0.5)]:
        V, _ = embeddings(n, alpha, seed=23)
        g = torch.Generator().manual_seed(29)
        # candidate k of clip i: a noisy copy of the clip's video embedding (un-normalised scale ~3 so that
        # the K paired logits differ visibly)
        A = (3.0 * V.repeat_interleave(k, dim=0) + torch.randn(n * k, V.shape[1], generator=g)).bfloat16().float()
        V = (3.0 * V).bfloat16().float()
        v = V.clone().requires_grad_(True)
        a = A.clone().requires_grad_(True)
        loss = pig.loss.MILNCELoss()(v, a)
        loss.backward()
        np.savez_compressed(os.path.join(OUT, f"milnce_n{n}_k{k}.npz"), V=bits(V), A=bits(A), k=np.int64(k),
                            loss=loss.detach().numpy(), dV=v.grad.numpy(), dA=a.grad.numpy())


def main():
    pig = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed summation order for the fixtures
    milnce_k_fixtures(pig)
    if len(sys.argv) > 1 and sys.argv[1] == "milnce_k":
        return

    # ---- loss + recall fixtures ------------------------------------------------------
    for n, alpha in [(8, 4.0), (8, 0.5), (64, 4.0), (100, 0.5), (257, 4.0)]:
        V, A = embeddings(n, alpha)
        rec = {"V": bits(V), "A": bits(A), "margin": np.float32(0.2)}
        for name, mod in (("hinge", pig.loss.TripletLoss(0.2)), ("milnce", pig.loss.MILNCELoss())):
            v = V.clone().requires_grad_(True)
            a = A.clone().requires_grad_(True)
            loss = mod(v, a)
            loss.backward()
            rec[f"{name}_loss"] = loss.detach().numpy()
            rec[f"{name}_dV"] = v.grad.numpy()
            rec[f"{name}_dA"] = a.grad.numpy()
        M = pig.util.cosine_matrix(V, A)
        rec["cosine_VA"] = M.numpy()
        rec["contrastive_M"] = pig.loss.contrastive(M, margin=0.2).numpy()
        eye = torch.eye(n)
        for k in (1, 5, 10):
            rec[f"recall_at_{k}"] = pig.metrics.recall_at_n(V, A, eye, n=k).numpy()
        rec["recall_at_1_to_10"] = pig.metrics.recall_at_1_to_n(V, A, eye, N=10).numpy()
        # general multi-target ``correct`` (every row has its own id plus (j*7+3) % n)
        multi = torch.eye(n)
        multi[torch.arange(n), (torch.arange(n) * 7 + 3) % n] = 1
        rec["correct_multi"] = multi.numpy()
        rec["recall_multi_at_5"] = pig.metrics.recall_at_n(V, A, multi, n=5).numpy()
        rec["recall_multi_1_to_10"] = pig.metrics.recall_at_1_to_n(V, A, multi, N=10).numpy()
        np.savez_compressed(os.path.join(OUT, f"sim_n{n}_a{alpha}.npz"), **rec)

    # ---- non-square retrieval (more candidates than queries) -----------------------------
    V, A = embeddings(96, 4.0, seed=7)
    Vq, Aq = V, A[:40]
    correct = torch.zeros(40, 96)
    correct[torch.arange(40), torch.arange(40)] = 1
    np.savez_compressed(os.path.join(OUT, "sim_rect_40x96.npz"), V=bits(Vq), A=bits(Aq),
                        correct=correct.numpy(),
                        recall_at_3=pig.metrics.recall_at_n(Vq, Aq, correct, n=3).numpy(),
                        recall_at_1_to_10=pig.metrics.recall_at_1_to_n(Vq, Aq, correct, N=10).numpy(),
                        cosine=pig.util.cosine_matrix(Aq, Vq).numpy())

    # ---- triplet fixtures ----------------------------------------------------------------
    for t, related in [(8, False), (8, True), (1194, True), (600, False)]:
        a, p, n_ = triplet_inputs(t, related=related)
        if t >= 600:  # edge cases the reference resolves to 0.5: exact ties and zero vectors
            p[3] = n_[3]
            a[5] = 0
            n_[9] = 0
            p[11] = 0
            n_[11] = 0
        np.savez_compressed(
            os.path.join(OUT, f"triplet_t{t}_{'rel' if related else 'rnd'}.npz"),
            anchor=bits(a), positive=bits(p), negative=bits(n_),
            discrete=pig.metrics.triplet_accuracy(a, p, n_).numpy(),
            gap=pig.metrics.triplet_accuracy(a, p, n_, discrete=False).numpy())

    # ---- resampled recall (torch global RNG, seed 666 as pig/evaluation.py:19) ------------
    V, A = embeddings(300, 4.0, seed=11)
    torch.manual_seed(666)
    rr = pig.metrics.resampled_recall(V, A, size=100, n_samples=6, n=10)
    torch.manual_seed(666)
    rr1n = pig.metrics.resampled_recall_at_1_to_n(V, A, size=100, n_samples=4, N=10)
    torch.manual_seed(666)
    ix = torch.stack([pig.metrics.sample_indices(V, 100) for _ in range(6)])
    np.savez_compressed(os.path.join(OUT, "resampled_g300.npz"), V=bits(V), A=bits(A),
                        resampled_recall_n10=rr.numpy(), resampled_1_to_10=rr1n.numpy(),
                        sample_indices=ix.numpy())

    # ---- duration-matched triplet sampler + comparative scores (random seed 666) ----------
    V, A = embeddings(240, 4.0, seed=13)
    V2, A2 = embeddings(240, 1.0, seed=17)
    g = torch.Generator().manual_seed(5)
    duration = torch.randint(20, 60, (240,), generator=g).float() / 10.0   # ~40 distinct values
    random.seed(666)
    draws = []
    for _ in range(5):
        pos, neg = zip(*pig.triplet._triplets(range(len(duration)), lambda i: duration[i]))
        draws.append(np.stack([np.array(pos), np.array(neg)]))
    random.seed(666)
    comp = pig.triplet.comparative_score_triplets([V, V2], [A, A2], duration, n_samples=5)
    # score_triplets raises NameError at reference HEAD (pig/triplet.py:93); record that fact and
    # the intended result (line 93 deleted) computed from the reference's own pieces.
    random.seed(666)
    try:
        pig.triplet.score_triplets(V, A, duration, n_samples=5)
        head_error = ""
    except NameError as e:  # expected
        head_error = repr(e)
    random.seed(666)
    acc, length = [], []
    for _ in range(5):
        pos, neg = zip(*pig.triplet._triplets(range(len(duration)), lambda i: duration[i]))
        pos, neg = torch.tensor(pos), torch.tensor(neg)
        acc.append(pig.metrics.triplet_accuracy(anchor=A[pos], positive=V[pos], negative=V[neg]).mean().item())
        length.append(duration[pos])
    np.savez_compressed(
        os.path.join(OUT, "triplet_sampler_g240.npz"), V=bits(V), A=bits(A), V2=bits(V2), A2=bits(A2),
        duration=duration.numpy(), draws=np.stack(draws),
        comp_success0=comp["success"][0].numpy(), comp_success1=comp["success"][1].numpy(),
        comp_duration=comp["duration"].numpy(), score_accuracy=np.array(acc, dtype=np.float32),
        score_duration=torch.cat(length).numpy(), head_error=np.array(head_error))
    print("golden fixtures written to", os.path.normpath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}  {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
