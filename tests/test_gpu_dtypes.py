"""GPU parity on inputs that are NOT bf16-representable: the reference's callers hand over fp32 tensors and, under
Lightning's ``precision: 16`` (hparams_base.yaml:45, pig/evaluation.py:70), fp16 ones.  Nothing may be rounded to
bf16 on the way in: fp16 rows feed ``tcgen05.mma.kind::f16`` natively, fp32 rows as their split-bf16 pair
(``pb2_split_f16``: the split-fp16 pair of the normalised rows, contraction length 3 D), and every row-wise kernel reads the true values.

Bars (BASELINE.json north_star): ranks identical to the oracle (fp32 math on the exact input values) except rows with
another candidate within 1e-6 of the positive; loss and gradients within 1e-3 -- norm-wise AND per row.
"""
import pytest
import torch

from conftest import rel_err, row_rel_err
from oracle import blockwise as B
from oracle import pig_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def generic(n, alpha, dtype, d=512, seed=666):
    """Unit-norm rows with full fp32 (or fp16) significands: NOT representable in bf16."""
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    V, A = V.to(dtype), A.to(dtype)
    assert not torch.equal(V.float(), V.bfloat16().float())
    return V, A


@pytest.fixture(scope="module")
def pb():
    import peppa_b200.loss as loss
    import peppa_b200.metrics as metrics
    import peppa_b200.util as util
    return type("PB", (), dict(loss=loss, metrics=metrics, util=util))


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("n,alpha", [(4096, 4.0), (4096, 0.5), (1000, 4.0)])
def test_ranks_on_unrounded_inputs(pb, dtype, n, alpha):
    """Round 1 rounded these inputs to bf16 and moved 164 of 4096 ranks; now every rank outside a 1e-6 near-tie
    equals the oracle's."""
    V, A = generic(n, alpha, dtype)
    ranks, near = O.ranks_identity(V.float(), A.float())
    got = pb.metrics.recall_at_1_to_n(V.cuda(), A.cuda(), None, N=10)
    for k in range(1, 11):
        bad = ((ranks < k).float() != got[k]) & ~near
        assert not bool(bad.any()), (k, int(bad.sum()))
    r = pb.metrics._pair_ranks(V.cuda(), A.cuda(), None)[0].cpu().long()
    assert bool(((r == ranks) | near).all())
    # the control: the same inputs rounded to bf16 first DO differ outside the near-ties (what silent rounding cost)
    if n == 4096 and alpha == 4.0:
        rb = pb.metrics._pair_ranks(V.bfloat16().cuda(), A.bfloat16().cuda(), None)[0].cpu().long()
        assert int(((rb != ranks) & ~near).sum()) > 20


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("n", [1024, 1000])
def test_losses_on_unrounded_inputs(pb, dtype, n):
    V, A = generic(n, 4.0, dtype)
    rl, rdv, rda = O.hinge_loss_and_grads(V.float(), A.float(), 0.2)
    ml, mdv, mda = O.milnce_loss_and_grads(V.float(), A.float())
    for mod, (l0, dv0, da0) in ((pb.loss.TripletLoss(0.2), (rl, rdv, rda)), (pb.loss.MILNCELoss(), (ml, mdv, mda))):
        v = V.cuda().requires_grad_(True)
        a = A.cuda().requires_grad_(True)
        out = mod(v, a)
        out.backward()
        assert rel_err(out.float().cpu(), l0) < TOL
        # an fp16 gradient tensor carries its own 2^-11 rounding per element; fp32 results meet the bar row by row
        tol = TOL if dtype == torch.float32 else 2e-3
        assert rel_err(v.grad.float().cpu(), dv0) < tol and rel_err(a.grad.float().cpu(), da0) < tol
        if dtype == torch.float32:      # fp16 gradients of ~1e-5 are fp16 SUBNORMALS (absolute step 6e-8): the row-wise
            # bar is checked on fp32 results here and on loss-scaled fp16 ones in test_fp16_gradients_survive_a_gradscaler
            if isinstance(mod, pb.loss.TripletLoss):     # hinge: entries within 1e-6 of a kink may fall on either side
                k = B.hinge_kink_counts(V, A, 0.2)
                assert B.hinge_rows_within(v.grad.cpu(), dv0, V, k, tol)[0] and B.hinge_rows_within(a.grad.cpu(), da0, A, k, tol)[0]
            else:
                assert row_rel_err(v.grad.cpu(), dv0) < tol and row_rel_err(a.grad.cpu(), da0) < tol
        assert v.grad.dtype == dtype


def test_blocked_loss_on_fp32_inputs(pb):
    """The block-walking path (sim_hinge / sim_lse_* on the split operands, sliced by rows) on fp32 rows."""
    old = pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS
    pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS = 512, 0
    try:
        V, A = generic(1200, 4.0, torch.float32)
        for mod, ref in ((pb.loss.TripletLoss(0.2), O.hinge_loss_and_grads(V, A, 0.2)),
                         (pb.loss.MILNCELoss(), O.milnce_loss_and_grads(V, A))):
            v, a = V.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
            out = mod(v, a)
            out.backward()
            assert rel_err(out.cpu(), ref[0]) < TOL
            if isinstance(mod, pb.loss.TripletLoss):
                k = B.hinge_kink_counts(V, A, 0.2)
                assert B.hinge_rows_within(v.grad.cpu(), ref[1], V, k, TOL)[0] and B.hinge_rows_within(a.grad.cpu(), ref[2], A, k, TOL)[0]
            else:
                assert row_rel_err(v.grad.cpu(), ref[1]) < TOL and row_rel_err(a.grad.cpu(), ref[2]) < TOL
    finally:
        pb.loss._MAX_BLOCK, pb.loss._LSE_BOTH_MIN_PAIRS = old


def test_cosine_matrix_fp32_is_not_rounded(pb):
    U, W = generic(300, 1.0, torch.float32, seed=5)
    M = pb.util.cosine_matrix(U.cuda(), W[:257].cuda()).cpu()
    ref = O.cosine_matrix(U, W[:257])
    assert (M - ref).abs().max() < 1e-6                      # split-bf16: ~2e-7; a bf16 rounding would be ~1e-4
    assert (pb.util.cosine_matrix(U.bfloat16().cuda(), W[:257].bfloat16().cuda()).float().cpu() - ref).abs().max() > 2e-5
    Uh, Wh = generic(300, 1.0, torch.float16, seed=5)
    Mh = pb.util.cosine_matrix(Uh.cuda(), Wh.cuda())
    assert Mh.dtype == torch.float16                          # dtype follows the input like the reference
    assert (Mh.float().cpu() - O.cosine_matrix(Uh.float(), Wh.float())).abs().max() < 1e-3


def test_mixed_dtypes_and_size_mismatch(pb):
    """bf16 video against fp16 audio meet in fp32 (exact for both); unequal embedding sizes raise like the
    reference's matmul instead of being zero-padded to a common width."""
    V, A = generic(512, 4.0, torch.float16)
    Vb = V.float().bfloat16()
    ranks, near = O.ranks_identity(Vb.float(), A.float())
    got = pb.metrics.recall_at_n(Vb.cuda(), A.cuda(), None, n=5)
    assert bool((((ranks < 5).float() == got) | near).all())
    v, a = Vb.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    pb.loss.TripletLoss(0.2)(v, a).backward()
    rl, rdv, rda = O.hinge_loss_and_grads(Vb.float(), A.float(), 0.2)
    assert v.grad.dtype == torch.bfloat16 and a.grad.dtype == torch.float16
    assert rel_err(a.grad.float().cpu(), rda) < 2e-3
    x100, x120 = torch.randn(64, 100).cuda(), torch.randn(64, 120).cuda()      # both would pad to 128 columns
    for call in (lambda: pb.metrics.recall_at_n(x100, x120, None), lambda: pb.loss.TripletLoss(0.2)(x100, x120),
                 lambda: pb.loss.MILNCELoss()(x100, x120), lambda: pb.util.cosine_matrix(x100, x120)):
        with pytest.raises(RuntimeError, match="cannot be multiplied"):
            call()


def test_fp16_gradients_survive_a_gradscaler(pb):
    """AMP (the reference trains with precision 16): autograd applies the loss scale BEFORE the cast to fp16.  The
    fused step keeps its gradients in fp32 until grad_output is applied, so fp16 gradients of ~1e-7 (1/N^2 is folded
    in) times 65536 keep their bits instead of flushing to zero / subnormals."""
    n = 1024
    V, A = generic(n, 4.0, torch.float16)
    v, a = V.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    (pb.loss.TripletLoss(0.2)(v, a) * 65536.0).backward()
    _, rdv, rda = O.hinge_loss_and_grads(V.float(), A.float(), 0.2)
    for got, ref in ((v.grad, rdv * 65536.0), (a.grad, rda * 65536.0)):
        got = got.float().cpu()
        assert got.abs().max() > 1e-4                                         # scaled into fp16's normal range
        big = ref.abs() > 0.05 * ref.abs().max()
        assert ((got - ref).abs()[big] / ref.abs()[big]).max() < 4e-3         # element-wise: fp16 rounding + 1e-3
        assert rel_err(got, ref) < 2e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_one_process(pb):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: the opt-in is tracked per (kernel, device),
    so cuda:1 after cuda:0 in ONE process launches (torchrun's one process per GPU never exercised this)."""
    V, A = generic(1024, 4.0, torch.float32)
    ref = O.hinge_loss_and_grads(V, A, 0.2)
    ranks, near = O.ranks_identity(V, A)
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        v, a = V.to(dev).requires_grad_(True), A.to(dev).requires_grad_(True)
        out = pb.loss.TripletLoss(0.2)(v, a)
        out.backward()
        assert out.device == torch.device(dev) and rel_err(out.cpu(), ref[0]) < TOL and rel_err(v.grad.cpu(), ref[1]) < TOL
        got = pb.metrics.recall_at_n(V.bfloat16().to(dev), A.bfloat16().to(dev), None, n=10)
        assert got.shape == (1024,)
        m = pb.loss.MILNCELoss()(V.to(dev), A.to(dev))
        assert torch.isfinite(m)


def test_step_workspace_is_per_stream(pb):
    """Two TripletLoss steps of one shape on two streams must not share scratch (the workspace cache is keyed by
    stream): run them concurrently and compare with the serial results."""
    V, A = generic(2048, 4.0, torch.float32)
    V2, A2 = generic(2048, 2.0, torch.float32, seed=9)
    def run(v_, a_):
        v, a = v_.cuda().requires_grad_(True), a_.cuda().requires_grad_(True)
        out = pb.loss.TripletLoss(0.2)(v, a)
        out.backward()
        return out.detach(), v.grad
    l1, g1 = run(V, A)
    l2, g2 = run(V2, A2)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            m1, h1 = run(V, A)
        with torch.cuda.stream(s2):
            m2, h2 = run(V2, A2)
    torch.cuda.synchronize()
    assert torch.equal(m1, l1) and torch.equal(h1, g1) and torch.equal(m2, l2) and torch.equal(h2, g2)
