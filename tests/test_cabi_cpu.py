"""CPU: the C-ABI library builds, loads and exports every symbol include/peppa_b200.h declares;
host-side argument checks that need no device; the product path fails loudly without CUDA."""
import ctypes as C

import pytest
import torch

import __graft_entry__ as entry
from peppa_b200 import _cabi


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _cabi.lib()


def test_every_declared_symbol_is_exported(lib):
    declared = _cabi.header_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_cabi.SIGNATURES), "ctypes table and header disagree"
    assert lib.pb2_version() == 2
    # the product library carries no debug switch (mutable global state); they live in the measurement build
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "pb2_sim_hinge" in syms and "pb2_debug" not in syms
    with _cabi.measurement_library() as mlib:
        for name in _cabi._DEBUG:
            assert hasattr(mlib, name), name
    assert _cabi.lib() is lib


def test_sass_has_no_hot_loop_fences_and_no_serialised_loader_loads(lib):
    """Two properties of the compiled kernels that cost 5-35 % when they broke and that no numerical test sees, read from
    the SASS of the product library (cuobjdump ships with the toolkit):
    (1) the similarity kernels' vector-loader warp issues every global load of a tile before the first use, i.e. no
        load destination register is written twice (a reused register serialises the round trips);
    (2) a CTA-pair kernel contains GPU-scope fences only for its two cluster barriers and its mbarrier initialisation,
        not in the accumulator hand-back of the tile loop (`mbarrier.arrive.release.cluster` = MEMBAR.ALL.GPU)."""
    import re
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    fn, loads, fences, n_sim = None, {}, {}, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            n_sim += "sim_kernel" in fn
            continue
        if fn is None:
            continue
        if "sim_kernel" in fn and " LDG." in line:
            dst = re.search(r"LDG\.\S+\s+(R\d+),", line).group(1)
            loads.setdefault(fn, []).append(dst)
        if "MEMBAR.ALL.GPU" in line and ("sim_kernel" in fn or "project_normalize" in fn):
            fences[fn] = fences.get(fn, 0) + 1
    assert n_sim >= 20 and loads, "no similarity kernels found in the SASS"
    for f, dst in loads.items():
        assert len(dst) == len(set(dst)), f"{f}: loader loads share a register: {dst}"
    for f, n in fences.items():
        assert n <= 2, f"{f}: {n} GPU-scope fences (expected the two cluster barriers only)"


def test_argument_errors_are_reported_without_touching_a_device(lib):
    null = C.c_void_p(0)
    assert lib.pb2_triplet_score(null, null, null, null, null, null, 4, 512, 512, 0, 1, null, null) == 1
    assert b"null" in lib.pb2_last_error()
    assert lib.pb2_triplet_score(null, null, null, null, null, null, 0, 512, 512, 0, 1, null, null) == 0   # empty ok
    assert lib.pb2_triplet_score(null, null, null, null, null, null, 4, 512, 512, 9, 1, null, null) == 1
    assert lib.pb2_sim_lse_parts(1000) == 16
    assert lib.pb2_grad_gemm(null, 1, 8, 8, 64, 0, null, 0, 512, 512, 1.0, 0, null, 512, null) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from peppa_b200 import loss, metrics
    x = torch.randn(8, 512)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss.TripletLoss(0.2)(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        metrics.recall_at_n(x, x, torch.eye(8))
    with pytest.raises(RuntimeError, match="CUDA"):
        metrics.triplet_accuracy(x, x, x)


def test_sampler_matches_golden_without_device():
    """The Python-side sampler of peppa_b200.triplet consumes `random` like pig/triplet.py:99-104."""
    import random

    import numpy as np

    from conftest import load_golden
    from peppa_b200 import triplet
    g = load_golden("triplet_sampler_g240.npz")
    dur = g["duration"]
    random.seed(666)
    for k in range(5):
        pos, neg = zip(*triplet._triplets(range(len(dur)), lambda i: dur[i]))
        assert np.array_equal(np.array(pos), g["draws"][k, 0].numpy())
        assert np.array_equal(np.array(neg), g["draws"][k, 1].numpy())
    assert triplet.pairs([1, 2, 3, 4, 5]) == [[1, 2], [3, 4]] and triplet.pairs([1]) == []
    # the batched sampler behind score_triplets / comparative_score_triplets (grouping hoisted out of the sample
    # loop, plain-number keys) draws the same triplets from the same seed
    random.seed(666)
    for k in range(5):
        list(triplet._triplets(range(len(dur)), lambda i: dur[i]))
    state = random.getstate()
    assert triplet._sample2_is_two_randbelow()      # CPython's sample(pair, 2) is the two rejection draws it unrolls
    for fast in (True, False):                      # ... and the plain random.sample path it falls back to otherwise
        triplet._FAST_SAMPLE2 = fast
        random.seed(666)
        pos, neg = triplet._sampled_index_pairs(dur, 5)
        assert pos.dtype == torch.int64 and tuple(pos.shape) == tuple(g["draws"][:, 0].shape)
        assert torch.equal(pos, g["draws"][:, 0].to(torch.int64)) and torch.equal(neg, g["draws"][:, 1].to(torch.int64))
        assert random.getstate() == state           # leaves the generator where the reference's loop leaves it
    triplet._FAST_SAMPLE2 = None
    # the loop above ran on the C++ replay of the draws (csrc/sampler.cu) for fast = True: it must have been active,
    # and it must equal the Python loops -- triplets and final generator state -- also on awkward group structures
    # (singletons, odd groups, a generator whose state index sits mid-block, one group of 1500 clips = the bucket sort's
    # long buckets, one of 2600 = the (key, position) pair sort) and over a long run
    assert triplet._native_sampler() is True
    gen = torch.Generator().manual_seed(7)
    for n_items, n_vals, n_samples, burn in ((2, 1, 3, 0), (9, 4, 7, 1), (200, 11, 40, 617), (1467, 40, 25, 5), (1500, 1, 3, 0), (2600, 1, 2, 3)):
        d = torch.randint(0, n_vals, (n_items,), generator=gen).float()
        if n_items == 9:
            d[0] = 99.0                                  # a clip alone in its group still costs one random() per sample
        out = []
        for native in (True, False):
            triplet._NATIVE_SAMPLER = native
            random.seed(4242)
            for _ in range(burn):
                random.random()
            out.append(triplet._sampled_index_pairs(d, n_samples) + (random.getstate(),))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]
        assert out[0][0].shape[0] == n_samples and out[0][0].numel() > 0
    triplet._NATIVE_SAMPLER = None
    null = C.c_void_p(0)
    lib = _cabi.lib()
    assert lib.pb2_host_sample_pairs(null, null, null, 3, 2, null, null) == 1 and lib.pb2_host_sample_pairs(null, null, null, 0, 2, null, null) == 0
    bad = (C.c_uint32 * 625)()
    bad[624] = 700
    assert lib.pb2_host_random_doubles(bad, 1, (C.c_double * 1)()) == 1        # state index out of range


def test_encoder_tail_module_is_state_dict_compatible_with_linear():
    """ProjectNormalize keeps nn.Linear's parameters (pig/models.py:96-98 `self.project = nn.Linear(...)`), so a
    checkpoint's `project.*` tensors load unchanged; without CUDA the forward raises (no CPU fallback)."""
    from peppa_b200 import encoder
    lin = torch.nn.Linear(28, 512)
    mod = encoder.ProjectNormalize.from_linear(lin)
    assert set(mod.state_dict()) == set(lin.state_dict()) == {"weight", "bias"}
    assert torch.equal(mod.weight, lin.weight) and torch.equal(mod.bias, lin.bias)
    nb = encoder.ProjectNormalize(512, 256, bias=False)
    assert nb.bias is None and nb.weight.shape == (256, 512)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            mod(torch.randn(4, 28))
    with pytest.raises(RuntimeError, match="multiple of 64"):
        encoder.project_normalize(torch.randn(4, 64), torch.randn(100, 64))


def test_new_entry_points_reject_bad_arguments(lib):
    null = C.c_void_p(0)
    assert lib.pb2_project_normalize(null, null, null, 8, 512, 512, 512, 512, 1e-12, null, 512, null, null, null) == 1
    assert lib.pb2_project_normalize(null, null, null, 0, 512, 512, 512, 512, 1e-12, null, 512, null, null, null) == 0
    assert lib.pb2_grad_gemm_dual(null, 1, 8, 8, 64, null, null, 1, 512, 512, 512, 1.0, null, null, 512, 512, null) == 1
    assert lib.pb2_milnce_finish_k(null, 512, null, 0, null, 8, 2, 1, 512, 512, 1.0, null, null, 512, null) == 1
    assert lib.pb2_grad_gemm_workspace() >= 148 * 128 * 512 * 4
    # one-pass log-sum-exp: null partials and a bound whose shift would underflow fp32 are refused before any launch
    one = C.c_void_p(256)
    assert lib.pb2_sim_lse_both(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, 1.0, null, null, null) == 1
    assert lib.pb2_sim_lse_both(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, 100.0, one, one, null) == 1
    assert lib.pb2_sim_lse_both(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, -1.0, one, one, null) == 1
    assert lib.pb2_sim_lse_both(one, one, null, null, 0, 8, 512, 0, 512, 512, 1.0, 1.0, null, null, null) == 0
    # the same pass with the rank counts fused in: the threshold and count vectors are required, same bound rule
    assert lib.pb2_sim_lse_both_rank(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, 1.0, one, one, null, null, null, 0, 0, one, null) == 1
    assert lib.pb2_sim_lse_both_rank(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, 1.0, one, one, null, null, one, 0, 0, null, null) == 1
    assert lib.pb2_sim_lse_both_rank(one, one, null, null, 8, 8, 512, 0, 512, 512, 1.0, 100.0, one, one, null, null, one, 0, 0, one, null) == 1
    assert lib.pb2_sim_lse_both_rank(one, one, null, null, 0, 8, 512, 0, 512, 512, 1.0, 1.0, null, null, null, null, null, 0, 0, null, null) == 0
    # operand dtypes: the tensor-core kernels take bf16 / fp16 (fp32 rows go through pb2_split_f16), the row-wise
    # kernels bf16 / fp16 / fp32; anything else is refused before a launch
    assert lib.pb2_sim_rank(one, one, null, null, one, one, 8, 8, 0, 512, 2, 512, 512, one, null) == 1
    assert b"split_f16" in lib.pb2_last_error()
    assert lib.pb2_row_norms(one, 7, 8, 512, 512, one, one, null) == 1
    assert lib.pb2_split_f16(one, null, 8, 512, 512, 2, one, 1536, one, null) == 1        # side is 0 or 1
    assert lib.pb2_split_f16(one, null, 8, 512, 512, 0, one, 512, one, null) == 1         # ld_out < 3 dim
    assert lib.pb2_split_f16(null, null, 0, 512, 512, 0, null, 1536, null, null) == 0
    # the one-byte gradient matrix pairs with the two-plane operand, and only with it
    assert lib.pb2_grad_gemm(one, 3, 8, 8, 128, 0, one, 1, 512, 1024, 1.0, 0, one, 512, null) == 1
    assert b"PB2_I8_PLANES" in lib.pb2_last_error()
    assert lib.pb2_grad_gemm(one, 3, 8, 8, 128, 0, one, 4, 320, 1024, 1.0, 0, one, 512, null) == 1      # dim % 256
    assert lib.pb2_rows_quant_i8(one, 0, null, 8, 512, 512, one, 512, null) == 1                        # ld_out < 2 dim
    assert lib.pb2_grad_gemm(one, 3, 40000, 8, 128, 1, one, 4, 512, 1024, 1.0, 0, one, 512, null) == 1  # contraction > 32768
    assert b"32768" in lib.pb2_last_error()
    assert lib.pb2_sim_hinge(one, one, one, one, one, one, 8, 8, 0, 0, 512, 0, 512, 512, 0.2, one, 1 << 20, one, one, one, 2,
                             128, null, null, null) == 1                                                # G is fp16 or u8
    # multi-GPU entry points: argument checks need neither a device nor a communicator
    arr = (C.c_void_p * 2)(256, 512)
    assert lib.pb2_peer_reduce(arr, 0, 64, one, null) == 1 and lib.pb2_peer_reduce(arr, 17, 64, one, null) == 1
    assert lib.pb2_peer_reduce(arr, 2, 63, one, null) == 1                                              # whole 16-byte vectors
    assert lib.pb2_peer_reduce(arr, 2, 0, one, null) == 0
    assert lib.pb2_ipc_export(null, null, null) == 1 and lib.pb2_ipc_open(null, 0, null, null) == 1 and lib.pb2_ipc_close(null) == 0
    assert lib.pb2_nccl_gallery_allgather(null, one, 8, 1024, one, null) == 1
    assert lib.pb2_nccl_gallery_allgather(null, one, 0, 1024, one, null) == 0
    assert lib.pb2_nccl_colstat_merge(null, one, 8, one, one, 11, null) == 1
    assert lib.pb2_nccl_dv_reduce_scatter(null, one, 8, 512, one, null) == 1
    assert lib.pb2_nccl_available() in (0, 1)
    assert lib.pb2_hinge_step_workspace(1024, 512, 2) > lib.pb2_hinge_step_workspace(1024, 512, 0) > 0
    assert lib.pb2_sim_lse_col_parts(1000) == 32 and lib.pb2_sim_lse_col_parts(128) == 4
    assert lib.pb2_lse_merge_const(null, 4, 8, 1.0, null, 0, null) == 1 and lib.pb2_lse_merge_const(null, 4, 0, 1.0, null, 0, null) == 0
    assert lib.pb2_scale_pair(null, null, 64, 0, null, null, null, null) == 1
    assert lib.pb2_scale_pair(one, one, 63, 2, one, one, one, null) == 1        # not whole 16-byte vectors
    assert lib.pb2_scale_pair(null, null, 0, 0, null, null, null, null) == 0
    # the training step as an autograd pair: sizes, nulls, alignment and too-small buffers are refused before any launch
    ws, st = lib.pb2_hinge_forward_workspace(1024, 512, 0), lib.pb2_hinge_state_bytes(1024, 512)
    assert 0 < ws and st >= 2 * 1024 * 512 * 4 and lib.pb2_hinge_step_workspace(1024, 512, 0) == ws + st
    assert lib.pb2_hinge_state_bytes(0, 512) == 0
    assert lib.pb2_hinge_forward(one, one, 0, 1024, 512, 512, 512, 0.2, null, ws, one, st, one, null, null, null) == 1   # null workspace
    assert lib.pb2_hinge_forward(one, one, 0, 1024, 512, 512, 512, 0.2, one, ws - 1, one, st, one, null, null, null) == 1
    assert b"workspace too small" in lib.pb2_last_error()
    assert lib.pb2_hinge_forward(one, one, 0, 1024, 512, 512, 512, 0.2, one, ws, one, st - 1, one, null, null, null) == 1
    assert b"state too small" in lib.pb2_last_error()
    assert lib.pb2_hinge_forward(one, one, 0, 1024, 512, 512, 512, 0.2, one, ws, C.c_void_p(264), st, one, null, null, null) == 1  # alignment
    assert lib.pb2_hinge_forward(one, one, 7, 1024, 512, 512, 512, 0.2, one, ws, one, st, one, null, null, null) == 1    # dtype
    assert lib.pb2_hinge_forward(one, one, 0, 0, 512, 512, 512, 0.2, one, ws, one, st, one, null, null, null) == 1       # empty batch
    assert lib.pb2_hinge_backward(null, st, one, one, 0, 1024, 512, 512, 512, null, one, one, 0, null) == 1
    assert lib.pb2_hinge_backward(one, st - 1, one, one, 0, 1024, 512, 512, 512, null, one, one, 0, null) == 1
    assert lib.pb2_hinge_backward(one, st, one, one, 0, 1024, 500, 512, 512, null, one, one, 0, null) == 1              # dim % 64


def test_triplet_scorer_keeps_the_reference_interface(monkeypatch):
    """TripletScorer (pig/triplet.py:31-61): constructor arguments go to the reference's own dataset class, _encode
    concatenates what trainer.predict yields; pig.data and Lightning are imported only on use."""
    import sys
    import types

    from peppa_b200 import triplet
    seen = {}

    class FakeDataset:
        def __init__(self, **kw):
            seen.update(kw)

    data = types.ModuleType("pig.data")
    data.PeppaPigDataset = FakeDataset
    data.collate = object()
    data.grouped_loader = lambda ds, key, collate, batch_size: [("loader", ds, batch_size)]
    pig = types.ModuleType("pig")
    pig.data = data
    monkeypatch.setitem(sys.modules, "pig", pig)
    monkeypatch.setitem(sys.modules, "pig.data", data)
    sc = triplet.TripletScorer("dialog", split=["val"], target_size=(90, 50), scrambled_video=True)
    assert seen == dict(target_size=(90, 50), split=["val"], fragment_type="dialog", duration=None, audio_sample_rate=44100,
                        scrambled_video=True)

    class Batch:
        def __init__(self, k):
            self.audio, self.video = torch.full((2, 4), float(k)), torch.full((2, 4), float(-k))
            self.audio_duration = torch.tensor([1.0 * k, 2.0 * k])

    class Trainer:
        def predict(self, model, loader):
            assert loader[0][0] == "loader" and loader[0][2] == 8
            return [Batch(1), Batch(2)]

    sc._encode(model=None, trainer=Trainer(), batch_size=8)
    assert tuple(sc._audio.shape) == (4, 4) and torch.equal(sc._duration, torch.tensor([1.0, 2.0, 2.0, 4.0]))
    assert torch.equal(sc._video, -sc._audio)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: exactly ONE line on stdout, the JSON record (anything a library prints goes to stderr).
    The reference arm is CPU-only, so the contract is checked here on a small bounded sample."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "train1024",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["unit"] == "pairs/s"
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/pig"), reason="the reference tree exists in the build container only")
def test_install_patches_the_real_pig_package():
    """peppa_b200.install() over the REAL reference package (imported from /root/reference with the same three
    stand-in modules oracle/make_golden.py uses for pig.triplet's moviepy / Lightning / pig.data imports): every
    hot-path name the reference's callers resolve (pig/models.py:14,25,228,262,297-317; pig/evaluation.py:128,159,
    167-193; evaluation_targeted_triplets.py:22,79) now points at the B200 implementation, signatures unchanged."""
    import inspect
    import subprocess
    import sys
    code = r'''
import inspect, sys, types
sys.path.insert(0, "/root/reference")
for name in ("moviepy", "moviepy.editor", "pytorch_lightning", "pig.data"):
    sys.modules.setdefault(name, types.ModuleType(name))
import pig, pig.util, pig.loss, pig.metrics, pig.triplet
ref = {m: {n: inspect.signature(getattr(getattr(pig, m), n)) for n in names} for m, names in {
    "loss": ["contrastive", "cosine_matrix"], "metrics": ["recall_at_n", "recall_at_1_to_n", "triplet_accuracy",
    "batch_triplet_accuracy", "resampled_recall", "resampled_recall_at_1_to_n", "sample_indices"],
    "triplet": ["score_triplets", "comparative_score_triplets", "_triplets", "triplets", "pairs"], "util": ["cosine_matrix"]}.items()}
ref_init = {"TripletLoss": inspect.signature(pig.loss.TripletLoss.__init__), "MILNCELoss.forward": inspect.signature(pig.loss.MILNCELoss.forward)}
import peppa_b200
from peppa_b200 import loss, metrics, triplet, util
peppa_b200.install(pig)
assert pig.loss is loss and pig.metrics is metrics and sys.modules["pig.loss"] is loss and sys.modules["pig.metrics"] is metrics
assert pig.util.cosine_matrix is util.cosine_matrix and pig.triplet.score_triplets is triplet.score_triplets
assert pig.triplet.comparative_score_triplets is triplet.comparative_score_triplets and pig.triplet.triplet_accuracy is metrics.triplet_accuracy
from pig.loss import TripletLoss, MILNCELoss                 # what pig/models.py:14,25 executes
from pig.metrics import recall_at_n, batch_triplet_accuracy  # pig/evaluation.py, evaluation_targeted_triplets.py:22
assert TripletLoss is loss.TripletLoss and recall_at_n is metrics.recall_at_n
for m, sigs in ref.items():
    for n, sig in sigs.items():
        assert inspect.signature(getattr(getattr(pig, m), n)) == sig, (m, n, sig)
assert inspect.signature(loss.TripletLoss.__init__) == ref_init["TripletLoss"]
assert list(inspect.signature(loss.MILNCELoss.forward).parameters) == list(ref_init["MILNCELoss.forward"].parameters)
assert pig.util.grouped is not None and hasattr(pig.util, "pad_audio_batch")      # the rest of pig.util is untouched
print("ok")
'''
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-3000:]
    assert inspect.isfunction(__import__("peppa_b200").install)


def test_cxx_autograd_glue_loads_and_refuses_cpu_tensors():
    """csrc/torch_fast.cpp (the training step's autograd node in C++) is built in-tree, loads without a GPU and -- like
    every other path -- has no CPU fallback: CPU tensors are refused, not computed."""
    import os

    from peppa_b200 import loss
    fast = _cabi.fast()
    assert fast is not None and os.path.exists(_cabi.FAST_PATH)
    with pytest.raises(RuntimeError, match="one CUDA device"):
        fast.triplet_loss(torch.zeros(4, 64), torch.zeros(4, 64), 0.2)
    with pytest.raises(RuntimeError):          # the public module on CPU tensors: the ctypes path's refusal
        loss.TripletLoss(0.2)(torch.zeros(4, 64, requires_grad=True), torch.zeros(4, 64))
