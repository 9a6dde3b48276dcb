#!/usr/bin/env python
"""Benchmark of the contrastive-scoring hot path (BASELINE.json metric: similarity pairs/sec for
loss fwd+bwd + recall@k at 1/2/4/8 B200, as a fraction of bf16 tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (``config.workload``):
  gallery (default)  BASELINE config 5, the only configuration the metric is quoted on at 1/2/4/8 GPUs and
                     one that fits a single B200 (2 GiB of embeddings): a 2^20 x 2^20 audio<->video gallery;
                     one step = symmetric hinge loss forward + gradients (dA, dV) AND recall@1..10 from a
                     single pass over the similarity matrix, rows sharded over the ranks (strong scaling).
  train1024          config 2: TripletLoss fwd+bwd at batch 1024 x 512 bf16 through the public API.
  retrieval16k       config 3: recall_at_1_to_n(N=10) on 16384 x 16384 through the public API.
  triplets1m         config 4: triplet_accuracy on 2^20 triplets x 512 bf16 (HBM bound).
The last three are single-GPU workloads; with the default workload they are also measured (N=1) and
reported under ``other_workloads`` of the same JSON line.

One JSON line on stdout (rank 0).  ``value`` is timed with device-resident inputs; ``e2e`` through the
public API with pinned-host inputs copied in and the result read back every step.  ``roofline`` is the
kernel with the largest summed device time, timed live with CUDA events around each of its launches.
``--impl reference`` times the CPU port of the reference (oracle/pig_oracle.py, torch-CPU, all host
threads) on a bounded sample of the same workload; /root/reference is pure Python and does not exist on
the GPU box, so the oracle port is the reference arm.  The gallery is drawn from ONE seed whatever the world size,
and ``check`` (loss, recall, an order-independent hash of all ranks, gradient checksums, 64 rows verified against
oracle/blockwise.py outside the timed region) is therefore the same at every ``--gpus N``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 512
MARGIN = 0.2
TOP_N = 10
METRIC = {"gallery": "similarity pairs/sec (loss fwd+bwd + recall@1..10)", "train1024": "similarity pairs/sec (loss fwd+bwd)",
          "retrieval16k": "similarity pairs/sec (recall@1..10)", "triplets1m": "triplets/sec (triplet_accuracy)",
          "encoder_tail": "rows/sec (Linear 512x512 + L2-normalise + bf16 + rinv)",
          "milnce64k": "similarity pairs/sec (MIL-NCE loss fwd+bwd, recall@k)",
          "eval1467": "subset pairs/sec (500 x recall@1..10 over 100-clip subsets + 500 triplet resamples)"}


UNIT = {"triplets1m": "triplets/s", "encoder_tail": "rows/s"}


# ----------------------------------------------------------------------------------- helpers
def peaks():
    """Measured roofline denominators (driver-written MEASURED_PEAKS.json) or the guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel):
    """(GB per launch, source) -- dram__bytes_read.sum + dram__bytes_write.sum of one launch of ``kernel``.  ncu
    cannot run inside the timed bench (it replays every kernel ~40 times), so the figure comes from the committed
    ``ncu --set full`` capture of the same kernel on the same shape; the source string names the file and its
    capture date so a stale number is visible as such.  (None, None) if there is no capture."""
    def gb(s):
        v, u = s.split()[:2]
        return float(v) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]

    try:
        with open(os.path.join(ROOT, "profiles", "traffic_index.json")) as f:
            idx = json.load(f)
        e = idx.get(kernel)
        if not e:
            return None, None
        with open(os.path.join(ROOT, "profiles", e["file"])) as f:
            d = json.load(f)
        for key in e["path"]:
            d = next(k for k in d if e["match"] in k["kernel"]) if key == "*" else d[key]
        return gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"]), f"profiles/{e['file']} ({e['captured']}; {e['shape']})"
    except (OSError, KeyError, ValueError, StopIteration, TypeError):
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up (NVML init) contends with the CUDA driver for ~100 ms: let it finish and
            # deliver its first sample BEFORE the timed region starts
            t0 = time.time()
            while not self.rows and time.time() - t0 < 2.0:
                time.sleep(0.02)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        rows = self.rows[1:] if len(self.rows) > 2 else self.rows      # the first sample predates the timed region
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_embeddings(n, seed, device, alpha=4.0):
    """SURVEY 8(d) synthetic inputs: V = normalize(randn), A = normalize(alpha V + randn), bf16."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    v = torch.nn.functional.normalize(torch.randn(n, DIM, generator=g, device=device), dim=1)
    a = torch.nn.functional.normalize(alpha * v + torch.randn(n, DIM, generator=g, device=device), dim=1)
    return a.bfloat16(), v.bfloat16()


def timed(fn, steps, warmup, sync):
    """W untimed + K timed calls bracketed by sync() (barrier + cudaDeviceSynchronize); ms per step."""
    import torch
    for _ in range(warmup):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def measure(fn, steps, warmup, sync, device, all_max=lambda x: x):
    """Timed region of the ``value`` leg: K steps, clocks sampled, every instrumented kernel launch timed with
    its own CUDA events (ops.EVENT_LOG), library launch counter read on both sides."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _cabi.lib()
    for _ in range(warmup):
        fn()
    sync()
    ops.EVENT_LOG = []
    l0 = lib.pb2_launch_count()
    with ClockSampler(device.index) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        sync()
    launches = lib.pb2_launch_count() - l0
    ms = all_max(e0.elapsed_time(e1) / steps)
    log, ops.EVENT_LOG = ops.EVENT_LOG, None
    kern = {}
    for name, work, s, e in log:
        k = kern.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
        k["launches"] += 1
        k["ms"] += s.elapsed_time(e)
        k["work"] += work
    return {"ms": ms, "kernels": kern, "launches": launches, "clocks": clk.summary(), "out": out}


def roofline_of(kern, bound, step_ms=None):
    """Dominant kernel (largest summed device time) against the measured peak of its bound.  Tensor-bound kernels
    are divided by the BURST cuBLAS figure when they run alone in a short step (the step lasts under 50 ms: the
    board has not reached its power cap) and by the SUSTAINED one inside a long step; both fractions are reported."""
    pk = peaks()
    if not kern:
        return None, {}
    name, k = max(((n, v) for n, v in kern.items() if bound == "hbm" or n != "peer_reduce"), key=lambda kv: kv[1]["ms"])
    extra = {}
    if bound == "hbm":
        achieved, peak, unit, src = k["work"] / (k["ms"] * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s", pk["source"]
    else:
        achieved, unit = k["work"] / (k["ms"] * 1e-3) / 1e12, "TFLOP/s"
        burst = step_ms is not None and step_ms < 50.0
        peak, src = (pk["tf_burst"], pk["source"] + " (burst: kernel timed in a short step)") if burst else \
            (pk["tf_sustained"], pk["source"] + " (sustained: kernel timed inside a long step)")
        extra = {"frac_of_burst": achieved / pk["tf_burst"], "frac_of_sustained": achieved / pk["tf_sustained"]}
    traffic, traffic_src = ncu_traffic(name)
    roof = {"bound": bound, "kernel": name, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak, **extra,
            "traffic": traffic, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)",
            "traffic_source": traffic_src,
            "peak_source": src, "launches": k["launches"], "avg_launch_ms": k["ms"] / k["launches"]}
    # (the peer-memory pull of the sharded step is instrumented in bytes, not flops)
    in_bytes = lambda n: bound == "hbm" or n == "peer_reduce"  # noqa: E731
    table = {n: {"launches": v["launches"], "ms_total": v["ms"],
                 ("gbs" if in_bytes(n) else "tflops"): (v["work"] / (v["ms"] * 1e-3) / (1e9 if in_bytes(n) else 1e12)) if v["ms"] else None}
             for n, v in kern.items()}
    for n, row in table.items():                  # every tensor-core kernel of the step against both measured bf16 rates
        if row.get("tflops"):
            row["frac_of_sustained"], row["frac_of_burst"] = row["tflops"] / pk["tf_sustained"], row["tflops"] / pk["tf_burst"]
    return roof, table


# ------------------------------------------------------------------------- reference (CPU) arm
def cpu_sample(workload, n_s, steps, warmup):
    """The reference's algorithm (oracle port, torch-CPU, all host threads) on a bounded sample of the
    workload; returns (metric value, seconds per step, description of the sample)."""
    import torch
    from oracle import pig_oracle as O
    if workload == "eval1467":
        import random
        a, v = synth_embeddings(1467, 666, "cpu")
        a, v = a.float(), v.float()
        gd = torch.Generator().manual_seed(5)
        dur = torch.randint(20, 60, (1467,), generator=gd).float() / 10.0
        n_smp = 50                                             # a tenth of the evaluation's 500 samples (bounded CPU time)

        def step():
            torch.manual_seed(666)
            random.seed(666)
            O.resampled_recall_at_1_to_n(v, a, size=100, n_samples=n_smp, N=10)
            O.score_triplets(v, a, dur, n_samples=n_smp)
        units, what = n_smp * 100.0 * 100.0, f"{n_smp} of the 500 subsets / resamples on the 1467-clip gallery"
    elif workload == "milnce64k":
        a, v = synth_embeddings(n_s, 666, "cpu")
        a, v = a.float(), v.float()

        def step():
            vv, aa = (v / 0.07).clone().requires_grad_(True), a.clone().requires_grad_(True)
            O.milnce_loss(aa, vv).backward()
        units, what = n_s * n_s, f"{n_s} x {n_s} sub-gallery, MILNCELoss fwd+bwd (logits / 0.07)"
    elif workload == "encoder_tail":
        t = n_s * 16
        g = torch.Generator().manual_seed(666)
        x = torch.randn(t, DIM, generator=g).bfloat16().float()
        lin = torch.nn.Linear(DIM, DIM)
        step, units = (lambda: torch.nn.functional.normalize(lin(x), p=2, dim=1)), t
        what = f"{t} rows: nn.Linear(512, 512) + F.normalize (pig/models.py:105-109), torch-CPU fp32"
    elif workload == "triplets1m":
        t = n_s * 32
        g = torch.Generator().manual_seed(666)
        a, p, n = (torch.randn(t, DIM, generator=g).bfloat16().float() for _ in range(3))
        step, units, what = (lambda: O.triplet_accuracy(a, p, n)), t, f"{t} triplets x {DIM} (fp32 values of the bf16 inputs)"
    else:
        if workload == "train1024":
            n_s = 1024
        a, v = synth_embeddings(n_s, 666, "cpu")
        a, v = a.float(), v.float()

        def loss_step():
            vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)
            O.triplet_loss(vv, aa, MARGIN).backward()

        def recall_step():
            O.recall_at_1_to_n(v, a, torch.eye(n_s), N=TOP_N)

        if workload == "train1024":
            step, what = loss_step, "the full 1024 x 1024 batch, TripletLoss fwd+bwd"
        elif workload == "retrieval16k":
            step, what = recall_step, f"{n_s} x {n_s} sub-gallery, recall_at_1_to_n"
        else:
            step, what = (lambda: (loss_step(), recall_step())), f"{n_s} x {n_s} sub-gallery, TripletLoss fwd+bwd + recall_at_1_to_n"
        units = n_s * n_s
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return units / dt, dt, what


def cpu_baseline(workload, n_s, gallery_n=None):
    """The reference's algorithm on the box's host cores, ONE step of a bounded sample (no warm-up: the step is tens of
    seconds of torch-CPU work).  For the gallery the sample is a 16384-clip sub-gallery; the full
    problem cannot exist on a host (N^2 fp32 = 4 TiB), so its time is given as an explicitly labelled ~N^2
    extrapolation of the measured sample."""
    import torch
    _all_host_threads()
    value, dt, what = cpu_sample(workload, n_s, 1, 0)
    out = {"value": value, "unit": UNIT.get(workload, "pairs/s"), "cores": torch.get_num_threads(),
           "kind": "port", "sample": f"{what}; reference algorithm (oracle port); {dt:.3f} s per step; host has {os.cpu_count()} cpus"}
    if workload == "gallery" and gallery_n:
        out["extrapolated"] = {"what": f"EXTRAPOLATION, not a measurement: one {gallery_n} x {gallery_n} step at the measured pairs/s "
                                       f"(cost ~ N^2; the reference's N x N fp32 temporaries would need {gallery_n * gallery_n * 4 / 2**40:.0f} TiB)",
                               "seconds_per_step": float(gallery_n) ** 2 / value}
    return out


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host core."""
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)


def run_reference(args):
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    _all_host_threads()
    # bounded so that the whole --steps K --warmup W run ends within a few minutes: a 16384-clip sample costs ~20-40 s a step
    n_s = args.cpu_sample
    if args.workload == "gallery" and n_s == 0:
        calls = args.steps + args.warmup
        n_s = 16384 if calls <= 4 else (8192 if calls <= 12 else 4096)
    value, dt, what = cpu_sample(args.workload, n_s or 4096, args.steps, args.warmup)
    unit = UNIT.get(args.workload, "pairs/s")
    line = {
        "impl": "reference", "metric": METRIC[args.workload], "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "gallery" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "dim": DIM, "margin": MARGIN, "top_n": TOP_N,
                   "note": "CPU port of the reference (pig.loss / pig.metrics algorithms) on a bounded sample: " + what},
        "cpu_baseline": {"value": value, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{what}; host has {os.cpu_count()} cpus"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------ GPU workloads
def rank_hash_terms(ranks, first_global_row=0):
    """Order-independent 64-bit fingerprint of (global row id, rank) pairs: the terms (rank_i + 1) * odd
    multiplier(i), summed with int64 wrap-around by the caller.  Row-sharded runs add their local sums (wrap-around
    addition is associative and commutative), so 1, 2, 4 and 8 GPUs print the same number iff every rank of every
    row agrees."""
    import torch
    r = ranks.to(torch.int64)
    i = torch.arange(r.numel(), device=r.device, dtype=torch.int64) + int(first_global_row)
    mult = (i * -7046029254386353131 + 7146057691288625177) | 1         # multiplicative hash constants, wraps mod 2^64
    return (r + 1) * mult


def verify_gallery(out, a_all, v_all, rank, world, n_rows=64):
    """Outside the timed region: ``n_rows`` seeded global rows of the step's result against the blockwise restatement
    of the reference's formulas (oracle/blockwise.py; plain torch on this GPU, never this repo's kernels) -- exact
    ranks against the WHOLE gallery (identical outside a 1e-6 near-tie, inside the tie window otherwise), and the
    dA / dV rows against the fp64 closed form.  Every rank checks the sampled rows it owns."""
    import torch
    from oracle import blockwise as B
    n = a_all.shape[0]
    nl = n // world
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(2026))[:n_rows].to(a_all.device)
    mine = rows[(rows >= rank * nl) & (rows < (rank + 1) * nl)]
    res = {"rows": 0, "rank_mismatch": 0, "near_tie_rows": 0, "dA_rel_err": 0.0, "dV_rel_err": 0.0}
    if mine.numel():
        loc = mine - rank * nl
        want, near, lo, hi = B.sampled_rank_bounds(v_all, a_all, mine)
        got = out["ranks"][loc].long()
        bad = ((got != want) & ~near) | (got < lo) | (got > hi)
        res.update(rows=int(mine.numel()), rank_mismatch=int(bad.sum()), near_tie_rows=int(near.sum()))
        if out["dA"] is not None:
            rel = lambda x, r: ((x.double() - r).abs().max() / r.abs().max()).item()  # noqa: E731
            res["dA_rel_err"] = rel(out["dA"][loc], B.hinge_grad_rows(a_all, v_all, mine, MARGIN))
            res["dV_rel_err"] = rel(out["dV"][loc], B.hinge_grad_rows(v_all, a_all, mine, MARGIN))
    return res


def bench_gallery(args, rank, world, device, sync, all_max):
    import torch
    import torch.distributed as dist
    from peppa_b200.gallery import GalleryStep
    n = args.gallery_n
    assert n % world == 0
    nl = n // world
    # ONE global seed: every world size scores the SAME gallery (each rank generates it and keeps its row block),
    # so the check block below is comparable across the N = 1, 2, 4, 8 lines of a scaling run
    a_all, v_all = synth_embeddings(n, 666, device)
    a_dev, v_dev = (a_all, v_all) if world == 1 else (a_all[rank * nl:(rank + 1) * nl].clone(), v_all[rank * nl:(rank + 1) * nl].clone())
    a_host, v_host = a_dev.cpu().pin_memory(), v_dev.cpu().pin_memory()
    step = GalleryStep(nl, DIM, margin=MARGIN, top_n=TOP_N, rank=rank, world=world, device=device)
    m = measure(lambda: step.run(a_dev, v_dev), args.steps, args.warmup, sync, device, all_max)
    a_in, v_in = torch.empty_like(a_dev), torch.empty_like(v_dev)

    def run_e2e():
        a_in.copy_(a_host, non_blocking=True)
        v_in.copy_(v_host, non_blocking=True)
        o = step.run(a_in, v_in)
        return o["loss"].item(), o["recall"].cpu()

    # the end-to-end leg repeats the same seconds-long step with the copies inside the timed region: 3-5 steps
    # resolve it to well under a percent, and the whole default run has to finish within minutes
    ms_e2e = all_max(timed(run_e2e, max(3, min(args.steps, 5)), 1, sync))
    roof, table = roofline_of(m["kernels"], "tensor", m["ms"])
    if roof and roof["kernel"] == "grad_gemm":
        roof["note"] = ("the hinge step's gradient products run on tcgen05 kind::i8 (one-byte gradient matrix x two 8-bit planes): "
                        "`achieved` counts the ALGORITHMIC 2*N*N*D flops once, the kernel executes twice the MACs at the int8 rate, "
                        "so the fraction of the measured bf16 (cuBLAS) peak can exceed 1; the step's bf16 kernel is sim_hinge+rank "
                        "(kernels[...].frac_of_sustained)")
    out = m["out"]
    # ---- check block (outside the timed region): world-size independent by construction
    sums = torch.stack([rank_hash_terms(out["ranks"], rank * nl).sum(),
                        (out["ranks"].to(torch.int64) * 0 + 1).sum()])                     # [hash, rows]
    grad_sums = torch.stack([out["dA"].double().abs().sum(), out["dV"].double().abs().sum()])
    ver = verify_gallery(out, a_all, v_all, rank, world)
    vt = torch.tensor([ver["rows"], ver["rank_mismatch"], ver["near_tie_rows"]], dtype=torch.int64, device=device)
    ve = torch.tensor([ver["dA_rel_err"], ver["dV_rel_err"]], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(sums)                   # int64 wrap-around sum
        dist.all_reduce(grad_sums)
        dist.all_reduce(vt)
        dist.all_reduce(ve, op=dist.ReduceOp.MAX)
    verified = {"rows": int(vt[0]), "rank_mismatches_outside_1e-6_ties": int(vt[1]), "near_tie_rows": int(vt[2]),
                "dA_rows_max_rel_err": float(ve[0]), "dV_rows_max_rel_err": float(ve[1]),
                "checker": "oracle/blockwise.py (reference formulas, torch fp32 ranks / fp64 gradients on the GPU), 64 seeded rows "
                           "against the whole gallery, outside the timed region",
                "ok": bool(int(vt[0]) == 64 and int(vt[1]) == 0 and float(ve[0]) < 1e-3 and float(ve[1]) < 1e-3)}
    rec = out["recall"]
    check = {"loss": out["loss"].item(), "recall_at_1": rec[1].item(), "recall_at_5": rec[5].item(), "recall_at_10": rec[TOP_N].item(),
             "rank_hash": f"{int(sums[0]) & 0xFFFFFFFFFFFFFFFF:016x}", "rows_hashed": int(sums[1]),
             "dA_abs_sum": float(grad_sums[0]), "dV_abs_sum": float(grad_sums[1]), "verified": verified,
             "note": "one global seed: loss / recall / rank_hash are the same numbers at every --gpus N (rank_hash bit for bit; "
                     "loss and the gradient sums to fp32 summation order)"}
    return {
        "units": float(n) * float(n), "unit": "pairs/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": roof, "kernels": table,
        "launches": m["launches"] * world, "clocks": m["clocks"], "h2d": 2 * n * DIM * 2, "d2h": 4 + 4 * (TOP_N + 1),
        "flops_per_unit": 6.0 * DIM, "scaling": "strong",
        "dtype_note": "scores: bf16 embeddings, fp32 accumulate in TMEM (tcgen05 kind::f16), fp32 row/column normalisation in the "
                      "epilogue; gradient products: the {0,1,2} gradient matrix as u8 x the normalised embeddings as two 8-bit "
                      "planes (16 bits), exact s32 accumulate (tcgen05 kind::i8), joined and normalised in fp32",
        "check": check,
        "config": {"workload": f"gallery (BASELINE config 5): {n} x {n} audio-video gallery, hinge loss fwd+bwd + recall@1..10 from one "
                               "similarity pass, rows sharded over ranks", "gallery": n, "dim": DIM, "margin": MARGIN, "top_n": TOP_N,
                   "rows_per_gpu": nl, "l2": "per-step working set (embeddings, fp16 copies, gradient-matrix blocks of 2 GiB) far exceeds the "
                                            "126 MB L2; no flush needed",
                   "parallelism": f"row-shard x{world}" + (" + NCCL all-gather/all-reduce/reduce-scatter" if world > 1 else "")}}


def bench_milnce64k(args, device, sync):
    """North-star kernels (a) + (b): MIL-NCE (pig/loss.py:13-26) forward + backward and recall@1..10 over a 65536-clip
    gallery with a temperature -- tensor-core logits with row and column log-sum-exp and the rank counts from ONE pass
    (the N x N logits never reach HBM) and the fused backward (recomputed logits -> fp16 gradient matrix -> two
    tensor-core GEMMs)."""
    import torch
    from peppa_b200.gallery import GalleryStep
    n = 65536
    a_dev, v_dev = synth_embeddings(n, 666, device)
    a_host, v_host = a_dev.cpu().pin_memory(), v_dev.cpu().pin_memory()
    step = GalleryStep(n, DIM, device=device, loss="milnce", temperature=0.07, with_recall=True)
    steps, warm = max(args.steps, 5), max(args.warmup, 3)
    m = measure(lambda: step.run(a_dev, v_dev), steps, warm, sync, device)
    a_in, v_in = torch.empty_like(a_dev), torch.empty_like(v_dev)

    def run_e2e():
        a_in.copy_(a_host, non_blocking=True)
        v_in.copy_(v_host, non_blocking=True)
        return step.run(a_in, v_in)["loss"].item()

    ms_e2e = timed(run_e2e, 3, 1, sync)
    roof, table = roofline_of(m["kernels"], "tensor", m["ms"])
    return {
        "units": float(n) * float(n), "unit": "pairs/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": roof, "kernels": table,
        "launches": m["launches"], "clocks": m["clocks"], "h2d": 2 * n * DIM * 2, "d2h": 4, "flops_per_unit": 6.0 * DIM,
        "scaling": "weak", "check": {"loss": m["out"]["loss"].item(), "recall_at_1": m["out"]["recall"][1].item(),
                                     "recall_at_10": m["out"]["recall"][10].item()}, "steps_used": steps,
        "config": {"workload": "milnce64k (north-star kernels a + b): MILNCELoss fwd+bwd + recall@1..10, 65536 x 65536 logits, temperature "
                               "0.07, row + column log-sum-exp AND the rank counts from one pass, fused backward", "gallery": n, "dim": DIM,
                   "temperature": 0.07,
                   "l2": "2 GiB gradient-matrix blocks and 128 MiB of embeddings per step exceed the 126 MB L2; no flush needed"}}


def bench_eval1467(args, device, sync):
    """The reference's evaluation call (pig/evaluation.py:159, pig/models.py:297-303): recall@1..10 over 500 random
    100-clip subsets of a validation gallery of 1467 clips (results/data_statistics.csv), plus score_triplets with
    500 duration-matched resamples (pig/models.py:311-317) -- 25.4 s + seconds on the CPU."""
    import random

    import torch
    from peppa_b200 import metrics, triplet
    n = 1467
    a, v = synth_embeddings(n, 666, device)
    g = torch.Generator().manual_seed(5)
    dur = torch.randint(20, 60, (n,), generator=g).float() / 10.0

    def step():
        torch.manual_seed(666)
        random.seed(666)
        rec = metrics.resampled_recall_at_1_to_n(v, a, size=100, n_samples=500, N=10)
        acc = triplet.score_triplets(v, a, dur, n_samples=500)["accuracy"]
        return rec, acc

    steps, warm = max(args.steps, 5), max(args.warmup, 3)
    m = measure(step, steps, warm, sync, device)
    a_host, v_host = a.cpu().pin_memory(), v.cpu().pin_memory()
    a_in, v_in = torch.empty_like(a), torch.empty_like(v)

    def run_e2e():
        a_in.copy_(a_host, non_blocking=True)
        v_in.copy_(v_host, non_blocking=True)
        torch.manual_seed(666)
        random.seed(666)
        rec = metrics.resampled_recall_at_1_to_n(v_in, a_in, size=100, n_samples=500, N=10)
        return rec.mean().item(), triplet.score_triplets(v_in, a_in, dur, n_samples=500)["accuracy"]

    ms_e2e = timed(run_e2e, 3, 1, sync)
    rec, acc = m["out"]
    pairs = 500.0 * 100 * 100
    return {
        "units": pairs, "unit": "pairs/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": None, "kernels": {}, "launches": m["launches"],
        "clocks": m["clocks"], "h2d": 2 * n * DIM * 2, "d2h": 500 * 11 * 100 * 4 + 500 * 4, "flops_per_unit": None, "scaling": "weak",
        "check": {"recall_at_10": rec[:, 10, :].mean().item(), "triplet_acc": float(torch.as_tensor(acc).float().mean())},
        "steps_used": steps,
        "config": {"workload": "eval1467 (SURVEY 8f rows 1-2): resampled_recall_at_1_to_n(size=100, n_samples=500, N=10) + "
                               "score_triplets(n_samples=500) on a 1467-clip gallery, public API; host-side: the reference's RNG draws "
                               "(500 randperm, 500 duration-matched pairings)", "gallery": n, "dim": DIM,
                   "l2": "latency / host bound: the whole gallery is 1.5 MB"}}


def bench_train1024(args, device, sync):
    """Config 2: TripletLoss fwd+bwd at batch 1024 x 512 bf16 through peppa_b200.loss.TripletLoss + autograd.
    The step is launch bound (4 library kernels + autograd's ones_like fill), so ``value`` replays it as a CUDA graph; the
    per-kernel events and ``e2e`` run it eagerly."""
    import torch
    from peppa_b200.loss import TripletLoss
    n = 1024
    a, v = synth_embeddings(n, 666, device)
    mod = TripletLoss(MARGIN)
    vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)

    def step():
        vv.grad = None
        aa.grad = None
        loss = mod(vv, aa)
        loss.backward()
        return loss

    steps, warm = max(args.steps, 50), max(args.warmup, 5)
    for _ in range(3):
        step()
    sync()
    g = torch.cuda.CUDAGraph()          # captured before any event-instrumented eager run
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        loss = step()
    ms_graph = timed(g.replay, steps, warm, sync)
    ms_eager = timed(step, steps, warm, sync)       # the eager public-API step as a user runs it (no per-launch events)
    m = measure(step, steps, warm, sync, device)
    a_host, v_host = a.cpu().pin_memory(), v.cpu().pin_memory()

    def run_e2e_eager():
        with torch.no_grad():
            vv.copy_(v_host, non_blocking=True)
            aa.copy_(a_host, non_blocking=True)
        return step().item()

    ms_e2e_eager = timed(run_e2e_eager, steps, warm, sync)
    # the same end-to-end step as ONE CUDA graph: H2D of both pinned host batches, TripletLoss fwd+bwd through the
    # public module, D2H of the loss into pinned memory; the host waits for the stream and reads the loss every step
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        with torch.no_grad():
            vv.copy_(v_host, non_blocking=True)
            aa.copy_(a_host, non_blocking=True)
        loss_host.copy_(step().detach(), non_blocking=True)

    def run_e2e():
        g2.replay()
        torch.cuda.current_stream().synchronize()
        return loss_host.item()

    ms_e2e = timed(run_e2e, steps, warm, sync)
    assert abs(run_e2e() - loss.item()) < 1e-6
    roof, table = roofline_of(m["kernels"], "tensor", m["ms"])
    return {
        "units": float(n) * n, "unit": "pairs/s", "ms": ms_graph, "ms_eager": ms_eager, "ms_e2e": ms_e2e, "ms_e2e_eager": ms_e2e_eager,
        "roofline": roof, "kernels": table,
        "launches": m["launches"], "clocks": m["clocks"], "h2d": 2 * n * DIM * 2, "d2h": 4, "flops_per_unit": 6.0 * DIM, "scaling": "weak",
        "check": {"loss": loss.item()}, "steps_used": steps,
        "config": {"workload": "train1024 (BASELINE config 2): TripletLoss(0.2) fwd+bwd, batch 1024 x 512 bf16, public API, CUDA-graph replay",
                   "batch": n, "dim": DIM, "margin": MARGIN,
                   "l2": "4 MiB working set is L2 resident by nature of the workload (one training step re-reads its own batch)"}}


def bench_retrieval16k(args, device, sync):
    import torch
    from peppa_b200 import metrics
    n = 16384
    a, v = synth_embeddings(n, 666, device)
    steps, warm = max(args.steps, 10), max(args.warmup, 3)
    m = measure(lambda: metrics._pair_ranks(v, a, None)[0], steps, warm, sync, device)
    a_host, v_host = a.cpu().pin_memory(), v.cpu().pin_memory()
    a_in, v_in = torch.empty_like(a), torch.empty_like(v)

    def run_e2e():
        a_in.copy_(a_host, non_blocking=True)
        v_in.copy_(v_host, non_blocking=True)
        return metrics.recall_at_1_to_n(v_in, a_in, None, N=TOP_N)          # CPU float32 result like the reference

    ms_e2e = timed(run_e2e, steps, warm, sync)
    r = run_e2e()
    roof, table = roofline_of(m["kernels"], "tensor", m["ms"])
    return {
        "units": float(n) * n, "unit": "pairs/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": roof, "kernels": table, "launches": m["launches"],
        "clocks": m["clocks"], "h2d": 2 * n * DIM * 2, "d2h": 4 * (TOP_N + 1) * n, "flops_per_unit": 2.0 * DIM, "scaling": "weak",
        "check": {"recall_at_10": r[TOP_N].mean().item()}, "steps_used": steps,
        "config": {"workload": "retrieval16k (BASELINE config 3): recall_at_1_to_n(N=10), 16384 audio x 16384 video, public API", "gallery": n,
                   "dim": DIM, "top_n": TOP_N, "l2": "32 MiB of embeddings are L2 resident (as in the real evaluation, which ranks one gallery)"}}


def bench_triplets1m(args, device, sync):
    import torch
    from peppa_b200 import metrics, ops
    t = 1 << 20
    g = torch.Generator(device=device).manual_seed(666)
    a, p, n = (torch.randn(t, DIM, generator=g, device=device).bfloat16() for _ in range(3))
    steps, warm = max(args.steps, 20), max(args.warmup, 3)
    m = measure(lambda: ops.triplet_score(a, p, n), steps, warm, sync, device)
    hosts = [x.cpu().pin_memory() for x in (a, p, n)]
    ins = [torch.empty_like(x) for x in (a, p, n)]

    def run_e2e():
        for d, h in zip(ins, hosts):
            d.copy_(h, non_blocking=True)
        return metrics.triplet_accuracy(*ins).float().mean().item()

    ms_e2e = timed(run_e2e, 3, 1, sync)
    roof, table = roofline_of(m["kernels"], "hbm")
    return {
        "units": float(t), "unit": "triplets/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": roof, "kernels": table, "launches": m["launches"],
        "clocks": m["clocks"], "h2d": 3 * t * DIM * 2, "d2h": 4, "flops_per_unit": None, "scaling": "weak",
        "check": {"mean_accuracy": m["out"].mean().item()}, "steps_used": steps,
        "config": {"workload": "triplets1m (BASELINE config 4): triplet_accuracy on 2^20 triplets x 512 bf16", "triplets": t, "dim": DIM,
                   "bytes_per_triplet": 3 * DIM * 2 + 4, "l2": "3.2 GB of inputs per step exceed the 126 MB L2; no flush needed"}}


def bench_encoder_tail(args, device, sync):
    """Linear(512, 512) + L2-normalise + bf16 + rinv for 2^20 rows (one modality of the
    gallery), fused in one tcgen05 kernel (pig/models.py:96-109, :130-150)."""
    import torch
    from peppa_b200 import encoder, ops
    n = 1 << 20
    g = torch.Generator(device=device).manual_seed(666)
    x = torch.randn(n, DIM, generator=g, device=device).bfloat16()
    w = (torch.randn(DIM, DIM, generator=g, device=device) / DIM ** 0.5).bfloat16()
    b = torch.randn(DIM, generator=g, device=device) * 0.1
    steps, warm = max(args.steps, 20), max(args.warmup, 3)
    m = measure(lambda: ops.project_normalize(x, w, b)[1], steps, warm, sync, device)
    host = x.cpu().pin_memory()
    xin = torch.empty_like(x)
    mod = encoder.ProjectNormalize(DIM, DIM).to(device)

    def run_e2e():
        xin.copy_(host, non_blocking=True)
        out, rinv = mod(xin, return_rinv=True)
        return rinv.sum().item()

    ms_e2e = timed(run_e2e, 3, 1, sync)
    roof, table = roofline_of(m["kernels"], "tensor", m["ms"])
    return {
        "units": float(n), "unit": "rows/s", "ms": m["ms"], "ms_e2e": ms_e2e, "roofline": roof, "kernels": table, "launches": m["launches"],
        "clocks": m["clocks"], "h2d": n * DIM * 2, "d2h": 4, "flops_per_unit": 2.0 * DIM * DIM, "scaling": "weak",
        "check": {"mean_rinv": m["out"].mean().item()}, "steps_used": steps,
        "config": {"workload": "encoder_tail (SURVEY 8f row 3): Linear(512,512) + L2-normalise + bf16 + rinv, 2^20 rows", "rows": n,
                   "dim": DIM, "hbm_bytes_per_row": 2 * DIM * 2 + 8,
                   "l2": "1 GiB in + 1 GiB out per step exceed the 126 MB L2; the 512 KiB weight is L2 resident by design"}}


def gpu_eager_baselines(device):
    """The reference's OWN GPU path -- its ATen / cuBLAS calls, i.e. the oracle port's functions run on cuda tensors
    ("that cuBLAS-based eager path is the bar to beat") -- timed beside ours for config 2, config 3 and a
    32768^2 loss.  fp32 inputs as the reference's code sees them without AMP, and under fp16 autocast as Lightning's
    ``precision: 16`` (hparams_base.yaml:45) runs it.  A reported baseline; nothing here is on the product path."""
    import torch
    from oracle import pig_oracle as O
    out = []

    def timeit(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for n, steps in ((1024, 50), (32768, 3)):
        a, v = synth_embeddings(n, 666, device)
        a, v = a.float(), v.float()

        def loss_step(amp):
            vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                loss = O.triplet_loss(vv, aa, MARGIN)
            loss.backward()
        for amp in (False, True):
            try:
                ms = timeit(lambda: loss_step(amp), steps, 2)
                out.append({"what": f"reference TripletLoss fwd+bwd (oracle port on cuda: torch.matmul / clamp / autograd), {n} x {n}, "
                                    + ("fp16 autocast (precision: 16)" if amp else "fp32"),
                            "ms_per_step": ms, "value": float(n) * n / (ms * 1e-3), "unit": "pairs/s"})
            except RuntimeError as e:      # e.g. out of memory at 32768^2 under a crowded GPU
                out.append({"what": f"reference TripletLoss {n} x {n} amp={amp}", "error": str(e)[:200]})
        del a, v
        torch.cuda.empty_cache()
    n = 16384
    a, v = synth_embeddings(n, 666, device)
    a, v = a.float(), v.float()
    eye = torch.eye(n, device=device)
    t0 = time.perf_counter()
    O.recall_at_1_to_n(v, a, eye, N=TOP_N)          # the per-row argsort + .item() loop of pig/metrics.py:23-40, on the GPU
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.append({"what": f"reference recall_at_1_to_n(N=10) (oracle port on cuda: matmul + per-row argsort loop with .item() syncs), {n} x {n}",
                "ms_per_step": dt * 1e3, "value": float(n) * n / dt, "unit": "pairs/s"})
    return out


def line_from(res, args, world, workload):
    pk = peaks()
    value = res["units"] / (res["ms"] * 1e-3)
    line = {
        "metric": METRIC[workload], "value": value, "unit": res["unit"], "n_gpus": world, "steps": res.get("steps_used", args.steps),
        "warmup": args.warmup, "ms_per_step": res["ms"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
        "dtype": "bf16", "dtype_note": res.get("dtype_note", "bf16 operands, fp32 accumulate (tcgen05 kind::f16)"),
        "data": "synthetic", "config": res["config"], "roofline": res["roofline"], "kernels": res["kernels"],
        "clocks": res["clocks"],
        "e2e": {"value": res["units"] / (res["ms_e2e"] * 1e-3), "unit": res["unit"], "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"], "ms_per_step": res["ms_e2e"]},
        "gpu_launches": res["launches"], "check": res["check"],
    }
    if res["flops_per_unit"]:
        tf = res["flops_per_unit"] * value / world / 1e12
        line["frac_of_bf16_peak"] = tf / pk["tf_sustained"]
        line["frac_of_bf16_peak_burst"] = tf / pk["tf_burst"]
        line["frac_of_bf16_peak_note"] = (f"algorithmic flops (SURVEY 8d) per GPU = {tf:.1f} TFLOP/s over the measured cuBLAS bf16 rates: "
                                          f"sustained {pk['tf_sustained']} (frac_of_bf16_peak: the step runs for seconds at the power cap) "
                                          f"and burst {pk['tf_burst']} (frac_of_bf16_peak_burst)")
    if "ms_eager" in res:
        line["ms_per_step_eager"] = res["ms_eager"]
    if "ms_e2e_eager" in res:
        line["e2e"]["ms_per_step_eager"] = res["ms_e2e_eager"]
        line["e2e"]["note"] = "one CUDA graph per step: H2D of both pinned batches + public-API fwd+bwd + D2H of the loss, host sync each step"
    return line


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gallery", choices=["gallery", "train1024", "retrieval16k", "triplets1m", "encoder_tail", "milnce64k", "eval1467"])
    ap.add_argument("--gallery-n", type=int, default=1 << 20)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="clips in the CPU arm's sub-gallery (0 = automatic: 16384 for cpu_baseline and short reference runs, "
                         "smaller when --steps + --warmup would otherwise take more than a few minutes)")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: libraries that print there (NCCL writes its version line to
    # stdout when NCCL_DEBUG is set on the box) are sent to stderr, and emit() below writes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    if local == 0:
        from peppa_b200 import build as _build
        _build.build()                  # no-op when the in-tree .so is current
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    single = {"train1024": bench_train1024, "retrieval16k": bench_retrieval16k, "triplets1m": bench_triplets1m,
              "encoder_tail": bench_encoder_tail, "milnce64k": bench_milnce64k, "eval1467": bench_eval1467}
    if args.workload == "gallery":
        line = line_from(bench_gallery(args, rank, world, device, sync, all_max), args, world, "gallery")
    else:
        # single-GPU workloads: every rank of a torchrun launch measures its own replica, rank 0 reports
        line = line_from(single[args.workload](args, device, torch.cuda.synchronize), args, 1, args.workload)
        line["n_gpus"] = world
        line["value"] *= world
        line["note"] = "replicas only: this workload does not shard; N independent replicas" if world > 1 else None
    if world == 1 and rank == 0:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_sample or (16384 if args.workload == "gallery" else 4096),
                                            args.gallery_n if args.workload == "gallery" else None)
        if args.workload == "gallery" and not args.no_extras:
            line["gpu_eager_baseline"] = gpu_eager_baselines(device)
            line["other_workloads"] = []
            for w, fn in single.items():
                o = line_from(fn(args, device, torch.cuda.synchronize), args, 1, w)
                line["other_workloads"].append({k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e",
                                                                  "gpu_launches", "check") if k in o}
                                               | ({"frac_of_bf16_peak": o["frac_of_bf16_peak"]} if "frac_of_bf16_peak" in o else {})
                                               | ({"ms_per_step_eager": o["ms_per_step_eager"]} if "ms_per_step_eager" in o else {}))
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
