#!/usr/bin/env python
"""Benchmark of the contrastive-scoring hot path (BASELINE.json metric: similarity pairs/sec for
loss fwd+bwd + recall@k at 1/2/4/8 B200, as a fraction of bf16 tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Workload (``config.workload``): ``gallery`` (default) = BASELINE config 5, the only configuration the
metric is quoted on at 1/2/4/8 GPUs and one that fits a single B200 (2 GiB of embeddings): a
2^20 x 2^20 audio<->video gallery, one step = symmetric hinge loss forward + gradients (dA, dV)
AND recall@1..10 from a single pass over the similarity matrix, rows sharded over the ranks
(strong scaling: the gallery is fixed, each rank owns N/P rows).  ``train1024`` (config 2),
``retrieval16k`` (config 3) and ``triplets1m`` (config 4) are measured too at N=1 and reported
under ``other_workloads`` of the same JSON line (or as the main line with --workload).

One JSON line on stdout (rank 0).  ``value`` is timed with device-resident inputs, ``e2e`` through
the public API with pinned-host inputs copied in and the loss/recall read back every step.
``--impl reference`` times the CPU port of the reference (oracle/pig_oracle.py, torch-CPU, all host
threads) on a bounded sample of the same workload; /root/reference is pure Python and does not
exist on the GPU box, so the oracle port is the reference arm.
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
        for r in self.rows:
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


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum (GB) of one launch on a 32768 x 32768 block, from the committed
    ncu --set full capture (profiles/r1_ncu_summary.json); None if unknown."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_summary.json")
    key = {"sim_hinge+rank": "r1_sim_hinge", "sim_hinge": "r1_sim_hinge", "grad_gemm": "r1_grad_gemm"}.get(kernel)
    if not key or not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f).get(key, {})

    def gb(s):
        v, u = s.split()[:2]
        return float(v) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]
    try:
        return gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
    except (KeyError, ValueError):
        return None


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


# ------------------------------------------------------------------------- reference (CPU) arm
def cpu_gallery_sample(n_s, steps, warmup):
    """Reference path on a bounded sample: TripletLoss fwd+bwd + recall_at_1_to_n(N=10) on an
    n_s x n_s sub-gallery with the port of pig/loss.py + pig/metrics.py (torch-CPU, all threads)."""
    import torch
    from oracle import pig_oracle as O
    a, v = synth_embeddings(n_s, 666, "cpu")
    a, v = a.float(), v.float()

    def step():
        vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)
        O.triplet_loss(vv, aa, MARGIN).backward()
        O.recall_at_1_to_n(v, a, torch.eye(n_s), N=TOP_N)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_s * n_s / dt, dt


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_s = args.cpu_sample
    value, dt = cpu_gallery_sample(n_s, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "similarity pairs/sec (loss fwd+bwd + recall@1..10)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gallery", "gallery": args.gallery_n, "dim": DIM, "margin": MARGIN, "top_n": TOP_N,
                   "note": "CPU port of pig.loss.TripletLoss fwd+bwd + pig.metrics.recall_at_1_to_n on a bounded sample"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n_s} x {n_s} sub-gallery of the same synthetic embeddings; host has {os.cpu_count()} cpus"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ GPU workloads
def bench_gallery(args, rank, world, device, sync, all_max):
    import torch
    from peppa_b200 import _cabi, ops
    from peppa_b200.gallery import GalleryStep
    n = args.gallery_n
    assert n % world == 0
    nl = n // world
    a_dev, v_dev = synth_embeddings(nl, 666 + rank, device)
    a_host, v_host = a_dev.cpu().pin_memory(), v_dev.cpu().pin_memory()
    step = GalleryStep(nl, DIM, margin=MARGIN, top_n=TOP_N, rank=rank, world=world, device=device)
    lib = _cabi.lib()

    def run_dev():
        return step.run(a_dev, v_dev)

    # kernel-only timing with device-resident inputs, per-kernel CUDA events for the roofline
    for _ in range(args.warmup):
        run_dev()
    sync()
    ops.EVENT_LOG = []
    launches0 = lib.pb2_launch_count()
    with ClockSampler(device.index) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = run_dev()
        e1.record()
        sync()
    launches = lib.pb2_launch_count() - launches0
    ms = all_max(e0.elapsed_time(e1) / args.steps)
    log, ops.EVENT_LOG = ops.EVENT_LOG, None
    kern = {}
    for name, flops, s, e in log:
        k = kern.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0})
        k["launches"] += 1
        k["ms"] += s.elapsed_time(e)
        k["flops"] += flops

    a_in, v_in = torch.empty_like(a_dev), torch.empty_like(v_dev)

    def run_e2e():
        a_in.copy_(a_host, non_blocking=True)
        v_in.copy_(v_host, non_blocking=True)
        o = step.run(a_in, v_in)
        return o["loss"].item(), o["recall"].cpu()

    ms_e2e = all_max(timed(run_e2e, args.steps, 1, sync))
    res = {"ms": ms, "ms_e2e": ms_e2e, "pairs": float(n) * float(n), "kernels": kern, "launches": launches * world,
           "clocks": clk.summary(), "h2d": 2 * n * DIM * 2, "d2h": 4 + 4 * (TOP_N + 1), "loss": out["loss"].item(),
           "recall10": out["recall"][TOP_N].item()}
    return res


def bench_train1024(args, device):
    """BASELINE config 2: TripletLoss fwd+bwd at batch 1024 x 512 through the public API
    (peppa_b200.loss.TripletLoss + autograd), replayed as a CUDA graph (the step is launch bound)."""
    import torch
    from peppa_b200.loss import TripletLoss
    n = 1024
    a, v = synth_embeddings(n, 666, device)        # bf16 leaves, bf16 gradients (config 2: "1024 x 512-d bf16")
    mod = TripletLoss(MARGIN)
    vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)

    def step():
        vv.grad = None
        aa.grad = None
        loss = mod(vv, aa)
        loss.backward()
        return loss

    sync = torch.cuda.synchronize
    eager = timed(step, 20, 5, sync)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        loss = step()
    graph = timed(g.replay, 50, 5, sync)
    return {"workload": "train1024 (config 2): TripletLoss fwd+bwd, batch 1024 x 512", "pairs_per_s_graph": n * n / graph * 1e3,
            "ms_graph": graph, "ms_eager": eager, "frac_of_bf16_peak": 6.0 * n * n * DIM / (graph * 1e-3) / (peaks()["tf_burst"] * 1e12),
            "loss": loss.item()}


def bench_retrieval16k(args, device):
    import torch
    from peppa_b200 import metrics
    n = 16384
    a, v = synth_embeddings(n, 666, device)
    sync = torch.cuda.synchronize
    ms = timed(lambda: metrics.recall_at_1_to_n(v, a, None, N=TOP_N), 10, 3, sync)
    r = metrics.recall_at_1_to_n(v, a, None, N=TOP_N)
    return {"workload": "retrieval16k (config 3): recall_at_1_to_n, 16384 x 16384, N=10, public API incl. D2H of the result",
            "pairs_per_s": n * n / ms * 1e3, "ms": ms, "frac_of_bf16_peak": 2.0 * n * n * DIM / (ms * 1e-3) / (peaks()["tf_burst"] * 1e12),
            "recall_at_10": r[TOP_N].mean().item()}


def bench_triplets1m(args, device):
    import torch
    from peppa_b200 import ops
    t = 1 << 20
    g = torch.Generator(device=device).manual_seed(666)
    a, p, n = (torch.randn(t, DIM, generator=g, device=device).bfloat16() for _ in range(3))
    sync = torch.cuda.synchronize
    ms = timed(lambda: ops.triplet_score(a, p, n), 20, 3, sync)     # 3.2 GB of inputs per call > 126 MB L2
    byt = t * (3 * DIM * 2 + 4)
    return {"workload": "triplets1m (config 4): triplet_accuracy, 2^20 triplets x 512 bf16", "triplets_per_s": t / ms * 1e3, "ms": ms,
            "roofline": {"bound": "hbm", "achieved": byt / ms / 1e6, "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                         "frac": byt / ms / 1e6 / peaks()["hbm_gbs"], "traffic": 3.2286, "algorithmic": byt / 1e9,
                         "traffic_unit": "GB per launch (ncu dram read+write, profiles/r1_ncu_summary.json)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gallery", choices=["gallery"])
    ap.add_argument("--gallery-n", type=int, default=1 << 20)
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
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

    res = bench_gallery(args, rank, world, device, sync, all_max)
    pk = peaks()
    value = res["pairs"] / (res["ms"] * 1e-3)
    # dominant kernel = the one with the largest summed device time on rank 0
    dom_name, dom = max(res["kernels"].items(), key=lambda kv: kv[1]["ms"])
    achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
    line = {
        "metric": "similarity pairs/sec (loss fwd+bwd + recall@1..10)", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"gallery (BASELINE config 5): {args.gallery_n} x {args.gallery_n} audio-video gallery, hinge loss "
                               "fwd+bwd + recall@1..10 from one similarity pass, rows sharded over ranks",
                   "gallery": args.gallery_n, "dim": DIM, "margin": MARGIN, "top_n": TOP_N, "rows_per_gpu": args.gallery_n // world,
                   "l2": "inputs (2 GiB bf16 + 4 GiB fp32 partials per step) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"row-shard x{world}" + (" + NCCL all-gather/all-reduce/reduce-scatter" if world > 1 else "")},
        "frac_of_bf16_peak": 6.0 * res["pairs"] * DIM / (res["ms"] * 1e-3) / world / (pk["tf_sustained"] * 1e12),
        "roofline": {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tf_sustained"], "traffic": ncu_traffic(dom_name),
                     "traffic_unit": "GB per launch (32768 x 32768 block, ncu --set full, profiles/r1_ncu_summary.json)",
                     "peak_source": pk["source"] + " (sustained)", "launches": dom["launches"]},
        "kernels": {k: {"launches": v["launches"], "ms_total": v["ms"], "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else None}
                    for k, v in res["kernels"].items()},
        "clocks": res["clocks"],
        "e2e": {"value": res["pairs"] / (res["ms_e2e"] * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"], "ms_per_step": res["ms_e2e"]},
        "gpu_launches": res["launches"],
        "check": {"loss": res["loss"], "recall_at_10": res["recall10"]},
    }
    if world == 1 and rank == 0:
        import torch as _t
        cpu_value, cpu_dt = cpu_gallery_sample(args.cpu_sample, 1, 1)
        line["cpu_baseline"] = {"value": cpu_value, "unit": "pairs/s", "cores": _t.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_sample} x {args.cpu_sample} sub-gallery, reference algorithm (oracle port) fwd+bwd + "
                                          f"recall_at_1_to_n; {cpu_dt:.2f} s per step; host has {os.cpu_count()} cpus"}
        if not args.no_extras:
            line["other_workloads"] = [bench_train1024(args, device), bench_retrieval16k(args, device), bench_triplets1m(args, device)]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
