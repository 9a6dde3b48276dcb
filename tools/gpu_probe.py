"""Bring-up probes for the B200 box: each step runs in its own process (a trap in one kernel
must not take the others down).  ``python tools/gpu_probe.py all`` runs every step under a
timeout and writes ``gpurun_out/probe_<step>.log``; ``python tools/gpu_probe.py <step>`` runs one.
Development tooling only -- the judged checks are tests/ and bench.py.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time


def _measure_lib():
    """The pb2_debug_* selectors live in the measurement build only."""
    from peppa_b200 import _cabi
    return _cabi.use_measurement_library()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def _t(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def emb(n, alpha=4.0, d=512, seed=666, device="cuda"):
    import torch
    g = torch.Generator().manual_seed(seed)
    V = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    A = torch.nn.functional.normalize(alpha * V + torch.randn(n, d, generator=g), dim=1)
    return V.bfloat16().to(device), A.bfloat16().to(device)


def step_triplet():
    import torch
    from peppa_b200 import metrics
    from oracle import pig_oracle as O
    for dt in (torch.bfloat16, torch.float32, torch.float16):
        g = torch.Generator().manual_seed(1)
        a, p, n = (torch.randn(1000, 512, generator=g).to(dt) for _ in range(3))
        got = metrics.triplet_accuracy(a.cuda(), p.cuda(), n.cuda(), discrete=False).float().cpu()
        ref = O.triplet_accuracy(a.float(), p.float(), n.float(), discrete=False)
        print(dt, "max abs gap err", (got - ref).abs().max().item())
        gd = metrics.triplet_accuracy(a.cuda(), p.cuda(), n.cuda()).float().cpu()
        rd = O.triplet_accuracy(a.float(), p.float(), n.float())
        print(dt, "discrete mismatches", int((gd != rd).sum()))
    from peppa_b200 import ops
    for dt, T in ((torch.bfloat16, 1 << 20), (torch.float32, 1 << 20)):
        a, p, n = (torch.randn(T, 512, device="cuda").to(dt) for _ in range(3))
        ms = _t(lambda: ops.triplet_score(a, p, n))
        byt = T * (3 * 512 * a.element_size() + 4)
        print(f"triplet {dt} T={T}: {ms:.3f} ms  {byt / ms / 1e6:.1f} GB/s  {T / ms / 1e6:.2f} Gtriplets/s")


def step_simmatrix():
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    torch.manual_seed(0)
    for (r, c, d) in [(128, 256, 64), (128, 256, 512), (128, 64, 512), (300, 500, 512), (1024, 1024, 512), (1000, 1536, 128)]:
        x = torch.randn(r, d, device="cuda").bfloat16()
        y = torch.randn(c, d, device="cuda").bfloat16()
        ref = x.float() @ y.float().T
        for bn in (64, 128, 256):
            lib.pb2_debug_force_bn(bn)
            got = ops.sim_matrix(x, y)
            torch.cuda.synchronize()
            err = (got - ref).abs().max().item()
            print(f"sim_matrix r={r} c={c} d={d} bn={bn}: max abs err {err:.3e} (ref max {ref.abs().max().item():.2f})",
                  "OK" if err < 1e-2 else "MISMATCH")
            if err >= 1e-2:
                bad = (got - ref).abs() > 1e-2
                idx = bad.nonzero()
                print("   first bad", idx[:5].tolist(), "n_bad", int(bad.sum()), "got", got[bad][:5].tolist(), "ref",
                      ref[bad][:5].tolist())
    lib.pb2_debug_force_bn(0)
    V, A = emb(16384)
    rv, _ = ops.row_norms(V)
    ra, _ = ops.row_norms(A)
    for bn in (128, 256):
        lib.pb2_debug_force_bn(bn)
        ms = _t(lambda: ops.sim_matrix(V, A, rv, ra), iters=5)
        print(f"sim_matrix 16384^2 bn={bn}: {ms:.3f} ms  {2 * 16384**2 * 512 / ms / 1e9:.1f} TFLOP/s")
    lib.pb2_debug_force_bn(0)


def step_rank():
    import torch
    from peppa_b200 import _cabi, metrics, ops
    from oracle import pig_oracle as O
    lib = _measure_lib()
    for n, alpha in ((8, 4.0), (100, 0.5), (1000, 4.0), (4096, 0.5)):
        V, A = emb(n, alpha)
        ranks_ref, near = O.ranks_identity(V.float().cpu(), A.float().cpu())
        for bn in (64, 128, 256):
            lib.pb2_debug_force_bn(bn)
            rank, *_ = metrics._pair_ranks(V, A, None)
            rank = rank.cpu().long()
            bad = (rank != ranks_ref) & ~near
            print(f"rank n={n} alpha={alpha} bn={bn}: mismatches {int(bad.sum())} (near-tie rows {int(near.sum())}),"
                  f" raw diff rows {int((rank != ranks_ref).sum())}")
    lib.pb2_debug_force_bn(0)
    V, A = emb(16384)
    rv, _ = ops.row_norms(V)
    ra, _ = ops.row_norms(A)
    idx = torch.arange(16384, device="cuda")
    _, pd = ops.pair_dot(A, V, rinv_x=ra, rinv_y=rv, want_dist=True)
    for bn in (128, 256):
        lib.pb2_debug_force_bn(bn)
        ms = _t(lambda: ops.sim_rank(A, V, ra, rv, pd, idx), iters=10)
        print(f"sim_rank 16384^2 bn={bn}: {ms:.3f} ms  {2 * 16384**2 * 512 / ms / 1e9:.1f} TFLOP/s  {16384**2 / ms / 1e6:.1f} Gpairs/s")
    lib.pb2_debug_force_bn(0)
    ms = _t(lambda: metrics.recall_at_1_to_n(V, A, None, N=10), iters=5)
    print(f"recall_at_1_to_n 16384^2 end-to-end API: {ms:.3f} ms")


def _gradgemm(gdt, zdt, perf=True):
    import torch
    from peppa_b200 import ops
    torch.manual_seed(0)
    for tr in (False, True):
        for (r, c, d) in [(128, 64, 256), (128, 128, 512), (256, 192, 512), (300, 500, 512), (1024, 1024, 512)]:
            g = torch.randn(r, c, device="cuda").to(gdt)
            gm, ld = ops.gmat_alloc(r, c, "cuda")
            gm = gm.view(torch.int16).view(gdt) if gdt != gm.dtype else gm
            gm.zero_()
            gm[:, :c] = g
            z = torch.randn(c if not tr else r, d, device="cuda").to(zdt)
            ref = (g.float() if not tr else g.float().T) @ z.float()
            got = ops.grad_gemm(gm, r, c, ld, z, transpose=tr)
            torch.cuda.synchronize()
            err = (got - ref).abs().max().item()
            print(f"grad_gemm {gdt} x {zdt} r={r} c={c} d={d} T={tr}: err {err:.3e} ref max {ref.abs().max():.1f}",
                  "OK" if err < 0.05 else "MISMATCH", flush=True)
    if not perf:
        return
    n = 16384
    gm, ld = ops.gmat_alloc(n, n, "cuda")
    gm = gm.view(torch.int16).view(gdt) if gdt != gm.dtype else gm
    gm.normal_()
    z = torch.randn(n, 512, device="cuda").to(zdt)
    for tr in (False, True):
        ms = _t(lambda: ops.grad_gemm(gm, n, n, ld, z, transpose=tr), iters=5)
        print(f"grad_gemm 16384 T={tr}: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s", flush=True)


def step_gg_f16_f16():
    import torch
    _gradgemm(torch.float16, torch.float16)


def step_gg_bf16_bf16():
    import torch
    _gradgemm(torch.bfloat16, torch.bfloat16, perf=False)


def step_gg_f16_bf16():
    import torch
    _gradgemm(torch.float16, torch.bfloat16, perf=False)


def step_gradgemm_sweep():
    """Only if the default MN-major descriptor is wrong: try the plausible alternatives."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    torch.manual_seed(0)
    r, c, d = 128, 128, 256
    g = torch.randn(r, c, device="cuda").half()
    gm, ld = ops.gmat_alloc(r, c, "cuda")
    gm.zero_()
    gm[:, :c] = g
    z = torch.randn(c, d, device="cuda").bfloat16()
    ref = g.float() @ z.float()
    for (lbo, sbo, ks) in [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 32), (1024, 8192, 32), (16, 1024, 2048),
                           (128, 1024, 2048), (8192, 128, 2048)]:
        lib.pb2_debug_set_mn_desc(lbo, sbo, ks)
        got = ops.grad_gemm(gm, r, c, ld, z.half(), transpose=False)
        torch.cuda.synchronize()
        print(f"desc=({lbo},{sbo},{ks}) err {(got - ref).abs().max().item():.3e}")


def step_loss():
    import torch
    from peppa_b200 import loss as L
    from oracle import pig_oracle as O
    for n, alpha in ((8, 4.0), (100, 0.5), (257, 4.0), (1024, 4.0), (1024, 0.5), (2048, 1.0)):
        V, A = emb(n, alpha)
        for name, mod, ref in (("hinge", L.TripletLoss(0.2), lambda v, a: O.hinge_loss_and_grads(v, a, 0.2)),
                               ("milnce", L.MILNCELoss(), O.milnce_loss_and_grads)):
            v = V.clone().requires_grad_(True)
            a = A.clone().requires_grad_(True)
            out = mod(v, a)
            out.backward()
            rl, rdv, rda = ref(V.float().cpu(), A.float().cpu())
            el = abs(out.item() - rl.item()) / abs(rl.item())
            ev = ((v.grad.float().cpu() - rdv).abs().max() / rdv.abs().max()).item()
            ea = ((a.grad.float().cpu() - rda).abs().max() / rda.abs().max()).item()
            print(f"{name} n={n} alpha={alpha}: loss {out.item():.6f} ref {rl.item():.6f} rel {el:.2e}  dV rel {ev:.2e}  dA rel {ea:.2e}",
                  "OK" if max(el, ev, ea) < 1e-3 else "MISMATCH")
    for n in (1024, 16384):
        V, A = emb(n)
        for name, mod in (("hinge", L.TripletLoss(0.2)), ("milnce", L.MILNCELoss())):
            def run():
                v = V.clone().requires_grad_(True)
                a = A.clone().requires_grad_(True)
                mod(v, a).backward()
            ms = _t(run, iters=5)
            print(f"{name} fwd+bwd n={n}: {ms:.3f} ms  {n * n / ms / 1e6:.2f} Gpairs/s  {6 * n * n * 512 / ms / 1e9:.1f} TFLOP/s(alg)")




def main():
    global STEPS
    STEPS = {k[5:]: v for k, v in list(globals().items()) if k.startswith("step_")}
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what != "all":
        STEPS[what]()
        return
    os.makedirs(OUT, exist_ok=True)
    order = sys.argv[2:] or ["triplet", "simmatrix", "rank", "gg_f16_f16", "gg_bf16_bf16", "gg_f16_bf16", "loss"]
    for s in order:
        t0 = time.time()
        log = os.path.join(OUT, f"probe_{s}.log")
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), s], stdout=f, stderr=subprocess.STDOUT,
                                   timeout=240, cwd=ROOT)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        print(f"=== {s}: rc={rc} {time.time() - t0:.1f}s")
        with open(log) as f:
            txt = f.read()
        print(txt[-6000:])



def step_hinge_perf():
    """sim_hinge(+rank) on a 32768 x 32768 block for each tile configuration (pb2_debug_force_bn)."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    n = 32768
    V, A = emb(n)
    rv, _ = ops.row_norms(V)
    ra, _ = ops.row_norms(A)
    diag, thr = ops.sim_diag(A, V, ra, rv)
    g, ld = ops.gmat_alloc(n, n, "cuda")
    rc = torch.zeros(n, dtype=torch.int32, device="cuda")
    cc = torch.zeros(n, dtype=torch.int32, device="cuda")
    rk = torch.zeros(n, dtype=torch.int32, device="cuda")
    for bn in (128, 192, 256):
        lib.pb2_debug_force_bn(bn)
        for with_rank in (True, False):
            fn = lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr if with_rank else None,
                                       rank=rk if with_rank else None)
            ms = _t(fn, iters=10)
            print(f"sim_hinge bn={bn} rank={with_rank}: {ms:.3f} ms  {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, None, 0), iters=10)
        print(f"sim_hinge bn={bn} no-G fwd only: {ms:.3f} ms  {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s", flush=True)
    for bits, what in ((0, "baseline"), (1, "no TMA store"), (3, "no TMA store, no STS"), (7, "no TMA store, no STS, no fence"),
                       (4, "no proxy fence (stores race: timing only)"), (0, "baseline again")):
        lib.pb2_debug_force_bn(256 | (bits << 16))
        ms = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk), iters=200, warm=20)
        print(f"sim_hinge bn=256 rank=True [{what}]: {ms:.3f} ms (200 back to back)", flush=True)
    lib.pb2_debug_force_bn(0)




def step_hinge_dim():
    """Is the fused hinge pass MMA-bound or epilogue-bound?  Time it at D = 128 .. 1024 (same epilogue work)."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    n = 32768
    for d in (128, 256, 512, 1024):
        V, A = emb(n, d=d)
        rv, _ = ops.row_norms(V)
        ra, _ = ops.row_norms(A)
        diag, thr = ops.sim_diag(A, V, ra, rv)
        g, ld = ops.gmat_alloc(n, n, "cuda")
        rc = torch.zeros(n, dtype=torch.int32, device="cuda")
        cc = torch.zeros(n, dtype=torch.int32, device="cuda")
        rk = torch.zeros(n, dtype=torch.int32, device="cuda")
        idx = torch.arange(n, device="cuda")
        for bn in (192, 256):
            lib.pb2_debug_force_bn(bn)
            ms = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk), iters=10)
            print(f"D={d} bn={bn} hinge+rank+G: {ms:.3f} ms", flush=True)
        lib.pb2_debug_force_bn(256)
        ms = _t(lambda: ops.sim_rank(A, V, ra, rv, thr, idx), iters=10)
        print(f"D={d} rank only: {ms:.3f} ms", flush=True)
    lib.pb2_debug_force_bn(0)


def step_simpair():
    """Similarity kernels on CTA pairs (cta_group::2) vs independent 128 x 256 tiles: same results, sustained time."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    n = 32768
    V, A = emb(n)
    rv, _ = ops.row_norms(V)
    ra, _ = ops.row_norms(A)
    diag, thr = ops.sim_diag(A, V, ra, rv)
    g, ld = ops.gmat_alloc(n, n, "cuda")
    idx = torch.arange(n, device="cuda")
    res = {}
    for pair in (0, 1, 2, 0):      # 0 independent CTAs, 1 shared-MMA pairs, 2 clusters of 2 with Y multicast
        lib.pb2_debug_sim_pair(pair)
        rc = torch.zeros(n, dtype=torch.int32, device="cuda")
        cc = torch.zeros(n, dtype=torch.int32, device="cuda")
        rk = torch.zeros(n, dtype=torch.int32, device="cuda")
        part = ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk)
        rk2 = ops.sim_rank(A, V, ra, rv, thr, idx)
        torch.cuda.synchronize()
        res[pair] = (rc.clone(), cc.clone(), rk.clone(), rk2.clone(), part.double().sum().item(), g[:, :n].float().sum().item())
        ms_h = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk), iters=200, warm=20)
        ms_r = _t(lambda: ops.sim_rank(A, V, ra, rv, thr, idx, rank=rk2), iters=200, warm=20)
        print(f"pair={pair}: hinge+rank+G {ms_h:.3f} ms ({2 * n * n * 512 / ms_h / 1e9:.0f} TF/s)   rank only {ms_r:.3f} ms "
              f"({2 * n * n * 512 / ms_r / 1e9:.0f} TF/s)   (200 back to back)", flush=True)
    for m in (1, 2):
        a, b = res[0], res[m]
        print(f"mode {m} vs 0: row counts equal", torch.equal(a[0], b[0]), "col counts equal", torch.equal(a[1], b[1]),
              "hinge ranks equal", torch.equal(a[2], b[2]), "rank-kernel ranks equal", torch.equal(a[3], b[3]),
              "loss partial sums", a[4], b[4], "G sums", a[5], b[5], flush=True)
    lib.pb2_debug_sim_pair(-1)


def step_milnce_perf():
    """MIL-NCE gallery step (pig/loss.py:13-26 fwd+bwd, K = 1) on a 65536-clip gallery: time per kernel."""
    import torch
    from peppa_b200 import ops
    from peppa_b200.gallery import GalleryStep
    n = 65536
    V, A = emb(n)
    step = GalleryStep(n, 512, loss="milnce", temperature=0.07)
    for _ in range(2):
        out = step.run(A, V)
    torch.cuda.synchronize()
    ops.EVENT_LOG = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = step.run(A, V)
    e1.record()
    torch.cuda.synchronize()
    log, ops.EVENT_LOG = ops.EVENT_LOG, None
    agg = {}
    for name, work, s, e in log:
        k = agg.setdefault(name, [0, 0.0, 0.0])
        k[0] += 1
        k[1] += s.elapsed_time(e)
        k[2] += work
    ms = e0.elapsed_time(e1) / 3
    print(f"milnce gallery {n}: {ms:.2f} ms/step  {n * n / ms / 1e6:.1f} Gpairs/s  loss {out['loss'].item():.5f}")
    for name, (cnt, t, w) in agg.items():
        print(f"  {name}: {cnt} launches, {t / cnt:.3f} ms avg, {w / t / 1e9:.0f} TF/s")


def step_gg_units():
    """Gradient GEMM (CTA pairs, 512-wide, stream-K) on fewer SM pairs: cuBLAS's own kernel runs 33 clusters of 4
    (132 of 148 SMs) and is ahead under the power cap -- does leaving SMs idle help ours?"""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    n = 32768
    gm, ld = ops.gmat_alloc(n, n, "cuda")
    gm.copy_(torch.randint(0, 3, (n, ld), device="cuda").half())
    z = (torch.randn(n, 512, device="cuda") * 0.05).half()
    out = torch.zeros(n, 512, device="cuda")
    for cap in (74, 70, 66, 62, 56, 74):
        lib.pb2_debug_gg_units(cap)
        for tr in (False, True):
            ms = _t(lambda: ops.grad_gemm(gm, n, n, ld, z, transpose=tr, out=out, accumulate=True), iters=300, warm=20)
            print(f"pairs={cap} T={tr}: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.0f} TF/s (300 back to back)", flush=True)
    lib.pb2_debug_gg_units(0)
    ms = _t(lambda: torch.matmul(gm[:, :n], z), iters=300, warm=20)
    print(f"torch.matmul: {ms:.3f} ms", flush=True)


def step_streamk():
    """grad_gemm: CTA pairs (cta_group::2) on/off x stream-K on/off: agreement with fp64, determinism, sustained time."""
    import torch
    from peppa_b200 import _cabi, ops
    lib = _measure_lib()
    torch.manual_seed(0)
    for tr in (False, True):
        for (r, c, d) in [(20480, 2048, 512), (2048, 20480, 512), (20000, 1000, 512), (9000, 4100, 256), (40960, 640, 256)]:
            gm, ld = ops.gmat_alloc(r, c, "cuda")
            gm.zero_()
            gm[:, :c] = torch.randint(0, 3, (r, c), device="cuda").half()
            z = (torch.randn(r if tr else c, d, device="cuda") * 0.05).half()
            ref = ((gm[:, :c].double().T if tr else gm[:, :c].double()) @ z.double())
            for pair in (0, 1, 2):
                lib.pb2_debug_gg_pair(pair)
                a = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=False)
                b = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=True)
                b2 = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, stream_k=True)
                base = torch.ones_like(a)
                acc = ops.grad_gemm(gm, r, c, ld, z, transpose=tr, out=base, accumulate=True, alpha=0.5)
                torch.cuda.synchronize()
                e_a = ((a - ref).abs().max() / ref.abs().max()).item()
                e_b = ((b - ref).abs().max() / ref.abs().max()).item()
                e_acc = ((acc - (1 + 0.5 * ref)).abs().max() / ref.abs().max()).item()
                print(f"T={tr} r={r} c={c} d={d} pair={pair}: tiles err {e_a:.2e}  stream-K err {e_b:.2e}  acc err {e_acc:.2e}  "
                      f"deterministic {bool((b == b2).all())}", "OK" if max(e_a, e_b, e_acc) < 3e-5 else "MISMATCH", flush=True)
    n = 32768
    gm, ld = ops.gmat_alloc(n, n, "cuda")
    gm.copy_(torch.randint(0, 3, (n, ld), device="cuda").half())
    z = (torch.randn(n, 512, device="cuda") * 0.05).half()
    out = torch.zeros(n, 512, device="cuda")
    for pair in (0, 1, 2):
        lib.pb2_debug_gg_pair(pair)
        for tr in (False, True):
            for sk in (False, True):
                ms = _t(lambda: ops.grad_gemm(gm, n, n, ld, z, transpose=tr, out=out, accumulate=True, stream_k=sk), iters=300, warm=20)
                print(f"grad_gemm 32768^2 pair={pair} T={tr} stream_k={sk}: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s "
                      f"(300 back to back)", flush=True)
    lib.pb2_debug_gg_pair(-1)
    # the library on the same products (fp16 in, fp32 accumulate, fp16 out): what 1 kW buys cuBLAS here
    for tr in (False, True):
        gmv = gm[:, :n]
        fn = (lambda: torch.matmul(gmv.T, z)) if tr else (lambda: torch.matmul(gmv, z))
        ms = _t(fn, iters=300, warm=20)
        print(f"torch.matmul fp16 32768^2 x 512 T={tr}: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s (300 back to back)", flush=True)
    # burst: 3 launches after a 2 s pause (clocks not yet power-throttled)
    import time
    for pair in (0, 1, 2):
        lib.pb2_debug_gg_pair(pair)
        for sk in (False, True):
            torch.cuda.synchronize()
            time.sleep(2.0)
            ms = _t(lambda: ops.grad_gemm(gm, n, n, ld, z, transpose=False, out=out, accumulate=True, stream_k=sk), iters=3, warm=1)
            print(f"burst grad_gemm pair={pair} stream_k={sk}: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s", flush=True)
    lib.pb2_debug_gg_pair(-1)
    torch.cuda.synchronize()
    time.sleep(2.0)
    ms = _t(lambda: torch.matmul(gm[:, :n], z), iters=3, warm=1)
    print(f"burst torch.matmul: {ms:.3f} ms {2 * n * n * 512 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
