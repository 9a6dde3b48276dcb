"""Where the batch-1024 training step's time goes INSIDE its chained launches (BASELINE config 2).

    python tools/timeline_train1024.py [n]

The measurement build stops pb2_hinge_step after its first k kernels (pb2_debug_step_stages); each truncated step is
captured as a CUDA graph and replayed 2000 times back to back, so the differences between k and k - 1 are what each
kernel adds to the chain with programmatic dependent launch in place (ncu's per-launch times are cold and serialised).
Also: the full step followed by the backward's scale launch, and the public module with autograd (+ ones_like).
"""
import os
import sys

os.environ["PEPPA_B200_NO_FAST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from peppa_b200 import _cabi
    _cabi.use_measurement_library()
    from bench import synth_embeddings, timed
    from peppa_b200 import ops
    from peppa_b200.loss import TripletLoss
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda", 0)
    a, v = synth_embeddings(n, 666, dev)
    one = torch.ones((), dtype=torch.float32, device=dev)

    def sync():
        torch.cuda.synchronize()

    def graph_of(fn):
        for _ in range(3):
            fn()
        sync()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            keep = fn()
        return g, keep

    names = {1: "hinge_prep", 2: "+ similarity pass", 3: "+ gradient products", 4: "+ finish (whole pb2_hinge_step)"}
    prev = 0.0
    for k in (1, 2, 3, 4):
        _cabi.lib().pb2_debug_step_stages(k)
        g, keep = graph_of(lambda: ops.hinge_step(v, a, 0.2))
        us = min(timed(g.replay, 2000, 200, sync) for _ in range(3)) * 1e3
        print(f"n={n} stages 1..{k} {names[k]:32s}: {us:6.2f} us/replay  (+{us - prev:5.2f})", flush=True)
        prev = us
    _cabi.lib().pb2_debug_step_stages(4)

    def with_scale():
        loss, grads = ops.hinge_step(v, a, 0.2)
        return ops.scale_pair(grads[0], grads[1], one, torch.bfloat16)

    g, keep = graph_of(with_scale)
    us = min(timed(g.replay, 2000, 200, sync) for _ in range(3)) * 1e3
    print(f"n={n} pb2_hinge_step + pb2_scale_pair (4 + 1 launches)    : {us:6.2f} us/replay  (+{us - prev:5.2f})", flush=True)
    # the autograd pair the public module uses: forward = prep, pass, products (+ the loss fold in their grid)
    g, keep = graph_of(lambda: ops.hinge_forward(v, a, 0.2))
    fwd = min(timed(g.replay, 2000, 200, sync) for _ in range(3)) * 1e3
    print(f"n={n} pb2_hinge_forward (3 launches)                       : {fwd:6.2f} us/replay", flush=True)

    def pair():
        loss, state = ops.hinge_forward(v, a, 0.2)
        return ops.hinge_backward(state, v, a, one, torch.bfloat16)

    g, keep = graph_of(pair)
    us = min(timed(g.replay, 2000, 200, sync) for _ in range(3)) * 1e3
    print(f"n={n} pb2_hinge_forward + pb2_hinge_backward (3 + 1)       : {us:6.2f} us/replay  (+{us - fwd:5.2f})", flush=True)
    prev = us
    mod = TripletLoss(0.2)
    vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)

    def step():
        vv.grad = None
        aa.grad = None
        loss = mod(vv, aa)
        loss.backward()
        return loss

    g, keep = graph_of(step)
    us = min(timed(g.replay, 2000, 200, sync) for _ in range(3)) * 1e3
    print(f"n={n} public module + autograd (ones_like fill)              : {us:6.2f} us/replay  (+{us - prev:5.2f})", flush=True)


if __name__ == "__main__":
    main()
