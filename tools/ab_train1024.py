"""A/B timing of the batch-1024 training step (BASELINE config 2) between two builds of the library.

    python tools/ab_train1024.py [path/to/other_lib.so]

TripletLoss(0.2) forward + backward on 1024 x 512 bf16 rows through the public module (ctypes glue, so that the
library under test can be swapped), captured as ONE CUDA graph and replayed 2000 times back to back, three rounds;
also the eager step.  Run it alternately with and without the argument inside ONE gpurun call so both builds see the
same board and clocks.  Prints the loss and gradient checksums so the builds can be compared.
"""
import os
import sys

os.environ["PEPPA_B200_NO_FAST"] = "1"      # the C++ glue is linked against the product library: ctypes for both arms
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from peppa_b200 import _cabi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _cabi.MEASURE_LIB_PATH = os.path.abspath(args.pop(0))
        _cabi.use_measurement_library()
        name = os.path.basename(_cabi.MEASURE_LIB_PATH)
    else:
        name = os.path.basename(_cabi.LIB_PATH)
    n = int(args[0]) if args else 1024
    from bench import synth_embeddings, timed
    from peppa_b200.loss import TripletLoss
    dev = torch.device("cuda", 0)
    a, v = synth_embeddings(n, 666, dev)
    mod = TripletLoss(0.2)
    vv, aa = v.clone().requires_grad_(True), a.clone().requires_grad_(True)

    def step():
        vv.grad = None
        aa.grad = None
        loss = mod(vv, aa)
        loss.backward()
        return loss

    def sync():
        torch.cuda.synchronize()

    for _ in range(3):
        step()
    sync()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        loss = step()
    g.replay()
    sync()
    print(f"{name} n={n}: loss {loss.item():.9f} |dV| {vv.grad.float().abs().sum().item():.9e} "
          f"|dA| {aa.grad.float().abs().sum().item():.9e}", flush=True)
    for rep in range(3):
        ms = timed(g.replay, 2000, 200, sync)
        print(f"{name} n={n}: graph replay {ms * 1e3:.2f} us/step", flush=True)
    ms = timed(step, 500, 50, sync)
    print(f"{name} n={n}: eager (ctypes glue) {ms * 1e3:.2f} us/step", flush=True)


if __name__ == "__main__":
    main()
