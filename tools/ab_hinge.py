"""A/B timing of the fused similarity passes between two builds of the library.

    python tools/ab_hinge.py [path/to/other_lib.so] [what ...]     what in {hinge, rank, hinge_nog, lse}; default hinge

Loads the given library instead of the in-tree one (measurement only: the product always loads
peppa_b200/csrc/libpeppa_b200.so), runs the pass on a 32768 x 32768 block 200 times back to back and prints the
average; run it alternately with and without the argument inside ONE gpurun call so both builds see the same
board and clocks.  Also prints a checksum of the counts so the two builds can be compared.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from peppa_b200 import _cabi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _cabi.MEASURE_LIB_PATH = os.path.abspath(args.pop(0))      # a variant built by tools/build_variant.sh
    _cabi.use_measurement_library()      # the pb2_debug_* selectors live in the measurement build only
    pair = -1
    for a_ in list(args):
        if a_.startswith("pair="):          # pb2_debug_sim_pair: 0 = independent CTAs, 1 = CTA pairs, 2 = multicast clusters (default -1: per policy)
            pair = int(a_[5:])
            args.remove(a_)
    what = args or ["hinge"]
    from peppa_b200 import ops
    from gpu_probe import _t, emb
    name = os.path.basename(_cabi.MEASURE_LIB_PATH) + (f" pair={pair}" if pair >= 0 else "")
    _cabi.lib().pb2_debug_sim_pair(pair)
    n = 32768
    V, A = emb(n)
    rv, _ = ops.row_norms(V)
    ra, _ = ops.row_norms(A)
    diag, thr = ops.sim_diag(A, V, ra, rv)
    rc = torch.zeros(n, dtype=torch.int32, device="cuda")
    cc = torch.zeros(n, dtype=torch.int32, device="cuda")
    rk = torch.zeros(n, dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda")
    flops = 2 * n * n * 512
    if "hinge" in what:
        g, ld = ops.gmat_alloc(n, n, "cuda", torch.uint8)        # the product's one-byte gradient matrix
        ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk)
        torch.cuda.synchronize()
        chk = (int(rc.sum()), int(cc.sum()), int(rk.sum()), float(g.float().sum()))
        for rep in range(2):
            ms = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, g, ld, pos_thr=thr, rank=rk), iters=200, warm=20)
            print(f"{name} hinge+rank+G 32768^2: {ms:.4f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
        print("checksums (row_cnt, col_cnt, rank, G):", chk, flush=True)
    if "hinge_nog" in what:
        for rep in range(2):
            ms = _t(lambda: ops.sim_hinge(A, V, ra, rv, diag, diag, 0.2, rc, cc, None, 0, pos_thr=thr, rank=rk), iters=200, warm=20)
            print(f"{name} hinge+rank, no G 32768^2: {ms:.4f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
    if "lse" in what:
        bound = ops.logit_bound(A, V, 1.0 / 0.07)
        for rep in range(2):
            ms = _t(lambda: ops.sim_lse_both(A, V, bound, scale=1.0 / 0.07), iters=100, warm=10)
            print(f"{name} lse_both 32768^2: {ms:.4f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
            ms = _t(lambda: ops.sim_lse_rows(A, V, scale=1.0 / 0.07), iters=100, warm=10)
            print(f"{name} lse_rows (one direction) 32768^2: {ms:.4f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
    if "rank" in what:
        for rep in range(2):
            ms = _t(lambda: ops.sim_rank(A, V, ra, rv, thr, idx), iters=200, warm=20)
            print(f"{name} rank 32768^2: {ms:.4f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
