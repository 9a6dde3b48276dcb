"""A/B of two library builds inside the real gallery step (hinge loss fwd + bwd + recall, 131072 clips = 4 x 4
gradient-matrix blocks, the kernels alternating as in the 2^20 step):
    python tools/ab_gallery.py [tools/ab/lib_X.so] [block]      (run alternately inside ONE gpurun call)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from peppa_b200 import _cabi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _cabi.LIB_PATH = os.path.abspath(args.pop(0))
    block = int(args[0]) if args else 32768          # gradient-matrix block edge
    from bench import synth_embeddings
    from gpu_probe import _t
    from peppa_b200.gallery import GalleryStep
    n = 131072
    dev = torch.device("cuda", 0)
    a, v = synth_embeddings(n, 666, dev)
    step = GalleryStep(n, 512, device=dev, block=block)
    out = step.run(a, v)
    for rep in range(3):
        ms = _t(lambda: step.run(a, v), iters=8, warm=2)
        print(f"{os.path.basename(_cabi.LIB_PATH)} block {block} gallery {n}: {ms:.2f} ms/step  {n * n / ms / 1e6:.1f} Gpairs/s  "
              f"loss {out['loss'].item():.6f}", flush=True)


if __name__ == "__main__":
    main()
