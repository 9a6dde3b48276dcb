"""CPU: the closed form behind ``rank_threshold`` in peppa_b200/csrc/sim.cu.

The rank kernels test ``s >= t`` instead of ``fl32(1 - s) < pd`` (pd = fl32(1 - s_pos), the comparison
argsort resolves in pig/metrics.py:8-12).  t is the smallest float with that property:
    q = pred(pd);  mid = (q + pd) / 2  (exact in double);  T = 1 - mid  (exact in double)
    fl32(1 - s) < pd  <=>  1 - s < mid, or 1 - s == mid and q has an even mantissa (round-half-even)
                      <=>  s > T, or s == T and q even
This restates it in numpy and brute-forces the equivalence on neighbourhoods of the threshold."""
import numpy as np


def rank_threshold(pd):
    pd = np.float32(pd)
    q = np.nextafter(pd, np.float32(-np.inf))
    mid = (np.float64(q) + np.float64(pd)) / 2
    T = 1.0 - mid
    tf = np.float32(T)
    if np.float64(tf) < T:                      # round up
        tf = np.nextafter(tf, np.float32(np.inf))
    if np.float64(tf) == T and (q.view(np.uint32) & 1):
        tf = np.nextafter(tf, np.float32(np.inf))
    return tf


def test_threshold_is_exact_on_neighbourhoods():
    rng = np.random.default_rng(0)
    one = np.float32(1.0)
    checked = 0
    for trial in range(4000):
        mode = trial % 4
        s_pos = np.float32([rng.uniform(-1, 1), rng.normal() * 1e-3, rng.uniform(0.9, 1.0), rng.normal() * 1e-6][mode])
        pd = one - s_pos
        t = rank_threshold(pd)
        s = t
        for _ in range(24):
            s = np.nextafter(s, np.float32(-np.inf))
        for _ in range(48):
            assert ((one - s) < pd) == (s >= t), (s_pos, s, t)
            s = np.nextafter(s, np.float32(np.inf))
            checked += 1
    assert checked == 4000 * 48


def test_the_positive_itself_is_never_closer():
    rng = np.random.default_rng(1)
    for _ in range(2000):
        s_pos = np.float32(rng.uniform(-1, 1))
        assert not (s_pos >= rank_threshold(np.float32(1.0) - s_pos))
