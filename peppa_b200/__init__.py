"""peppa_b200 -- B200-native drop-in for the contrastive-scoring hot path of gchrupala/peppa.

``peppa_b200.loss``, ``peppa_b200.metrics``, ``peppa_b200.triplet`` and ``peppa_b200.util`` keep
the call signatures of ``pig.loss``, ``pig.metrics``, ``pig.triplet`` (scoring half) and
``pig.util.cosine_matrix``; they call hand-written sm_100a CUDA kernels through the C ABI in
``include/peppa_b200.h``.  ``install()`` aliases them over an importable ``pig`` package so the
reference's ``run.py`` / ``evaluate.py`` pick them up unchanged.
"""
from __future__ import annotations

__version__ = "0.1.0"


def install(pig_package=None):
    """Replace the hot-path names of an imported ``pig`` package with the B200 implementations."""
    import importlib
    import sys

    from . import loss, metrics, triplet, util
    pig = pig_package or importlib.import_module("pig")
    for name, mod in (("loss", loss), ("metrics", metrics)):
        sys.modules[f"{pig.__name__}.{name}"] = mod
        setattr(pig, name, mod)
    # pig.util / pig.triplet contain out-of-scope helpers too: patch only the hot-path names
    for target_name, src, names in (
            ("util", util, ["cosine_matrix"]),
            ("triplet", triplet, ["score_triplets", "comparative_score_triplets", "_triplets", "triplets", "pairs",
                                  "triplet_accuracy"])):
        try:
            target = importlib.import_module(f"{pig.__name__}.{target_name}")
        except Exception:  # noqa: BLE001 -- pig.triplet needs moviepy/lightning; nothing to patch then
            continue
        for n in names:
            setattr(target, n, getattr(src, n))
    return pig
