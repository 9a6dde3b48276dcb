"""Drop-in for the scoring half of ``pig/triplet.py`` (lines 17-29 and 63-121).

The duration-matched sampler consumes ``random`` in exactly the reference's order (seeded runs draw identical
triplets and leave the generator in the same state), with the duration grouping hoisted out of the sample loop
 and the draws themselves replayed in C++ on a copy of the generator's state (csrc/sampler.cu); the arithmetic of *all* samples runs in one launch of the fused gather + cosine-gap
kernel, so ``audio[pos]``, ``video[pos]``, ``video[neg]`` are never materialised.
``TripletScorer`` (pig/triplet.py:31-61) is the encode-side caller: the class below keeps its interface and
imports the dataset (``pig.data``) and Lightning only when it is used -- the encoders and the data pipeline are
not part of this package; its ``_score`` is ``score_triplets`` below.
"""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass

import torch

from . import _cabi, ops
from .metrics import triplet_accuracy  # noqa: F401
from .util import grouped, shuffled


@dataclass
class Triplet:
    anchor: ...
    positive: ...
    negative: ...


@dataclass
class TripletBatch:
    anchor: ...
    positive: ...
    negative: ...


class TripletScorer:
    """pig/triplet.py:31-61 with the same constructor, ``_encode`` / ``_score`` / ``evaluate``.  The dataset and the
    trainer are the reference's own (``pig.data.PeppaPigDataset``, ``pig.data.grouped_loader``, ``pl.Trainer``),
    imported on use; only the scoring runs here."""

    def __init__(self, fragment_type, split=['val'], target_size=(180, 100), audio_sample_rate=44100, scrambled_video=False):
        from pig.data import PeppaPigDataset      # the reference's data pipeline (not part of this package)
        self.dataset = PeppaPigDataset(target_size=target_size, split=split, fragment_type=fragment_type, duration=None,
                                       audio_sample_rate=audio_sample_rate, scrambled_video=scrambled_video)

    def _encode(self, model, trainer, batch_size):
        """Encode the whole split once (batches grouped by audio duration, like pig/triplet.py:44-51)."""
        import pig.data as data
        loader = data.grouped_loader(self.dataset, lambda clip: clip.audio_duration, data.collate, batch_size=batch_size)
        encoded = list(trainer.predict(model, loader))
        self._audio = torch.cat([b.audio for b in encoded])
        self._video = torch.cat([b.video for b in encoded])
        self._duration = torch.cat([b.audio_duration for b in encoded])

    def _score(self, n_samples=100):
        return score_triplets(self._video, self._audio, self._duration, n_samples=n_samples)

    def evaluate(self, model, batch_size, n_samples=100, trainer=None):
        if trainer is None:                       # pig/triplet.py:58-59
            from pytorch_lightning import Trainer
            trainer = Trainer(gpus=1, logger=False)
        self._encode(model, trainer, batch_size)
        return self._score(n_samples=n_samples)


def _gather_scores(audio_b, video_b, pos_idx, neg_idx, discrete):
    dev = audio_b.device
    pos = pos_idx.to(device=dev, dtype=torch.int64)
    neg = neg_idx.to(device=dev, dtype=torch.int64)
    return ops.triplet_score(audio_b, video_b, video_b, ia=pos, ip=pos, in_=neg, discrete=discrete)


def _as_rows(x, device=None):
    dev = ops.require_cuda(device if device is not None else x.device)
    dt = x.dtype if x.dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.float32
    return x.detach().to(device=dev, dtype=dt).contiguous()


def _duration_keys(duration):
    """The sort / group keys of ``lambda idx: duration[idx]`` (pig/triplet.py:69,87) as plain Python numbers.
    The reference indexes the tensor once per comparison key (2 x len(duration) 0-d tensors per sample: 47 ms per
    sample at 1467 clips, more than all the arithmetic); the ordering and the equality classes of the float32
    values and of their exact Python doubles are the same."""
    if isinstance(duration, torch.Tensor):
        return duration.detach().cpu().tolist()
    if hasattr(duration, "tolist"):
        return duration.tolist()
    return [duration[i] for i in range(len(duration))]


_FAST_SAMPLE2 = None


def _sample2_is_two_randbelow():
    """``random.sample(p, 2)`` on a pair is, in CPython, ``j = _randbelow(2); _randbelow(1)`` -> ``[p[j], p[1 - j]]``
    with ``_randbelow(n)`` = rejection sampling on ``getrandbits(n.bit_length())``.  Checked once against the
    interpreter that is running (results and generator state over 256 draws on private generators); if it ever
    stops holding, the sampler below keeps calling ``random.sample`` itself."""
    global _FAST_SAMPLE2
    if _FAST_SAMPLE2 is None:
        a, b = random.Random(20211), random.Random(20211)
        ok = True
        for i in range(256):
            p = [2 * i, 2 * i + 1]
            r = b.getrandbits(2)
            while r >= 2:
                r = b.getrandbits(2)
            q = b.getrandbits(1)
            while q:
                q = b.getrandbits(1)
            ok = ok and a.sample(p, 2) == [p[r], p[1 - r]]
        _FAST_SAMPLE2 = bool(ok and a.getstate() == b.getstate())
    return _FAST_SAMPLE2


_NATIVE_SAMPLER = None


def _native_sampler():
    """The C++ replay of the sampler's draws (csrc/sampler.cu: ``pb2_host_random_doubles`` / ``pb2_host_sample_pairs``
    on a copy of the generator's MT19937 state), or None.  Taken only after it has reproduced THIS interpreter's
    ``random.random()`` and ``random.sample(pair, 2)`` -- values and final generator state -- on a private generator;
    if another interpreter ever draws differently, the Python loops below keep drawing."""
    global _NATIVE_SAMPLER
    if _NATIVE_SAMPLER is None:
        lib = _cabi.lib()           # no library, no sampler: raises like every other entry of this package
        a, b = random.Random(20211), random.Random(20211)
        st = b.getstate()
        ok = st[0] == 3 and len(st[1]) == 625 and _sample2_is_two_randbelow()
        if ok:
            mt = (C.c_uint32 * 625)(*st[1])
            out = (C.c_double * 1500)()
            ok = lib.pb2_host_random_doubles(mt, 1500, out) == 0 and list(out) == [a.random() for _ in range(1500)]
            items = (C.c_int64 * 7)(*range(7))
            start = (C.c_int64 * 4)(0, 1, 3, 7)
            pos, neg = (C.c_int64 * 9)(), (C.c_int64 * 9)()
            ok = ok and lib.pb2_host_sample_pairs(mt, items, start, 3, 3, pos, neg) == 0
            want = [a.sample(p, 2) for _ in range(3) for grp in ([0], [1, 2], [3, 4, 5, 6])
                    for p in pairs(sorted(grp, key=lambda _: a.random()))]
            ok = ok and [list(t) for t in zip(pos, neg)] == want and tuple(mt) == a.getstate()[1]
        _NATIVE_SAMPLER = ok
    return _NATIVE_SAMPLER


def _sampled_index_pairs(duration, n_samples):
    """``n_samples`` draws of ``zip(*_triplets(range(len(duration)), lambda idx: duration[idx]))`` as two int64
    tensors [n_samples, pairs].  Sorting and grouping by duration consume no randomness, so they are done once; the
    per-group ``shuffled`` / ``random.sample`` draws run in the reference's order (pig/triplet.py:99-104), one sample
    after the other, on the global ``random`` generator, so seeded runs draw identical triplets and leave the
    generator in the same state."""
    keys = _duration_keys(duration)
    groups = [list(items) for _, items in grouped(range(len(keys)), key=keys.__getitem__)]
    if n_samples > 0 and not any(len(items) > 1 for items in groups):
        pos_idx, neg_idx = zip(*[])      # no two clips share a duration: the reference's unpack raises ValueError
    per_sample = sum(len(items) // 2 for items in groups)
    if n_samples > 0 and _native_sampler():
        # the same draws on a copy of the global generator's state, handed back afterwards (csrc/sampler.cu)
        version, words, gauss = random.getstate()
        mt = (C.c_uint32 * 625)(*words)
        flat = [i for items in groups for i in items]
        starts = [0]
        for items in groups:
            starts.append(starts[-1] + len(items))
        pos_t = torch.empty(n_samples, per_sample, dtype=torch.int64)
        neg_t = torch.empty(n_samples, per_sample, dtype=torch.int64)
        _cabi.check(_cabi.lib().pb2_host_sample_pairs(mt, (C.c_int64 * len(flat))(*flat), (C.c_int64 * len(starts))(*starts),
                                                      len(groups), n_samples, pos_t.data_ptr(), neg_t.data_ptr()),
                    "host_sample_pairs")
        random.setstate((version, tuple(mt), gauss))
        return pos_t, neg_t
    pos, neg = [], []
    if _sample2_is_two_randbelow():
        getrandbits = random.getrandbits
        for i in range(n_samples):
            for items in groups:
                xs = shuffled(items)
                for k in range(0, len(xs) - 1, 2):          # pairs(xs), random.sample(pair, 2) unrolled
                    r = getrandbits(2)
                    while r >= 2:
                        r = getrandbits(2)
                    q = getrandbits(1)
                    while q:
                        q = getrandbits(1)
                    pos.append(xs[k + r])
                    neg.append(xs[k + 1 - r])
    else:
        sample = random.sample
        for i in range(n_samples):
            for items in groups:
                for p in pairs(shuffled(items)):
                    target, distractor = sample(p, 2)
                    pos.append(target)
                    neg.append(distractor)
    return (torch.tensor(pos, dtype=torch.int64).view(n_samples, per_sample),
            torch.tensor(neg, dtype=torch.int64).view(n_samples, per_sample))


def _durations_of(duration, pos):
    """``torch.cat([duration[p] for p in pos])`` (pig/triplet.py:77-79, :94-96): one gather instead of one per sample when
    ``duration`` is a 1-D tensor; anything else, and the empty case with its error, stays the reference's expression."""
    if isinstance(duration, torch.Tensor) and duration.dim() == 1 and pos.shape[0] > 0:
        return duration[pos.reshape(-1)]
    return torch.cat([duration[p] for p in pos])


def comparative_score_triplets(video_set, audio_set, duration, n_samples=100):
    vids = [_as_rows(v) for v in video_set]
    auds = [_as_rows(a, device=vids[k].device).to(vids[k].dtype) for k, a in enumerate(audio_set)]
    pos, neg = _sampled_index_pairs(duration, n_samples)
    # every sample's triplets of a model in one launch; torch.cat(success_i) of the reference is the flat order
    success = [_gather_scores(auds[k], vids[k], pos.reshape(-1), neg.reshape(-1), discrete=False)
               .to(device=video_set[k].device, dtype=video_set[k].dtype) for k in range(len(video_set))]
    return {'success': success,
            'duration': _durations_of(duration, pos)}


def score_triplets(video, audio, duration, n_samples=100):
    # pig/triplet.py:82-96 without the stray line :93 (a NameError at the reference's HEAD)
    vid = _as_rows(video)
    aud = _as_rows(audio, device=vid.device).to(vid.dtype)
    pos, neg = _sampled_index_pairs(duration, n_samples)
    if n_samples > 0:
        acc = _gather_scores(aud, vid, pos.reshape(-1), neg.reshape(-1), discrete=True)
        # per-sample means of values in {0, 0.5, 1}: the fp32 sums are exact in any order
        accuracy = acc.view(n_samples, -1).to(video.dtype).mean(dim=1).cpu()
    else:
        accuracy = torch.tensor([])
    return {'accuracy': accuracy,
            'duration': _durations_of(duration, pos)}


def _triplets(clips, criterion):
    for size, items in grouped(clips, key=criterion):
        paired = pairs(shuffled(items))
        for p in paired:
            target, distractor = random.sample(p, 2)
            yield (target, distractor)


def triplets(clips):
    """Generates triplets of (a, v1, v2) where a is an audio clip, v1
       matching video and v2 a distractor video, matched by duration."""
    items = _triplets(clips, lambda x: x.duration)
    for target, distractor in items:
        yield Triplet(anchor=target.audio, positive=target.video, negative=distractor.video)


def pairs(xs):
    return [xs[i:i + 2] for i in range(0, len(xs) - 1, 2)]
