"""Drop-in for the scoring half of ``pig/triplet.py`` (lines 17-29 and 63-121).

The duration-matched sampler stays in Python and consumes ``random`` in exactly the reference's
order (seeded runs draw identical triplets); the arithmetic runs in the fused gather + cosine-gap
kernel, so ``audio[pos]``, ``video[pos]``, ``video[neg]`` are never materialised.
``TripletScorer`` (pig/triplet.py:31-61) is the encode-side caller and is out of scope: it needs the
dataset and Lightning; its ``_score`` is ``score_triplets`` below.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import torch

from . import ops
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


def _gather_scores(audio_b, video_b, pos_idx, neg_idx, discrete):
    dev = audio_b.device
    pos = pos_idx.to(device=dev, dtype=torch.int64)
    neg = neg_idx.to(device=dev, dtype=torch.int64)
    return ops.triplet_score(audio_b, video_b, video_b, ia=pos, ip=pos, in_=neg, discrete=discrete)


def _as_rows(x, device=None):
    dev = ops.require_cuda(device if device is not None else x.device)
    dt = x.dtype if x.dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.float32
    return x.detach().to(device=dev, dtype=dt).contiguous()


def comparative_score_triplets(video_set, audio_set, duration, n_samples=100):
    vids = [_as_rows(v) for v in video_set]
    auds = [_as_rows(a, device=vids[k].device).to(vids[k].dtype) for k, a in enumerate(audio_set)]
    success = [[] for i in range(len(video_set))]
    length = []
    for i in range(n_samples):
        pos_idx, neg_idx = zip(*_triplets(range(len(duration)), lambda idx: duration[idx]))
        pos_idx = torch.tensor(pos_idx)
        neg_idx = torch.tensor(neg_idx)
        for k in range(len(video_set)):
            acc = _gather_scores(auds[k], vids[k], pos_idx, neg_idx, discrete=False)
            success[k].append(acc.to(device=video_set[k].device, dtype=video_set[k].dtype))
        length.append(duration[pos_idx])
    return {'success': [torch.cat(success_i) for success_i in success],
            'duration': torch.cat(length)}


def score_triplets(video, audio, duration, n_samples=100):
    # pig/triplet.py:82-96 without the stray line :93 (a NameError at the reference's HEAD)
    vid = _as_rows(video)
    aud = _as_rows(audio, device=vid.device).to(vid.dtype)
    accuracy = []
    length = []
    for i in range(n_samples):
        pos_idx, neg_idx = zip(*_triplets(range(len(duration)), lambda idx: duration[idx]))
        pos_idx = torch.tensor(pos_idx)
        neg_idx = torch.tensor(neg_idx)
        acc = _gather_scores(aud, vid, pos_idx, neg_idx, discrete=True)
        accuracy.append(acc.to(video.dtype).mean())
        length.append(duration[pos_idx])
    return {'accuracy': torch.stack(accuracy).cpu() if accuracy else torch.tensor([]),
            'duration': torch.cat(length)}


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
