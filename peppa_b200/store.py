"""Sharded on-disk embedding store and the reference's results layout.

The reference re-encodes every clip for every evaluation (pig/evaluation.py:131-163 runs
``trainer.predict`` and concatenates ``batch.video`` / ``batch.audio``) and keeps only the final
score tensors (``torch.save(add_condition(rows), "results/full_scores_v{version}.pt")``,
pig/evaluation.py:103-110,261).  At gallery scale (2^20 clips = 1 GiB of bf16 per modality) the
embeddings themselves are the asset: this module stores them once, row-sharded, in the exact layout the
scoring kernels read -- row-major bf16, 16-byte aligned rows -- so that a gallery can be scored on 1..P
GPUs without touching an encoder:

* ``EmbeddingStoreWriter`` / ``EmbeddingStore``: a directory with ``manifest.json`` and raw shard files
  ``{video,audio}-00000.bf16`` (little-endian bf16, ``rows x dim``; a clip's duration rides along as fp32).
  Reading is ``np.memmap`` -> pinned staging -> ``cudaMemcpyAsync``; ``load_rank_rows`` hands rank r its
  contiguous row block, which is what ``GalleryStep.run`` consumes.
* ``score_store``: loss + recall@1..N over a stored gallery through ``GalleryStep`` (1 GPU, or one process
  per GPU with an initialised ``torch.distributed`` group).
* ``evaluation_row`` / ``save_full_scores`` / ``load_full_scores``: the ``results/full_scores_v*.pt`` row
  dictionaries of pig/evaluation.py:103-110 built from stored embeddings with the fused metrics
  (``resampled_recall_at_1_to_n`` with size=100, n_samples=500, N=10 and ``score_triplets`` with
  n_samples=500, as pig/evaluation.py:159,172 and pig/triplet.py:55-61 call them).

Host-side Python like the code it stands in for; the arithmetic stays in the CUDA kernels.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

MANIFEST = "manifest.json"
FORMAT = "peppa_b200.embedding_store/1"
MODALITIES = ("video", "audio")


def _as_bf16_bits(x: torch.Tensor) -> np.ndarray:
    """[n, d] float tensor -> uint16 bf16 bit patterns on the host (round-to-nearest-even)."""
    if x.dim() != 2:
        raise ValueError(f"expected [N, D] embeddings, got {tuple(x.shape)}")
    return x.detach().to(device="cpu", dtype=torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


class EmbeddingStoreWriter:
    """Append-only writer: ``append(video, audio[, duration])`` any number of times, then ``close()``."""

    def __init__(self, path: str, dim: int, rows_per_shard: int = 1 << 18, meta: Optional[dict] = None):
        if dim <= 0 or dim % 8 != 0:
            raise ValueError("dim must be a positive multiple of 8 (16-byte rows)")
        if rows_per_shard <= 0:
            raise ValueError("rows_per_shard must be positive")
        os.makedirs(path, exist_ok=True)
        if os.path.exists(os.path.join(path, MANIFEST)):
            raise FileExistsError(f"{path} already holds an embedding store")
        self.path, self.dim, self.rows_per_shard = path, dim, rows_per_shard
        self.meta = dict(meta or {})
        self._buf: Dict[str, List[np.ndarray]] = {m: [] for m in MODALITIES}
        self._dur: List[np.ndarray] = []
        self._buffered = 0
        self._shards: List[dict] = []
        self._rows = 0
        self._has_duration: Optional[bool] = None
        self._closed = False

    def append(self, video: torch.Tensor, audio: torch.Tensor, duration: Optional[torch.Tensor] = None):
        if self._closed:
            raise RuntimeError("store already closed")
        if video.shape != audio.shape or video.shape[1] != self.dim:
            raise ValueError(f"video/audio must both be [n, {self.dim}]; got {tuple(video.shape)}, {tuple(audio.shape)}")
        if self._has_duration is None:
            self._has_duration = duration is not None
        if (duration is not None) != self._has_duration:
            raise ValueError("either every append carries durations or none does")
        self._buf["video"].append(_as_bf16_bits(video))
        self._buf["audio"].append(_as_bf16_bits(audio))
        if duration is not None:
            d = duration.detach().to("cpu", torch.float32).reshape(-1).numpy()
            if d.shape[0] != video.shape[0]:
                raise ValueError("one duration per clip")
            self._dur.append(d)
        self._buffered += video.shape[0]
        while self._buffered >= self.rows_per_shard:
            self._flush(self.rows_per_shard)

    def _take(self, chunks: List[np.ndarray], n: int) -> np.ndarray:
        out, need = [], n
        while need:
            head = chunks[0]
            if head.shape[0] <= need:
                out.append(head)
                need -= head.shape[0]
                chunks.pop(0)
            else:
                out.append(head[:need])
                chunks[0] = head[need:]
                need = 0
        return np.concatenate(out) if len(out) > 1 else out[0]

    def _flush(self, n: int):
        idx = len(self._shards)
        rec = {"rows": int(n), "first_row": int(self._rows), "files": {}, "sha256": {}}
        for m in MODALITIES:
            arr = np.ascontiguousarray(self._take(self._buf[m], n))
            name = f"{m}-{idx:05d}.bf16"
            arr.tofile(os.path.join(self.path, name))
            rec["files"][m] = name
            rec["sha256"][m] = hashlib.sha256(arr.tobytes()).hexdigest()
        if self._has_duration:
            d = np.ascontiguousarray(self._take(self._dur, n))
            name = f"duration-{idx:05d}.f32"
            d.tofile(os.path.join(self.path, name))
            rec["files"]["duration"] = name
        self._shards.append(rec)
        self._rows += n
        self._buffered -= n

    def close(self) -> str:
        if self._closed:
            return self.path
        if self._buffered:
            self._flush(self._buffered)
        manifest = {"format": FORMAT, "dim": self.dim, "dtype": "bf16", "byteorder": "little", "rows": self._rows,
                    "modalities": list(MODALITIES), "has_duration": bool(self._has_duration), "shards": self._shards,
                    "meta": self.meta}
        tmp = os.path.join(self.path, MANIFEST + ".tmp")
        with open(tmp, "w") as f:
            json.dump(manifest, f, indent=1)
        os.replace(tmp, os.path.join(self.path, MANIFEST))      # the manifest appears last, atomically
        self._closed = True
        return self.path

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None:
            self.close()


class EmbeddingStore:
    """Read side: memory-mapped shards, row-range reads, rank-sharded loads onto a GPU."""

    def __init__(self, path: str):
        with open(os.path.join(path, MANIFEST)) as f:
            self.manifest = json.load(f)
        if self.manifest.get("format") != FORMAT:
            raise ValueError(f"{path}: not a {FORMAT} store")
        self.path = path
        self.dim = int(self.manifest["dim"])
        self.rows = int(self.manifest["rows"])
        self.meta = self.manifest.get("meta", {})
        self.has_duration = bool(self.manifest.get("has_duration"))
        self._shards = self.manifest["shards"]
        self._maps: Dict[tuple, np.memmap] = {}

    def __len__(self):
        return self.rows

    def _map(self, modality: str, k: int) -> np.memmap:
        key = (modality, k)
        if key not in self._maps:
            rec = self._shards[k]
            fn = os.path.join(self.path, rec["files"][modality])
            if modality == "duration":
                self._maps[key] = np.memmap(fn, dtype=np.float32, mode="r", shape=(rec["rows"],))
            else:
                self._maps[key] = np.memmap(fn, dtype=np.uint16, mode="r", shape=(rec["rows"], self.dim))
        return self._maps[key]

    def verify(self) -> bool:
        """Recompute every shard's SHA-256 against the manifest."""
        for k, rec in enumerate(self._shards):
            for m in MODALITIES:
                if hashlib.sha256(np.ascontiguousarray(self._map(m, k)).tobytes()).hexdigest() != rec["sha256"][m]:
                    return False
        return True

    def _pieces(self, start: int, stop: int):
        """(shard index, first row inside the shard, rows) covering [start, stop)."""
        if not (0 <= start <= stop <= self.rows):
            raise IndexError(f"rows [{start}, {stop}) outside a store of {self.rows}")
        for k, rec in enumerate(self._shards):
            lo, hi = rec["first_row"], rec["first_row"] + rec["rows"]
            a, b = max(start, lo), min(stop, hi)
            if a < b:
                yield k, a - lo, b - a

    def read(self, modality: str, start: int = 0, stop: Optional[int] = None) -> torch.Tensor:
        """Rows [start, stop) of ``video`` / ``audio`` as a bf16 CPU tensor (``duration``: fp32)."""
        stop = self.rows if stop is None else stop
        if modality == "duration":
            if not self.has_duration:
                raise KeyError("this store carries no durations")
            out = np.empty(stop - start, dtype=np.float32)
        else:
            if modality not in MODALITIES:
                raise KeyError(modality)
            out = np.empty((stop - start, self.dim), dtype=np.uint16)
        at = 0
        for k, first, n in self._pieces(start, stop):
            out[at:at + n] = self._map(modality, k)[first:first + n]
            at += n
        t = torch.from_numpy(out)
        return t if modality == "duration" else t.view(torch.int16).view(torch.bfloat16)

    def to_device(self, modality: str, start: int, stop: int, device, chunk_rows: int = 1 << 16) -> torch.Tensor:
        """Rows [start, stop) on ``device`` as bf16: memmap -> two pinned staging buffers -> async copies, so
        the page-cache read of chunk k+1 overlaps the host-to-device copy of chunk k."""
        device = torch.device(device)
        out = torch.empty(stop - start, self.dim, dtype=torch.bfloat16, device=device)
        if stop == start:
            return out
        if device.type != "cuda":
            out.copy_(self.read(modality, start, stop))
            return out
        chunk_rows = max(1, min(chunk_rows, stop - start))
        stage = [torch.empty(chunk_rows, self.dim, dtype=torch.int16).pin_memory() for _ in range(2)]
        done = [None, None]
        stream = torch.cuda.current_stream(device)
        at, i = 0, 0
        for k, first, n in self._pieces(start, stop):
            src = self._map(modality, k)
            for c in range(0, n, chunk_rows):
                m = min(chunk_rows, n - c)
                b = i & 1
                if done[b] is not None:
                    done[b].synchronize()               # the staging buffer's previous copy has drained
                stage[b][:m].numpy().view(np.uint16)[...] = src[first + c:first + c + m]
                out[at:at + m].view(torch.int16).copy_(stage[b][:m], non_blocking=True)
                done[b] = torch.cuda.Event()
                done[b].record(stream)
                at += m
                i += 1
        stream.synchronize()
        return out

    def load_rank_rows(self, rank: int, world: int, device):
        """This rank's contiguous row block of both modalities (``GalleryStep`` shards rows evenly, so the
        store's row count must divide by ``world``).  Returns (audio, video) bf16 on ``device``."""
        if self.rows % world != 0:
            raise ValueError(f"{self.rows} rows do not divide over {world} ranks")
        nl = self.rows // world
        r0 = rank * nl
        return (self.to_device("audio", r0, r0 + nl, device), self.to_device("video", r0, r0 + nl, device))


def score_store(path: str, margin: float = 0.2, top_n: int = 10, with_grad: bool = False, device=None, group=None,
                loss: str = "hinge"):
    """Loss (+ recall@1..top_n for the hinge loss) of a stored gallery.  Single process: one GPU.  Under
    ``torch.distributed`` (one process per GPU, group initialised by the caller): rows sharded by rank."""
    import torch.distributed as dist

    from .gallery import GalleryStep
    store = EmbeddingStore(path)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    a_loc, v_loc = store.load_rank_rows(rank, world, device)
    step = GalleryStep(store.rows // world, store.dim, margin=margin, top_n=top_n, rank=rank, world=world, group=group,
                       device=device, with_grad=with_grad, loss=loss)
    return step.run(a_loc, v_loc)


# ------------------------------------------------------------------ results layout of pig/evaluation.py
def evaluation_row(fragment_type: str, scrambled_video: bool, fixed: str, jitter: str, n_samples: int = 500,
                   size: int = 100, N: int = 10, device=None) -> dict:
    """One row of ``full_score`` (pig/evaluation.py:78-110) from two stored embedding sets of the same clips
    (fixed-duration and jittered fragments): resampled recall@1..N for both, triplet accuracy from the
    fixed set's durations.  RNG use follows the reference: ``torch.randperm`` draws for the recall subsets,
    Python ``random`` for the duration-matched triplets (seed both with 666 like pig/evaluation.py:18-19)."""
    from . import metrics, triplet
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    out = {"fragment_type": fragment_type, "scrambled_video": scrambled_video}
    rec = {}
    for name, path in (("fixed", fixed), ("jitter", jitter)):
        st = EmbeddingStore(path)
        V = st.to_device("video", 0, st.rows, dev)
        A = st.to_device("audio", 0, st.rows, dev)
        rec[name] = metrics.resampled_recall_at_1_to_n(V, A, size=size, n_samples=n_samples, N=N)
        if name == "fixed":
            if not st.has_duration:
                raise ValueError("the fixed-duration store needs clip durations for the triplet score")
            dur = st.read("duration")
            out["triplet_acc"] = torch.as_tensor(triplet.score_triplets(V, A, dur, n_samples=n_samples)["accuracy"])
    out["recall_fixed"], out["recall_jitter"] = rec["fixed"], rec["jitter"]
    out["recall_at_10_fixed"] = rec["fixed"][:, 10, :] if N >= 10 else None
    out["recall_at_10_jitter"] = rec["jitter"][:, 10, :] if N >= 10 else None
    return out


def save_full_scores(rows: Sequence[dict], path: str, version=None, checkpoint_path=None, hparams_path=None):
    """``torch.save`` of the row list, with the bookkeeping keys ``full_run`` adds (pig/evaluation.py:255-261)."""
    out = []
    for row in rows:
        r = dict(row)
        if version is not None:
            r.setdefault("version", version)
        if checkpoint_path is not None:
            r.setdefault("checkpoint_path", checkpoint_path)
        if hparams_path is not None:
            r.setdefault("hparams_path", hparams_path)
        out.append(r)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(out, path)
    return out


def load_full_scores(path: str):
    return torch.load(path, weights_only=False)
