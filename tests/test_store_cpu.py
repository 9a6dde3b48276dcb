"""CPU: the sharded embedding store -- write / manifest / memory-mapped reads / integrity /
rank-sharded loads -- and the reference's results-row bookkeeping.  No device needed."""
import json
import os

import numpy as np
import pytest
import torch

from peppa_b200 import store


def _emb(n, d=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g), torch.randn(n, d, generator=g), torch.rand(n, generator=g) * 4 + 1


def test_round_trip_across_ragged_shards(tmp_path):
    V, A, dur = _emb(1000)
    p = str(tmp_path / "st")
    with store.EmbeddingStoreWriter(p, 64, rows_per_shard=256, meta={"split": "val"}) as w:
        for a, b in [(0, 100), (100, 101), (101, 700), (700, 1000)]:      # appends do not align with shards
            w.append(V[a:b], A[a:b], dur[a:b])
    st = store.EmbeddingStore(p)
    assert len(st) == 1000 and st.dim == 64 and st.meta == {"split": "val"} and st.has_duration
    assert [s["rows"] for s in st.manifest["shards"]] == [256, 256, 256, 232]
    assert st.verify()
    assert torch.equal(st.read("video"), V.bfloat16()) and torch.equal(st.read("audio"), A.bfloat16())
    assert torch.equal(st.read("duration"), dur)
    for a, b in [(0, 0), (0, 1), (255, 257), (300, 900), (999, 1000)]:     # ranges crossing shard boundaries
        assert torch.equal(st.read("video", a, b), V[a:b].bfloat16())
        assert torch.equal(st.to_device("audio", a, b, "cpu"), A[a:b].bfloat16())
    with pytest.raises(IndexError):
        st.read("video", 0, 1001)
    # files are raw little-endian bf16: readable without this module
    raw = np.fromfile(os.path.join(p, "video-00001.bf16"), dtype="<u2").reshape(256, 64)
    assert np.array_equal(raw, V[256:512].bfloat16().view(torch.int16).numpy().view(np.uint16))


def test_rank_blocks_and_errors(tmp_path):
    V, A, _ = _emb(512, seed=1)
    p = str(tmp_path / "st")
    w = store.EmbeddingStoreWriter(p, 64, rows_per_shard=200)
    w.append(V, A)
    w.close()
    st = store.EmbeddingStore(p)
    assert not st.has_duration
    with pytest.raises(KeyError):
        st.read("duration")
    for world in (1, 2, 4):
        got_a = torch.cat([st.load_rank_rows(r, world, "cpu")[0] for r in range(world)])
        got_v = torch.cat([st.load_rank_rows(r, world, "cpu")[1] for r in range(world)])
        assert torch.equal(got_a, A.bfloat16()) and torch.equal(got_v, V.bfloat16())
    with pytest.raises(ValueError):
        st.load_rank_rows(0, 3, "cpu")
    with pytest.raises(FileExistsError):
        store.EmbeddingStoreWriter(p, 64)
    with pytest.raises(ValueError):
        store.EmbeddingStoreWriter(str(tmp_path / "bad"), 60)
    # a flipped byte is caught by verify()
    fn = os.path.join(p, "audio-00000.bf16")
    b = bytearray(open(fn, "rb").read())
    b[10] ^= 0xFF
    open(fn, "wb").write(bytes(b))
    assert not store.EmbeddingStore(p).verify()
    assert json.load(open(os.path.join(p, store.MANIFEST)))["format"] == store.FORMAT


def test_full_scores_rows_round_trip(tmp_path):
    """pig/evaluation.py:103-110,255-261: the row dictionaries and the bookkeeping keys full_run adds."""
    row = dict(fragment_type="narration", scrambled_video=False, triplet_acc=torch.rand(5),
               recall_fixed=torch.rand(4, 11, 100), recall_jitter=torch.rand(4, 11, 100))
    row["recall_at_10_fixed"] = row["recall_fixed"][:, 10, :]
    row["recall_at_10_jitter"] = row["recall_jitter"][:, 10, :]
    p = str(tmp_path / "results" / "full_scores_v68.pt")
    store.save_full_scores([row], p, version=68, checkpoint_path="x.ckpt", hparams_path="hparams.yaml")
    back = store.load_full_scores(p)
    assert len(back) == 1 and back[0]["version"] == 68 and back[0]["hparams_path"] == "hparams.yaml"
    assert set(back[0]) == {"fragment_type", "scrambled_video", "triplet_acc", "recall_fixed", "recall_jitter",
                            "recall_at_10_fixed", "recall_at_10_jitter", "version", "checkpoint_path", "hparams_path"}
    assert torch.equal(back[0]["recall_at_10_fixed"], row["recall_fixed"][:, 10, :])
