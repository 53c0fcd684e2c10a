"""Batched shard reader (b200ctr.data.BatchedShardReader) against the reference's own ShardedDataset + collate_sharded
(src/data/dataset.py:45-124) on a synthetic App.-D cache: bit-for-bit the same batch dicts.  CPU only; the reference classes
come from baseline/_ref when the copy is present (build container), otherwise from a restatement of the seven lines that
matter (stack rows, cast to the contract dtypes)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _reference_batch(manifest, gidx, train):
    if os.path.isdir(os.path.join(REF, "src", "data")):
        sys.path.insert(0, REF)
        from src.data.dataset import ShardedDataset, collate_sharded       # the real reference
        ds = ShardedDataset(manifest, np.asarray(gidx), train=train, cat_cols=[])
        return collate_sharded([ds[i] for i in range(len(gidx))])
    import json
    man = json.load(open(manifest))
    ends = np.array([m["end"] for m in man["shards"]])
    starts = np.array([m["start"] for m in man["shards"]])
    arrs = [{k: np.load(m[k]["path"], mmap_mode="r" if k != "ids" else None) for k in ("X_num", "X_mask", "X_cat", "seq", "y", "groups", "ids")}
            for m in man["shards"]]
    rows = []
    for g in gidx:
        s = int(np.searchsorted(ends, g, side="right"))
        rows.append((s, int(g - starts[s])))
    st = lambda k, dt: torch.from_numpy(np.ascontiguousarray(np.stack([arrs[s][k][i] for s, i in rows])).astype(dt))   # noqa: E731
    out = {"X_num": st("X_num", np.float32), "X_mask": st("X_mask", np.float32), "X_cat": st("X_cat", np.int64), "seq": st("seq", np.int64)}
    if train:
        out["y"] = torch.from_numpy(np.asarray([arrs[s]["y"][i] for s, i in rows], dtype=np.float32))
    else:
        out["ids"] = np.asarray([arrs[s]["ids"][i] for s, i in rows])
    out["groups"] = torch.from_numpy(np.asarray([arrs[s]["groups"][i] for s, i in rows], dtype=np.int64))
    return out


@pytest.fixture(scope="module")
def cache(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "toss-next-ctr-prediction_b200"))
    from b200ctr import configs, synth
    cfg = configs.dare_qnn_next()
    cfg["sequence"]["max_len"] = 48
    root = str(tmp_path_factory.mktemp("cache"))
    return synth.write_cache(root, cfg, n_train=2500, n_test=700, n_num=11, shard_rows=900, seed=3)


@pytest.mark.parametrize("train", [True, False])
def test_reader_reproduces_collate_sharded_bit_for_bit(cache, train):
    from b200ctr import BatchedShardReader
    manifest = cache[0] if train else cache[1]
    n = 2500 if train else 700
    rng = np.random.default_rng(1)
    index = rng.permutation(n)[: n - 37]                       # a K-fold style subset, in scrambled order, crossing shard borders
    rd = BatchedShardReader(manifest, index, train=train, batch_size=256, shuffle=False, n_threads=3)
    assert len(rd) == (len(index) + 255) // 256
    seen = 0
    for bi, b in enumerate(rd):
        g = index[bi * 256:(bi + 1) * 256]
        ref = _reference_batch(manifest, g, train)
        assert set(b.keys()) == set(ref.keys())
        for k, v in ref.items():
            if k == "ids":
                assert b[k].dtype == v.dtype and (b[k] == v).all()
            else:
                assert b[k].dtype == v.dtype and b[k].shape == v.shape and torch.equal(b[k], v), k
        seen += len(g)
    assert seen == len(index)


def test_reader_shuffle_is_a_seeded_permutation_per_epoch(cache):
    from b200ctr import BatchedShardReader
    index = np.arange(1000)
    rd = BatchedShardReader(cache[0], index, train=True, batch_size=128, shuffle=True, seed=9, drop_last=True)
    e0 = torch.cat([b["groups"] for b in rd])
    e1 = torch.cat([b["groups"] for b in rd])
    rd2 = BatchedShardReader(cache[0], index, train=True, batch_size=128, shuffle=True, seed=9, drop_last=True)
    f0 = torch.cat([b["groups"] for b in rd2])
    assert len(rd) == 7 and e0.shape[0] == 7 * 128
    assert torch.equal(e0, f0) and not torch.equal(e0, e1)     # same seed -> same epoch order; next epoch differs
