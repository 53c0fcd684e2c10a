"""GPU parity of the on-GPU cache builder (b200ctr/cache_builder.py + csrc/cache_build.cu) against the CPU oracle
(oracle/cache_oracle.py): bit-exact shard arrays, shard cutting, manifest, and the edge cases the reference's rules cover
(empty / null / over-long sequences, double commas, null categoricals, all-missing numeric columns, +-inf)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cache_oracle as co

pytestmark = pytest.mark.gpu

L = 24


def _table(rows, seed):
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    seqs = []
    for i in range(rows):
        mode = i % 9
        if mode == 0:
            seqs.append("")
        elif mode == 1:
            seqs.append(None)
        else:
            n = int(rng.integers(1, 3 * L)) if mode != 2 else L
            toks = [str(int(t)) for t in rng.integers(0, 2_000_000_000 if mode == 3 else 10_000_000, n)]
            s = ",".join(toks)
            if mode == 4:
                s = s.replace(",", ",,", 2) + ","
            seqs.append(s)
    f32 = rng.standard_normal(rows).astype(np.float32)
    f32[rng.random(rows) < 0.1] = np.nan
    f64 = rng.standard_normal(rows) * 1e3
    f64[:3] = [np.inf, -np.inf, np.nan]
    i64 = rng.integers(-5, 50, rows)
    cat_s = [None if rng.random() < 0.05 else f"v{int(v)}" for v in rng.integers(0, 40, rows)]
    cat_i = [None if rng.random() < 0.05 else int(v) for v in rng.integers(-3, 1000, rows)]
    return pa.table({
        "ID": [f"TEST_{i:07d}" for i in range(rows)],
        "clicked": pa.array((rng.random(rows) < 0.02).astype(np.int64)),
        "seq": pa.array(seqs, type=pa.string()),
        "feat_a_1": pa.array(f32, from_pandas=False),
        "feat_a_2": pa.array(f64),
        "feat_b_1": pa.array(i64, mask=rng.random(rows) < 0.2),
        "feat_b_2": pa.array([None] * rows, type=pa.float32()),
        "gender": pa.array(cat_s, type=pa.string()),
        "inventory_id": pa.array(cat_i, type=pa.int64()),
        "hour": pa.array(rng.integers(0, 24, rows).astype(np.int32)),
    })


def _oracle(tbl, cat_cols, buckets, margin, num_cols):
    cols = {c: tbl.column(c).to_pylist() for c in tbl.column_names}
    x = np.stack([np.array([np.nan if v is None else v for v in cols[c]], dtype=np.float64).astype(np.float32) for c in num_cols], axis=1)
    med = [co.median_ignore_missing(x[:, j]) for j in range(len(num_cols))]
    X, M = co.numeric_block(x, med)
    xcat = np.stack([co.hash_column(cols[c], buckets.get(c, 1000003) + margin) for c in cat_cols], axis=1).astype(np.int32)
    return {"X_num": X, "X_mask": M, "X_cat": xcat, "seq": co.parse_seq(cols["seq"], L, 0),
            "y": np.asarray(cols["clicked"], dtype=np.int8), "groups": co.hash_column(cols["inventory_id"], 2 ** 31 - 1),
            "ids": np.asarray(cols["ID"], dtype="U64")}, med


def test_build_sharded_cache_matches_oracle(tmp_path):
    import pyarrow.parquet as pq
    from b200ctr.cache_builder import build_sharded_cache
    rows = 5003
    tbl = _table(rows, 1)
    path = str(tmp_path / "train.parquet")
    pq.write_table(tbl, path, row_group_size=700)
    cat_cols, buckets = ["gender", "inventory_id", "hour", "missing_col"], {"gender": 1009, "inventory_id": 2000003}
    man = build_sharded_cache(path, str(tmp_path / "cache"), is_train=True, target_col="clicked", seq_col="seq", cat_cols=cat_cols,
                              hash_buckets=buckets, hash_buckets_margin=500, num_patterns=["feat_*"], max_len=L, pad_id=0,
                              group_key="inventory_id", shard_rows=1500, batch_rows=640)
    m = json.load(open(man))
    assert m["rows"] == rows and [s["rows"] for s in m["shards"]] == [1500, 1500, 1500, 503]
    assert m["num_cols"] == ["feat_a_1", "feat_a_2", "feat_b_1", "feat_b_2"]
    want, med = _oracle(tbl, cat_cols[:3], buckets, 500, m["num_cols"])
    assert [m["medians"][c] for c in m["num_cols"]] == [float(np.float32(v)) for v in med]
    for s in m["shards"]:
        assert s["end"] - s["start"] == s["rows"]
        for k, dt in (("X_num", np.float32), ("X_mask", np.uint8), ("X_cat", np.int32), ("seq", np.int32), ("y", np.int8),
                      ("groups", np.int64)):
            a = np.load(s[k]["path"])
            assert a.dtype == dt and list(a.shape) == s[k]["shape"]
            w = want[k][s["start"]:s["end"]]
            if k == "X_cat":
                assert np.array_equal(a[:, :3], w) and not a[:, 3].any()       # absent column -> zeros (build_cache_v1.py:130-131)
            else:
                assert np.array_equal(a, w), k
        assert np.array_equal(np.load(s["ids"]["path"]), want["ids"][s["start"]:s["end"]])
    # the shards feed the batched reader of the hot path unchanged
    from b200ctr.data import BatchedShardReader
    idx = np.array([0, 1499, 1500, 5002])
    batch = next(iter(BatchedShardReader(man, idx, train=True, batch_size=4)))
    assert np.array_equal(batch["seq"].numpy().astype(np.int32), want["seq"][idx])
    assert np.array_equal(batch["X_cat"].numpy()[:, :3].astype(np.int32), want["X_cat"][idx])
    assert np.array_equal(batch["X_num"].numpy(), want["X_num"][idx])


def test_parse_seq_kernel_long_rows_and_rejects():
    import pyarrow as pa
    from b200ctr.cache_builder import CacheKernels
    K = CacheKernels("cuda:0")
    rng = np.random.default_rng(3)
    rows = []
    for n in (1, 31, 32, 33, 400, 1000, 0):
        rows.append(",".join(str(int(t)) for t in rng.integers(0, 10 ** int(rng.integers(1, 10)), n)))
    for max_len in (1, 7, 400):
        got = K.parse_seq(pa.array(rows, type=pa.large_string()), max_len, 3).cpu().numpy()
        assert np.array_equal(got, co.parse_seq(rows, max_len, 3))
    with pytest.raises(ValueError):
        K.parse_seq(pa.array(["1,2,a"]), 4, 0)
    with pytest.raises(ValueError):
        K.parse_seq(pa.array(["99999999999"]), 4, 0)


def test_median_radix_select_matches_sort():
    from b200ctr.cache_builder import CacheKernels
    K = CacheKernels("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 20001, generator=g)
    x[1] = x[1].abs() * -1e-30                       # negative denormal-range values
    x[2, ::3] = float("nan")
    x[3] = float("nan")
    x[4, :10] = torch.tensor([float("inf")] * 5 + [float("-inf")] * 5)
    med = K.medians(x.cuda().contiguous(), x.shape[1])
    want = np.array([co.median_ignore_missing(x[j].numpy()) for j in range(5)], dtype=np.float32)
    assert np.array_equal(med, want)
