"""Synthetic Toss-shaped batches honouring the collate_sharded contract (reference src/data/dataset.py:98-124):
X_num f32 (B,Fn), X_mask f32 (B,Fm), X_cat i64 (B,Fc), seq i64 (B,L) left-padded with pad_id, y f32 (B).
Distributions follow SURVEY.md section 8d: N(0,1) numerics with 5% NaN->median(0)+mask, Zipf(1.1) categoricals mapped
through a fixed permutation of the bucket range (stands in for the unpinned Polars hash), lognormal(5.3, 0.9) sequence
lengths with 1% empty and 1% shorter than K, Zipf(1.05) tokens over [1, 1_000_000], CTR 1.9%.
Pure numpy (host); the caller pins / uploads.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

TOKEN_RANGE = 1_000_000


class ToySchema:
    def __init__(self, cfg: dict, n_num: int, n_mask: int, seed: int = 777):
        self.cfg, self.n_num, self.n_mask = cfg, n_num, n_mask
        d = cfg["data"]
        self.cols = d["cat_cols"]
        margin = d.get("hash_buckets_margin", 0)
        self.cards = [d["hash_buckets"].get(c, 1000003) + margin for c in self.cols]
        rng = np.random.default_rng(seed)
        self.perm, self.cdf = [], []
        for c, card in zip(self.cols, self.cards):
            b = d["hash_buckets"].get(c, 1000003)
            n_unique = int(min(card, max(2, b - 1002 if b > 1002 else b // 2)))
            n_unique = min(n_unique, 200_000)
            w = 1.0 / np.arange(1, n_unique + 1) ** 1.1
            self.cdf.append(np.cumsum(w / w.sum()))
            self.perm.append(rng.choice(card, size=n_unique, replace=False).astype(np.int64))
        w = 1.0 / np.arange(1, TOKEN_RANGE + 1) ** 1.05
        self.tok_cdf = np.cumsum(w / w.sum())
        self.tok_perm = rng.permutation(TOKEN_RANGE).astype(np.int64) + 1       # tokens 1..1e6, popularity shuffled
        self.L = int(cfg["sequence"]["max_len"])
        self.K = int(cfg["sequence"]["top_k"])
        self.pad = int(cfg["sequence"]["pad_id"])

    def batch(self, B: int, seed: int) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
        rng = np.random.default_rng(seed)
        X_num = rng.standard_normal((B, self.n_num), dtype=np.float32)
        nan = rng.random((B, self.n_num)) < 0.05
        X_num[nan] = 0.0                                          # imputed to the column median
        X_mask = np.zeros((B, self.n_mask), dtype=np.float32)
        X_mask[:, :min(self.n_num, self.n_mask)] = nan[:, :min(self.n_num, self.n_mask)]
        X_cat = np.empty((B, len(self.cols)), dtype=np.int64)
        for i in range(len(self.cols)):
            r = np.searchsorted(self.cdf[i], rng.random(B), side="right")
            X_cat[:, i] = self.perm[i][np.minimum(r, len(self.perm[i]) - 1)]
        L, K = self.L, self.K
        lens = np.minimum(L, np.floor(rng.lognormal(5.3, 0.9, size=B))).astype(np.int64)
        u = rng.random(B)
        lens[u < 0.01] = 0
        short = (u >= 0.01) & (u < 0.02)
        lens[short] = rng.integers(1, max(2, K), size=int(short.sum()))
        lens = np.minimum(lens, L)
        tot = int(lens.sum())
        r = np.searchsorted(self.tok_cdf, rng.random(tot), side="right")
        toks = self.tok_perm[np.minimum(r, TOKEN_RANGE - 1)]
        seq = np.full((B, L), self.pad, dtype=np.int64)
        # right-aligned (left-padded) rows: build_cache_v1.py:149-156
        col = np.arange(L)[None, :]
        valid = col >= (L - lens)[:, None]
        seq[valid] = toks
        y = (rng.random(B) < 0.019).astype(np.float32)
        return {"X_num": X_num, "X_mask": X_mask, "X_cat": X_cat, "seq": seq}, y


def planted_labels(b: Dict[str, np.ndarray], cols, rng: np.random.Generator, query_key: str = "inventory_id") -> np.ndarray:
    """Labels with a learnable signal (the default `batch()` labels are pure Bernoulli(0.019) noise, useless for AP / WLL):
    a linear + one pairwise numeric term, a per-bucket effect of the query-key categorical, and a target-aware sequence
    term -- the share of the 64 most recent tokens whose id is congruent to the query-key bucket mod 8 -- which is what
    DARE's query . key retrieval can pick up.  Base rate ~ 7 %."""
    x = b["X_num"]
    qi = cols.index(query_key) if query_key in cols else 0
    inv = b["X_cat"][:, qi]
    eff = np.array([0.9, -0.6, 0.3, -0.2, 0.7, -0.8, 0.1, 0.4])[inv % 8]
    recent = b["seq"][:, -64:]
    nz = recent != 0
    match = ((recent % 8) == (inv % 8)[:, None]) & nz
    frac = match.sum(1) / np.maximum(1, nz.sum(1))
    z = -3.3 + 0.9 * x[:, 0] - 0.6 * x[:, 1] + 0.5 * x[:, 2] * x[:, 3] + eff + 6.0 * (frac - 0.125)
    return (rng.random(x.shape[0]) < 1.0 / (1.0 + np.exp(-z))).astype(np.float32)


def write_cache(root: str, cfg: dict, n_train: int, n_test: int, n_num: int, shard_rows: int = 50_000, seed: int = 777,
                planted: bool = True) -> Tuple[str, str]:
    """Write a synthetic cache in the reference's on-disk format (build_cache_v1.py:158-177,223-306):
    <root>/{train,test}/shard_%03d/{X_num f32, X_mask u8, X_cat i32, seq i32, y i8, groups i64, ids <U64}.npy + manifest.json.
    Returns (manifest_train, manifest_test).  Deterministic in (cfg, sizes, seed)."""
    import json
    import os
    schema = ToySchema(cfg, n_num, n_num, seed=seed)
    cols = cfg["data"]["cat_cols"]
    qk = cfg["sequence"].get("query_key", cols[0])
    out = []
    for split, n_rows, is_train in (("train", n_train, True), ("test", n_test, False)):
        shards, start, si = [], 0, 0
        while start < n_rows:
            rows = min(shard_rows, n_rows - start)
            b, y = schema.batch(rows, seed=seed + 1000 * (1 + is_train) + si)
            if planted:
                y = planted_labels(b, cols, np.random.default_rng(seed + 7 + si), qk)
            d = os.path.join(root, split, f"shard_{si:03d}")
            os.makedirs(d, exist_ok=True)
            inv = b["X_cat"][:, cols.index(qk) if qk in cols else 0]
            dow = b["X_cat"][:, cols.index("day_of_week")] if "day_of_week" in cols else 0
            arrs = {"X_num": b["X_num"].astype(np.float32), "X_mask": b["X_mask"].astype(np.uint8),
                    "X_cat": b["X_cat"].astype(np.int32), "seq": b["seq"].astype(np.int32),
                    "y": (y if is_train else np.zeros_like(y)).astype(np.int8),
                    "groups": (inv * 7 + dow).astype(np.int64) % (2 ** 31 - 1),
                    "ids": np.array([f"{'TRAIN' if is_train else 'TEST'}_{start + i:07d}" for i in range(rows)], dtype="<U64")}
            meta = {"rows": rows, "index": si, "start": start, "end": start + rows}
            for k, a in arrs.items():
                p = os.path.join(d, f"{k}.npy")
                np.save(p, a)
                meta[k] = {"path": p, "shape": list(a.shape), "dtype": str(a.dtype)}
            shards.append(meta)
            start += rows
            si += 1
        man = {"parquet": "synthetic", "is_train": is_train, "rows": n_rows, "num_cols": [f"num_{i}" for i in range(n_num)],
               "cat_cols": list(cols), "group_key": qk, "seq_col": "seq", "shards": shards}
        mp = os.path.join(root, split, "manifest.json")
        with open(mp, "w") as fh:
            json.dump(man, fh)
        out.append(mp)
    return out[0], out[1]
