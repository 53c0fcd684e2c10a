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
