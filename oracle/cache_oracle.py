"""TEST INFRASTRUCTURE -- CPU restatement of the cache builder's per-batch arithmetic, plain Python /
numpy.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this; the product (b200ctr/cache_builder.py +
csrc/cache_build.cu) never does.

What is restated, with the reference lines it follows (src/data/build_cache_v1.py):
  * parse_seq        149-158  `[int(x) for x in str(st).split(",") if x][-max_len:]`, right-aligned, pad_id on the left
  * numeric_block    135-146  mask = isnan; NaN -> column median; np.nan_to_num(nan=0, posinf=1e6, neginf=-1e6)
  * medians          55-66    pl.col(c).median(): missing values ignored, mean of the two middle values for an even count,
                              0.0 when the column has no value
  * hash_value       123-132 / 97-113: `cast(pl.Utf8).fill_null("NA").hash(seed=2025) % buckets` -- the Polars hash is
                              unpinned and absent from this image, so the PINNED replacement of cache format v2
                              is restated instead: splitmix64_finalize(fnv1a64(utf8 bytes, basis ^ seed)).

Parity is pinned by known-answer vectors (the published FNV-1a 64 test vectors with seed 0 before the finalizer, hand-computed
splitmix64 outputs) in tests/test_cache_oracle.py; the reference itself cannot run here (`import polars` fails), which the
header of that test states as well.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

MASK64 = (1 << 64) - 1
FNV_BASIS = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
SEED = 2025


def fnv1a64(data: bytes, basis: int = FNV_BASIS) -> int:
    h = basis
    for b in data:
        h = ((h ^ b) * FNV_PRIME) & MASK64
    return h


def splitmix64_finalize(z: int) -> int:
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def hash_bytes(data: bytes, seed: int = SEED) -> int:
    return splitmix64_finalize(fnv1a64(data, FNV_BASIS ^ seed))


def hash_value(v, seed: int = SEED) -> int:
    """Hash of one categorical / group value: None -> "NA", integers -> decimal string, str -> its UTF-8 bytes."""
    if v is None:
        s = "NA"
    elif isinstance(v, (int, np.integer)):
        s = str(int(v))
    else:
        s = str(v)
    return hash_bytes(s.encode("utf-8"), seed)


def parse_seq(strings: Sequence[Optional[str]], max_len: int, pad_id: int) -> np.ndarray:
    out = np.full((len(strings), max_len), pad_id, dtype=np.int32)
    for i, st in enumerate(strings):
        if not st:
            continue
        toks = [int(x) for x in str(st).split(",") if x]
        toks = toks[-max_len:]
        if toks:
            out[i, -len(toks):] = np.asarray(toks, dtype=np.int32)
    return out


def median_ignore_missing(col: np.ndarray) -> float:
    v = np.asarray(col, dtype=np.float32)
    v = np.sort(v[~np.isnan(v)])
    if v.size == 0:
        return 0.0
    lo, hi = v[(v.size - 1) // 2], v[v.size // 2]
    m = float((np.float64(lo) + np.float64(hi)) * 0.5)
    return 0.0 if math.isnan(m) else m


def numeric_block(x: np.ndarray, med: Sequence[float]):
    """x: (rows, Fn) float32 with NaN for missing -> (X_num float32, X_mask uint8)."""
    X = np.array(x, dtype=np.float32, copy=True)
    mask = np.isnan(X).astype(np.uint8)
    for j in range(X.shape[1]):
        if mask[:, j].any():
            X[mask[:, j].astype(bool), j] = np.float32(med[j])
    np.nan_to_num(X, copy=False, nan=0.0, posinf=1e6, neginf=-1e6)
    return X, mask


def hash_column(values: Sequence, buckets: int, seed: int = SEED) -> np.ndarray:
    return np.asarray([hash_value(v, seed) % buckets for v in values], dtype=np.int64)
