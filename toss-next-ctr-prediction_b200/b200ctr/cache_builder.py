"""On-GPU cache builder: Parquet -> App. D shards, mirroring the reference's
`src/data/build_cache_v1.py::build_sharded_cache` (same keyword arguments, same shard / manifest layout, same two scans:
medians first, then the batches) with the per-row work on the device (csrc/cache_build.cu):

  * numeric columns are uploaded once into a resident column-major float32 matrix (nulls = NaN); exact column medians come
    from a 4-pass radix select over it; X_num / X_mask rows are cut from it per batch (NaN -> median, nan_to_num, transpose);
  * `seq` strings are parsed by one warp per row straight from the Arrow (bytes, offsets) buffers;
  * categorical columns and the group key are hashed from their Arrow buffers.

The reference hashes with polars.Series.hash (unpinned across Polars versions, not installed here), so this builder writes
cache format v2 with a pinned hash -- see include/b200ctr.h.  Everything else follows the reference bit for bit.
Parquet decoding itself is pyarrow's (host); `ids` and `y` never touch the device.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check

HASH_SEED = 2025
_DT = {"float": 0, "double": 1, "int32": 2, "int64": 3, "uint8": 4}
_WIDTH = {"float": 4, "double": 8, "int32": 4, "int64": 8, "uint8": 1}


def _match_patterns(cols: List[str], patterns: List[str]) -> List[str]:
    """build_cache_v1.py:11-16"""
    out: List[str] = []
    for p in patterns:
        rx = re.compile("^" + p.replace("*", ".*") + "$")
        out += [c for c in cols if rx.match(c)]
    return sorted(list(dict.fromkeys(out)))


def _p(t):
    return None if t is None else t.data_ptr()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _dev_bytes(buf, dev, nbytes=None) -> torch.Tensor:
    """One Arrow buffer -> uint8 device tensor (H2D); never empty so that its pointer is valid."""
    a = np.frombuffer(buf, dtype=np.uint8) if buf is not None else np.zeros(0, np.uint8)
    if nbytes is not None:
        a = a[:nbytes]
    if a.size == 0:
        return torch.zeros(8, dtype=torch.uint8, device=dev)
    return torch.from_numpy(a.copy()).to(dev)


def _rebased(arr):
    """A contiguous, offset-0 copy of a (possibly chunked / sliced) Arrow array."""
    import pyarrow as pa
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    if arr.offset != 0:
        arr = pa.concat_arrays([arr, arr.slice(0, 0)])
        if arr.offset != 0:
            arr = pa.array(arr.to_pylist(), type=arr.type)
    return arr


def _validity(arr, dev) -> Optional[torch.Tensor]:
    if arr.null_count == 0:
        return None
    return _dev_bytes(arr.buffers()[0], dev)


def _as_string_array(arr):
    import pyarrow as pa
    import pyarrow.compute as pc
    if pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type):
        return arr
    return _rebased(pc.cast(arr, pa.string()))


class CacheKernels:
    """Thin host wrappers of the cache_build.cu entry points, one Arrow column at a time."""

    def __init__(self, device=None):
        self.dev = torch.device(device if device is not None else "cuda:0")
        if self.dev.type != "cuda":
            raise _lib.B200CtrError("the cache builder runs on a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()

    def _strings(self, arr):
        import pyarrow as pa
        s = _as_string_array(_rebased(arr))
        n = len(s)
        off64 = pa.types.is_large_string(s.type)
        bufs = s.buffers()
        offs = _dev_bytes(bufs[1], self.dev, (n + 1) * (8 if off64 else 4))
        data = _dev_bytes(bufs[2], self.dev)
        return s, n, int(off64), offs, data, _validity(s, self.dev)

    def hash_column(self, arr, mod: int, out: torch.Tensor, stride: int, out32: bool) -> None:
        """out[r * stride] = hash(value r) % mod for one Arrow column (string or integer typed)."""
        import pyarrow as pa
        import pyarrow.compute as pc
        arr = _rebased(arr)
        n = len(arr)
        if n == 0:
            return
        if pa.types.is_integer(arr.type):
            a64 = _rebased(pc.cast(arr, pa.int64()))
            vals = _dev_bytes(a64.buffers()[1], self.dev, n * 8)
            val = _validity(a64, self.dev)
            check(self.lib.b200ctr_cache_hash_int64(_p(vals), _p(val), n, mod, HASH_SEED, _p(out), stride, int(out32), _stream(self.dev)))
            return
        _, n, off64, offs, data, val = self._strings(arr)
        check(self.lib.b200ctr_cache_hash_strings(_p(data), _p(offs), off64, _p(val), n, mod, HASH_SEED, _p(out), stride, int(out32),
                                                  _stream(self.dev)))

    def parse_seq(self, arr, max_len: int, pad_id: int) -> torch.Tensor:
        _, n, off64, offs, data, val = self._strings(arr)
        out = torch.empty((n, max_len), dtype=torch.int32, device=self.dev)
        if n == 0:
            return out
        bad = torch.zeros(1, dtype=torch.int64, device=self.dev)
        check(self.lib.b200ctr_cache_parse_seq(_p(data), _p(offs), off64, _p(val), n, max_len, pad_id, _p(out), _p(bad),
                                               _stream(self.dev)))
        nb = int(bad.item())
        if nb:
            raise ValueError(f"seq column: {nb} bytes / tokens that int() would reject (only digits and ',' are accepted, "
                             f"tokens must fit int32)")
        return out

    def col_to_f32(self, arr, out_col: torch.Tensor) -> None:
        """One numeric Arrow column -> float32 (nulls = NaN) written into out_col (a contiguous 1-D device view)."""
        import pyarrow as pa
        import pyarrow.compute as pc
        arr = _rebased(arr)
        n = len(arr)
        if n == 0:
            return
        t = str(arr.type)
        if t not in _DT:
            arr = _rebased(pc.cast(arr, pa.float64()))
            t = "double"
        vals = _dev_bytes(arr.buffers()[1], self.dev, n * _WIDTH[t])
        val = _validity(arr, self.dev)
        check(self.lib.b200ctr_cache_col_to_f32(_p(vals), _DT[t], _p(val), n, _p(out_col), _stream(self.dev)))

    def medians(self, x: torch.Tensor, rows: int) -> np.ndarray:
        """Exact medians (missing ignored, mean of the two middle values) of the rows of x (Fn, ld) by radix select: two
        rank queries per column, four 8-bit digit passes, the 256-bin histograms come back to the host between passes."""
        Fn, ld = x.shape
        if Fn == 0:
            return np.zeros((0,), np.float32)
        nq = 2 * Fn
        q_col = torch.arange(Fn, dtype=torch.int32, device=self.dev).repeat_interleave(2).contiguous()
        prefix = np.zeros(nq, dtype=np.uint32)
        k = np.zeros(nq, dtype=np.int64)
        n_valid = None
        for shift in (24, 16, 8, 0):
            hist = torch.zeros((nq, 256), dtype=torch.int64, device=self.dev)
            qp = torch.from_numpy(prefix.view(np.int32).copy()).to(self.dev)
            check(self.lib.b200ctr_cache_select_hist(_p(x), ld, rows, _p(q_col), _p(qp), nq, shift, _p(hist), _stream(self.dev)))
            h = hist.cpu().numpy()
            if n_valid is None:
                n_valid = h.sum(axis=1)
                k[0::2] = np.maximum(n_valid[0::2] - 1, 0) // 2
                k[1::2] = n_valid[1::2] // 2
            cum = np.cumsum(h, axis=1)
            for q in range(nq):
                if n_valid[q] == 0:
                    continue
                b = min(int(np.searchsorted(cum[q], k[q], side="right")), 255)
                k[q] -= cum[q][b - 1] if b > 0 else 0
                prefix[q] |= np.uint32(b << shift)
        neg = (prefix & np.uint32(0x80000000)) == 0          # keys below 2^31 are negative floats (stored complemented)
        bits = np.where(neg, ~prefix, prefix & np.uint32(0x7FFFFFFF)).astype(np.uint32)
        vals = bits.view(np.float32).astype(np.float64)
        med = (vals[0::2] + vals[1::2]) * 0.5
        med[n_valid[0::2] == 0] = 0.0
        med[np.isnan(med)] = 0.0
        return med.astype(np.float32)

    def numeric(self, x: torch.Tensor, r0: int, rows: int, med_dev: torch.Tensor):
        Fn, ld = x.shape
        X = torch.empty((rows, Fn), dtype=torch.float32, device=self.dev)
        M = torch.empty((rows, Fn), dtype=torch.uint8, device=self.dev)
        if Fn and rows:
            check(self.lib.b200ctr_cache_numeric(_p(x), ld, r0, rows, Fn, _p(med_dev), _p(X), _p(M), _stream(self.dev)))
        return X, M


def _save_shard(shard_dir: str, arrays: Dict[str, np.ndarray]) -> Dict:
    """build_cache_v1.py:160-169"""
    os.makedirs(shard_dir, exist_ok=True)
    meta: Dict = {}
    for k, v in arrays.items():
        path = os.path.join(shard_dir, f"{k}.npy")
        np.save(path, v)
        meta[k] = {"path": path, "shape": list(v.shape), "dtype": str(v.dtype)}
    meta["rows"] = int(arrays["seq"].shape[0])
    return meta


_KEYS = ["X_num", "X_mask", "X_cat", "seq", "y", "groups", "ids"]


def build_sharded_cache(parquet_path: str, out_dir: str, *, is_train: bool, target_col: Optional[str], seq_col: str,
                        cat_cols: List[str], hash_buckets: Dict[str, int], hash_buckets_margin: int, num_patterns: List[str],
                        max_len: int, pad_id: int, group_key: str, time_key: Optional[str] = None, composite_group: bool = False,
                        shard_rows: int = 2_000_000, impute_strategy: str = "median",
                        num_cols_explicit: Optional[List[str]] = None, remove_cols: Optional[List[str]] = None,
                        device=None, batch_rows: int = 200_000) -> str:
    """Parquet -> shards + manifest.json; returns the manifest path (build_cache_v1.py:171-306, same arguments).
    `device` / `batch_rows` are additions (CUDA device, Arrow batch size; the reference streams 200,000-row batches too)."""
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.parquet as pq

    K = CacheKernels(device)
    os.makedirs(out_dir, exist_ok=True)
    pf = pq.ParquetFile(parquet_path)
    cols = list(pf.schema_arrow.names)
    remove = set(remove_cols or [])
    # ---- schema (build_cache_v1.py:44-53) ----
    if num_cols_explicit:
        num_cols = [c for c in num_cols_explicit if c in cols]
    else:
        num_cols = [c for c in _match_patterns(cols, num_patterns)
                    if c not in cat_cols and c not in [target_col, seq_col, group_key, "ID"] and c in cols]
    num_cols = [c for c in num_cols if c not in remove]
    total_n = int(pf.metadata.num_rows)
    Fn = len(num_cols)
    # ---- scan 1: numeric columns -> resident column-major matrix, medians (build_cache_v1.py:55-66) ----
    x = torch.empty((Fn, max(total_n, 1)), dtype=torch.float32, device=K.dev)
    if Fn:
        r0 = 0
        for rb in pf.iter_batches(batch_size=batch_rows, columns=num_cols):
            names = rb.schema.names
            for j, c in enumerate(num_cols):
                K.col_to_f32(rb.column(names.index(c)), x[j, r0:r0 + rb.num_rows])
            r0 += rb.num_rows
    med = K.medians(x, total_n) if (impute_strategy == "median" and Fn) else np.zeros((Fn,), np.float32)
    med_dev = torch.from_numpy(med.copy()).to(K.dev) if Fn else torch.zeros(1, device=K.dev)
    # ---- scan 2: everything else, batch by batch (build_cache_v1.py:78-173, 244-289) ----
    other = [c for c in cols if c not in num_cols and c not in remove]
    manifest = {"parquet": parquet_path, "is_train": is_train, "rows": 0, "shards": [], "num_cols": num_cols, "cat_cols": cat_cols,
                "group_key": group_key, "seq_col": seq_col, "format": "b200ctr-v2 (pinned fnv1a64+splitmix64 hash, seed 2025)",
                "medians": {c: float(m) for c, m in zip(num_cols, med)}}
    acc: Dict[str, list] = {k: [] for k in _KEYS}
    shard_idx, row_buf, r0 = 0, 0, 0
    Fc = len(cat_cols)

    def emit(arrays):
        nonlocal shard_idx
        meta = _save_shard(os.path.join(out_dir, f"shard_{shard_idx:03d}"), arrays)
        meta["index"] = shard_idx
        meta["start"] = manifest["rows"]
        meta["end"] = manifest["rows"] + meta["rows"]
        manifest["shards"].append(meta)
        manifest["rows"] += meta["rows"]
        shard_idx += 1

    for rb in pf.iter_batches(batch_size=batch_rows, columns=other if other else None):
        n = rb.num_rows
        names = rb.schema.names
        col = lambda c: rb.column(names.index(c))
        y = (np.asarray(pc.cast(col(target_col), pa.int8()).to_numpy(zero_copy_only=False), dtype=np.int8)
             if (is_train and target_col in names) else np.zeros((n,), np.int8))
        # groups (97-113)
        g_dev = torch.zeros((n,), dtype=torch.int64, device=K.dev)
        if composite_group and group_key in names and time_key is not None and time_key in names:
            gs = pc.fill_null(pc.cast(col(group_key), pa.string()), "NA")
            ts = pc.fill_null(pc.cast(col(time_key), pa.string()), "NA")
            K.hash_column(pc.binary_join_element_wise(gs, ts, "\x1f"), 2 ** 31 - 1, g_dev, 1, False)
        elif group_key in names:
            K.hash_column(col(group_key), 2 ** 31 - 1, g_dev, 1, False)
        # ids (115-120)
        if "ID" in names:
            ids = np.asarray(pc.fill_null(pc.cast(col("ID"), pa.string()), "").to_numpy(zero_copy_only=False), dtype="U64")
        else:
            ids = np.arange(n, dtype=np.int64).astype(str).astype("U64")
        # categorical hash (121-132)
        xcat = torch.zeros((n, max(Fc, 1)), dtype=torch.int32, device=K.dev)
        for j, c in enumerate(cat_cols):
            if c in names:
                K.hash_column(col(c), int(hash_buckets.get(c, 1000003)) + int(hash_buckets_margin), xcat[:, j], max(Fc, 1), True)
        # numeric + mask (134-146)
        X_num, X_mask = K.numeric(x, r0, n, med_dev)
        # seq (148-158)
        if seq_col in names:
            seq = K.parse_seq(col(seq_col), max_len, pad_id)
        else:
            seq = torch.full((n, max_len), pad_id, dtype=torch.int32, device=K.dev)
        batch = {"X_num": X_num.cpu().numpy(), "X_mask": X_mask.cpu().numpy(),
                 "X_cat": xcat[:, :Fc].contiguous().cpu().numpy() if Fc else np.zeros((n, 0), np.int32),
                 "seq": seq.cpu().numpy(), "y": y, "groups": g_dev.cpu().numpy(), "ids": ids}
        for k in _KEYS:
            acc[k].append(batch[k])
        row_buf += n
        r0 += n
        while row_buf >= shard_rows:
            cat = {k: np.concatenate(acc[k], axis=0) for k in _KEYS}
            emit({k: cat[k][:shard_rows] for k in _KEYS})
            acc = {k: [cat[k][shard_rows:]] for k in _KEYS}
            row_buf = acc["seq"][0].shape[0]
    if row_buf > 0:
        emit({k: np.concatenate(acc[k], axis=0) for k in _KEYS})
    man_path = os.path.join(out_dir, "manifest.json")
    with open(man_path, "w") as f:
        json.dump(manifest, f, indent=2)
    return man_path
