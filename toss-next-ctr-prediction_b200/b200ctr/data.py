"""Batched shard reader: ShardedDataset + DataLoader(collate_fn=collate_sharded) of the reference (src/data/dataset.py:45-124,
src/train.py:100-110) as ONE object that yields whole batches.

The reference walks a batch row by row in Python (`ShardedDataset.__getitem__` -> `ShardedNPY.get_row` over memory-mapped .npy
files) and stacks the rows in `collate_sharded`: 10^3-10^4 samples/s per worker, three orders of magnitude below one B200 step.
Here a batch's global indices are mapped to (shard, local row) with one searchsorted, sorted by local row inside each shard
(block-sequential reads of the mmap) and gathered by host threads (libb200ctr `b200ctr_gather_rows`) into pinned staging
buffers in the ON-DISK dtypes (f32 X_num, u8 X_mask, i32 X_cat / seq, i8 y, i64 groups).  Two output forms:

  * `device=None`  -> the collate contract verbatim: CPU tensors X_num f32, X_mask f32, X_cat i64, seq i64, y f32 (train),
                      groups i64, ids numpy (test) -- bit-identical to collate_sharded on the same indices (tests).
  * `device=cuda`  -> the narrow buffers go over PCIe as they are (17.6 MB instead of 33.9 MB per 8192-row batch) on a copy
                      stream, double-buffered against the consumer; the mask / labels are widened on the device
                      (`b200ctr_widen`), ids stay int32 (every b200ctr kernel reads int32 or int64 ids).
Shuffling reproduces nothing of torch's RandomSampler bit-for-bit (it is a different sampler); it is a seeded permutation
per epoch.  `shuffle=False` walks `index` in order exactly like the reference's validation / inference loaders.
"""
from __future__ import annotations

import ctypes
import json
import threading
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from . import _lib

_KEYS = ("X_num", "X_mask", "X_cat", "seq", "y", "groups")


class BatchedShardReader:
    def __init__(self, manifest_path: str, index: np.ndarray, train: bool, batch_size: int, shuffle: bool = False, seed: int = 0,
                 drop_last: bool = False, device: Optional[torch.device] = None, n_threads: int = 8):
        with open(manifest_path) as fh:
            man = json.load(fh)
        self.manifest = man
        self.index = np.asarray(index, dtype=np.int64)
        self.train, self.batch_size, self.shuffle, self.seed, self.drop_last = bool(train), int(batch_size), bool(shuffle), int(seed), bool(drop_last)
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and self.device.type != "cuda":
            raise _lib.B200CtrError("BatchedShardReader(device=...) must be a CUDA device; use device=None for the CPU collate contract")
        self.n_threads = int(n_threads)
        self.starts = np.array([m["start"] for m in man["shards"]], dtype=np.int64)
        self.ends = np.array([m["end"] for m in man["shards"]], dtype=np.int64)
        self.arrs: List[Dict[str, np.ndarray]] = []
        for m in man["shards"]:
            d = {k: np.load(m[k]["path"], mmap_mode="r") for k in _KEYS}
            d["ids"] = np.load(m["ids"]["path"], allow_pickle=False)
            self.arrs.append(d)
        a0 = self.arrs[0]
        self.row_shape = {k: tuple(a0[k].shape[1:]) for k in _KEYS}
        self.dtypes = {k: a0[k].dtype for k in _KEYS}
        self.epoch = 0
        self._lib = _lib.load()

    def __len__(self) -> int:
        n = self.index.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    # ---- host side: one batch of global indices -> narrow arrays (pinned when a device is set) ----
    def _alloc(self, B: int) -> Dict[str, torch.Tensor]:
        pin = self.device is not None
        out = {}
        for k in _KEYS:
            if k == "y" and not self.train:
                continue
            t = torch.empty((B,) + self.row_shape[k], dtype=torch.from_numpy(np.empty(0, self.dtypes[k])).dtype)
            out[k] = t.pin_memory() if pin else t
        return out

    def _gather(self, gidx: np.ndarray, out: Dict[str, torch.Tensor]):
        B = gidx.shape[0]
        sid = np.searchsorted(self.ends, gidx, side="right")                  # dataset.py:70-75 (_locate)
        local = gidx - self.starts[sid]
        ids = None if self.train else np.empty(B, dtype=self.arrs[0]["ids"].dtype)
        for s in np.unique(sid):
            pos = np.nonzero(sid == s)[0]
            order = np.argsort(local[pos], kind="stable")                     # block-sequential reads inside the shard
            dst = np.ascontiguousarray(pos[order].astype(np.int64))
            src = np.ascontiguousarray(local[pos][order].astype(np.int64))
            for k, t in out.items():
                a = self.arrs[s][k]
                row_bytes = int(a.dtype.itemsize * int(np.prod(a.shape[1:], dtype=np.int64)))
                _lib.check(self._lib.b200ctr_gather_rows(ctypes.c_void_p(a.ctypes.data), row_bytes, ctypes.c_void_p(src.ctypes.data),
                                                         ctypes.c_void_p(dst.ctypes.data), int(src.shape[0]),
                                                         ctypes.c_void_p(t.data_ptr()), self.n_threads), "gather_rows")
            if ids is not None:
                ids[dst] = self.arrs[s]["ids"][src]
        if ids is not None:
            ids = np.asarray(ids.tolist())          # collate_sharded builds the array from str scalars: itemsize = longest id
        return ids

    def _batches(self) -> List[np.ndarray]:
        idx = self.index
        if self.shuffle:
            idx = idx[np.random.default_rng(self.seed + self.epoch).permutation(idx.shape[0])]
        n = len(self)
        return [idx[i * self.batch_size:(i + 1) * self.batch_size] for i in range(n)]

    # ---- the two output forms ----
    def _contract(self, narrow: Dict[str, torch.Tensor], ids) -> dict:
        """collate_sharded's dtypes (dataset.py:111-123)."""
        out = {"X_num": narrow["X_num"].float(), "X_mask": narrow["X_mask"].float(), "X_cat": narrow["X_cat"].long(),
               "seq": narrow["seq"].long()}
        if self.train:
            out["y"] = narrow["y"].float()
        if ids is not None:
            out["ids"] = ids
        out["groups"] = narrow["groups"].long()
        return out

    def _to_device(self, narrow: Dict[str, torch.Tensor], ids, stream: torch.cuda.Stream) -> dict:
        dev = self.device
        with torch.cuda.stream(stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in narrow.items() if k != "groups"}
            B = d["seq"].shape[0]
            out = {"X_num": d["X_num"], "X_cat": d["X_cat"], "seq": d["seq"]}
            mask = torch.empty(d["X_mask"].shape, device=dev, dtype=torch.float32)
            _lib.check(self._lib.b200ctr_widen(0, d["X_mask"].data_ptr(), d["X_mask"].numel(), mask.data_ptr(), stream.cuda_stream), "widen")
            out["X_mask"] = mask
            if self.train:
                y = torch.empty((B,), device=dev, dtype=torch.float32)
                _lib.check(self._lib.b200ctr_widen(1, d["y"].data_ptr(), B, y.data_ptr(), stream.cuda_stream), "widen")
                out["y"] = y
            ev = torch.cuda.Event()
            ev.record(stream)
        out["groups"] = narrow["groups"].long().clone()          # the staging slot is refilled two batches later
        if ids is not None:
            out["ids"] = ids
        out["_ready"] = ev
        out["_keep"] = d                      # the narrow device buffers live until the consumer drops the batch
        return out

    def __iter__(self) -> Iterator[dict]:
        batches = self._batches()
        self.epoch += 1
        if self.device is None:
            for g in batches:
                narrow = self._alloc(g.shape[0])
                ids = self._gather(g, narrow)
                yield self._contract(narrow, ids)
            return
        # device form: staging set i % 2 is filled by a worker thread while the consumer computes on batch i - 1
        stream = torch.cuda.Stream(self.device)
        staging = [self._alloc(self.batch_size), self._alloc(self.batch_size)]
        free = [threading.Event(), threading.Event()]
        for e in free:
            e.set()
        results: List[Optional[dict]] = [None] * len(batches)
        ready = [threading.Event() for _ in batches]
        err: List[BaseException] = []

        def produce():
            try:
                torch.cuda.set_device(self.device)
                for i, g in enumerate(batches):
                    slot = i % 2
                    free[slot].wait()
                    free[slot].clear()
                    narrow = {k: v[:g.shape[0]] for k, v in staging[slot].items()}
                    ids = self._gather(g, narrow)
                    results[i] = self._to_device(narrow, ids, stream)
                    ready[i].set()
            except BaseException as e:          # surfaced in the consumer
                err.append(e)
                for r in ready:
                    r.set()
        th = threading.Thread(target=produce, daemon=True)
        th.start()
        for i in range(len(batches)):
            ready[i].wait()
            if err:
                raise err[0]
            b = results[i]
            results[i] = None
            ev = b.pop("_ready")
            torch.cuda.current_stream(self.device).wait_event(ev)     # the batch's H2D + widening are done before it is read
            ev.synchronize()                                          # ... and the pinned staging slot may be refilled
            free[i % 2].set()
            yield b
        th.join()
