"""Kernel-level throughput of the on-GPU cache builder (csrc/cache_build.cu) on Toss-shaped synthetic columns, with the CPU
oracle (oracle/cache_oracle.py = the reference's per-row Python rules) timed beside it on a bounded sample.
Algorithmic bytes: parse = string bytes + offsets in, 4 L per row out; numeric = 4 Fn in, 5 Fn out per row; hash = value bytes +
offsets in, 4 out.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "toss-next-ctr-prediction_b200"))
sys.path.insert(0, ROOT)


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import pyarrow as pa
    from b200ctr import _lib
    from b200ctr.cache_builder import CacheKernels, _dev_bytes, _p, _stream, HASH_SEED
    from oracle import cache_oracle as co
    rows, L, Fn = int(os.environ.get("ROWS", 400_000)), 400, 82
    rng = np.random.default_rng(777)
    K = CacheKernels("cuda:0")
    lib = K.lib
    # seq strings: length min(L, lognormal(5.3, 0.9)) tokens over [1, 1e6]
    lens = np.minimum(L, rng.lognormal(5.3, 0.9, rows).astype(np.int64))
    toks = rng.integers(1, 1_000_000, int(lens.sum()))
    strs, pos = [], 0
    for n in lens:
        strs.append(",".join(map(str, toks[pos:pos + n])))
        pos += n
    arr = pa.array(strs, type=pa.large_string())
    bufs = arr.buffers()
    offs, data = _dev_bytes(bufs[1], K.dev), _dev_bytes(bufs[2], K.dev)
    out = torch.empty((rows, L), dtype=torch.int32, device=K.dev)
    bad = torch.zeros(1, dtype=torch.int64, device=K.dev)
    ms = ev(lambda: lib.b200ctr_cache_parse_seq(_p(data), _p(offs), 1, None, rows, L, 0, _p(out), _p(bad), _stream(K.dev)))
    nbytes = data.numel() + offs.numel() + out.numel() * 4
    res = {"rows": rows, "parse_seq": {"ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1), "Mrows_s": round(rows / ms / 1e3, 2),
                                       "string_bytes": int(data.numel())}}
    t0 = time.perf_counter()
    ns = 4000
    ref = co.parse_seq(strs[:ns], L, 0)
    dt = time.perf_counter() - t0
    assert np.array_equal(out[:ns].cpu().numpy(), ref)
    res["parse_seq"]["cpu_oracle_rows_s"] = round(ns / dt)
    # numeric block
    x = torch.randn(Fn, rows, device=K.dev)
    x[:, ::20] = float("nan")
    med = torch.zeros(Fn, device=K.dev)
    X = torch.empty((rows, Fn), device=K.dev)
    M = torch.empty((rows, Fn), dtype=torch.uint8, device=K.dev)
    ms = ev(lambda: lib.b200ctr_cache_numeric(_p(x), rows, 0, rows, Fn, _p(med), _p(X), _p(M), _stream(K.dev)))
    res["numeric"] = {"ms": round(ms, 3), "GBps": round(rows * Fn * 9 / ms / 1e6, 1)}
    t0 = time.perf_counter()
    med_h = K.medians(x, rows)
    torch.cuda.synchronize()
    res["medians_82_cols"] = {"ms": round((time.perf_counter() - t0) * 1e3, 1)}
    assert np.array_equal(med_h[:3], np.array([co.median_ignore_missing(x[j].cpu().numpy()) for j in range(3)], dtype=np.float32))
    # hash of an int64 column
    vals = torch.from_numpy(rng.integers(0, 5000, rows)).to(K.dev)
    oc = torch.empty(rows, dtype=torch.int32, device=K.dev)
    ms = ev(lambda: lib.b200ctr_cache_hash_int64(_p(vals), None, rows, 2000003, HASH_SEED, _p(oc), 1, 1, _stream(K.dev)))
    res["hash_int64"] = {"ms": round(ms, 3), "GBps": round(rows * 12 / ms / 1e6, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
