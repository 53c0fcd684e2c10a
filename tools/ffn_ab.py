"""Timing of the fused FFN block kernels (csrc/ffn_fused.cu) at the bench shape: T = 8192 x 60 tokens, hidden 384, dropout 0.1."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "toss-next-ctr-prediction_b200"))
from b200ctr import ops  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


def main():
    dev = torch.device("cuda:0")
    T, hidden = 8192 * 60, 384
    g = torch.Generator(device=dev).manual_seed(1)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    x1, w1, b1, w2, b2, n2 = r(T, 64), r(hidden, 64) * 0.1, r(hidden) * 0.1, r(64, hidden) * 0.05, r(64) * 0.1, 1 + 0.1 * r(64)
    dy = r(T, 64) * 1e-6
    for p in (0.1, 0.0):
        dr = ops.drop(p, 5, 9)
        x2, s2, r2, keep = ops.ffn_fwd(x1, w1, b1, w2, b2, n2, dr, save_keep=True)
        z = lambda *s: torch.zeros(*s, device=dev)
        dw1, db1, dw2, db2, dn2 = z(hidden, 64), z(hidden), z(64, hidden), z(64), z(64)
        res = {"p": p,
               "fwd_ms": timeit(lambda: ops.ffn_fwd(x1, w1, b1, w2, b2, n2, dr, save_keep=True)),
               "bwd_ms": timeit(lambda: ops.ffn_bwd(dy, s2, r2, n2, x1, w1, b1, w2, dw1, db1, dw2, db2, dn2, dr, keep=keep))}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
