"""A/B timing of the tensor-core attention flavours (packed-half csrc/attn_h16.cu vs TF32 csrc/attn_tc.cu) at the bench shapes:
B = 8192 samples, 8 heads of 8, top_k in {60, 100, 148}, dropout 0.1, bf16 dqkv.  CUDA events around `reps` back-to-back launches
(each launch streams > 1 GB, far above the 126 MB L2).  Prints one JSON line per (K, impl)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "toss-next-ctr-prediction_b200"))
from b200ctr import ops  # noqa: E402


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, nargs="*", default=[60, 100, 148])
    ap.add_argument("--impl", nargs="*", default=["tf32", "h16"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--in16", action="store_true", help="h16 only: feed the fp16 qkv tensor of inproj_fwd(h16=True)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, reps = 8192, a.reps
    out = []
    for K in a.K:
        g = torch.Generator(device=dev).manual_seed(K)
        qkv = torch.randn(B * K, 192, device=dev, generator=g) * 0.8
        relmean = torch.randn(2 * K + 1, device=dev, generator=g) * 0.3
        d_o = torch.randn(B * K, 64, device=dev, generator=g) * 1e-6
        dr = ops.drop(0.1, 7, 3)
        for impl in a.impl:
            res = {}
            if a.in16 and impl == "h16":
                qkv = torch.cat([(qkv[:, :64] * 0.51006972789764404297).half(), qkv[:, 64:].half()], dim=1).contiguous()
            for name in ("fwd", "bwd"):
                o, lse, keep = ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl=impl)
                drel = torch.zeros_like(relmean)

                def run():
                    if name == "fwd":
                        ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl=impl)
                    else:
                        ops.attn_bwd(qkv, relmean, K, o, lse, d_o, B, K, 8, 64, None, dr, tc=True, keep=keep, out_bf16=True, impl=impl)
                for _ in range(a.warm):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                res[name + "_ms"] = round(e0.elapsed_time(e1) / reps, 4)
            rec = {"K": K, "impl": impl, "in16": bool(a.in16 and impl == "h16"), **res}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    return out


if __name__ == "__main__":
    main()
