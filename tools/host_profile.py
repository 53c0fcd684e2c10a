"""Development aid: cProfile of the HOST side of FusedTrainer.step (the step is launch-bound on the host when Python needs as
long to enqueue it as the GPU needs to run it)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "toss-next-ctr-prediction_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    env = bench.Env()
    cfg = bench.build_cfg(os.environ.get("CFG", "dare_qnn_next"))
    cfg["amp"] = "bf16"
    B, NB = cfg["train"]["batch_size"], 8
    host, dev_b = bench.make_batches(cfg, B, NB, 0, env.dev)
    model, tr = bench.new_trainer(env, cfg, 10_000_000)
    for i in range(4):
        tr.step(*dev_b[i % NB])
    torch.cuda.synchronize()
    # enqueue time with an idle GPU in front (no back-pressure), 3 separate steps
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.step(*dev_b[i % NB])
        print("enqueue ms", round((time.perf_counter() - t0) * 1e3, 3), file=sys.stderr)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(5):
        tr.step(*dev_b[i % NB])
    pr.disable()
    torch.cuda.synchronize()
    for key in ("tottime", "cumtime"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
        print(s.getvalue())


if __name__ == "__main__":
    main()
