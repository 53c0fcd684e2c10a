"""Development aid: where does the end-to-end step lose time against the device-resident one?  Variants of the host loop
(feeder / serial copy / device batches) x (no read-back / float(loss) per step), GPU event time and host wall time per step.
Finding (r02ah): with no per-step read-back the host runs many batches ahead of the GPU and the DeviceFeeder loop degrades
(23 ms/step); with float(loss) per step -- the reference loop's own pattern -- it is 9.90 ms against 10.18 ms for copy-then-step."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "toss-next-ctr-prediction_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import b200ctr
    env = bench.Env()
    cfg = bench.build_cfg("dare_qnn_next")
    cfg["amp"] = "bf16"
    B, NB, steps = cfg["train"]["batch_size"], 16, 10
    host, dev_b = bench.make_batches(cfg, B, NB, 0, env.dev)
    model, tr = bench.new_trainer(env, cfg, 10_000_000)
    cur = [0]

    def nxt(src):
        item = src[cur[0] % NB]
        cur[0] += 1
        return item

    def loop(kind, read):
        def run(n):

            def consume(loss):
                if read == "float":
                    float(loss)
            if kind == "feeder":
                for bb, yy in b200ctr.DeviceFeeder((nxt(host) for _ in range(n)), env.dev):
                    consume(tr.step(bb, yy))
            else:
                for _ in range(n):
                    bb, yy = nxt(host if kind == "serial" else dev_b)
                    consume(tr.step(bb, yy))
        return run
    for _ in range(4):
        tr.step(*nxt(dev_b))
    torch.cuda.synchronize()
    out = {}
    for kind in ("device", "serial", "feeder"):
        for read in ("none", "float"):
            fn = loop(kind, read)
            fn(2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ms = env.timed(fn, steps)
            wall = (time.perf_counter() - t0) * 1e3
            # host-only cost of launching: time to ENQUEUE the steps (no final sync inside fn for read == none)
            out[f"{kind}/{read}"] = {"gpu_ms_per_step": round(ms / steps, 3), "wall_ms_per_step": round(wall / steps, 3)}
    # pure host launch time of one step while the GPU is idle at the start
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.step(*nxt(dev_b))
    out["host_enqueue_ms_one_step"] = round((time.perf_counter() - t0) * 1e3, 3)
    torch.cuda.synchronize()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
