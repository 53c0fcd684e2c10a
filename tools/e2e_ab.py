"""Development aid: where does the end-to-end step lose time against the device-resident one?  Host-loop variants
(device batches / pinned batch handed to step() / next batch's H2D on a copy stream, issued before or after the step is
enqueued) x (no read-back / float(loss) per step); GPU event time and host wall time per step.
Findings on the B200 boxes of this pool (r02ah / r02ak / r02ap):
  * device 9.18-9.40 ms, + float(loss) 9.44-9.58 ms, pinned batch through step() ("serial") 9.86-10.18 ms;
  * copy stream with the device tensors ALLOCATED PER BATCH on the side stream ("overlap_*"): 9.6-9.9 ms in one process, 12.0 ms
    in another, 18.8 ms with the copy issued once the step is enqueued, 14-23 ms without any read-back;
  * copy stream into two PREALLOCATED staging sets ("static_*"): 9.27-9.38 ms whichever way the copy is issued.
The slow cases were allocator traffic on the side stream (cudaMalloc while the GPU is busy, deferred reuse of blocks shared between
streams), not DMA contention.  FusedTrainer.stage() is the static form; bench.py's e2e uses it."""
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
            if kind.startswith("overlap"):
                cs = torch.cuda.Stream(env.dev)
                cur_s = torch.cuda.current_stream(env.dev)

                def issue():
                    hb, hy = nxt(host)
                    with torch.cuda.stream(cs):
                        d = {k: v.to(env.dev, non_blocking=True) for k, v in hb.items()}
                        yd = hy.to(env.dev, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(cs)
                    return d, yd, ev
                ready = issue()
                for i in range(n):
                    d, yd, ev = ready
                    ready = None
                    if kind == "overlap_before" and i + 1 < n:
                        cur_s.synchronize()
                        ready = issue()
                    cur_s.wait_event(ev)
                    for v in list(d.values()) + [yd]:
                        v.record_stream(cur_s)
                    loss = tr.step(d, yd)
                    if kind == "overlap_after" and i + 1 < n:
                        ready = issue()           # the next batch's copy is issued once this step is enqueued
                    consume(loss)
            elif kind.startswith("static"):
                # two preallocated device staging sets (no allocator traffic, no record_stream), copies on a copy stream
                cs = stat["cs"]
                cur_s = torch.cuda.current_stream(env.dev)

                def issue(slot):
                    hb, hy = nxt(host)
                    cs.wait_event(stat["used"][slot])               # the step that read this slot is done
                    with torch.cuda.stream(cs):
                        for k, v in hb.items():
                            stat["b"][slot][k].copy_(v, non_blocking=True)
                        stat["y"][slot].copy_(hy, non_blocking=True)
                        stat["ready"][slot].record(cs)
                issue(0)
                for i in range(n):
                    slot = i & 1
                    if kind == "static_before" and i + 1 < n:
                        issue(slot ^ 1)
                    cur_s.wait_event(stat["ready"][slot])
                    loss = tr.step(stat["b"][slot], stat["y"][slot])
                    stat["used"][slot].record(cur_s)
                    if kind == "static_after" and i + 1 < n:
                        issue(slot ^ 1)                             # the next batch's copy is issued once this step is enqueued
                    consume(loss)
            else:
                for _ in range(n):
                    bb, yy = nxt(host if kind == "serial" else dev_b)
                    consume(tr.step(bb, yy))
        return run
    for _ in range(4):
        tr.step(*nxt(dev_b))
    torch.cuda.synchronize()
    stat = {"cs": torch.cuda.Stream(env.dev),
            "b": [{k: torch.empty_like(v, device=env.dev) for k, v in host[0][0].items()} for _ in range(2)],
            "y": [torch.empty_like(host[0][1], device=env.dev) for _ in range(2)],
            "ready": [torch.cuda.Event(), torch.cuda.Event()], "used": [torch.cuda.Event(), torch.cuda.Event()]}
    for e in stat["used"]:
        e.record(torch.cuda.current_stream(env.dev))
    torch.cuda.synchronize()
    out = {}
    for kind in ("device", "serial", "static_before", "static_after", "overlap_before"):
        for read in ("none", "float"):
            if read == "none" and (kind.startswith("overlap") or kind == "device" or kind == "serial"):
                continue                          # known to degrade (host many batches ahead)
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
