#!/usr/bin/env python
"""bench.py — train samples/s (fwd + bwd + clip + AdamW + EMA) of the DARE + QNN-alpha hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference itself on the host CPU

Workload (BASELINE.json configs[1]): cfgs/dare_qnn_next.yaml — B=8192, L=400, K=60, 3 encoder layers, QNN-alpha
(6 heads, rank 16, proj 192, MLP 512-256), 35 hashed-categorical tables, 82 numeric + 82 mask features, two
10,000,000 x 64 sequence tables, dropout ON (train mode), EMA 0.999 — on synthetic Toss-shaped data (b200ctr.synth).
The headline `value` is the `amp: bf16` mode (BASELINE configs[1] names fp32/bf16); the same run also times the `amp: none`
mode (`fp32` key: tcgen05 split-precision contractions, the mode every shipped cfg selects) and BASELINE configs[2]
(`scale_config` key: dare_qnn_next_k100_s1, the config the 1/2/4/8 scaling is quoted on) so every N carries both.
One process per GPU; N > 1 = data parallel over the batch with the two sequence tables row-sharded across the GPUs, B per GPU
fixed (weak scaling).

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM, timed over K plain steps (no
instrumentation inside the region); `e2e` = the same step through FusedTrainer.step() with pinned HOST batches (H2D inside the
timed region) and a D2H read of the loss every step.  `roofline` describes the dominant kernel family of the step, timed live
with a CUDA event pair around every C-ABI call on the launch stream in a SECOND timed region of K steps (`ms_per_step_profiled`:
the ~350 event records per step cost 0.7 ms, which is why they are not inside the region `value` is quoted on).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "toss-next-ctr-prediction_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train samples/sec (fwd+bwd+opt)"
UNIT = "samples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_cfg(name: str):
    from b200ctr import configs
    return getattr(configs, name)()


def workload_name(cfg, B):
    s = cfg["sequence"]
    return (f"cfgs/{cfg['exp_name']}.yaml: B={B}/GPU, L={s['max_len']}, top_k={s['top_k']}, query={s['query_mode']}, "
            f"{s.get('tfm', {}).get('n_layers', 0) if s['transformer_block'] else 0} encoder layers, QNN-alpha "
            f"{'on' if cfg['model']['qnn_alpha']['enabled'] else 'off'}, seq_vocab=10,000,000 x 64 x 2 tables, "
            f"amp={cfg.get('amp', 'none')} (fp32 master weights + optimizer), dropout on, "
            f"EMA {'on' if (cfg.get('ema') or {}).get('enabled') else 'off'}")


# ---------------------------------------------------------------------------------------------------------------
# CPU reference arm: the UNMODIFIED reference (baseline/_ref: CTRModel + the literal step body of src/train.py:152-199) on
# the host cores; the oracle port only if the copy is absent.  bench.py may execute oracle/ / baseline/_ref here only.
# ---------------------------------------------------------------------------------------------------------------
def _ref_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "src", "models"))


def _port_state(cfg, spec, vocab):
    """Random parameters in the reference's state_dict layout for the oracle-port fallback arm."""
    from b200ctr import configs
    g = torch.Generator().manual_seed(cfg["seed"])
    cols = cfg["data"]["cat_cols"]
    D = spec.D
    sd = {}
    qa = cfg["model"]["qnn_alpha"]
    fe = int(qa.get("feature_embed_dim", 16))
    sd["num_embed.weight"] = torch.randn(configs.N_NUM, fe, generator=g) * 0.02
    sd["num_embed.bias"] = torch.zeros(configs.N_NUM, fe)
    sd["num_embed.out_proj.weight"] = torch.randn(D, fe, generator=g) * fe ** -0.5
    sd["mask_embed.weight"] = torch.randn(configs.N_NUM, fe, generator=g) * 0.02
    sd["mask_embed.out_proj.weight"] = torch.randn(D, fe, generator=g) * fe ** -0.5
    card = configs.cardinals(cfg)
    dims = cfg["model"].get("cat_embedding_dims", {})
    for c in cols:
        dc = int(dims.get(c, D))
        sd[f"cat_embs.{c}.weight"] = torch.randn(card[c], dc, generator=g)
        sd[f"cat_proj.{c}.weight"] = torch.randn(D, dc, generator=g) * dc ** -0.5
    sd["ctx_mlp.0.weight"] = torch.randn(D, 3 * D, generator=g) * (3 * D) ** -0.5
    sd["ctx_mlp.0.bias"] = torch.zeros(D)
    for n in ("dare.emb_att.weight", "dare.emb_rep.weight"):
        sd[n] = torch.randn(vocab, D, generator=g)
        sd[n][0] = 0
    ffn = cfg["sequence"]["tfm"]["ffn_hidden"] if spec.transformer else 0
    K = spec.top_k
    for i in range(spec.n_layers):
        p = f"dare.layers.{i}."
        sd[p + "mha.in_proj_weight"] = torch.randn(3 * D, D, generator=g) * D ** -0.5
        sd[p + "mha.in_proj_bias"] = torch.zeros(3 * D)
        sd[p + "mha.out_proj.weight"] = torch.randn(D, D, generator=g) * D ** -0.5
        sd[p + "mha.out_proj.bias"] = torch.zeros(D)
        sd[p + "norm1.w"] = torch.ones(D)
        sd[p + "ffn.0.weight"] = torch.randn(ffn, D, generator=g) * D ** -0.5
        sd[p + "ffn.0.bias"] = torch.zeros(ffn)
        sd[p + "ffn.3.weight"] = torch.randn(D, ffn, generator=g) * ffn ** -0.5
        sd[p + "ffn.3.bias"] = torch.zeros(D)
        sd[p + "norm2.w"] = torch.ones(D)
        sd[p + "pbias.rel.weight"] = torch.randn(2 * K + 1, spec.n_heads, generator=g)
    sd["dare.aux_head.weight"] = torch.randn(1, D, generator=g) * D ** -0.5
    sd["dare.aux_head.bias"] = torch.zeros(1)
    if spec.use_qnn:
        F = 1 + 2 * configs.N_NUM + len(cols)
        H, r, Pd = qa["heads"], qa["rank"], qa["proj_dim"]
        sd["qnn.pre_norm.w"] = torch.ones(F * D)
        sd["qnn.U"] = torch.randn(H, D, r, generator=g) * 0.02
        sd["qnn.V"] = torch.randn(H, r, Pd, generator=g) * 0.02
        HP = H * Pd
        sd["qnn.se.fc.0.weight"] = torch.randn(HP // 8, HP, generator=g) * HP ** -0.5
        sd["qnn.se.fc.0.bias"] = torch.zeros(HP // 8)
        sd["qnn.se.fc.2.weight"] = torch.randn(HP, HP // 8, generator=g) * (HP // 8) ** -0.5
        sd["qnn.se.fc.2.bias"] = torch.zeros(HP)
        ind = HP + F * D
        for li, h in enumerate(list(qa["mlp_hidden"]) + [1]):
            sd[f"qnn.mlp.{3 * li}.weight"] = torch.randn(h, ind, generator=g) * ind ** -0.5
            sd[f"qnn.mlp.{3 * li}.bias"] = torch.zeros(h)
            ind = h
    return sd


def cpu_reference(cfg_name: str, steps: int, warmup: int, sample_B: int = 512, vocab: int = 1_000_001):
    from b200ctr import configs, synth
    cfg = build_cfg(cfg_name)
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cols = cfg["data"]["cat_cols"]
    schema = synth.ToySchema(cfg, configs.N_NUM, configs.N_NUM, seed=cfg["seed"])
    batches = [schema.batch(sample_B, seed=1000 + s) for s in range(warmup + steps)]
    times = []
    if _ref_available():
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        from src.models.wrapper import CTRModel            # the unmodified reference
        from src.train import bce_wll_style
        from src.utils.ema import build_ema
        from src.utils.sched import cosine_warmup_lr
        torch.manual_seed(cfg["seed"])
        model = CTRModel(cfg, vocab, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cols).train()
        ema = build_ema(model, cfg)
        opt = torch.optim.AdamW(model.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
        aux_w = float(cfg["model"]["qnn_alpha"].get("aux_head_weight", 0.0))
        for s, (b, y) in enumerate(batches):
            bt = {k: torch.from_numpy(v) for k, v in b.items()}
            yt = torch.from_numpy(y)
            t0 = time.perf_counter()
            opt.param_groups[0]["lr"] = cosine_warmup_lr(0, s, 1000, cfg["train"]["lr"], cfg["train"]["warmup_epochs"], cfg["train"]["epochs"])
            opt.zero_grad(set_to_none=True)
            logits, prob, aux = model(bt)
            loss = bce_wll_style(logits, yt)
            if aux_w > 0:
                loss = loss + aux_w * bce_wll_style(aux, yt)
            loss.backward()
            if cfg["train"]["grad_clip_norm"] > 0:
                torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["train"]["grad_clip_norm"])
            opt.step()
            if ema is not None:
                ema.update(model, s + 1)
            float(loss.detach())
            if s >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
        how = (f"the unmodified reference (baseline/_ref: src.models.wrapper.CTRModel + the literal step body of src/train.py:152-199 "
               f"incl. clip_grad_norm_, torch AdamW, ModelEMA), fp32 eager, dropout on as shipped")
    else:
        from oracle import ctr_oracle as O
        spec = O.ModelSpec(cfg, configs.N_NUM, configs.N_NUM, cols)
        sd = _port_state(cfg, spec, vocab)
        ema = cfg.get("ema") or {}
        tr = O.OracleTrainer(sd, spec, cfg["train"]["lr"], cfg["train"]["weight_decay"], cfg["train"]["grad_clip_norm"],
                             ema_decay=ema.get("decay") if ema.get("enabled") else None)
        for s, (b, y) in enumerate(batches):
            bt = {k: torch.from_numpy(v) for k, v in b.items()}
            t0 = time.perf_counter()
            tr.step(bt, torch.from_numpy(y))
            if s >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
        how = "the oracle port (torch CPU ops), fp32, dropout off"
    dt = sum(times)
    return dict(value=sample_B * len(times) / dt, unit=UNIT, cores=cores, kind=kind,
                sample=f"{len(times)} steps x B={sample_B} of the same workload through {how}, on {cores} threads; seq_vocab={vocab} "
                       f"(covers the synthetic token range; the shipped 10,000,000-row tables add ~10 s/step of dense AdamW+EMA on the "
                       f"CPU, so this choice favours the CPU arm)"), dt / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, sec = cpu_reference(args.cfg, max(1, args.steps), max(0, args.warmup), sample_B=args.ref_batch)
    cfg = build_cfg(args.cfg)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, cfg["train"]["batch_size"]), "reference_step_batch": args.ref_batch},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.pg = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.pg = dist.group.WORLD

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, loop_fn, n):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop_fn(n)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def make_batches(cfg, B, NB, rank, dev, uniform_tokens=False, vocab=10_000_000):
    """NB distinct synthetic batches: (pinned host copies, device copies).  uniform_tokens: token ids uniform over the whole
    vocabulary (every row a cold L2 miss) instead of the Zipf(1.05) stream over 1 M ids."""
    from b200ctr import configs, synth
    schema = synth.ToySchema(cfg, configs.N_NUM, configs.N_NUM, seed=cfg["seed"])

    def one(i):
        b, y = schema.batch(B, seed=cfg["seed"] + 100_000 * rank + i)
        if uniform_tokens:
            rng = np.random.default_rng(cfg["seed"] + 7_000_000 + 100_000 * rank + i)
            nz = b["seq"] != 0
            b["seq"][nz] = rng.integers(1, vocab, size=int(nz.sum()))
        return b, y
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        raw = list(ex.map(one, range(NB)))
    host, dev_b = [], []
    for b, y in raw:
        hb = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
        hy = torch.from_numpy(y).pin_memory()
        host.append((hb, hy))
        dev_b.append(({k: v.to(dev) for k, v in hb.items()}, hy.to(dev)))
    return host, dev_b


def gather_bytes(seq: torch.Tensor, L, Kk, D=64) -> float:
    """Algorithmic bytes of retrieve_fwd for one batch (every logical row use counted once)."""
    ell = (seq != 0).sum(1).double()
    kv = torch.clamp(ell, max=Kk)
    return float((8 * L + 4 * D * ell + 4 * D * kv + 4 * D + 4 * D * Kk + 12 * Kk).sum())


def new_trainer(env, cfg, vocab, replicate=False):
    import b200ctr
    from b200ctr import configs
    torch.manual_seed(cfg["seed"])
    with torch.device(env.dev):
        model = b200ctr.CTRModel(cfg, vocab, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cfg["data"]["cat_cols"])
    model.train()
    tr = b200ctr.FusedTrainer(model, cfg, steps_per_epoch=1000, process_group=env.pg, shard_tables=(env.world > 1 and not replicate))
    return model, tr


def measure(env, cfg, vocab, dev_b, host, steps, warmup, profile=False, e2e=False, replicate=False):
    """Warm-up, then K device-resident steps (optionally per-call CUDA-event profile + collective timing), then K host-fed steps."""
    from b200ctr import ops
    model, tr = new_trainer(env, cfg, vocab, replicate)
    NB = len(dev_b)
    B = dev_b[0][1].shape[0]
    cursor = [0]

    def dev_loop(n):
        for _ in range(n):
            bb, yy = dev_b[cursor[0] % NB]
            cursor[0] += 1
            tr.step(bb, yy)

    def host_loop_serial(n):
        """The pinned host batch handed to FusedTrainer.step itself: H2D on the step's stream in front of its first kernel."""
        for _ in range(n):
            hb, hy = host[cursor[0] % NB]
            cursor[0] += 1
            float(tr.step(hb, hy))             # D2H read of the step's loss (train.py:202)

    def host_loop(n):
        """The call a user makes: every pinned host batch goes through FusedTrainer.stage (H2D into one of the trainer's two
        preallocated staging sets on its copy stream) -- batch t + 1 is staged right after step t is enqueued, so it crosses
        PCIe under step t -- and the loss is read back after every step (train.py:202).  Exactly n copies and n reads."""
        def take():
            item = host[cursor[0] % NB]
            cursor[0] += 1
            return item
        tr.stage(*take())
        for i in range(n):
            loss = tr.step()
            if i + 1 < n:
                tr.stage(*take())
            float(loss)

    dev_loop(max(3, warmup))
    torch.cuda.synchronize()
    out = {}
    sampler = ClockSampler(env.local) if (env.rank == 0 and profile) else None
    # timed region 1: the K plain steps `value` is quoted on.  Timed region 2 (profile=True): the same K-step loop with a CUDA
    # event pair around every C-ABI call (and every collective) -- the live per-kernel durations behind `roofline`; its own
    # ms/step is reported as `ms_per_step_profiled` (the ~350 event records per step are not free)
    l0 = ops.LAUNCHES
    ms = env.timed(dev_loop, steps)
    launches = ops.LAUNCHES - l0
    first = cursor[0]
    if profile:
        ops.profile_start()
        tr.comm_prof = []
        out["ms_profiled"] = env.timed(dev_loop, steps)
        out["prof"] = ops.profile_stop()
        if env.world > 1:          # per-rank view: which rank's kernels make the others wait inside the collectives
            fam = {}
            for tag, (n, t) in out["prof"].items():
                f = tag.split("[")[0]
                fam[f] = fam.get(f, 0.0) + t / steps
            mine = {k: round(v, 4) for k, v in fam.items() if v > 0.1}
            per = {}
            for tag, a, b in tr.comm_prof:
                per[tag] = per.get(tag, 0.0) + a.elapsed_time(b) / steps
            mine.update({"comm:" + k: round(v, 4) for k, v in per.items()})
            allr = [None] * env.world
            env.dist.all_gather_object(allr, mine)
            out["per_rank_ms"] = allr
        out["comm_ms_per_step"] = sum(a.elapsed_time(b) for _, a, b in tr.comm_prof) / steps
        per = {}
        for tag, a, b in tr.comm_prof:
            per[tag] = per.get(tag, 0.0) + a.elapsed_time(b) / steps
        out["comm_ms_by_call"] = {k: round(v, 4) for k, v in per.items()}
        out["comm_calls_per_step"] = len(tr.comm_prof) / steps
        tr.comm_prof = None
        out["clocks"] = sampler.stop() if sampler else None
    out.update(ms=ms, launches=launches, batches_used=[(first + i) % NB for i in range(steps)],
               value=env.world * B * steps / (ms * 1e-3))
    if e2e:
        host_loop(2)
        ms2 = env.timed(host_loop, steps)
        out.update(ms_e2e=ms2, e2e=env.world * B * steps / (ms2 * 1e-3))
        host_loop_serial(2)
        out["ms_e2e_serial"] = env.timed(host_loop_serial, steps)      # the same without staging: copy, then step, on one stream
    return model, tr, out


def rowopt_report(env, tr, dev_b, cursor0, prof, steps):
    """How old are the rows the table optimizer has to catch up?  Gap = steps since a row (with Adam moments) was last brought
    up to date.  `stress`: every row with moments is aged by 2,000 steps (the revisit distance of the Zipf tail over a real
    epoch of ~1,000 steps) and one more batch is stepped: the catch-up kernel's time with the bounded replay."""
    from b200ctr import ops
    rep = {"catchup_ms_per_step": round(prof.get("rowopt_catchup", (0, 0.0))[1] / steps, 4),
           "update_ms_per_step": round(prof.get("rowopt_update", (0, 0.0))[1] / steps, 4),
           "sqnorm_ms_per_step": round(prof.get("rows_sqnorm", (0, 0.0))[1] / steps, 4)}
    if not tr.lazy or tr.shard:
        return rep
    name = "dare.emb_att.weight"
    st = tr.tables[name]["state"]
    bb, yy = dev_b[cursor0 % len(dev_b)]
    rows = torch.unique(bb["seq"][bb["seq"] != 0]).long()
    s = st[rows]
    mom = (s & 1) == 1
    gaps = (tr.global_step - (s >> 1))[mom].float()
    if gaps.numel():
        q = torch.quantile(gaps, torch.tensor([0.5, 0.9, 0.99, 1.0], device=gaps.device)).tolist()
        rep["next_batch"] = {"distinct_rows": int(rows.numel()), "rows_with_moments": int(mom.sum()),
                             "gap_steps_p50_p90_p99_max": [int(x) for x in q], "replayed_row_steps": int(gaps.sum())}
    # stress: age everything by 2000 steps in which no table row is touched (pure decay), then step one batch
    AGE = 2000
    lr = tr.lr_at(tr.global_step + 1)
    d = tr.ema_decay if tr.ema_on else 1.0
    for _ in range(AGE):
        tr.global_step += 1
        tr.hist.push(tr.global_step, lr, tr.wd, d)
    ops.profile_start()
    tr.step(bb, yy)
    p2 = ops.profile_stop()
    rep["stress_gap_2000"] = {"catchup_ms": round(p2.get("rowopt_catchup", (0, 0.0))[1], 4),
                              "note": "rows with moments replay 256 steps exactly, then the closed-form tail (csrc/rowopt.cu)"}
    return rep


def dp_parity(env):
    """N > 1: the reference's own 3-step result (tests/golden/model_qnn_concat.npz `final/*`, produced by the real reference on
    ONE global batch) against this job's data-parallel FusedTrainer with the batch split over the ranks and the tables
    row-sharded.  Returns the max scaled error over all parameters (rank 0)."""
    import b200ctr
    z = np.load(os.path.join(ROOT, "tests", "golden", "model_qnn_concat.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    g = cfg["golden"]
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    B = int(z["step0/y"].shape[0])
    if B % env.world:
        return {"skipped": f"golden batch {B} not divisible by {env.world} ranks"}
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cols)
    m.load_state_dict({k[5:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith("init/")}, strict=True)
    m = m.to(env.dev).train()
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=g["steps_per_epoch"], process_group=env.pg, shard_tables=True)
    lo, hi = env.rank * B // env.world, (env.rank + 1) * B // env.world
    losses = []
    for s in range(3):
        batch = {k: torch.from_numpy(np.array(z[f"step{s}/batch/{k}"]))[lo:hi] for k in ("X_num", "X_mask", "X_cat", "seq")}
        y = torch.from_numpy(np.array(z[f"step{s}/y"]))[lo:hi]
        losses.append(float(tr.step(batch, y)))
    sd = m.state_dict()                                  # collective: re-assembles the sharded tables
    worst, worst_name = 0.0, ""
    for k in z.files:
        if k.startswith("final/"):
            ref = torch.from_numpy(np.array(z[k])).double()
            e = float((sd[k[6:]].double().cpu() - ref).abs().max() / (ref.abs().max() + 1e-30))
            if e > worst:
                worst, worst_name = e, k[6:]
    loss_err = max(abs(losses[s] - float(z[f"step{s}/loss"])) for s in range(3))
    env._keep = getattr(env, "_keep", []) + [(m, tr)]    # peers hold IPC mappings of these shards: keep them alive
    return {"max_param_err_vs_reference_after_3_steps": worst, "worst_param": worst_name, "max_loss_err": loss_err,
            "ranks": env.world, "tolerance": 1e-3, "ok": bool(worst < 1e-3 and loss_err < 2e-4)}


def latest_traffic():
    """Newest committed ncu summary (profiles/r*_traffic.json): DRAM bytes per launch by kernel family."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return {}, None
    return json.load(open(files[-1])), os.path.basename(files[-1])


def rooflines(cfg, B, steps, prof, host, used, args):
    from b200ctr import configs
    pk = peaks()
    tot_ms = sum(v[1] for v in prof.values())
    fam = {}
    for tag, (n, t) in prof.items():
        f = tag.split("[")[0]
        a = fam.setdefault(f, [0, 0.0])
        a[0] += n; a[1] += t
    top = sorted(fam.items(), key=lambda kv: -kv[1][1])
    sq = cfg["sequence"]
    L, Kk, D = sq["max_len"], sq["top_k"], 64
    hid = sq.get("tfm", {}).get("ffn_hidden", 0) if sq["transformer_block"] else 0
    T = B * Kk
    F_all = 1 + 2 * configs.N_NUM + len(cfg["data"]["cat_cols"])
    gbytes = sum(gather_bytes(host[i][0]["seq"], L, Kk) for i in used)
    gemm_flops = {"gemm": 0.0, "gemm_bf16": 0.0, "gemm_tma": 0.0, "gemm_x3": 0.0}
    for tag, (n, t) in prof.items():
        for fn_ in gemm_flops:
            if tag.startswith(fn_ + "["):
                M_, N_, K_ = tag[len(fn_) + 1:-3].split("x")
                gemm_flops[fn_] += 2.0 * int(M_) * int(N_) * int(K_) * n
    keep_b = 4.0 * B * 8 * ((Kk + 15) // 16) * ((2 * ((Kk + 15) // 16) + 7) // 8) * 32      # attention keep-bit words per launch
    # bytes per qkv element on the wire between the in-projection and the attention kernels: fp16 when the packed-half attention
    # runs (ops.attn_takes_h16), fp32 otherwise
    from b200ctr import ops as _ops
    with _ops.precision("bf16" if cfg.get("amp") == "bf16" else "fp32"):
        qkv_b = 2.0 if _ops.attn_takes_h16(Kk, 8, D) else 4.0
    models = {   # family -> (bound, work per launch [FLOP or byte], description); "both": the BINDING roof is reported
        "attn_tc_bwd": ("both", (10.0 * Kk * Kk * D * B, T * (qkv_b * 3 * D + 4.0 * D + 4.0 * D + 2.0 * 3 * D) + 4.0 * B * 8 * Kk + keep_b),
                        "attn_h16_bwd_kernel (packed-half attention backward, head_dim 8: fp16-accumulate score MMAs, half2 exp / mask / dS, "
                        "fp32-accumulate dQ / dK / dV MMAs; one pass = 10 K^2 D FLOP/sample; fp16 qkv, o, d_o, lse, keep bits in, bf16 dqkv out)"),
        "attn_tc_fwd": ("both", (4.0 * Kk * Kk * D * B, T * (qkv_b * 3 * D + 4.0 * D) + 4.0 * B * 8 * Kk + keep_b),
                        "attn_h16_fwd_kernel (packed-half attention forward: QK^T + bias in one fp16-accumulate MMA, half2 softmax / dropout, "
                        "PV and the row sums on fp32-accumulate MMAs = 4 K^2 D FLOP/sample; fp16 qkv in, o, lse, keep bits out)"),
        "ffn_fwd": ("both", (4.0 * T * D * hid, T * (4.0 * D * 3 + 4.0 + hid / 8.0)),
                    "ffn_fwd_kernel (tcgen05 fused FFN block: 2 contractions of T x 64 x ffn_hidden, half2 GELU / dropout epilogue, fp16 "
                    "activation operand; x1 in, s2, x2, r2, keep bits out)"),
        "ffn_bwd": ("both", (8.0 * T * D * hid, T * (4.0 * D * 4 + 4.0 + hid / 8.0)),
                    "ffn_bwd1/ffn_bwd2 (tcgen05 fused FFN backward: 4 algorithmic contractions, the recompute and the internal bf16 "
                    "dpre spill are not counted; dy, s2, x1, r2, keep bits in, dx1 out)"),
        "inproj_fwd": ("hbm", T * (4.0 * D + qkv_b * 3 * D), "lin_fwd_kernel<192> (x in, qkv out -- fp16 for the packed-half attention)"),
        "outproj_norm_fwd": ("hbm", 4.0 * T * D * 4 + 4.0 * T, "lin_fwd_kernel<64,norm> (o, x in; s1, x1, r1 out)"),
        "outproj_norm_bwd": ("hbm", 4.0 * T * D * 5, "lin_bwd_kernel<64,norm> (dx1, s1, o in; d_s1, d_o out)"),
        "inproj_bwd": ("hbm", 2.0 * T * 3 * D + 4.0 * T * 3 * D, "lin_bwd_kernel<192> (bf16 dqkv, x, d_s1 in; dx out)"),
        "fm_fused_fwd": ("hbm", 2.0 * B * F_all * D, "fm_fwd_kernel (bf16 z in)"),
        "fm_fused_bwd": ("hbm", B * F_all * D * (2.0 + 4.0 + 4.0), "fm_bwd_kernel (bf16 z, fp32 dz in / out)"),
        "retrieve_fwd": ("hbm", gbytes / max(1, len(used)), "retrieve_fwd_warp_kernel (8L + 4D*l + 4D*min(K,l) + 4D + 4D*K + 12K bytes/sample; every logical row use counted once)"),
    }
    traffic, traffic_src = latest_traffic()
    if not (args.cfg == "dare_qnn_next" and B == 8192 and cfg.get("amp") == "bf16"):
        traffic = {}
    gdesc = {"gemm_bf16": "gemm_tc_kernel / gemm_tc_ws_kernel (tcgen05.mma kind::f16, fp32 TMEM accumulators); the remaining small dense "
                          "projections of the step",
             "gemm_tma": "gemm_tma_kernel (cp.async.bulk.tensor SWIZZLE_128B ring -> tcgen05.mma, double-buffered TMEM): the deep-K first "
                         "QNN MLP layer, forward and both backward products",
             "gemm_x3": "gemm_tc_kernel<X3> / gemm_tc_ws_kernel<X3> (tcgen05 split precision: 3 bf16 MMAs per fp32 product; FLOPs counted once)",
             "gemm": "gemm_kernel (fp32 CUDA-core contraction)"}

    def roof_of(name):
        n, t = fam[name]
        if name in gemm_flops:
            ach = gemm_flops[name] / (t * 1e-3) / 1e12
            return {"kernel": gdesc[name], "bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"],
                    "traffic": None, "peak_source": f"{pk['src']} bf16 sustained", "launches": n, "avg_launch_ms": t / n,
                    "share_of_step": t / tot_ms}
        bound, work, desc = models[name]
        other = None
        if bound == "both":          # pick the binding roof: the one with the larger time floor
            fl, by = work
            t_fl, t_by = fl / (pk["tf_sus"] * 1e12), by / (pk["hbm"] * 1e9)
            if t_by >= t_fl:
                bound, work = "hbm", by
                other = {"bound": "tensor", "frac": fl * n / (t * 1e-3) / 1e12 / pk["tf_sus"]}
            else:
                bound, work = "tensor", fl
                other = {"bound": "hbm", "frac": by * n / (t * 1e-3) / 1e9 / pk["hbm"]}
        tr = traffic.get("ffn_bwd1" if name == "ffn_bwd" else name, {}).get("traffic_bytes_per_launch")
        if name == "ffn_bwd" and tr is not None:
            tr += traffic.get("ffn_bwd2", {}).get("traffic_bytes_per_launch", 0)
        common = {"kernel": desc, "traffic": tr, "traffic_source": traffic_src if tr is not None else None, "launches": n,
                  "avg_launch_ms": t / n, "share_of_step": t / tot_ms}
        if bound == "tensor":
            ach = work * n / (t * 1e-3) / 1e12
            out = dict(common, bound="tensor", achieved=ach, peak=pk["tf_sus"], unit="TFLOP/s", frac=ach / pk["tf_sus"],
                       peak_source=f"{pk['src']} bf16 sustained")
        else:
            ach = work * n / (t * 1e-3) / 1e9
            out = dict(common, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                       frac_of_nominal_8TBs=ach / 8000.0, peak_source=f"{pk['src']} HBM copy")
            if tr is not None:
                out["dram_actual_GBs"] = tr / (t / n * 1e-3) / 1e9
        if other is not None:
            out["non_binding_roof"] = other
        return out

    modelled = [k for k, _ in top if k in models or k in gemm_flops]
    roof = roof_of(modelled[0]) if modelled else None
    if roof is not None and modelled[0].startswith("attn_tc"):
        roof["note"] = ("HBM is the binding roof (head_dim 8: the tensor floor is 9x lower); the kernel is bound by the per-score work "
                        "and, since the packed-half rewrite, by the latency of its per-sample staging phase (ncu r02: issue 43-57 %)")
    return {"roofline": roof, "gather_roofline": roof_of("retrieve_fwd") if "retrieve_fwd" in fam else None,
            "kernel_rooflines": [dict(roof_of(k), family=k) for k in modelled[:14]],
            "kernel_time_share": {k: round(v[1] / tot_ms, 4) for k, v in top[:14]},
            "kernel_ms_per_step": {k: round(v[1] / steps, 4) for k, v in top[:16]},
            "gemm_shapes_ms_per_step": {k: [v[0] // steps, round(v[1] / steps, 4)] for k, v in
                                        sorted(((t, v) for t, v in prof.items() if t.startswith("gemm")), key=lambda kv: -kv[1][1])[:10]}}, top


def run_cuda(args):
    import b200ctr
    from b200ctr import ops
    env = Env()
    world, rank, dev = env.world, env.rank, env.dev
    b200ctr._lib.check(b200ctr.load().b200ctr_check_device(), "device check")
    cfg = build_cfg(args.cfg)
    cfg["amp"] = args.amp
    B = args.batch or cfg["train"]["batch_size"]
    vocab = args.vocab
    NB = args.batches
    keep = []                         # N > 1: peers hold IPC mappings of every trainer's shards -> nothing is freed before exit

    parity = dp_parity(env) if world > 1 else None

    host, dev_b = make_batches(cfg, B, NB, rank, dev)
    h2d = sum(v.numel() * v.element_size() for v in host[0][0].values()) + host[0][1].numel() * 4
    model, tr, r = measure(env, cfg, vocab, dev_b, host, args.steps, args.warmup, profile=True, e2e=True, replicate=args.replicate_tables)
    line = None
    if rank == 0:
        roofs, top = rooflines(cfg, B, args.steps, r["prof"], host, r["batches_used"], args)
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.amp == "bf16" else "fp32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, B), "global_batch": B * world,
                       "parallelism": (f"dp{world}" if world == 1 else
                                       f"dp{world} over the batch; seq tables " + ("replicated" if args.replicate_tables else
                                       f"row-sharded {world}-way, gathered/scattered through NVLink peer memory inside the kernels")),
                       "l2_policy": f"{NB} distinct batches, each used at most once per {NB} steps; every step touches >10 GB (two 2.56 GB "
                                    f"tables + optimizer state + activations), far above the 126 MB L2",
                       "seq_vocab": vocab},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e"], "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": r["ms_e2e"] / args.steps,
                    "api": "tr.stage(batch_0); for t: loss = tr.step(); tr.stage(batch_t+1); float(loss) -- FusedTrainer.stage copies the "
                           "pinned host batch into one of two preallocated device staging sets on the trainer's copy stream (batch t+1 "
                           "crosses PCIe under step t), the loss is read back after every step",
                    "ms_per_step_serial_copy": r["ms_e2e_serial"] / args.steps},
            "gpu_launches": int(r["launches"]),
            "ms_per_step_profiled": r["ms_profiled"] / args.steps,
        }
        line.update(roofs)
        if world > 1:
            line["comm_ms_per_step"] = round(r["comm_ms_per_step"], 4)
            line["comm_ms_by_call"] = r["comm_ms_by_call"]
            line["comm_calls_per_step"] = r["comm_calls_per_step"]
            line["per_rank_ms"] = r["per_rank_ms"]
            line["dp_parity"] = parity
        if os.environ.get("B200CTR_BENCH_DUMP"):      # full per-op table (development aid): every family and GEMM shape
            with open(os.environ["B200CTR_BENCH_DUMP"], "w") as fh:
                json.dump({"families_ms_per_step": {k: [v[0] // args.steps, round(v[1] / args.steps, 4)] for k, v in top},
                           "ops_ms_per_step": {k: [v[0] // args.steps, round(v[1] / args.steps, 4)]
                                               for k, v in sorted(r["prof"].items(), key=lambda kv: -kv[1][1])}}, fh, indent=1)
    # ---- the table optimizer's real cost: gap histogram of the next batch + the aged-rows stress (N = 1) ----
    if world == 1 and not args.quick:
        rep = rowopt_report(env, tr, dev_b, 2 * args.steps + max(3, args.warmup) + 2 * (2 + args.steps), r["prof"], args.steps)
        if rank == 0:
            line["rowopt"] = rep
    # ---- cold gather: token ids uniform over all 10 M rows (no Zipf-hot L2 hits) ----
    if world == 1 and not args.quick:
        _, cold_dev = make_batches(cfg, B, 2, rank, dev, uniform_tokens=True, vocab=vocab)
        tr.step(*cold_dev[0])
        ops.profile_start()
        tr.step(*cold_dev[1])
        p = ops.profile_stop()
        n, t = p["retrieve_fwd"]
        gb = gather_bytes(cold_dev[1][0]["seq"].cpu(), cfg["sequence"]["max_len"], cfg["sequence"]["top_k"])
        pk = peaks()
        traffic, tsrc = latest_traffic()
        cold = {"kernel": "retrieve_fwd_warp_kernel on token ids uniform over all 10,000,000 rows (every row a cold miss; ~1.7 M distinct "
                          "rows per batch)", "bound": "hbm", "achieved": gb / (t / n * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": gb / (t / n * 1e-3) / 1e9 / pk["hbm"], "avg_launch_ms": t / n,
                "traffic": traffic.get("retrieve_fwd_cold", {}).get("traffic_bytes_per_launch"), "traffic_source": tsrc}
        line["gather_roofline_cold"] = cold
        del cold_dev
    if world == 1:
        del model, tr
        torch.cuda.empty_cache()
    else:
        keep.append((model, tr))
    # ---- the `amp: none` mode (every shipped cfg) on the same workload ----
    if not args.quick and args.amp == "bf16":
        cfg32 = build_cfg(args.cfg)
        cfg32["amp"] = "none"
        n32 = max(2, min(4, args.steps))
        m32, t32, r32 = measure(env, cfg32, vocab, dev_b[:16], host[:16], n32, 3, profile=True, replicate=args.replicate_tables)
        if rank == 0:
            fam = {}
            for tag, (n, t) in r32["prof"].items():
                f = tag.split("[")[0]
                fam[f] = fam.get(f, 0.0) + t
            line["fp32"] = {"value": r32["value"], "unit": UNIT, "ms_per_step": r32["ms"] / n32, "steps": n32,
                            "ratio_to_bf16": r32["value"] / line["value"], "contractions": f"FP32_GEMM={ops.FP32_GEMM} (x3 = tcgen05 split "
                            f"precision, 3 bf16 MMAs per fp32 product; the fused bf16 kernels are not used in this mode)",
                            "kernel_ms_per_step": {k: round(v / n32, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:10]}}
        if world == 1:
            del m32, t32
            torch.cuda.empty_cache()
        else:
            keep.append((m32, t32))
    # ---- BASELINE configs[2] (the scaling config) at the same N ----
    if not args.quick and args.cfg == "dare_qnn_next":
        cfg2 = build_cfg("dare_qnn_next_k100_s1")
        cfg2["amp"] = args.amp
        n2 = max(3, min(6, args.steps))
        m2, t2, r2 = measure(env, cfg2, vocab, dev_b[:16], host[:16], n2, 3, profile=world > 1, replicate=args.replicate_tables)
        if rank == 0:
            line["scale_config"] = {"workload": workload_name(cfg2, B), "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms"] / n2,
                                    "steps": n2, "n_gpus": world}
            if world > 1:
                line["scale_config"]["comm_ms_per_step"] = round(r2["comm_ms_per_step"], 4)
        keep.append((m2, t2))
    if rank == 0:
        # newest record first: r02z was re-measured with the packed-half attention / half2 FFN epilogues (fp32 and bf16 arms; the
        # CPU-reference arm is in r02_amp_delta.json)
        ad = next((p_ for p_ in (os.path.join(ROOT, "profiles", f) for f in ("r02z_amp_delta.json", "r02_amp_delta.json")) if os.path.exists(p_)), None)
        if ad is not None:
            d = json.load(open(ad))
            line["amp_delta_one_epoch"] = {"source": f"profiles/{os.path.basename(ad)} (tools/amp_delta.py)", "bf16_minus_fp32": d["delta"]["bf16_minus_fp32"],
                                           "fp32_seed_noise": d["delta"]["fp32_seed_noise"], "stated_tolerance": d["stated_tolerance"],
                                           "within_tolerance": d["within_tolerance"]}
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference(args.cfg, 2, 1, sample_B=args.ref_batch)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200ctr", choices=["b200ctr", "reference"])
    ap.add_argument("--cfg", default="dare_qnn_next")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug only; invalidates the headline)")
    ap.add_argument("--batches", type=int, default=64, help="distinct synthetic batches the steps walk through")
    ap.add_argument("--vocab", type=int, default=10_000_000)
    ap.add_argument("--ref-batch", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + e2e only (no fp32 / scale_config / rowopt / cold-gather legs)")
    ap.add_argument("--replicate-tables", action="store_true",
                    help="N > 1: keep full replicas of the two sequence tables on every GPU (all-gather of gradient rows) "
                         "instead of the default row-sharded tables gathered / scattered through peer memory")
    ap.add_argument("--amp", default="bf16", choices=["none", "bf16"],
                    help="matmul precision of the headline step: bf16 = the reference's `amp: bf16` autocast mode on the tcgen05 tensor "
                         "cores (fp32 accumulate, fp32 master weights/optimizer); none = the fp32 mode (split-precision tcgen05)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: libraries that write to fd 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set) are sent to stderr, the JSON line goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
