#!/usr/bin/env python
"""north_star acceptance check for the bf16 mode: "bf16 AMP within a stated AP/WLL delta after one epoch".

    python tools/amp_delta.py [--rows 98304] [--val-rows 16384] [--batch 1024] [--out profiles/r02_amp_delta.json]

One epoch of cfgs/dare_qnn_next.yaml (all shipped widths, dropout ON as shipped) on synthetic Toss-shaped data whose labels
carry a planted signal (b200ctr.synth.planted_labels), from the same seeded initial weights and the same batch order, with
  (a) FusedTrainer `amp: none`  (fp32 kernels)                  -- twice, with two dropout seeds = the run-to-run noise floor
  (b) FusedTrainer `amp: bf16`  (tcgen05 / tensor-core kernels) -- the mode bench.py reports
  (c) the UNMODIFIED reference CTRModel + the literal step body of src/train.py:152-199 on the host CPU (baseline/_ref),
then AP and the 50:50 weighted log-loss (src/utils/metrics.py:8-29) on held-out rows, evaluated in fp32 like the reference's
validation loop (train.py:212-226).  Deviations from the shipped cfg, all stated in the output: B = 1024 and lr = 1e-3 with a
quarter-epoch warm-up (96 steps must learn something), seq_vocab = 1,000,001 (covers the synthetic token range; the CPU arm's
dense AdamW over 2 x 10 M rows would take 10 s per step), EMA off (decay 0.999 over 96 steps would evaluate the initial weights).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "toss-next-ctr-prediction_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def weighted_logloss(y, p):
    """src/utils/metrics.py:8-24 (50:50 class-weighted log-loss)."""
    p = np.clip(p.astype(np.float64), 1e-15, 1 - 1e-15)
    pos, neg = y == 1, y == 0
    lp = -np.log(p[pos]).mean() if pos.any() else 0.0
    ln = -np.log(1 - p[neg]).mean() if neg.any() else 0.0
    return 0.5 * lp + 0.5 * ln


def metrics(y, p):
    from sklearn.metrics import average_precision_score
    return {"AP": float(average_precision_score(y, p)), "WLL": float(weighted_logloss(y, p))}


def make_data(cfg, n_rows, n_val, B, seed):
    from b200ctr import configs, synth
    schema = synth.ToySchema(cfg, configs.N_NUM, configs.N_NUM, seed=seed)
    cols = cfg["data"]["cat_cols"]

    def chunk(n, s):
        b, _ = schema.batch(n, seed=s)
        y = synth.planted_labels(b, cols, np.random.default_rng(s + 1), cfg["sequence"]["query_key"])
        return b, y
    train = [chunk(B, seed + 10 + i) for i in range(n_rows // B)]
    val = [chunk(B, seed + 100_000 + i) for i in range(n_val // B)]
    return train, val


def to_t(b):
    return {k: torch.from_numpy(v) for k, v in b.items()}


def run_fused(cfg, amp, drop_seed, train, val, vocab, dev):
    import b200ctr
    from b200ctr import configs
    c = copy.deepcopy(cfg)
    c["amp"] = amp
    torch.manual_seed(c["seed"])
    model = b200ctr.CTRModel(c, vocab, configs.N_NUM, configs.N_NUM, configs.cardinals(c), c["data"]["cat_cols"]).to(dev).train()
    tr = b200ctr.FusedTrainer(model, c, steps_per_epoch=len(train), seed=drop_seed)
    t0 = time.time()
    losses = []
    for b, y in train:
        losses.append(tr.step(to_t(b), torch.from_numpy(y)))
    torch.cuda.synchronize()
    dt = time.time() - t0
    ps = []
    for b, _ in val:                                   # fp32 evaluation, like the reference's validation loop
        _, p, _ = tr.predict(to_t(b))
        ps.append(p.cpu().numpy())
    yv = np.concatenate([y for _, y in val])
    out = metrics(yv, np.concatenate(ps))
    out.update(train_loss_mean=float(torch.stack(losses).mean().cpu()), train_seconds=dt)
    del tr, model
    torch.cuda.empty_cache()
    return out


def run_reference(cfg, train, val, vocab):
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "src")):
        return {"unavailable": "baseline/_ref missing (run oracle/install_ref.py in the build container)"}
    sys.path.insert(0, ref)
    from src.models.wrapper import CTRModel            # the unmodified reference
    from src.train import bce_wll_style
    from src.utils.sched import cosine_warmup_lr
    from b200ctr import configs
    torch.set_num_threads(os.cpu_count() or 1)
    c = copy.deepcopy(cfg)
    torch.manual_seed(c["seed"])
    model = CTRModel(c, vocab, configs.N_NUM, configs.N_NUM, configs.cardinals(c), c["data"]["cat_cols"]).train()
    opt = torch.optim.AdamW(model.parameters(), lr=c["train"]["lr"], weight_decay=c["train"]["weight_decay"])
    aux_w = float(c["model"]["qnn_alpha"].get("aux_head_weight", 0.0))
    spe = len(train)
    t0 = time.time()
    losses = []
    for step, (b, y) in enumerate(train):              # src/train.py:152-199, literally
        yt = torch.from_numpy(y)
        opt.param_groups[0]["lr"] = cosine_warmup_lr(0, step, spe, c["train"]["lr"], c["train"]["warmup_epochs"], c["train"]["epochs"])
        opt.zero_grad(set_to_none=True)
        logits, prob, aux = model(to_t(b))
        loss = bce_wll_style(logits, yt)
        if aux_w > 0:
            loss = loss + aux_w * bce_wll_style(aux, yt)
        loss.backward()
        if c["train"]["grad_clip_norm"] > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), c["train"]["grad_clip_norm"])
        opt.step()
        losses.append(float(loss.detach()))
    dt = time.time() - t0
    model.eval()
    ps = []
    with torch.no_grad():
        for b, _ in val:
            ps.append(model(to_t(b))[1].numpy())
    yv = np.concatenate([y for _, y in val])
    out = metrics(yv, np.concatenate(ps))
    out.update(train_loss_mean=float(np.mean(losses)), train_seconds=dt, cores=torch.get_num_threads())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=98_304)
    ap.add_argument("--val-rows", type=int, default=16_384)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--vocab", type=int, default=1_000_001)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_amp_delta.json"))
    args = ap.parse_args()
    from b200ctr import configs
    cfg = configs.dare_qnn_next()
    cfg["train"].update(batch_size=args.batch, lr=1e-3, warmup_epochs=0.25, epochs=1)
    cfg["ema"] = {"enabled": False}
    dev = torch.device("cuda:0")
    train, val = make_data(cfg, args.rows, args.val_rows, args.batch, seed=cfg["seed"])
    yv = np.concatenate([y for _, y in val])
    res = {"workload": f"cfgs/dare_qnn_next.yaml widths, one epoch = {len(train)} steps x B={args.batch} ({args.rows} rows), "
                       f"held-out {args.val_rows} rows, planted-signal synthetic labels (base rate {float(yv.mean()):.4f}), dropout on, "
                       f"lr 1e-3 with a quarter-epoch warm-up, seq_vocab {args.vocab}, EMA off, fp32 evaluation",
           "prior_AP": float(yv.mean())}
    res["fp32_seed_a"] = run_fused(cfg, "none", 777, train, val, args.vocab, dev)
    res["fp32_seed_b"] = run_fused(cfg, "none", 4242, train, val, args.vocab, dev)
    res["bf16_seed_a"] = run_fused(cfg, "bf16", 777, train, val, args.vocab, dev)
    res["bf16_seed_b"] = run_fused(cfg, "bf16", 4242, train, val, args.vocab, dev)
    if not args.no_reference:
        res["reference_cpu_fp32"] = run_reference(cfg, train, val, args.vocab)
    f32 = [res["fp32_seed_a"], res["fp32_seed_b"]]
    b16 = [res["bf16_seed_a"], res["bf16_seed_b"]]
    mean = lambda rs, k: float(np.mean([r[k] for r in rs]))     # noqa: E731
    res["delta"] = {
        "bf16_minus_fp32": {k: mean(b16, k) - mean(f32, k) for k in ("AP", "WLL")},
        "fp32_seed_noise": {k: abs(f32[0][k] - f32[1][k]) for k in ("AP", "WLL")},
        "bf16_seed_noise": {k: abs(b16[0][k] - b16[1][k]) for k in ("AP", "WLL")},
    }
    if "AP" in res.get("reference_cpu_fp32", {}):
        r = res["reference_cpu_fp32"]
        res["delta"]["fp32_minus_reference"] = {k: mean(f32, k) - r[k] for k in ("AP", "WLL")}
        res["delta"]["bf16_minus_reference"] = {k: mean(b16, k) - r[k] for k in ("AP", "WLL")}
    res["stated_tolerance"] = {"AP": 0.01, "WLL": 0.01,
                               "meaning": "|bf16 - fp32| (means over two dropout seeds) after one epoch; the seed-to-seed spread of "
                                          "the fp32 path itself is reported beside it"}
    res["within_tolerance"] = bool(abs(res["delta"]["bf16_minus_fp32"]["AP"]) <= 0.01 and abs(res["delta"]["bf16_minus_fp32"]["WLL"]) <= 0.01)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
