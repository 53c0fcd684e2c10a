#!/usr/bin/env python
"""BASELINE configs[4] end to end on the b200ctr fast tier: K-fold training + EMA + fold `logit_mean` inference.

    python tools/kfold_pipeline.py [--rows 2000000] [--test-rows 500000] [--folds 5] [--epochs 1] [--cfg dare_qnn_next]

Mirrors the reference's drivers with the fused pieces in place of the per-row loader, the eager step and the per-fold loop:
  src/train.py:333-352   StratifiedGroupKFold(max(5, n_splits), shuffle, seed) over (y, groups) of the cache  -> same call here
  src/train.py:100-110   ShardedDataset + DataLoader(collate_sharded)                                          -> BatchedShardReader (device form)
  src/train.py:152-203   step body (lr schedule, fwd, bce_wll, bwd, clip, AdamW, EMA)                         -> FusedTrainer.step
  src/train.py:206-226   validation with the EMA weights, AP / WLL (utils/metrics.py:8-29)                     -> EMA shadows copied into an eval
                                                                                                                   CTRModel, FoldEnsemble of one
  src/infer.py:82-162    per-fold models with EMA weights, clamp, ensemble_probs("logit_mean")                 -> FoldEnsemble.predict
The synthetic cache is written in the reference's App.-D format (b200ctr.synth.write_cache, planted-signal labels).
Prints one JSON line with per-fold AP / WLL, train samples/s (whole epoch, loader + H2D + step) and inference samples/s.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "toss-next-ctr-prediction_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def wll(y, p):
    p = np.clip(p.astype(np.float64), 1e-15, 1 - 1e-15)
    pos, neg = y == 1, y == 0
    return 0.5 * (-np.log(p[pos]).mean() if pos.any() else 0.0) + 0.5 * (-np.log(1 - p[neg]).mean() if neg.any() else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--test-rows", type=int, default=500_000)
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--cfg", default="dare_qnn_next")
    ap.add_argument("--amp", default="bf16")
    ap.add_argument("--cache", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_kfold_pipeline.json"))
    args = ap.parse_args()
    import b200ctr
    from b200ctr import configs, synth
    from sklearn.metrics import average_precision_score
    from sklearn.model_selection import StratifiedGroupKFold
    cfg = getattr(configs, args.cfg)()
    cfg["amp"] = args.amp
    cfg["train"]["epochs"] = args.epochs
    cfg["train"]["warmup_epochs"] = min(cfg["train"]["warmup_epochs"], 0.25)
    dev = torch.device("cuda:0")
    B = cfg["train"]["batch_size"]
    cache = args.cache or tempfile.mkdtemp(prefix="kfold_cache_")
    t0 = time.time()
    mtr, mte = synth.write_cache(cache, cfg, args.rows, args.test_rows, configs.N_NUM, shard_rows=500_000, seed=cfg["seed"])
    t_cache = time.time() - t0
    man = json.load(open(mtr))
    y_all = np.concatenate([np.load(m["y"]["path"]) for m in man["shards"]])
    g_all = np.concatenate([np.load(m["groups"]["path"]) for m in man["shards"]])
    sgkf = StratifiedGroupKFold(n_splits=max(5, args.folds), shuffle=True, random_state=cfg["seed"])      # train.py:334-336
    cols = cfg["data"]["cat_cols"]
    folds, fold_models = [], []
    for fold, (tr_idx, va_idx) in enumerate(sgkf.split(np.zeros_like(y_all), y_all, g_all)):
        if fold >= args.folds:
            break
        torch.manual_seed(cfg["seed"] + fold)
        with torch.device(dev):
            model = b200ctr.CTRModel(cfg, configs.SEQ_VOCAB, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cols)
        model.train()
        rd = b200ctr.BatchedShardReader(mtr, tr_idx, train=True, batch_size=B, shuffle=True, seed=cfg["seed"] + fold, drop_last=True, device=dev)
        trn = b200ctr.FusedTrainer(model, cfg, steps_per_epoch=len(rd))
        torch.cuda.synchronize()
        t0 = time.time()
        n_seen = 0
        loss = None
        for _ in range(args.epochs):
            for b in rd:
                yb = b.pop("y")
                loss = trn.step(b, yb)
                n_seen += yb.shape[0]
        last_loss = float(loss)                                 # one host sync per fold, not per step
        torch.cuda.synchronize()
        t_train = time.time() - t0
        # validation with the EMA weights (train.py:206-226): an eval model that holds the shadows
        ema = trn.ema_state() if trn.ema_on else None
        trn.sync_tables()
        with torch.device(dev):
            ev = b200ctr.CTRModel(cfg, configs.SEQ_VOCAB, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cols)
        src = ema if ema else {n: p for n, p in model.named_parameters()}
        with torch.no_grad():
            for n, p in ev.named_parameters():
                p.copy_(src[n] if n in src else dict(model.named_parameters())[n])
        ev.eval()
        ens = b200ctr.FoldEnsemble([ev])
        vr = b200ctr.BatchedShardReader(mtr, va_idx, train=True, batch_size=B, shuffle=False, device=dev)
        ps, ys = [], []
        t0 = time.time()
        for b in vr:
            ys.append(b.pop("y").cpu().numpy())
            ps.append(ens.predict(b).cpu().numpy())
        t_val = time.time() - t0
        yv, pv = np.concatenate(ys), np.concatenate(ps)
        folds.append({"fold": fold, "train_rows": int(len(tr_idx)), "val_rows": int(len(va_idx)), "steps": int(trn.global_step),
                      "last_loss": last_loss, "val_AP": float(average_precision_score(yv, pv)), "val_WLL": float(wll(yv, pv)),
                      "train_samples_per_s": n_seen / t_train, "val_samples_per_s": len(va_idx) / t_val})
        fold_models.append(ev)
        del trn, model, rd
        torch.cuda.empty_cache()
    # inference: all folds, one H2D per batch, clamp + logit_mean fused (infer.py:104-158)
    ens = b200ctr.FoldEnsemble(fold_models)
    te = b200ctr.BatchedShardReader(mte, np.arange(args.test_rows), train=False, batch_size=B, shuffle=False, device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    preds, ids = [], []
    for b in te:
        ids.append(b["ids"])
        preds.append(ens.predict(b))
    p = torch.cat(preds).cpu().numpy()
    torch.cuda.synchronize()
    t_inf = time.time() - t0
    res = {"workload": f"cfgs/{cfg['exp_name']}.yaml widths, amp={args.amp}, {args.folds}-fold StratifiedGroupKFold over {args.rows} synthetic rows "
                       f"(App.-D cache, planted-signal labels, base rate {float(y_all.mean()):.4f}), {args.epochs} epoch(s) per fold at B={B}, EMA "
                       f"{'on' if (cfg.get('ema') or {}).get('enabled') else 'off'}, fold logit_mean inference over {args.test_rows} test rows; "
                       f"BASELINE configs[4] names ~10 M rows: this run is the same pipeline on a smaller cache",
           "cache_write_seconds": t_cache, "folds": folds,
           "inference": {"rows": int(args.test_rows), "folds": len(fold_models), "samples_per_s": args.test_rows / t_inf,
                         "fold_forwards_per_s": args.test_rows * len(fold_models) / t_inf, "mean_p": float(p.mean()),
                         "first_ids": [str(x) for x in np.concatenate(ids)[:3]]}}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
