"""Shared set-up of the drop-in integration runs (test infrastructure): the reference's own `python -m src.train` and
`python -m src.infer` (src/train.py:319-366, src/infer.py:10-180), unmodified, on a synthetic App.-D cache.

Two arms use it:
  * oracle/make_e2e_golden.py -- the UNMODIFIED reference alone on the CPU (build container) -> tests/golden/e2e_*.json
  * tests/test_gpu_dropin_e2e.py -- the same two commands on the B200 box with this repo's overlay in front of the
    reference copy on PYTHONPATH, so `src.models.*` resolve to b200ctr and everything else to the reference.
Both arms see identical shards, the same YAML and the same seed; with dropout off they see identical batches too.
"""
from __future__ import annotations

import csv
import json
import os
import subprocess
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "toss-next-ctr-prediction_b200")
OVERLAY = os.path.join(PKG, "overlay")
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
N_NUM = 82

# name -> (reference yaml, rows train, rows test, overrides)
CASES = {
    # BASELINE configs[0]: dare_base + the keys the shipped yaml lacks (manifest_*, ensemble).  Its FC head has
    # a hard-coded Dropout(0.1) (wrapper.py:98), so this case is compared statistically, not bit-for-bit.
    "dare_base": ("dare_base.yaml", 20_000, 4_000, {
        "cv.n_splits": 1, "train.epochs": 1, "train.batch_size": 2048, "num_workers": 2,
        "ensemble": {"method": "logit_mean", "trim_ratio": 0.2, "weights": []}, "logging.tb": False,
        "model.dare_dropout": 0.0}),
    # BASELINE configs[1] shrunk in rows only (the widths are the shipped ones): every dropout is configurable -> 0, so
    # the overlay run and the reference run are the same computation up to float rounding.
    "dare_qnn_next": ("dare_qnn_next.yaml", 12_000, 3_000, {
        "cv.n_splits": 1, "train.epochs": 1, "train.batch_size": 1024, "num_workers": 2, "logging.tb": False,
        "train.lr": 1e-3, "train.warmup_epochs": 1,
        "model.dare_dropout": 0.0, "model.qnn_alpha.dropout": 0.0, "sequence.tfm.mha_dropout": 0.0,
        "sequence.tfm.ffn_dropout": 0.0}),
}


def _set(cfg: dict, dotted: str, value):
    node = cfg
    parts = dotted.split(".")
    for p in parts[:-1]:
        node = node.setdefault(p, {})
    node[parts[-1]] = value


def prepare(case: str, work: str, ref_root: str, device: str) -> str:
    """Write the synthetic cache + the cfg yaml under `work`; returns the cfg path."""
    sys.path.insert(0, PKG)
    from b200ctr import synth
    yaml_name, n_train, n_test, over = CASES[case]
    with open(os.path.join(ref_root, "cfgs", yaml_name)) as fh:
        cfg = yaml.safe_load(fh)
    for k, v in over.items():
        _set(cfg, k, v)
    cfg["device"] = device
    cache = os.path.join(work, "cache")
    mtr, mte = synth.write_cache(cache, cfg, n_train, n_test, N_NUM, shard_rows=8_000, seed=int(cfg["seed"]))
    _set(cfg, "data.manifest_train", mtr)
    _set(cfg, "data.manifest_test", mte)
    _set(cfg, "logging.log_dir", os.path.join(work, "runs"))
    path = os.path.join(work, f"{case}.yaml")
    with open(path, "w") as fh:
        yaml.safe_dump(cfg, fh)
    return path


def run_module(mod: str, cfg_path: str, pythonpath: list, timeout: int, log_path: str) -> None:
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(pythonpath)
    env.setdefault("TQDM_DISABLE", "1")
    with open(log_path, "w") as log:
        r = subprocess.run([sys.executable, "-m", mod, "--cfg", cfg_path], env=env, stdout=log, stderr=subprocess.STDOUT,
                           timeout=timeout, cwd=os.path.dirname(cfg_path))
    if r.returncode != 0:
        tail = open(log_path).read()[-4000:]
        raise RuntimeError(f"`python -m {mod}` failed (rc={r.returncode}); log tail:\n{tail}")


def collect(case: str, work: str) -> dict:
    """What both arms are compared on: the train_log.csv rows, the submission, the checkpoint's layout."""
    import torch
    cfg = yaml.safe_load(open(os.path.join(work, f"{case}.yaml")))
    out_dir = os.path.join(cfg["logging"]["log_dir"], cfg["exp_name"])
    rows = list(csv.DictReader(open(os.path.join(out_dir, "train_log.csv"))))
    log = [{k: r[k] for k in ("fold", "epoch", "split", "loss", "AP", "WLL", "Score", "lr", "bs", "K", "tau")} for r in rows]
    sub = open(os.path.join(out_dir, "submission.csv")).read().strip().splitlines()
    ids = [ln.split(",")[0] for ln in sub[1:]]
    probs = [float(ln.split(",")[1]) for ln in sub[1:]]
    sys.path.insert(0, REF_COPY if os.path.isdir(os.path.join(REF_COPY, "src")) else "/root/reference")   # unpickles Calibrator
    ck = torch.load(os.path.join(out_dir, "ckpt_folds_0.pt"), map_location="cpu", weights_only=False)
    st = ck["state"]
    layout = {k: list(v.shape) for k, v in st["model"].items()}
    ema_keys = sorted((st.get("ema") or {}).get("shadow_params", {}).keys()) if st.get("ema") else []
    return {"case": case, "train_log": log, "header": sub[0], "ids_head": ids[:4], "n_sub": len(ids), "probs": probs,
            "ckpt_model_layout": layout, "ckpt_state_keys": sorted(st.keys()), "ckpt_ema_keys": ema_keys,
            "ckpt_epoch": st["epoch"], "ckpt_global_step": st["global_step"], "score": float(ck["score"]),
            "calibrator_T": (float(torch.exp(st["calibrator"].temp_scaler.log_temp.detach())) if getattr(st.get("calibrator"), "temp_scaler", None)
                             is not None else None)}


def golden_path(case: str) -> str:
    return os.path.join(GOLDEN, f"e2e_{case}.json")
