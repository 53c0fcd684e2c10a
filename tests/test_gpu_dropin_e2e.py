"""The reference's own training and inference drivers, UNMODIFIED, running on the drop-in (run with -m gpu).

    PYTHONPATH=<pkg>/overlay:<pkg>:baseline/_ref  python -m src.train --cfg X ;  python -m src.infer --cfg X

`src` is a namespace package, so `src.models.{wrapper,dare,qnn_alpha,feature_embed}` resolve to the overlay (b200ctr on the
B200) and `src.train`, `src.infer`, `src.data`, `src.utils` to the reference copy (reference: src/train.py:92-366 --
K-fold split, DataLoader + collate_sharded, torch AdamW + clip_grad_norm_, ModelEMA update / store / copy_to / restore,
Calibrator, checkpoint; src/infer.py:10-173 -- checkpoint load, load_state_dict(strict=True), EMA copy-in, submission).
Compared with tests/golden/e2e_<case>.json = the same two commands run by the reference ALONE on the CPU
(oracle/make_e2e_golden.py), same shards / yaml / seed.

Stated tolerances: `dare_qnn_next` (all dropout 0: the two runs see identical batches and differ only by float rounding fed
through 10 AdamW steps) mean train loss 2e-3, AP / WLL 5e-3, submission probabilities 5e-3 max-abs; `dare_base` (its FC head
has a hard-coded train-mode Dropout(0.1), wrapper.py:98, so the runs are different random draws) mean train loss 5e-2,
AP 5e-2, WLL 5e-2, probabilities 0.1 max-abs.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest
import torch

from tests import e2e_common as E

pytestmark = pytest.mark.gpu

TOL = {"dare_qnn_next": dict(loss=2e-3, ap=5e-3, wll=5e-3, prob=5e-3),
       "dare_base": dict(loss=5e-2, ap=5e-2, wll=5e-2, prob=0.1)}


def _need_ref():
    if not os.path.isdir(os.path.join(E.REF_COPY, "src")):
        pytest.skip("baseline/_ref (copy of the reference made by oracle/install_ref.py) is not in this snapshot")


@pytest.mark.parametrize("case", list(E.CASES))
def test_unmodified_train_and_infer_run_on_the_dropin(case):
    _need_ref()
    gold = json.load(open(E.golden_path(case)))
    work = tempfile.mkdtemp(prefix=f"e2e_b200_{case}_")
    try:
        cfg_path = E.prepare(case, work, E.REF_COPY, device="cuda")
        pp = [E.OVERLAY, E.PKG, E.REF_COPY]
        E.run_module("src.train", cfg_path, pp, 1500, os.path.join(work, "train.log"))
        E.run_module("src.infer", cfg_path, pp, 900, os.path.join(work, "infer.log"))
        got = E.collect(case, work)
        keep = os.environ.get("B200CTR_E2E_KEEP")
        if keep:
            os.makedirs(keep, exist_ok=True)
            json.dump({k: v for k, v in got.items() if k != "probs"}, open(os.path.join(keep, f"e2e_{case}_b200.json"), "w"))
            for f in ("train.log", "infer.log"):
                shutil.copy(os.path.join(work, f), os.path.join(keep, f"e2e_{case}_{f}"))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    tol = TOL[case]
    # train_log.csv: same rows (val [+ val_cal]), same bookkeeping columns, metrics within tolerance
    assert [(r["fold"], r["epoch"], r["split"]) for r in got["train_log"]] == [(r["fold"], r["epoch"], r["split"]) for r in gold["train_log"]]
    for a, b in zip(got["train_log"], gold["train_log"]):
        for k in ("bs", "K", "tau"):
            assert a[k] == b[k]
        assert abs(float(a["lr"]) - float(b["lr"])) < 1e-12
        if a["loss"] not in ("", "--"):
            assert abs(float(a["loss"]) - float(b["loss"])) < tol["loss"], (a, b)
        assert abs(float(a["AP"]) - float(b["AP"])) < tol["ap"], (a, b)
        assert abs(float(a["WLL"]) - float(b["WLL"])) < tol["wll"], (a, b)
    # checkpoint: the reference's layout (state_dict keys AND shapes, EMA shadow names, bookkeeping)
    assert got["ckpt_model_layout"] == gold["ckpt_model_layout"]
    assert got["ckpt_state_keys"] == gold["ckpt_state_keys"] and got["ckpt_ema_keys"] == gold["ckpt_ema_keys"]
    assert got["ckpt_epoch"] == gold["ckpt_epoch"] and got["ckpt_global_step"] == gold["ckpt_global_step"]
    # submission.csv: header, ids, row count, probabilities
    assert got["header"] == gold["header"] == "ID,clicked"
    assert got["n_sub"] == gold["n_sub"] and got["ids_head"] == gold["ids_head"]
    d = np.abs(np.array(got["probs"]) - np.array(gold["probs"]))
    assert float(d.max()) < tol["prob"], float(d.max())


def test_state_dict_interop_with_the_reference_module_both_ways():
    """A reference-made state_dict loads strict=True into the drop-in and vice versa, at the shipped widths (seq_vocab shrunk);
    eval-mode forward of the drop-in (CUDA) equals the reference module (CPU) on the same weights."""
    _need_ref()
    import yaml
    sys.path.insert(0, E.PKG)
    import b200ctr
    from b200ctr import configs, synth
    sys.path.insert(0, E.REF_COPY)
    from src.models.wrapper import CTRModel as RefModel          # the real reference class, CPU
    cfg = yaml.safe_load(open(os.path.join(E.REF_COPY, "cfgs", "dare_qnn_next_k100_s1.yaml")))
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"].get(c, 1000003) + cfg["data"].get("hash_buckets_margin", 0) for c in cols}
    torch.manual_seed(3)
    ref = RefModel(cfg, 20_000, E.N_NUM, E.N_NUM, card, cols).eval()
    torch.manual_seed(4)
    mine = b200ctr.CTRModel(cfg, 20_000, E.N_NUM, E.N_NUM, card, cols)
    missing = mine.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    mine = mine.to("cuda:0").eval()
    back = RefModel(cfg, 20_000, E.N_NUM, E.N_NUM, card, cols)
    back.load_state_dict({k: v.cpu() for k, v in mine.state_dict().items()}, strict=True)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), back.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    schema = synth.ToySchema(cfg, E.N_NUM, E.N_NUM, seed=5)
    b, _ = schema.batch(256, seed=6)
    b["seq"] = np.where(b["seq"] > 0, b["seq"] % 19_999 + 1, 0)
    batch = {k: torch.from_numpy(v) for k, v in b.items()}
    with torch.no_grad():
        z_ref, p_ref, a_ref = ref(batch)
        z, p, a = mine(batch)
    assert float((z.cpu() - z_ref).abs().max() / z_ref.abs().max()) < 1e-4
    assert float((p.cpu() - p_ref).abs().max()) < 1e-5
    assert float((a.cpu() - a_ref).abs().max() / a_ref.abs().max()) < 1e-4
