"""GPU parity of the BENCHMARKED code path against the reference at the BASELINE shapes (run with -m gpu).

Fixtures: tests/golden/full_*.npz, generated from the real reference by oracle/make_golden_full.py -- F = 200 feature rows
with the real cat_embedding_dims, ffn_hidden 384, H = 6 / r = 16 / P = 192, MLP 512-256, L = 400, K = 60 / 100 / 148,
3 or 4 encoder layers, B = 128, dropout 0.  At these widths the dispatch predicates select the same fused tcgen05 / TMA /
tensor-core-attention kernels bench.py times (asserted below from the per-call profile tags), so a wrong leading
dimension or hand-off in the full-size integration fails here.

Stated tolerances (max-abs error / max-abs of the reference tensor, plus relative error of the whole-tensor L2 norm):
  fp32 path (`amp: none`)   : 1e-4 logits / prob / aux / every gradient; 1e-3 parameters + EMA after 2 AdamW steps
  bf16 path (`amp: bf16`)   : 2e-2 logits / aux, loss within 2e-3; every gradient: relative L2 error of the stored subsample 6e-2,
                              its max-abs error 1.2e-1 of max|ref| (single elements of gradients that crossed the whole bf16
                              network), whole-tensor norm 3e-2;
                              the top-k selection runs in fp32 in both modes (vals <= 1e-5).
"""
import json
import os

import pytest
import torch

from tests.helpers import FULL_CASES, full_batch, full_init, load_full, max_err_scaled, packed_err, packed_keys

pytestmark = pytest.mark.gpu

FUSED_TAGS = ("ffn_fwd", "ffn_bwd", "fm_fused_fwd", "fm_fused_bwd", "gemm_tma", "attn_tc_fwd", "attn_tc_bwd", "inproj_fwd",
              "inproj_bwd", "outproj_norm_fwd", "outproj_norm_bwd", "retrieve_fwd", "retrieve_bwd", "headproj_fwd", "lin1_fwd")


def _model(name, dev):
    import b200ctr
    z, cfg = load_full(name)
    m = full_init(b200ctr.CTRModel, cfg, z).to(dev).train()
    return m, z, cfg


def _fwd_bwd(m, z, dev, prec):
    from b200ctr import ops
    batch, y = full_batch(z, 0)
    ops.profile_start()
    with ops.precision(prec):
        logits, prob, aux = m(batch)
        loss, dz, da, _, _ = ops.bce_wll(logits.detach(), aux.detach(), y.to(dev), m.aux_weight)
        torch.autograd.backward([logits, aux], [dz, da])
    prof = ops.profile_stop()
    fams = {t.split("[")[0] for t in prof}
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    return logits.detach().cpu(), prob.cpu(), aux.detach().cpu(), float(loss.cpu()), grads, fams


@pytest.mark.parametrize("name", FULL_CASES)
def test_fp32_path_matches_reference_at_full_width(name):
    dev = torch.device("cuda:0")
    m, z, cfg = _model(name, dev)
    logits, prob, aux, loss, grads, fams = _fwd_bwd(m, z, dev, "fp32")
    assert max_err_scaled(logits, torch.from_numpy(z["step0/logits"])) < 1e-4
    assert max_err_scaled(prob, torch.from_numpy(z["step0/prob"])) < 1e-4
    assert max_err_scaled(aux, torch.from_numpy(z["step0/aux"])) < 1e-4
    assert abs(loss - float(z["step0/loss"])) < 1e-5
    ref = packed_keys(z, "grad0/")
    assert set(grads) == set(ref), set(grads) ^ set(ref)
    for n in ref:
        e_sub, e_norm = packed_err(z, "grad0/" + n, grads[n])
        assert e_sub < 1e-4 and e_norm < 1e-4, (n, e_sub, e_norm)


@pytest.mark.parametrize("name", FULL_CASES)
def test_retrieval_at_full_width(name):
    """DARE.topk_select at L = 400, K = 60 / 100 / 148 on the reference's own query: sorted values and gathered rows."""
    dev = torch.device("cuda:0")
    m, z, cfg = _model(name, dev)
    batch, _ = full_batch(z, 0)
    q = torch.from_numpy(z["step0/query"]).to(dev)
    with torch.no_grad():
        sel, vals = m.dare.topk_select(batch["seq"], q)
    ref_vals = torch.from_numpy(z["step0/vals"])
    valid = ref_vals > -1e8
    assert bool(((vals.cpu() > -1e8) == valid).all())                      # same number of real (non-pad) slots per row
    err = ((vals.cpu() - ref_vals).abs() / ref_vals.abs().clamp_min(1.0))[valid].max()
    assert float(err) < 1e-5
    assert bool((vals.cpu()[:, :-1] >= vals.cpu()[:, 1:]).all())           # sorted descending
    e_sub, e_norm = packed_err(z, "step0/sel_rep", sel)
    assert e_sub < 1e-6 and e_norm < 1e-6                                   # gathered rows are copies: exact unless a near-tie flips


@pytest.mark.parametrize("name", FULL_CASES)
def test_bf16_fused_path_matches_reference_at_full_width(name):
    """The mode bench.py reports: every fused kernel family must have run, and the result must track the fp32 reference."""
    dev = torch.device("cuda:0")
    m, z, cfg = _model(name, dev)
    logits, prob, aux, loss, grads, fams = _fwd_bwd(m, z, dev, "bf16")
    missing = [t for t in FUSED_TAGS if t not in fams]
    assert not missing, f"fused kernels not dispatched at BASELINE widths: {missing}; ran {sorted(fams)}"
    assert max_err_scaled(logits, torch.from_numpy(z["step0/logits"])) < 2e-2
    assert max_err_scaled(aux, torch.from_numpy(z["step0/aux"])) < 2e-2
    assert abs(loss - float(z["step0/loss"])) < 2e-3
    worst = {n: packed_err(z, "grad0/" + n, grads[n], l2=True) for n in packed_keys(z, "grad0/")}
    if os.environ.get("B200CTR_DUMP_ERRS"):
        with open(os.path.join(os.environ["B200CTR_DUMP_ERRS"], f"bf16_grad_err_{name}.json"), "w") as fh:
            json.dump({"logits": max_err_scaled(logits, torch.from_numpy(z["step0/logits"])), "grads(max_abs,norm,l2)": worst}, fh, indent=1)
    bad = {n: e for n, e in worst.items() if e[0] >= 1.2e-1 or e[1] >= 3e-2 or e[2] >= 6e-2}
    assert not bad, bad


@pytest.mark.parametrize("prec,tol", [("none", 1e-3), ("bf16", 3e-2)])
def test_fused_trainer_two_steps_at_full_width(prec, tol):
    """FusedTrainer (flat buffers, fused clip / AdamW / EMA, row-lazy table optimizer) on the K = 60 concat case: loss and
    global gradient norm of both steps, parameters and EMA shadows after the second step."""
    import b200ctr
    dev = torch.device("cuda:0")
    m, z, cfg = _model("full_k60_concat", dev)
    cfg = dict(cfg, amp=prec)
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
    ltol = 2e-4 if prec == "none" else 3e-3
    for s in range(2):
        batch, y = full_batch(z, s)
        assert abs(tr.lr_at(s + 1) - float(z[f"step{s}/lr"])) < 1e-15
        loss = tr.step(batch, y)
        assert abs(float(loss.cpu()) - float(z[f"step{s}/loss"])) < ltol, (s, float(loss.cpu()), float(z[f"step{s}/loss"]))
        tn = float(z[f"step{s}/total_norm"])
        assert abs(float(tr.total_norm.cpu()) - tn) < (2e-4 if prec == "none" else 3e-2) * tn
    sd = m.state_dict()
    for n in packed_keys(z, "final/"):
        e_sub, e_norm = packed_err(z, "final/" + n, sd[n])
        assert e_sub < tol and e_norm < tol, ("final", n, e_sub, e_norm)
    ema = tr.ema_state()
    for n in packed_keys(z, "ema/"):
        e_sub, e_norm = packed_err(z, "ema/" + n, ema[n])
        assert e_sub < tol and e_norm < tol, ("ema", n, e_sub, e_norm)
