"""GPU parity of the BENCHMARKED code path against the reference at the BASELINE shapes (run with -m gpu).

Fixtures: tests/golden/full_*.npz, generated from the real reference by oracle/make_golden_full.py -- F = 200 feature rows
with the real cat_embedding_dims, ffn_hidden 384, H = 6 / r = 16 / P = 192, MLP 512-256, L = 400, K = 60 / 100 / 148,
3 or 4 encoder layers, B = 128, dropout 0.  At these widths the dispatch predicates select the same fused tcgen05 / TMA /
tensor-core-attention kernels bench.py times (asserted below from the per-call profile tags), so a wrong leading
dimension or hand-off in the full-size integration fails here.

Stated tolerances (max-abs error / max-abs of the reference tensor, plus relative error of the whole-tensor L2 norm):
  fp32 path (`amp: none`)   : 1e-4 logits / prob / aux / every gradient; 2e-3 parameters + EMA after 2 AdamW steps at lr 3e-3
                              (bf16: 1.5e-1 -- sign flips of near-zero gradients become full lr-sized steps under Adam)
  bf16 path (`amp: bf16`)   : 2e-2 logits / aux, loss within 2e-3, whole-tensor gradient norms 3e-2; elementwise, every gradient's
                              stored subsample: relative L2 error 1.5e-1, max-abs error 3.5e-1 of max|ref| (measured: <= 0.13 /
                              0.28, largest on embedding-table rows that collect a handful of samples' gradients after they
                              crossed every bf16 contraction of the network).  The yardstick is the reference's OWN autocast
                              error, measured on the same GPU by test_bf16_error_is_no_larger_than_the_reference_autocast_error;
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
    bad = {n: e for n, e in worst.items() if e[0] >= 3.5e-1 or e[1] >= 3e-2 or e[2] >= 1.5e-1}
    assert not bad, bad


def test_bf16_error_is_no_larger_than_the_reference_autocast_error():
    """The yardstick for the bf16 tolerances above: the REFERENCE module itself on this GPU under torch.autocast(bfloat16)
    (exactly what its train loop does for `amp: bf16`, src/train.py:158-161) against its own fp32 result.  The b200ctr bf16
    path keeps activations / softmax / norms in fp32 between the tensor-core contractions, so its distance to the fp32
    reference must not exceed the reference's own autocast distance (factor 1.5 + a 1e-2 floor for run-to-run atomics)."""
    import sys
    from tests import e2e_common as E
    if not os.path.isdir(os.path.join(E.REF_COPY, "src")):
        pytest.skip("baseline/_ref is not in this snapshot")
    sys.path.insert(0, E.REF_COPY)
    from src.models.wrapper import CTRModel as RefModel
    from src.train import bce_wll_style
    dev = torch.device("cuda:0")
    name = "full_k60_concat"
    z, cfg = load_full(name)
    ref = full_init(RefModel, cfg, z).to(dev).train()
    batch, y = full_batch(z, 0)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, prob, aux = ref(batch)
        loss = bce_wll_style(logits, y.to(dev)) + float(cfg["model"]["qnn_alpha"]["aux_head_weight"]) * bce_wll_style(aux, y.to(dev))
    loss.backward()
    ref_err = {n: packed_err(z, "grad0/" + n, p.grad.float(), l2=True) for n, p in ref.named_parameters() if p.grad is not None}
    ref_logit_err = max_err_scaled(logits.detach().float().cpu(), torch.from_numpy(z["step0/logits"]))
    m, z, cfg = _model(name, dev)
    logits2, _, _, _, grads, _ = _fwd_bwd(m, z, dev, "bf16")
    my_err = {n: packed_err(z, "grad0/" + n, grads[n], l2=True) for n in packed_keys(z, "grad0/")}
    my_logit_err = max_err_scaled(logits2, torch.from_numpy(z["step0/logits"]))
    if os.environ.get("B200CTR_DUMP_ERRS"):
        with open(os.path.join(os.environ["B200CTR_DUMP_ERRS"], "bf16_vs_reference_autocast.json"), "w") as fh:
            json.dump({"logits": {"b200ctr_bf16": my_logit_err, "reference_autocast_bf16": ref_logit_err},
                       "grad_l2_err(b200ctr, reference_autocast)": {n: (my_err[n][2], ref_err[n][2]) for n in my_err}}, fh, indent=1)
    assert my_logit_err <= 1.5 * ref_logit_err + 1e-2
    worse = {n: (my_err[n][2], ref_err[n][2]) for n in my_err if my_err[n][2] > 1.5 * ref_err[n][2] + 1e-2}
    assert not worse, worse


@pytest.mark.parametrize("prec,tol", [("none", 2e-3), ("bf16", 1.5e-1)])
def test_fused_trainer_two_steps_at_full_width(prec, tol):
    """FusedTrainer (flat buffers, fused clip / AdamW / EMA, row-lazy table optimizer) on the K = 60 concat case: loss and
    global gradient norm of both steps, parameters and EMA shadows after the second step."""
    import b200ctr
    dev = torch.device("cuda:0")
    m, z, cfg = _model("full_k60_concat", dev)
    cfg = dict(cfg, amp=prec)
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
    ltol = 2e-4 if prec == "none" else 3e-3
    # global gradient norm: step 0 sees the reference's parameters; step 1 sees parameters one AdamW step later, where Adam's
    # g / sqrt(v) normalisation has turned rounding-level gradient differences into lr-sized parameter differences
    # (the fixture's own step-0 value is 4.6e-4 BELOW the norm of the gradients stored beside it -- sqrt(sum of the per-tensor
    # sums of squares) = 0.916448 against clip_grad_norm_'s 0.916026: torch accumulates the 7 M squares of qnn.mlp.0.weight in
    # fp32 on the CPU; the kernels here accumulate in double and land on 0.916446 -- so the bound is 1e-3, not 2e-4)
    ntol = {"none": (1e-3, 2e-3), "bf16": (1.5e-2, 1e-1)}[prec]
    for s in range(2):
        batch, y = full_batch(z, s)
        assert abs(tr.lr_at(s + 1) - float(z[f"step{s}/lr"])) < 1e-15
        loss = tr.step(batch, y)
        assert abs(float(loss.cpu()) - float(z[f"step{s}/loss"])) < ltol * (1 + 4 * s), (s, float(loss.cpu()), float(z[f"step{s}/loss"]))
        tn = float(z[f"step{s}/total_norm"])
        assert abs(float(tr.total_norm.cpu()) - tn) < ntol[s] * tn, (prec, s, float(tr.total_norm.cpu()), tn)
    # Parameters after two AdamW steps at lr 3e-3.  Adam moves EVERY element by about lr per step whatever its gradient's size,
    # so an element whose gradient is at the rounding level of the two implementations (|g| <~ 1e-5 max|g|) may move +lr in
    # one and -lr in the other: a max-abs bound per element is ill-conditioned by construction (observed: single elements of
    # qnn.mlp.0.weight off by 2 lr = 0.66 max|w| with the tensor's norm within 8e-5).  Checked instead: the whole-tensor norm,
    # the MEDIAN element error, and the share of elements further off than 10 % of max|w|.
    from oracle.make_golden_full import sample_stride
    sd = m.state_dict()
    ema = tr.ema_state()
    med_tol, frac_tol = (1e-4, 0.02) if prec == "none" else (2e-2, 0.10)
    report = {}
    for grp, src in (("final", sd), ("ema", ema)):
        for n in packed_keys(z, grp + "/"):
            e_sub, e_norm = packed_err(z, f"{grp}/{n}", src[n])
            ref = torch.from_numpy(z[f"{grp}/{n}"]).double()
            a = src[n].detach().double().cpu().reshape(-1)
            d = (a[::sample_stride(a.numel())] - ref).abs() / (ref.abs().max() + 1e-30)
            med, frac = float(d.median()), float((d > 0.1).double().mean())
            report[f"{grp}/{n}"] = (e_sub, e_norm, med, frac)
            if a.numel() < 64:            # a handful of elements: no distribution to speak of, plain bound of 2 lr-sized steps
                assert e_sub < max(tol, 0.15), (grp, n, e_sub)
            else:
                assert e_norm < tol and med < med_tol and frac < frac_tol, (grp, n, e_sub, e_norm, med, frac)
    if os.environ.get("B200CTR_DUMP_ERRS"):
        with open(os.path.join(os.environ["B200CTR_DUMP_ERRS"], f"trainer_two_steps_{prec}.json"), "w") as fh:
            json.dump({"(max_abs, norm, median, share>0.1)": report}, fh, indent=1)
