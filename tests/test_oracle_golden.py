"""Pin oracle/ctr_oracle.py against fixtures produced by the REAL reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import CASES, GOLDEN, batch_of, group, load_case, max_err_scaled, rel_err, spec_of

TOL = 1e-4      # north_star: fp32 rel <= 1e-4 on logits and gradients


@pytest.mark.parametrize("kind", CASES)
def test_forward_matches_reference(kind):
    z, cfg = load_case(kind)
    sd = group(z, "init/")
    spec = spec_of(cfg)
    batch, y = batch_of(z, 0)
    logits, prob, aux, c = O.model_forward(sd, spec, batch, keep=True)
    assert max_err_scaled(logits, torch.from_numpy(z["step0/logits"])) < TOL
    assert max_err_scaled(prob, torch.from_numpy(z["step0/prob"])) < TOL
    assert max_err_scaled(aux, torch.from_numpy(z["step0/aux"])) < TOL
    assert max_err_scaled(c["qv"], torch.from_numpy(z["step0/query"])) < 1e-5
    # valid prefix of vals must match; pad slots carry -1e9 in both
    assert torch.allclose(c["vals"], torch.from_numpy(z["step0/vals"]), rtol=1e-5, atol=1e-5)
    assert max_err_scaled(c["sel_rep"], torch.from_numpy(z["step0/sel_rep"])) < 1e-6
    assert max_err_scaled(c["u"], torch.from_numpy(z["step0/u_seq"])) < TOL


@pytest.mark.parametrize("kind", CASES)
def test_loss_and_gradients_match_reference_autograd(kind):
    z, cfg = load_case(kind)
    sd = group(z, "init/")
    spec = spec_of(cfg)
    batch, y = batch_of(z, 0)
    loss, logits, prob, aux, G, _ = O.loss_and_grads(sd, spec, batch, y)
    assert abs(float(loss) - float(z["step0/loss"])) < 1e-5 * max(1.0, abs(float(z["step0/loss"])))
    ref = group(z, "grad0/")
    assert set(ref.keys()) == set(G.keys()), (set(ref) ^ set(G))
    for k, g in ref.items():
        assert G[k].shape == g.shape, k
        assert max_err_scaled(G[k], g) < TOL, (k, max_err_scaled(G[k], g))
    if f"step0/total_norm" in z.files:
        _, tot = O.clip_coef(G, cfg["train"]["grad_clip_norm"])
        assert abs(tot - float(z["step0/total_norm"])) < 1e-4 * tot


@pytest.mark.parametrize("kind", CASES)
def test_three_optimizer_steps_match_reference(kind):
    """lr schedule + clip + AdamW (+EMA) over 3 steps: parameters and EMA shadows equal the reference's."""
    z, cfg = load_case(kind)
    sd = group(z, "init/")
    spec = spec_of(cfg)
    ema = cfg.get("ema", {})
    tr = O.OracleTrainer(sd, spec, cfg["train"]["lr"], cfg["train"]["weight_decay"], cfg["train"]["grad_clip_norm"],
                         ema_decay=ema.get("decay") if ema.get("enabled") else None,
                         ema_warmup_steps=int(ema.get("warmup_steps", 0)), ema_warmup_type=ema.get("warmup_type", "linear"))
    spe = cfg["golden"]["steps_per_epoch"]
    for s in range(3):
        batch, y = batch_of(z, s)
        epoch, step = divmod(s, spe)
        lr = O.cosine_warmup_lr(epoch, step, spe, cfg["train"]["lr"], cfg["train"]["warmup_epochs"], cfg["train"]["epochs"])
        assert abs(lr - float(z[f"step{s}/lr"])) < 1e-15
        loss, logits, _ = tr.step(batch, y, lr)
        assert abs(loss - float(z[f"step{s}/loss"])) < 2e-4 * max(1.0, abs(float(z[f"step{s}/loss"])))
        assert max_err_scaled(logits, torch.from_numpy(z[f"step{s}/logits"])) < 5e-4
    fin = group(z, "final/")
    # Adam's m/sqrt(v) turns fp32 rounding noise on mathematically-zero gradients (e.g. the key bias of
    # softmax attention) into O(lr) steps, so post-step parameters get a looser bound than gradients.
    for k, v in fin.items():
        assert max_err_scaled(sd[k], v) < 1e-3, (k, max_err_scaled(sd[k], v))
    if tr.shadow is not None:
        for k, v in group(z, "ema/").items():
            assert max_err_scaled(tr.shadow[k], v) < 1e-3, k


def test_retrieval_dyadic_bit_exact():
    """On the dyadic grid, scores are order-independent: the oracle's top-k values and selected rows are
    bit-identical to the reference's, and its index order is the canonical (score desc, position asc) one."""
    z = np.load(os.path.join(GOLDEN, "retrieval_dyadic.npz"))
    K = int(z["K"])
    seq = torch.from_numpy(z["seq"])
    q = torch.from_numpy(z["q"])
    att, rep = torch.from_numpy(z["emb_att"]), torch.from_numpy(z["emb_rep"])
    rec = O.recency_table(seq.shape[1], float(z["tau"]))
    assert torch.equal(rec, torch.from_numpy(z["rec"]))
    sel, vals, idx = O.retrieve_fwd(seq, q, att, rep, rec, K)
    assert torch.equal(vals, torch.from_numpy(z["vals"]))                  # bit-exact
    scores = torch.from_numpy(z["scores"])
    # canonical order: stable sort of the reference's own score matrix
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :K]
    assert torch.equal(idx, order)
    ref_sel = torch.from_numpy(z["sel_rep"])
    valid = vals > -1e8
    # selected rows: identical wherever the score is not tied with a neighbour (ties may be permuted by torch.topk)
    same = (sel == ref_sel).all(-1)
    tied = torch.zeros_like(valid)
    tied[:, 1:] |= vals[:, 1:] == vals[:, :-1]
    tied[:, :-1] |= vals[:, :-1] == vals[:, 1:]
    assert bool((same | tied | ~valid).all())
    # multiset of selected rows is identical per sample even under ties (sum is exact on the grid)
    assert torch.equal(sel.sum(1), ref_sel.sum(1))
    assert bool((sel[~valid] == 0).all())


def test_bce_edge_cases():
    zt = torch.tensor([0.3, -1.2, 25.0, -30.0])
    for y in (torch.zeros(4), torch.ones(4), torch.tensor([1.0, 0, 1, 0])):
        zr = zt.clone().requires_grad_(True)
        pos = y > 0.5
        lp = torch.nn.functional.softplus(-zr[pos]).mean() if pos.any() else torch.zeros(())
        ln = torch.nn.functional.softplus(zr[~pos]).mean() if (~pos).any() else torch.zeros(())
        ref = 0.5 * (lp + ln)
        ref.backward()
        loss, dz = O.bce_wll(zt, y)
        assert abs(float(loss) - float(ref)) < 1e-6
        assert torch.allclose(dz, zr.grad, atol=1e-7)
