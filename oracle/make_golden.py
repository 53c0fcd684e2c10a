"""Generate golden fixtures from the REAL reference (build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Imports the unmodified reference modules from /root/reference (never copied
into this repo), runs them on seeded inputs and records inputs, outputs,
autograd gradients and post-step parameters.  The fixtures pin oracle/ctr_oracle.py
(tests/test_oracle_golden.py) and, through it and directly, the CUDA path
(tests/test_gpu_*.py).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("B200CTR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def small_cfg(kind: str) -> dict:
    cat_cols = ["gender", "inventory_id", "l_feat_5", "l_feat_14", "l_feat_12"]
    dims = {"gender": 8, "inventory_id": 17, "l_feat_5": 33, "l_feat_14": 57, "l_feat_12": 64}
    buckets = {"gender": 11, "inventory_id": 40, "l_feat_5": 97, "l_feat_14": 150, "l_feat_12": 211}
    cfg = {
        "seed": 777,
        "data": {"cat_cols": cat_cols, "hash_buckets": buckets, "hash_buckets_margin": 3},
        "sequence": {"max_len": 40, "pad_id": 0, "top_k": 12, "recency_tau": 16, "query_mode": "concat",
                     "query_key": "inventory_id", "transformer_block": True,
                     "tfm": {"n_layers": 2, "n_heads": 8, "mha_dropout": 0.0, "ffn_hidden": 96,
                             "ffn_dropout": 0.0, "norm": "rms", "gating": "softmax",
                             "add_positional_bias": True}},
        "model": {"emb_dim": 64, "dare_dropout": 0.0, "embedding_dropout": 0.0, "cat_embedding_dims": dims,
                  "qnn_alpha": {"enabled": True, "feature_embed_dim": 12, "heads": 3, "rank": 8, "proj_dim": 24,
                                "mlp_hidden": [24, 16], "dropout": 0.0, "use_se": True, "se_reduction": 8,
                                "use_residual": True, "norm": "rms", "pair_grouping": "all",
                                "aux_head_weight": 0.1}},
        "train": {"batch_size": 24, "epochs": 3, "lr": 3e-3, "weight_decay": 1e-2, "warmup_epochs": 1,
                  "grad_clip_norm": 0.5},
        "ema": {"enabled": True, "decay": 0.9, "warmup_steps": 5, "warmup_type": "linear"},
        "golden": {"seq_vocab": 200, "num_dim": 7, "mask_dim": 7, "steps_per_epoch": 2},
    }
    if kind == "qnn_concat":
        pass
    elif kind == "base_s1":
        cfg["sequence"].update(query_mode="S1", transformer_block=False, top_k=16, tfm=None, recency_tau=8)
        cfg["model"]["qnn_alpha"] = {"enabled": False}
        cfg["model"]["dare_dropout"] = 0.0
        cfg["data"]["cat_cols"] = ["inventory_id", "l_feat_14"]      # keeps the hard-coded 512-wide FC small
        cfg["train"]["grad_clip_norm"] = 1.0
        cfg["ema"] = {"enabled": False}
    elif kind == "s2_relu":
        cfg["sequence"].update(query_mode="S2", top_k=9)
        cfg["sequence"]["tfm"].update(n_layers=1, gating="relu", n_heads=8)
        cfg["model"]["qnn_alpha"].update(use_se=False, aux_head_weight=0.0, heads=2, rank=4, mlp_hidden=[16])
        cfg["train"]["grad_clip_norm"] = 0.0
        cfg["ema"] = {"enabled": True, "decay": 0.95}
    else:
        raise ValueError(kind)
    return cfg


def make_batch(cfg: dict, rng: np.random.Generator, B: int):
    g = cfg["golden"]
    L = cfg["sequence"]["max_len"]
    K = cfg["sequence"]["top_k"]
    V = g["seq_vocab"]
    X_num = rng.standard_normal((B, g["num_dim"])).astype(np.float32)
    X_mask = (rng.random((B, g["mask_dim"])) < 0.2).astype(np.float32)
    X_num[X_mask > 0] = 0.0
    cards = [cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cfg["data"]["cat_cols"]]
    X_cat = np.stack([rng.integers(0, c, size=B) for c in cards], axis=1).astype(np.int64)
    seq = np.zeros((B, L), dtype=np.int64)
    lens = rng.integers(0, L + 1, size=B)
    lens[0] = 0                      # all-pad row
    lens[1] = 1
    lens[2] = K - 1
    lens[3] = K
    lens[4] = K + 1
    lens[5] = L
    for b in range(B):
        ell = int(lens[b])
        if ell:
            toks = rng.integers(1, V, size=ell)
            if ell > 3:
                toks[1] = toks[0]    # duplicate tokens in one row
            seq[b, L - ell:] = toks
    y = (rng.random(B) < 0.3).astype(np.float32)
    y[0], y[1] = 1.0, 0.0
    return {"X_num": X_num, "X_mask": X_mask, "X_cat": X_cat, "seq": seq}, y


def run_case(kind: str):
    sys.path.insert(0, REF)
    from src.models.wrapper import CTRModel            # noqa: E402  (the real reference)
    from src.train import bce_wll_style                # noqa: E402
    from src.utils.ema import build_ema                # noqa: E402
    from src.utils.sched import cosine_warmup_lr       # noqa: E402

    cfg = small_cfg(kind)
    g = cfg["golden"]
    torch.manual_seed(cfg["seed"])
    rng = np.random.default_rng(cfg["seed"])
    cat_cols = cfg["data"]["cat_cols"]
    cardinals = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cat_cols}
    model = CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], cardinals, cat_cols)
    # perturb the all-ones / all-zeros initialisations so every gradient path is exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".w") or n.endswith("bias") or "pbias" in n:
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    if hasattr(model, "fc"):
        model.fc[2].p = 0.0          # the FC head's Dropout(0.1) is hard-coded (wrapper.py:98): disable for exact parity
    ema = build_ema(model, cfg)
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
    out = {"cfg_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for n, p in model.state_dict().items():
        out["init/" + n] = p.detach().clone().numpy()
    aux_w = float(cfg["model"]["qnn_alpha"].get("aux_head_weight", 0.0))
    B = cfg["train"]["batch_size"]
    n_steps = 3
    spe = g["steps_per_epoch"]
    gstep = 0
    for s in range(n_steps):
        gstep += 1
        epoch, step = divmod(s, spe)
        batch_np, y_np = make_batch(cfg, rng, B)
        batch = {k: torch.from_numpy(v) for k, v in batch_np.items()}
        y = torch.from_numpy(y_np)
        lr = cosine_warmup_lr(epoch, step, spe, cfg["train"]["lr"], cfg["train"]["warmup_epochs"],
                              cfg["train"]["epochs"])
        opt.param_groups[0]["lr"] = lr
        opt.zero_grad(set_to_none=True)
        logits, prob, aux = model(batch)
        loss = bce_wll_style(logits, y)
        if aux_w > 0:
            loss = loss + aux_w * bce_wll_style(aux, y)
        loss.backward()
        for k, v in batch_np.items():
            out[f"step{s}/batch/{k}"] = v
        out[f"step{s}/y"] = y_np
        out[f"step{s}/lr"] = np.float64(lr)
        out[f"step{s}/logits"] = logits.detach().numpy().copy()
        out[f"step{s}/prob"] = prob.detach().numpy().copy()
        out[f"step{s}/aux"] = aux.detach().numpy().copy()
        out[f"step{s}/loss"] = np.float64(loss.item())
        if s == 0:
            # DARE intermediates of the first forward (retrieval + pooling), from the reference's own methods
            with torch.no_grad():
                X_cat = batch["X_cat"]
                cat_embs = model._embed_cats(X_cat)
                num_e = model.num_embed(batch["X_num"])
                mask_e = model.mask_embed(batch["X_mask"])
                ctx = model._context_vector(num_e, mask_e, cat_embs)
                qv = model._make_query(model.dare.query_mode, model.query_key, ctx, cat_embs)
                sel_rep, vals = model.dare.topk_select(batch["seq"], qv)
                u_seq, _ = model.dare(batch["seq"], qv)
            out["step0/query"] = qv.numpy().copy()
            out["step0/vals"] = vals.numpy().copy()
            out["step0/sel_rep"] = sel_rep.numpy().copy()
            out["step0/u_seq"] = u_seq.numpy().copy()
            for n, p in model.named_parameters():
                if p.grad is not None:
                    out["grad0/" + n] = p.grad.detach().numpy().copy()
        if cfg["train"]["grad_clip_norm"] > 0:
            tn = torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["train"]["grad_clip_norm"])
            out[f"step{s}/total_norm"] = np.float64(float(tn))
        opt.step()
        if ema is not None:
            ema.update(model, gstep)
    for n, p in model.state_dict().items():
        out["final/" + n] = p.detach().numpy().copy()
    if ema is not None:
        for n, p in ema.shadow_params.items():
            out["ema/" + n] = p.detach().numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"model_{kind}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def run_retrieval_dyadic():
    """DARE.topk_select on a dyadic grid (values k/64 in [-2,2]): every 64-term dot product is exactly
    representable in fp32, so scores are summation-order independent and top-k can be compared bit-for-bit
    (SURVEY section 4).  Includes heavy ties (duplicate tokens) and rows of every length class."""
    sys.path.insert(0, REF)
    from src.models.dare import DARE                   # noqa: E402
    torch.manual_seed(1234)
    rng = np.random.default_rng(1234)
    V, D, L, K, tau = 300, 64, 64, 20, 16
    dare = DARE(seq_vocab=V, emb_dim=D, dropout=0.0, query_mode="S1", transformer_block=False, top_k=K,
                recency_tau=tau, pad_id=0, tfm_cfg=None)
    with torch.no_grad():
        for emb in (dare.emb_att, dare.emb_rep):
            w = torch.from_numpy(rng.integers(-128, 129, size=(V, D)).astype(np.float32) / 64.0)
            w[0] = 0
            emb.weight.copy_(w)
    B = 48
    q = torch.from_numpy(rng.integers(-128, 129, size=(B, D)).astype(np.float32) / 64.0)
    seq = np.zeros((B, L), dtype=np.int64)
    lens = rng.integers(0, L + 1, size=B)
    lens[:8] = [0, 1, K - 1, K, K + 1, L, 2, L - 1]
    for b in range(B):
        ell = int(lens[b])
        if ell:
            seq[b, L - ell:] = rng.integers(1, V if b % 3 else 6, size=ell)   # every 3rd row: only 5 tokens -> ties
    seq_t = torch.from_numpy(seq)
    with torch.no_grad():
        sel_rep, vals = dare.topk_select(seq_t, q)
        # reference scores, for canonical-order reconstruction in the tests
        att = dare.emb_att(seq_t)
        pos = torch.arange(L).view(1, L).expand(B, L)
        decay = torch.exp(-(L - 1 - pos).float() / max(1.0, float(tau))).unsqueeze(-1)
        scores = (att * q.unsqueeze(1)).sum(-1, keepdim=True) + torch.log(decay + 1e-8)
        scores = scores.masked_fill((seq_t == 0).unsqueeze(-1), -1e9).squeeze(-1)
        rec = torch.log(decay + 1e-8)[0, :, 0]
    out = dict(emb_att=dare.emb_att.weight.detach().numpy(), emb_rep=dare.emb_rep.weight.detach().numpy(),
               q=q.numpy(), seq=seq, vals=vals.numpy(), sel_rep=sel_rep.numpy(), scores=scores.numpy(),
               rec=rec.numpy(), K=np.int64(K), tau=np.float64(tau))
    path = os.path.join(OUT, "retrieval_dyadic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF}: fixtures can only be regenerated in the build container")
    torch.set_num_threads(4)
    for kind in ("qnn_concat", "base_s1", "s2_relu"):
        run_case(kind)
    run_retrieval_dyadic()
