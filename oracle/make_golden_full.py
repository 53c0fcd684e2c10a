"""Full-width golden fixtures from the REAL reference at the BASELINE shapes (build container only).

    python oracle/make_golden_full.py        # writes tests/golden/full_*.npz

The three cases load the reference's own YAMLs (cfgs/dare_qnn_next.yaml, dare_qnn_next_k100_s1.yaml, v3_k148_s1.yaml)
and keep every width the CUDA dispatch predicates look at: F = 1 + 82 + 82 + 35 = 200 feature rows with the real
`cat_embedding_dims`, ffn_hidden 384, 8 heads, H = 6 / r = 16 / P = 192, MLP 512-256, L = 400, K in {60, 100, 148},
3 or 4 encoder layers.  Only what has no influence on the code path is shrunk: B = 128 (the smallest batch the bf16
first-layer kernels accept), seq_vocab = 5,000, dropout 0 (exact parity; train-mode dropout is statistical only).

To keep the fixtures small the INITIAL parameters are not stored: both sides rebuild them from the seed (the drop-in's
constructor consumes the RNG in the reference's order -- tests/test_host_logic.py -- and the fixture carries a
checksum per tensor so a drift is detected, not silently compared).  Large tensors are stored as a strided subsample
(<= CAP elements, prime stride) plus the fp64 sum and sum of squares of the whole tensor.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import yaml

REF = os.environ.get("B200CTR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CAP = 2048
N_NUM = 82

CASES = {
    # name: (yaml, steps, groups stored)
    "full_k60_concat": ("dare_qnn_next.yaml", 2, ("grad0", "final", "ema")),
    "full_k100_s1": ("dare_qnn_next_k100_s1.yaml", 1, ("grad0",)),
    "full_k148_s1": ("v3_k148_s1.yaml", 1, ("grad0",)),
}


def _next_prime(n: int) -> int:
    n = max(2, n)
    while True:
        if all(n % p for p in range(2, int(n ** 0.5) + 1)):
            return n
        n += 1


def sample_stride(numel: int) -> int:
    return 1 if numel <= CAP else _next_prime((numel + CAP - 1) // CAP)


def pack(out: dict, key: str, t: torch.Tensor):
    """key -> subsample, key+'#stats' -> [sum, sumsq, numel] in fp64."""
    a = t.detach().double().reshape(-1)
    out[key] = t.detach().reshape(-1)[::sample_stride(a.numel())].float().numpy().copy()
    out[key + "#stats"] = np.array([float(a.sum()), float((a * a).sum()), float(a.numel())], dtype=np.float64)


def full_cfg(yaml_name: str) -> dict:
    with open(os.path.join(REF, "cfgs", yaml_name)) as fh:
        cfg = yaml.safe_load(fh)
    cfg["sequence"]["tfm"].update(mha_dropout=0.0, ffn_dropout=0.0)
    cfg["model"]["dare_dropout"] = 0.0
    cfg["model"]["embedding_dropout"] = 0.0
    cfg["model"]["qnn_alpha"]["dropout"] = 0.0
    cfg["train"].update(batch_size=128, lr=3e-3, weight_decay=1e-2, epochs=2, warmup_epochs=1)
    if cfg["ema"].get("enabled"):
        cfg["ema"].update(decay=0.9)
    cfg["golden"] = {"seq_vocab": 5000, "num_dim": N_NUM, "mask_dim": N_NUM, "steps_per_epoch": 2}
    return cfg


def build_init(model_cls, cfg: dict):
    """The seeded initial state both sides rebuild: torch.manual_seed(seed) -> ctor -> perturb the all-ones / all-zero
    parameters (norm gains, biases, relative-position tables) so every gradient path carries signal."""
    g = cfg["golden"]
    cat_cols = cfg["data"]["cat_cols"]
    cardinals = {c: cfg["data"]["hash_buckets"].get(c, 1000003) + cfg["data"].get("hash_buckets_margin", 0) for c in cat_cols}
    torch.manual_seed(cfg["seed"])
    model = model_cls(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], cardinals, cat_cols)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".w") or n.endswith("bias") or "pbias" in n:
                p.add_(0.1 * torch.randn_like(p))
    return model


def make_batch(cfg: dict, rng: np.random.Generator, B: int):
    g = cfg["golden"]
    L, K, V = cfg["sequence"]["max_len"], cfg["sequence"]["top_k"], g["seq_vocab"]
    X_num = rng.standard_normal((B, g["num_dim"])).astype(np.float32)
    X_mask = (rng.random((B, g["mask_dim"])) < 0.05).astype(np.float32)
    X_num[X_mask > 0] = 0.0
    margin = cfg["data"].get("hash_buckets_margin", 0)
    cards = [cfg["data"]["hash_buckets"].get(c, 1000003) + margin for c in cfg["data"]["cat_cols"]]
    X_cat = np.stack([rng.integers(0, c, size=B) for c in cards], axis=1).astype(np.int64)
    seq = np.zeros((B, L), dtype=np.int64)
    lens = np.minimum(L, np.floor(rng.lognormal(5.3, 0.9, size=B))).astype(np.int64)
    lens[:6] = [0, 1, K - 1, K, K + 1, L]
    for b in range(B):
        ell = int(lens[b])
        if ell:
            toks = rng.integers(1, V, size=ell)
            if ell > 3:
                toks[1] = toks[0]
            seq[b, L - ell:] = toks
    y = (rng.random(B) < 0.3).astype(np.float32)
    y[0], y[1] = 1.0, 0.0
    return {"X_num": X_num, "X_mask": X_mask, "X_cat": X_cat, "seq": seq}, y


def run_case(name: str):
    sys.path.insert(0, REF)
    from src.models.wrapper import CTRModel            # noqa: E402  (the real reference)
    from src.train import bce_wll_style                # noqa: E402
    from src.utils.ema import build_ema                # noqa: E402
    from src.utils.sched import cosine_warmup_lr       # noqa: E402

    yaml_name, n_steps, groups = CASES[name]
    cfg = full_cfg(yaml_name)
    g = cfg["golden"]
    model = build_init(CTRModel, cfg)
    model.train()
    ema = build_ema(model, cfg) if cfg.get("ema", {}).get("enabled", False) else None
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
    out = {"cfg_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for n, p in model.state_dict().items():
        a = p.detach().double().reshape(-1)
        out["initsum/" + n] = np.array([float(a.sum()), float((a * a).sum()), float(a.numel())])
    aux_w = float(cfg["model"]["qnn_alpha"].get("aux_head_weight", 0.0))
    B = cfg["train"]["batch_size"]
    rng = np.random.default_rng(cfg["seed"])
    spe = g["steps_per_epoch"]
    for s in range(n_steps):
        epoch, step = divmod(s, spe)
        batch_np, y_np = make_batch(cfg, rng, B)
        batch = {k: torch.from_numpy(v) for k, v in batch_np.items()}
        y = torch.from_numpy(y_np)
        lr = cosine_warmup_lr(epoch, step, spe, cfg["train"]["lr"], cfg["train"]["warmup_epochs"], cfg["train"]["epochs"])
        opt.param_groups[0]["lr"] = lr
        opt.zero_grad(set_to_none=True)
        logits, prob, aux = model(batch)
        loss = bce_wll_style(logits, y)
        if aux_w > 0:
            loss = loss + aux_w * bce_wll_style(aux, y)
        loss.backward()
        out[f"step{s}/batch/X_num"] = batch_np["X_num"]
        out[f"step{s}/batch/X_mask"] = batch_np["X_mask"].astype(np.uint8)
        out[f"step{s}/batch/X_cat"] = batch_np["X_cat"].astype(np.int32)
        out[f"step{s}/batch/seq"] = batch_np["seq"].astype(np.int32)
        out[f"step{s}/y"] = y_np
        out[f"step{s}/lr"] = np.float64(lr)
        out[f"step{s}/logits"] = logits.detach().numpy().copy()
        out[f"step{s}/prob"] = prob.detach().numpy().copy()
        out[f"step{s}/aux"] = aux.detach().numpy().copy()
        out[f"step{s}/loss"] = np.float64(loss.item())
        if s == 0:
            with torch.no_grad():
                cat_embs = model._embed_cats(batch["X_cat"])
                num_e = model.num_embed(batch["X_num"])
                mask_e = model.mask_embed(batch["X_mask"])
                ctx = model._context_vector(num_e, mask_e, cat_embs)
                qv = model._make_query(model.dare.query_mode, model.query_key, ctx, cat_embs)
                sel_rep, vals = model.dare.topk_select(batch["seq"], qv)
                u_seq, _ = model.dare(batch["seq"], qv)
            out["step0/query"] = qv.numpy().copy()
            out["step0/vals"] = vals.numpy().copy()
            out["step0/u_seq"] = u_seq.numpy().copy()
            pack(out, "step0/sel_rep", sel_rep)
            if "grad0" in groups:
                for n, p in model.named_parameters():
                    if p.grad is not None:
                        pack(out, "grad0/" + n, p.grad)
        if cfg["train"]["grad_clip_norm"] > 0:
            tn = torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["train"]["grad_clip_norm"])
            out[f"step{s}/total_norm"] = np.float64(float(tn))
        opt.step()
        if ema is not None:
            ema.update(model, s + 1)
    if "final" in groups:
        for n, p in model.state_dict().items():
            pack(out, "final/" + n, p)
    if "ema" in groups and ema is not None:
        for n, p in ema.shadow_params.items():
            pack(out, "ema/" + n, p)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF}: fixtures can only be regenerated in the build container")
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)
