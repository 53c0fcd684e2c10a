"""The BASELINE.json configurations, restated as plain dicts (the values of the reference's cfgs/*.yaml that the
model and the step body read; file:line cites are to the reference).  bench.py and the tests build models from
these so nothing needs /root/reference at run time; a user of the reference passes their own yaml dict instead.
"""
from __future__ import annotations

import copy

# cfgs/dare_qnn_next.yaml:16-53 (cat_cols), 135-170 (hash_buckets), 198-233 (cat_embedding_dims)
_CATS = ["gender", "age_group", "inventory_id", "day_of_week", "hour"] + [f"l_feat_{i}" for i in range(1, 28)] + \
        ["feat_a_2", "feat_a_8", "feat_a_9"]
_BUCKETS = dict(gender=1005, age_group=1011, day_of_week=1009, hour=1026, inventory_id=1020, l_feat_1=1004, l_feat_2=1005,
                l_feat_3=1005, l_feat_4=1028, l_feat_5=2078, l_feat_6=1902, l_feat_7=1315, l_feat_8=1005, l_feat_9=1478,
                l_feat_10=1264, l_feat_11=2526, l_feat_12=6052, l_feat_13=1004, l_feat_14=4239, l_feat_15=3584,
                l_feat_16=1004, l_feat_17=1478, l_feat_18=1010, l_feat_19=1005, l_feat_20=1004, l_feat_21=1005,
                l_feat_22=1005, l_feat_23=1004, l_feat_24=1005, l_feat_25=2790, l_feat_26=1020, l_feat_27=1007,
                feat_a_2=1008, feat_a_8=1010, feat_a_9=1012)
_DIMS = dict(gender=8, age_group=8, day_of_week=8, hour=8, inventory_id=16, l_feat_1=8, l_feat_2=8, l_feat_3=8, l_feat_4=8,
             l_feat_5=33, l_feat_6=30, l_feat_7=18, l_feat_8=8, l_feat_9=22, l_feat_10=17, l_feat_11=40, l_feat_12=64,
             l_feat_13=8, l_feat_14=57, l_feat_15=51, l_feat_16=8, l_feat_17=22, l_feat_18=8, l_feat_19=8, l_feat_20=8,
             l_feat_21=8, l_feat_22=8, l_feat_23=8, l_feat_24=8, l_feat_25=43, l_feat_26=8, l_feat_27=8, feat_a_2=8,
             feat_a_8=8, feat_a_9=8)
N_NUM = 82            # len(num_cols_explicit) in dare_qnn_next.yaml; X_mask has the same width
SEQ_VOCAB = 10_000_000  # hard-coded in the reference: src/train.py:116, src/infer.py:77


def dare_qnn_next() -> dict:
    """cfgs/dare_qnn_next.yaml — BASELINE configs[1], the single-GPU headline workload."""
    return {
        "exp_name": "dare_qnn_next", "seed": 777, "device": "cuda", "amp": "none",
        "data": {"cat_cols": list(_CATS), "hash_buckets": dict(_BUCKETS)},
        "sequence": {"col": "seq", "max_len": 400, "pad_id": 0, "top_k": 60, "recency_tau": 512, "query_mode": "concat",
                     "query_key": "inventory_id", "transformer_block": True,
                     "tfm": {"n_layers": 3, "n_heads": 8, "mha_dropout": 0.1, "ffn_hidden": 384, "ffn_dropout": 0.1,
                             "norm": "rms", "gating": "softmax", "add_positional_bias": True}},
        "model": {"emb_dim": 64, "dare_dropout": 0.2, "cat_embedding_dims": dict(_DIMS),
                  "qnn_alpha": {"enabled": True, "feature_embed_dim": 32, "heads": 6, "rank": 16, "proj_dim": 192,
                                "mlp_hidden": [512, 256], "dropout": 0.2, "use_se": True, "se_reduction": 8,
                                "use_residual": True, "norm": "rms", "pair_grouping": "all", "aux_head_weight": 0.1}},
        "train": {"batch_size": 8192, "epochs": 8, "lr": 3e-4, "weight_decay": 1e-4, "warmup_epochs": 2,
                  "grad_clip_norm": 0.5, "early_stop_patience": 3},
        "ema": {"enabled": True, "decay": 0.999, "eval_with_ema": True},
    }


def dare_qnn_next_k100_s1() -> dict:
    """cfgs/dare_qnn_next_k100_s1.yaml — configs[2]: top_k=100, S1 query."""
    c = dare_qnn_next()
    c["exp_name"] = "dare_qnn_next_k100_s1"
    c["sequence"].update(top_k=100, query_mode="S1")
    return c


def v3_k148_s1() -> dict:
    """cfgs/v3_k148_s1.yaml — configs[3]: top_k=148, 4 encoder layers, B=4096, margin 500, EMA off."""
    c = dare_qnn_next()
    c["exp_name"] = "v3_k148_s1"
    c["data"]["hash_buckets_margin"] = 500
    c["sequence"].update(top_k=148, query_mode="S1")
    c["sequence"]["tfm"].update(n_layers=4, ffn_dropout=0.15)
    c["model"]["dare_dropout"] = 0.25
    c["model"]["cat_embedding_dims"]["inventory_id"] = 32          # cfgs/v3_k148_s1.yaml:206
    c["train"]["batch_size"] = 4096
    c["ema"] = {"enabled": False}
    return c


def v3_k148_s1_big() -> dict:
    """BASELINE configs[3]: cfgs/v3_k148_s1.yaml with the enlarged hash_buckets of cfgs/v4_k120_s1.yaml:142,156,204,218
    (inventory_id 2,000,003 x 32, l_feat_14 4,000,013 x 57; + the 500-row margin): 1.17 GB of hashed-categorical tables, kept
    under the row-lazy optimizer (FusedTrainer: tables with >= 500,000 rows)."""
    c = v3_k148_s1()
    c["exp_name"] = "v3_k148_s1_big"
    c["data"]["hash_buckets"].update(inventory_id=2_000_003, l_feat_14=4_000_013)
    c["model"]["cat_embedding_dims"].update(inventory_id=32, l_feat_14=57)
    return c


def dare_base() -> dict:
    """cfgs/dare_base.yaml — configs[0]: S1, no encoder, FC head, two large hashed tables (2,000,000 and 1,000,000 rows)."""
    return {
        "exp_name": "dare_base", "seed": 777, "device": "cuda", "amp": "none",
        "data": {"cat_cols": ["gender", "age_group", "inventory_id", "day_of_week", "hour", "l_feat_14"],
                 "hash_buckets": dict(gender=16, age_group=64, day_of_week=8, hour=32, inventory_id=2_000_000,
                                      l_feat_14=1_000_000)},
        "sequence": {"col": "seq", "max_len": 400, "pad_id": 0, "top_k": 80, "recency_tau": 256, "query_mode": "S1",
                     "query_key": "inventory_id", "transformer_block": False},
        "model": {"emb_dim": 64, "dare_dropout": 0.1, "qnn_alpha": {"enabled": False}},
        "train": {"batch_size": 16384, "epochs": 10, "lr": 1e-3, "weight_decay": 1e-4, "warmup_epochs": 1,
                  "grad_clip_norm": 1.0, "early_stop_patience": 3},
    }


def cardinals(cfg: dict) -> dict:
    """train.py:119."""
    d = cfg["data"]
    return {c: d["hash_buckets"].get(c, 1000003) + d.get("hash_buckets_margin", 0) for c in d["cat_cols"]}


def shrink(cfg: dict, **over) -> dict:
    c = copy.deepcopy(cfg)
    for k, v in over.items():
        node = c
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return c
