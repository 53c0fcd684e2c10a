"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

from oracle import ctr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("qnn_concat", "base_s1", "s2_relu")


def load_case(kind):
    z = np.load(os.path.join(GOLDEN, f"model_{kind}.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    return z, cfg


def group(z, prefix, dtype=torch.float32):
    out = {}
    for k in z.files:
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(z[k])).to(dtype) if z[k].dtype.kind == "f" \
                else torch.from_numpy(np.array(z[k]))
    return out


def spec_of(cfg):
    g = cfg["golden"]
    return O.ModelSpec(cfg, g["num_dim"], g["mask_dim"], cfg["data"]["cat_cols"])


def batch_of(z, s):
    b = {k: torch.from_numpy(np.array(z[f"step{s}/batch/{k}"])) for k in ("X_num", "X_mask", "X_cat", "seq")}
    y = torch.from_numpy(np.array(z[f"step{s}/y"]))
    return b, y


def rel_err(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err_scaled(a, b):
    """max |a-b| / (max|b| + tiny): the fp32 'rel <= 1e-4' criterion used throughout."""
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
