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
    """max |a-b| / (max|b| + tiny).  This is the criterion behind every "<= 1e-4" in the suite: a PER-TENSOR max-norm bound
    (the largest absolute deviation relative to the tensor's largest magnitude), not a per-element relative error -- elements
    much smaller than the tensor's maximum may differ by more than 1e-4 of their own value."""
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ---- full-width fixtures (oracle/make_golden_full.py): seeded init + strided subsamples --------------------------------
FULL_CASES = ("full_k60_concat", "full_k100_s1", "full_k148_s1")


def load_full(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    return z, cfg


def full_batch(z, s):
    """collate_sharded dtypes (dataset.py:98-124): f32 X_num / X_mask, i64 X_cat / seq (stored narrow in the fixture)."""
    b = {"X_num": torch.from_numpy(np.array(z[f"step{s}/batch/X_num"])),
         "X_mask": torch.from_numpy(np.array(z[f"step{s}/batch/X_mask"]).astype(np.float32)),
         "X_cat": torch.from_numpy(np.array(z[f"step{s}/batch/X_cat"]).astype(np.int64)),
         "seq": torch.from_numpy(np.array(z[f"step{s}/batch/seq"]).astype(np.int64))}
    return b, torch.from_numpy(np.array(z[f"step{s}/y"]))


def full_init(model_cls, cfg, z):
    """Rebuild the fixture's seeded initial parameters with `model_cls` (CPU) and hold them to the stored checksums."""
    from oracle.make_golden_full import build_init
    m = build_init(model_cls, cfg)
    for n, p in m.state_dict().items():
        a = p.detach().double().reshape(-1)
        ref = z["initsum/" + n]
        assert int(ref[2]) == a.numel(), n
        assert abs(float(a.sum()) - ref[0]) <= 1e-9 * max(1.0, abs(ref[0])) + 1e-9, f"seeded init drifted: {n}"
        assert abs(float((a * a).sum()) - ref[1]) <= 1e-9 * max(1.0, ref[1]), f"seeded init drifted: {n}"
    return m


def packed_err(z, key, t, l2=False):
    """(max-abs error of the stored subsample / max|ref|, relative error of the whole-tensor L2 norm[, relative L2 error of the
    subsample])."""
    from oracle.make_golden_full import sample_stride
    ref = torch.from_numpy(np.array(z[key])).double()
    st = z[key + "#stats"]
    a = t.detach().double().cpu().reshape(-1)
    assert a.numel() == int(st[2]), (key, a.numel(), st[2])
    sub = a[::sample_stride(a.numel())]
    e_sub = float((sub - ref).abs().max() / (ref.abs().max() + 1e-30))
    nref = float(st[1]) ** 0.5
    e_norm = abs(float(a.norm()) - nref) / (nref + 1e-30)
    if l2:
        return e_sub, e_norm, float((sub - ref).norm() / (ref.norm() + 1e-30))
    return e_sub, e_norm


def packed_keys(z, prefix):
    return [k[len(prefix):] for k in z.files if k.startswith(prefix) and not k.endswith("#stats")]
