"""Fold-ensemble inference: the inner loop of the reference's src/infer.py:82-162 as one object.

The reference keeps one CTRModel per fold checkpoint, and for every batch calls each model (every `forward` re-copies the
batch host -> device, wrapper.py:139-143), runs the fold's Calibrator on the CPU in numpy (infer.py:112-116: a D2H of the logits,
a sigmoid, an H2D back), clamps, stacks and calls `ensemble_probs("logit_mean", ...)` (utils/metrics.py:68-71).  Here the batch is
uploaded ONCE, every fold's eval forward (hotpath.model_fwd on that fold's parameters) writes its logits into one (folds, B)
buffer, and a single kernel (`b200ctr_ensemble_logit_mean`) does temperature scaling (1/T per fold), the reference's clipping,
logit, mean over folds and the final sigmoid.  Isotonic calibrators (CPU sklearn objects) and the other ensemble methods are
host-side per-batch math the reference keeps in numpy / torch; they are passed through unchanged (`method` != logit_mean or a
calibrator with an isotonic stage falls back to per-fold probabilities + the caller's own ensemble function).

Fold-parallel over GPUs: rank r evaluates folds r, r + N, ...; the (folds, B) logit rows are all-gathered before the
ensemble kernel, which reproduces the single-process result exactly (the SE batch statistics are per model, untouched).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import hotpath as hp
from . import ops
from ._lib import B200CtrError
from .modules import CTRModel


def _temperature_of(cal) -> Optional[float]:
    """1 fold's temperature: Calibrator.temp_scaler (utils/calibration.py:11-21) -> clamp(exp(log_temp), clamp_T)."""
    if cal is None:
        return None
    ts = getattr(cal, "temp_scaler", None)
    if ts is None:
        return None
    T = float(torch.exp(ts.log_temp.detach()))
    lo, hi = getattr(cal, "clamp_T", (0.2, 5.0)) or (0.0, float("inf"))
    return min(max(T, lo), hi)


class FoldEnsemble:
    def __init__(self, models: Sequence[CTRModel], calibrators: Optional[Sequence[object]] = None, eps: float = 1e-7,
                 process_group=None):
        if not models:
            raise B200CtrError("FoldEnsemble needs at least one fold model")
        self.models = list(models)
        for m in self.models:
            m.eval()
        self.spec = self.models[0]._spec
        self.dev = next(self.models[0].parameters()).device
        if self.dev.type != "cuda":
            raise B200CtrError("FoldEnsemble: models must live on a CUDA device (no CPU fallback)")
        cals = list(calibrators) if calibrators is not None else [None] * len(self.models)
        self.fusable = all(c is None or getattr(c, "iso", None) is None for c in cals)
        self.calibrators = cals
        Ts = [_temperature_of(c) for c in cals]
        self.inv_T = None
        if any(t is not None for t in Ts):
            self.inv_T_local = [1.0 / t if t is not None else 1.0 for t in Ts]
        else:
            self.inv_T_local = None
        self.eps = float(eps)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self._P = [dict(m.named_parameters()) for m in self.models]
        self._cache: List[dict] = [dict() for _ in self.models]

    def device_batch(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """ONE host -> device copy per batch (the reference copies once per fold model)."""
        return self.models[0]._device_batch(batch)

    @torch.no_grad()
    def fold_logits(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """(n_local_folds, B) logits of this rank's folds on one shared device batch."""
        db = batch if batch["seq"].is_cuda else self.device_batch(batch)
        B = db["seq"].shape[0]
        out = torch.empty((len(self.models), B), device=self.dev, dtype=torch.float32)
        dc = hp.DropCtl(False)
        for i, P in enumerate(self._P):
            logits, _, _ = hp.model_fwd(P, self.spec, db, dc, self._cache[i])
            out[i].copy_(logits)
        return out

    @torch.no_grad()
    def predict(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Ensembled probability (B,) for `ensemble.method: logit_mean` -- infer.py:104-158 for one batch."""
        if not self.fusable:
            raise B200CtrError("FoldEnsemble.predict fuses temperature calibration only; a fold with an isotonic stage needs the "
                               "reference's host-side Calibrator.predict_proba on fold_logits()")
        z = self.fold_logits(batch)
        inv_T = None
        if self.inv_T_local is not None:
            inv_T = torch.tensor(self.inv_T_local, device=self.dev, dtype=torch.float32)
        if self.world > 1:                                            # C6: gather every rank's fold rows
            zs = [torch.empty_like(z) for _ in range(self.world)]
            torch.distributed.all_gather(zs, z, group=self.pg)
            z = torch.cat(zs, 0)
            if inv_T is not None:
                ts = [torch.empty_like(inv_T) for _ in range(self.world)]
                torch.distributed.all_gather(ts, inv_T, group=self.pg)
                inv_T = torch.cat(ts, 0)
        if z.shape[0] == 1 and inv_T is None:
            # single model: the reference skips the ensemble and returns the clamped probability (infer.py:126-127)
            p = ops.sigmoid(z[0])
            return torch.clamp(p, self.eps, 1.0 - self.eps)
        return ops.ensemble_logit_mean(z, inv_T, self.eps)
