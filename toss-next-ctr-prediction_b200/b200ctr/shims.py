"""Optimizer / EMA shims for the UNMODIFIED reference loop (Tier A of the drop-in).

`src/train.py` builds `torch.optim.AdamW(model.parameters(), ...)` (train.py:138) and `ModelEMA` (utils/ema.py) itself; with
138 separate parameter tensors that is ~550 elementwise launches per step and, around every validation pass, three full
`state_dict` copies (store / copy_to / restore: 3 x 5.2 GB with the two 10 M-row tables).  The overlay replaces both:

  * `ModelEMA` (drop-in for utils/ema.py:11-198, same constructor, schedule, skip rules, state_dict format): the shadows are
    updated by ONE multi-tensor kernel (b200ctr_multi_adamw_ema, mode 1); `store` / `copy_to` / `restore` exchange the tensors
    behind `param.data` instead of copying them (valid because the loop only evaluates between copy_to and restore).
  * `AdamW` (subclass of torch.optim.AdamW, same state keys `step` / `exp_avg` / `exp_avg_sq`): one multi-tensor launch per
    param group for dense fp32 CUDA parameters, torch's own step for anything else.  Installed over `torch.optim.AdamW` when the
    overlay's `src.utils.ema` is imported (B200CTR_OPTIM_SHIM=0 leaves torch's class in place).

Dense on purpose: inside the reference loop the table gradients are dense `.grad` tensors that `clip_grad_norm_` reads in full,
so the ROW-LAZY optimizer stays with `FusedTrainer` (Tier B), whose step owns the whole train.py:152-203 body.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check

CHUNK = 1 << 16


class _MultiTensor:
    """Chunk table + record table for one fixed list of tensor sizes; the pointer records are refreshed per call."""

    def __init__(self, sizes: List[int], device: torch.device):
        self.device = device
        self.n = len(sizes)
        ct, co = [], []
        for t, n in enumerate(sizes):
            k = (n + CHUNK - 1) // CHUNK
            ct.append(np.full(k, t, dtype=np.int32))
            co.append(np.arange(k, dtype=np.int64) * CHUNK)
        self.chunk_tensor = torch.from_numpy(np.concatenate(ct) if ct else np.zeros(0, np.int32)).to(device)
        self.chunk_off = torch.from_numpy(np.concatenate(co) if co else np.zeros(0, np.int64)).to(device)
        self.sizes = sizes
        self.host = torch.zeros((max(self.n, 1), 6), dtype=torch.int64).pin_memory()
        self.recs = torch.zeros((max(self.n, 1), 6), dtype=torch.int64, device=device)

    def launch(self, p, g, m, v, ema, *, lr=0.0, wd=0.0, betas=(0.9, 0.999), eps=1e-8, step=1, ema_decay=0.0, mode=0):
        h = self.host.numpy()
        z = [None] * self.n
        for i, (a, b, c, d, e) in enumerate(zip(p, g or z, m or z, v or z, ema or z)):
            h[i] = (a.data_ptr(), b.data_ptr() if b is not None else 0, c.data_ptr() if c is not None else 0,
                    d.data_ptr() if d is not None else 0, e.data_ptr() if e is not None else 0, self.sizes[i])
        self.recs.copy_(self.host, non_blocking=True)
        lib = _lib.load()
        check(lib.b200ctr_multi_adamw_ema(self.recs.data_ptr(), self.chunk_tensor.data_ptr(), self.chunk_off.data_ptr(),
                                          self.chunk_tensor.numel(), CHUNK, float(lr), float(wd), float(betas[0]), float(betas[1]),
                                          float(eps), int(step), None, float(ema_decay), int(mode),
                                          torch.cuda.current_stream(self.device).cuda_stream))


def _fusable(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and not t.is_sparse


class AdamW(torch.optim.AdamW):
    """torch.optim.AdamW with the dense fp32 CUDA parameters of a group stepped by one multi-tensor kernel."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)
        self._mt: Dict[tuple, _MultiTensor] = {}

    @torch.no_grad()
    def step(self, closure=None):
        plain = any(g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable") for g in self.param_groups)
        live = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        if plain or not live or not all(_fusable(p) and _fusable(p.grad) for p in live):
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            by_step: Dict[int, List[torch.nn.Parameter]] = {}
            for p in ps:
                by_step.setdefault(int(self.state[p]["step"].item() if torch.is_tensor(self.state[p]["step"]) else self.state[p]["step"]), []).append(p)
            for s0, grp in by_step.items():
                key = (gi, tuple(id(p) for p in grp))
                mt = self._mt.get(key)
                if mt is None:
                    mt = self._mt[key] = _MultiTensor([p.numel() for p in grp], grp[0].device)
                mt.launch([p.data for p in grp], [p.grad for p in grp], [self.state[p]["exp_avg"] for p in grp],
                          [self.state[p]["exp_avg_sq"] for p in grp], None, lr=group["lr"], wd=group["weight_decay"],
                          betas=group["betas"], eps=group["eps"], step=s0 + 1, mode=0)
                for p in grp:
                    self.state[p]["step"] += 1
        return loss


@torch.no_grad()
def _to_fp32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32)


class ModelEMA:
    """Drop-in for the reference's ModelEMA (utils/ema.py:11-198)."""

    def __init__(self, model: torch.nn.Module, base_decay: float = 0.999, warmup_steps: int = 0, warmup_type: str = "linear",
                 update_after_step: int = 0, update_interval: int = 1, ema_on_buffers: str = "copy", offload_to_cpu: bool = False,
                 pin_memory: bool = False, param_filter: Optional[Iterable[str]] = None):
        self.base_decay = float(base_decay)
        self.warmup_steps = int(max(0, warmup_steps))
        self.warmup_type = warmup_type
        self.update_after_step = int(max(0, update_after_step))
        self.update_interval = int(max(1, update_interval))
        self.ema_on_buffers = ema_on_buffers
        self.offload_to_cpu = bool(offload_to_cpu)
        self.pin_memory = bool(pin_memory)
        self.param_filter = set(param_filter or [])
        self.num_updates = 0
        self._param_names = [n for n, p in model.named_parameters() if p.requires_grad]
        self._buffer_names = [n for n, _ in model.named_buffers()]
        # FP32 shadows of the values AT CONSTRUCTION (utils/ema.py:52-66): the first update blends them with the stepped weights
        self.shadow_params: Dict[str, torch.Tensor] = {n: self._place(p) for n, p in model.named_parameters() if p.requires_grad}
        self.shadow_buffers: Dict[str, torch.Tensor] = {n: self._place(b) for n, b in model.named_buffers()}
        self._saved = None
        self._saved_data = None
        self._mt: Optional[_MultiTensor] = None
        self._mt_names: List[str] = []

    def _place(self, t: torch.Tensor) -> torch.Tensor:
        t = _to_fp32(t)
        if self.offload_to_cpu:
            t = t.cpu().pin_memory() if self.pin_memory else t.cpu()
        return t.clone()

    def _decay_at(self, step: int) -> float:
        """utils/ema.py:72-88"""
        d = self.base_decay
        if self.warmup_steps <= 0 or self.warmup_type == "none":
            return d
        t = min(1.0, (step + 1) / self.warmup_steps)
        if self.warmup_type == "linear":
            decay = 1.0 - (1.0 - d) * t
        elif self.warmup_type == "cosine":
            c = 0.5 * (1 + math.cos(math.pi * (1 - t)))
            decay = 1.0 - (1.0 - d) * c
        else:
            decay = d
        return float(max(0.0, min(1.0, decay)))

    @torch.no_grad()
    def update(self, model: torch.nn.Module, global_step: int):
        """utils/ema.py:91-131.  Call AFTER optimizer.step()."""
        if global_step < self.update_after_step:
            return
        if (global_step - self.update_after_step) % self.update_interval != 0:
            return
        decay = self._decay_at(self.num_updates)
        one_minus = 1.0 - decay
        sp = self.shadow_params
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in self.param_filter]
        fused = [(n, p) for n, p in named if _fusable(p.data) and _fusable(sp[n])]
        if fused:
            names = [n for n, _ in fused]
            if self._mt is None or self._mt_names != names:
                self._mt = _MultiTensor([p.numel() for _, p in fused], fused[0][1].device)
                self._mt_names = names
            self._mt.launch([p.data for _, p in fused], None, None, None, [sp[n] for n in names], ema_decay=decay, mode=1)
        done = {n for n, _ in fused}
        for n, p in named:                          # CPU-offloaded shadows / non-fp32 parameters: the reference's own arithmetic
            if n in done:
                continue
            cur = _to_fp32(p)
            if self.offload_to_cpu:
                cur = cur.cpu()
            sp[n].mul_(decay).add_(cur, alpha=one_minus)
        for n, b in model.named_buffers():
            sb = self.shadow_buffers[n]
            cur = _to_fp32(b)
            if self.offload_to_cpu:
                cur = cur.cpu()
            if self.ema_on_buffers == "copy":
                sb.copy_(cur)
            else:
                sb.mul_(decay).add_(cur, alpha=one_minus)
        self.num_updates += 1

    # ---- swap-in / swap-out for eval: tensor exchanges, no copies (utils/ema.py:135-165 copies the state_dict 3 times) ----
    @torch.no_grad()
    def store(self, model: torch.nn.Module):
        self._saved = {n: t for n, t in self._tensors(model).items()}
        self._saved_data = {n: t.data for n, t in self._saved.items()}

    @staticmethod
    def _tensors(model: torch.nn.Module) -> Dict[str, torch.Tensor]:
        out = dict(model.named_parameters())
        out.update(dict(model.named_buffers()))
        return out

    @torch.no_grad()
    def copy_to(self, model: torch.nn.Module):
        for n, tgt in self._tensors(model).items():
            if n in self.shadow_params:
                src = self.shadow_params[n]
            elif n in self.shadow_buffers:
                src = self.shadow_buffers[n]
            else:
                continue
            src = src.to(device=tgt.device, dtype=tgt.dtype, non_blocking=True)      # no-op (an alias) for fp32 shadows on the device
            if src.shape != tgt.shape:
                raise RuntimeError(f"EMA shadow {n}: shape {tuple(src.shape)} does not match the model's {tuple(tgt.shape)}")
            tgt.data = src

    @torch.no_grad()
    def restore(self, model: torch.nn.Module):
        if self._saved is None:
            return
        for n, t in self._saved.items():
            t.data = self._saved_data[n]
        self._saved = None
        self._saved_data = None

    # ---- serialization (utils/ema.py:168-198: same keys) ----
    def state_dict(self) -> Dict:
        return {
            "base_decay": self.base_decay, "warmup_steps": self.warmup_steps, "warmup_type": self.warmup_type,
            "update_after_step": self.update_after_step, "update_interval": self.update_interval,
            "ema_on_buffers": self.ema_on_buffers, "offload_to_cpu": self.offload_to_cpu, "pin_memory": self.pin_memory,
            "param_filter": list(self.param_filter), "num_updates": self.num_updates,
            "shadow_params": {k: v.cpu() if v.is_cuda else v for k, v in self.shadow_params.items()},
            "shadow_buffers": {k: v.cpu() if v.is_cuda else v for k, v in self.shadow_buffers.items()},
        }

    def load_state_dict(self, state: Dict):
        self.base_decay = float(state["base_decay"])
        self.warmup_steps = int(state["warmup_steps"])
        self.warmup_type = state["warmup_type"]
        self.update_after_step = int(state["update_after_step"])
        self.update_interval = int(state["update_interval"])
        self.ema_on_buffers = state["ema_on_buffers"]
        self.offload_to_cpu = bool(state.get("offload_to_cpu", False))
        self.pin_memory = bool(state.get("pin_memory", False))
        self.param_filter = set(state.get("param_filter", []))
        self.num_updates = int(state.get("num_updates", 0))
        self.shadow_params = {k: v.clone().detach() for k, v in state["shadow_params"].items()}
        self.shadow_buffers = {k: v.clone().detach() for k, v in state["shadow_buffers"].items()}
        self._mt = None


def build_ema(model, cfg):
    """utils/ema.py:200-216"""
    if not cfg.get("ema", {}).get("enabled", False):
        return None
    ec = cfg["ema"]
    return ModelEMA(model=model, base_decay=float(ec.get("decay", 0.999)), warmup_steps=int(ec.get("warmup_steps", 0)),
                    warmup_type=str(ec.get("warmup_type", "linear")), update_after_step=int(ec.get("update_after_step", 0)),
                    update_interval=int(ec.get("update_interval", 1)), ema_on_buffers=str(ec.get("ema_on_buffers", "copy")),
                    offload_to_cpu=bool(ec.get("offload_to_cpu", False)), pin_memory=bool(ec.get("pin_memory", False)),
                    param_filter=ec.get("param_filter", []))


def install_optimizer_shim() -> bool:
    """Put the fused AdamW where `torch.optim.AdamW(...)` in train.py:138 finds it (B200CTR_OPTIM_SHIM=0 opts out)."""
    if os.environ.get("B200CTR_OPTIM_SHIM", "1") == "0":
        return False
    if torch.optim.AdamW is not AdamW:
        torch.optim.AdamW = AdamW
    return True
