"""Namespace-package overlay for `src.utils.ema` (reference src/utils/ema.py): the same `ModelEMA` / `build_ema` surface on the
fused kernels -- one multi-tensor EMA launch per step, tensor exchanges instead of state_dict copies around validation -- and,
on import, the fused AdamW behind `torch.optim.AdamW` (b200ctr/shims.py).  Every other `src.utils.*` module
still resolves to the reference."""
from b200ctr.shims import ModelEMA, build_ema, install_optimizer_shim  # noqa: F401

install_optimizer_shim()
