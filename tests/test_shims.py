"""Host logic of the optimizer / EMA shims for the unmodified reference loop (b200ctr/shims.py) on the CPU:
the EMA schedule and skip rules of utils/ema.py:72-131, the exchange-not-copy store / copy_to / restore, the state_dict format,
and the AdamW shim delegating to torch for parameters it cannot fuse."""
import math

import torch

from b200ctr import shims


def _ref_decay(d, warm, kind, step):                      # utils/ema.py:72-88
    if warm <= 0 or kind == "none":
        return d
    t = min(1.0, (step + 1) / warm)
    if kind == "linear":
        return 1.0 - (1.0 - d) * t
    return 1.0 - (1.0 - d) * 0.5 * (1 + math.cos(math.pi * (1 - t)))


def test_ema_schedule_skip_rules_and_values():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.BatchNorm1d(4))
    e = shims.ModelEMA(m, base_decay=0.99, warmup_steps=4, warmup_type="cosine", update_after_step=2, update_interval=3,
                       param_filter=["0.bias"])
    want = {n: p.detach().clone() for n, p in m.named_parameters()}
    n_upd = 0
    for g in range(12):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.1 * (g + 1))
        e.update(m, g)
        if g >= 2 and (g - 2) % 3 == 0:
            d = _ref_decay(0.99, 4, "cosine", n_upd)
            for n, p in m.named_parameters():
                if n != "0.bias":
                    want[n] = want[n] * d + p.detach() * (1 - d)
            n_upd += 1
    assert e.num_updates == n_upd == 4
    for n in want:
        assert torch.allclose(e.shadow_params[n], want[n], atol=1e-6), n
    assert torch.equal(e.shadow_buffers["1.running_mean"], m[1].running_mean)          # ema_on_buffers="copy"


def test_store_copy_to_restore_exchange_tensors():
    m = torch.nn.Linear(3, 2)
    e = shims.ModelEMA(m, base_decay=0.5)
    e.update(m, 0)
    with torch.no_grad():
        m.weight.add_(1.0)
    train_ptr, train_val = m.weight.data_ptr(), m.weight.detach().clone()
    e.store(m)
    e.copy_to(m)
    assert m.weight.data_ptr() == e.shadow_params["weight"].data_ptr()               # an alias, not a copy
    assert not torch.equal(m.weight.detach(), train_val)
    e.restore(m)
    assert m.weight.data_ptr() == train_ptr and torch.equal(m.weight.detach(), train_val)
    e.restore(m)                                                                     # second restore is a no-op (ema.py:160-162)


def test_state_dict_round_trip():
    m = torch.nn.Linear(3, 2)
    a = shims.ModelEMA(m, base_decay=0.9, warmup_steps=7, param_filter=["bias"])
    a.update(m, 0)
    sd = a.state_dict()
    assert set(sd) == {"base_decay", "warmup_steps", "warmup_type", "update_after_step", "update_interval", "ema_on_buffers",
                       "offload_to_cpu", "pin_memory", "param_filter", "num_updates", "shadow_params", "shadow_buffers"}
    b = shims.ModelEMA(torch.nn.Linear(3, 2), base_decay=0.1)
    b.load_state_dict(sd)
    assert b.base_decay == 0.9 and b.warmup_steps == 7 and b.num_updates == 1 and b.param_filter == {"bias"}
    assert torch.equal(b.shadow_params["weight"], a.shadow_params["weight"])


def test_adamw_shim_is_torch_adamw_on_cpu_parameters():
    torch.manual_seed(1)
    w0 = torch.randn(4, 3)
    outs = []
    for cls in (torch.optim.AdamW if torch.optim.AdamW is not shims.AdamW else shims.AdamW.__mro__[1], shims.AdamW):
        p = torch.nn.Parameter(w0.clone())
        opt = cls([p], lr=1e-2, weight_decay=0.1)
        for s in range(3):
            p.grad = torch.full_like(p, 0.5 * (s + 1))
            opt.step()
        outs.append(p.detach().clone())
        assert set(opt.state[p]) == {"step", "exp_avg", "exp_avg_sq"}
    assert torch.equal(outs[0], outs[1])
    assert issubclass(shims.AdamW, shims.AdamW.__mro__[1])
