"""GPU parity of the fused optimizer / EMA shims (b200ctr/shims.py + b200ctr_multi_adamw_ema) against torch.optim.AdamW and the
reference's EMA arithmetic (utils/ema.py:91-131) on ragged tensor lists, through the C ABI."""
import pytest
import torch

from b200ctr import shims

pytestmark = pytest.mark.gpu

SHAPES = [(7,), (64, 64), (1, 1), (100003, 8), (70000,), (3, 5, 2)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in SHAPES]


def test_fused_adamw_matches_torch_adamw():
    base = torch.optim.AdamW.__mro__[1] if torch.optim.AdamW is shims.AdamW else torch.optim.AdamW
    pa, pb = _params(0), _params(0)
    oa = base(pa, lr=3e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    ob = shims.AdamW(pb, lr=3e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    g = torch.Generator().manual_seed(5)
    for step in range(6):
        lr = 3e-3 * (step + 1) / 6
        for o in (oa, ob):
            for grp in o.param_groups:
                grp["lr"] = lr                                # train.py sets the lr of every group each step
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.0 if step == 3 else 1.0)
            a.grad, b.grad = gr.clone(), gr.clone()
        pb[2].grad = None if step == 1 else pb[2].grad        # a parameter without a gradient is skipped, like torch does
        pa[2].grad = None if step == 1 else pa[2].grad
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
        assert torch.allclose(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-12)      # fma vs mul + addcmul
        assert float(oa.state[a]["step"]) == float(ob.state[b]["step"])
    assert oa.state_dict()["param_groups"][0].keys() == ob.state_dict()["param_groups"][0].keys()


def test_fused_ema_update_and_aliasing_swap():
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList(_params(3))
            self.frozen = torch.nn.Parameter(torch.ones(4).cuda(), requires_grad=False)
    m = M()
    e = shims.ModelEMA(m, base_decay=0.999, warmup_steps=10, param_filter=["ps.1"])
    want = {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
    for step in range(4):
        with torch.no_grad():
            for p in m.ps:
                p.mul_(1.01).add_(0.01)
        e.update(m, step)
        d = 1.0 - (1.0 - 0.999) * min(1.0, (step + 1) / 10)
        for n, p in m.named_parameters():
            if p.requires_grad and n != "ps.1":
                want[n] = want[n] * d + p.detach() * (1 - d)
    assert "frozen" not in e.shadow_params
    for n in want:
        assert torch.allclose(e.shadow_params[n], want[n], rtol=1e-6, atol=1e-7), n
    live = {n: (p.data_ptr(), p.detach().clone()) for n, p in m.named_parameters()}
    e.store(m); e.copy_to(m)
    assert all(p.data_ptr() == e.shadow_params[n].data_ptr() for n, p in m.named_parameters() if p.requires_grad)
    e.restore(m)
    assert all(p.data_ptr() == live[n][0] and torch.equal(p.detach(), live[n][1]) for n, p in m.named_parameters())
    sd = e.state_dict()
    assert all(not t.is_cuda for t in sd["shadow_params"].values())
    e2 = shims.ModelEMA(m)
    e2.load_state_dict(sd)
    e2.copy_to(m)                                            # infer.py:90-92: CPU shadows from a checkpoint land on the model's device
    assert all(p.is_cuda and torch.allclose(p.detach(), want[n], rtol=1e-6, atol=1e-7) for n, p in m.named_parameters() if p.requires_grad)
