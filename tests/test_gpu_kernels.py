"""GPU parity tests of individual libb200ctr kernels against the oracle / exact arithmetic (run with -m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import GOLDEN, max_err_scaled

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: fp32 rel <= 1e-4


@pytest.fixture(scope="module")
def dev():
    from b200ctr import _lib
    _lib.check(_lib.load().b200ctr_check_device(), "device")
    return torch.device("cuda:0")


def _rand(gen, *shape):
    return torch.randn(*shape, generator=gen)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 5, 19), (128, 64, 16), (300, 70, 129), (257, 200, 64), (64, 513, 300),
                                   (1000, 1, 256), (1, 256, 1000)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_all_layouts(dev, M, N, K, ta, tb):
    from b200ctr import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = _rand(g, K, M) if ta else _rand(g, M, K)
    b = _rand(g, N, K) if tb else _rand(g, K, N)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    out = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb)
    assert max_err_scaled(out.cpu(), ref) < 1e-5


def test_gemm_strided_views_and_splitk(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(5)
    big = _rand(g, 4096, 96).to(dev)
    w = _rand(g, 40, 50).to(dev)
    a = big[:, 16:56]                       # (4096, 40) ld 96
    ref = a.double().cpu().t() @ big[:, 3:53].double().cpu()
    for sk in (1, 4, 16):
        out = ops.gemm(a, big[:, 3:53], trans_a=True, split_k=sk, tc=False)     # CUDA-core kernel; unaligned second operand -> scalar path
        assert max_err_scaled(out.cpu(), ref) < 1e-5
    outv = torch.zeros(4096, 128, device=dev)
    ops.gemm(a, w, out=outv[:, 64:114])
    assert max_err_scaled(outv[:, 64:114].cpu(), a.double().cpu() @ w.double().cpu()) < 1e-5
    assert float(outv[:, :64].abs().max()) == 0.0 and float(outv[:, 114:].abs().max()) == 0.0


def test_gemm_epilogues(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(11)
    M, N, K = 200, 96, 64
    a, w, bias, res, acc = _rand(g, M, K), _rand(g, N, K), _rand(g, N), _rand(g, M, N), _rand(g, M, N)
    ad, wd, bd, rd, accd = (t.to(dev) for t in (a, w, bias, res, acc))
    lin = a.double() @ w.double().t()
    # bias + gelu + saved pre-activation + residual
    pre = torch.empty(M, N, device=dev)
    out = ops.gemm(ad, wd, trans_b=True, bias=bd, act=ops.ACT_GELU, pre_out=pre, residual=rd)
    pre_ref = lin + bias.double()
    assert max_err_scaled(pre.cpu(), pre_ref) < 1e-5
    assert max_err_scaled(out.cpu(), O.gelu_fwd(pre_ref) + res.double()) < 1e-5
    # accum (pre-activation) + relu
    out = ops.gemm(ad, wd, trans_b=True, accum=accd, bias=bd, act=ops.ACT_RELU)
    assert max_err_scaled(out.cpu(), torch.relu(lin + acc.double() + bias.double())) < 1e-5
    # masks
    out = ops.gemm(ad, wd, trans_b=True, mask_mode=ops.MASK_NONZERO, mask_src=torch.relu(rd), mask_scale=1.25)
    assert max_err_scaled(out.cpu(), lin * (res > 0).double() * 1.25) < 1e-5
    out = ops.gemm(ad, wd, trans_b=True, mask_mode=ops.MASK_GELU_GRAD, mask_src=rd)
    assert max_err_scaled(out.cpu(), lin * O.gelu_grad(res.double())) < 1e-5
    # split-k goes through the same epilogue
    out = ops.gemm(ad, wd, trans_b=True, bias=bd, act=ops.ACT_RELU, split_k=4)
    assert max_err_scaled(out.cpu(), torch.relu(lin + bias.double())) < 1e-5


def test_gemm_dropout_is_regenerable_and_unbiased(dev):
    from b200ctr import ops
    M, N, K = 512, 384, 8
    a = torch.ones(M, K, device=dev)
    w = torch.ones(N, K, device=dev) / K
    d = ops.drop(0.25, 1234567, 77)
    o1 = ops.gemm(a, w, trans_b=True, dropout=d, tc=True)
    o2 = ops.gemm(a, w, trans_b=True, dropout=d)
    assert torch.equal(o1, o2)                                   # stateless: backward regenerates the same mask
    vals = torch.unique(o1)
    assert set(np.round(vals.cpu().numpy(), 5)) <= {0.0, np.float32(round(1 / 0.75, 5))}
    keep = float((o1 != 0).float().mean())
    assert abs(keep - 0.75) < 0.01
    o3 = ops.gemm(a, w, trans_b=True, dropout=ops.drop(0.25, 1234567, 78))
    assert not torch.equal(o1, o3)


def test_colsum(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(3)
    for M, N in [(1, 7), (1000, 64), (5000, 130)]:
        x = _rand(g, M, N)
        assert max_err_scaled(ops.colsum(x.to(dev)).cpu(), x.double().sum(0)) < 1e-5


def test_retrieval_dyadic_bit_exact_against_reference(dev):
    """Golden fixture from the real DARE.topk_select on a dyadic grid: values bit-equal, index order canonical."""
    from b200ctr import ops
    z = np.load(os.path.join(GOLDEN, "retrieval_dyadic.npz"))
    K = int(z["K"])
    seq = torch.from_numpy(z["seq"])
    rec = O.recency_table(seq.shape[1], float(z["tau"]))
    scores = torch.from_numpy(z["scores"])
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :K]
    for seq_dev in (seq.to(dev), seq.int().to(dev)):               # int64 (collate contract) and int32 (on-disk dtype)
        sel, vals, idx, tok = ops.retrieve_fwd(seq_dev, torch.from_numpy(z["q"]).to(dev), torch.from_numpy(z["emb_att"]).to(dev),
                                               torch.from_numpy(z["emb_rep"]).to(dev), rec.to(dev), K, 0)
        assert torch.equal(vals.cpu(), torch.from_numpy(z["vals"]))                 # bit-exact scores of the reference
        assert torch.equal(idx.cpu().long(), order)                                # canonical (score desc, pos asc)
        assert torch.equal(tok.cpu().long(), torch.gather(seq, 1, order))
        osel, ovals, oidx = O.retrieve_fwd(seq, torch.from_numpy(z["q"]), torch.from_numpy(z["emb_att"]),
                                           torch.from_numpy(z["emb_rep"]), rec, K)
        assert torch.equal(sel.cpu(), osel)                                        # gathered rows bit-identical
        assert torch.equal(sel.cpu().sum(1), torch.from_numpy(z["sel_rep"]).sum(1))  # same multiset as the reference


@pytest.mark.parametrize("L,K", [(1, 1), (7, 7), (40, 12), (64, 20), (100, 33), (400, 60), (600, 80), (400, 100), (400, 148), (600, 148)])
def test_retrieval_random_vs_oracle(dev, L, K):
    """Generic inputs: scores agree to fp32 rounding; selections agree except at K-boundary near-ties (counted)."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(L * 131 + K)
    B, V, D = 33, 500, 64
    att, rep = _rand(g, V, D), _rand(g, V, D)
    att[0] = 0; rep[0] = 0
    q = _rand(g, B, D)
    seq = torch.randint(1, V, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0] = 0; lens[1] = L
    for b in range(B):
        seq[b, :L - int(lens[b])] = 0
    rec = O.recency_table(L, 16.0)
    sel, vals, idx, tok = ops.retrieve_fwd(seq.to(dev), q.to(dev), att.to(dev), rep.to(dev), rec.to(dev), K, 0)
    osel, ovals, oidx = O.retrieve_fwd(seq, q, att, rep, rec, K)
    assert torch.allclose(vals.cpu(), ovals, rtol=1e-5, atol=1e-4)
    # descending, ties broken by ascending position
    v = vals.cpu(); ix = idx.cpu()
    assert bool((v[:, 1:] <= v[:, :-1]).all())
    tie = v[:, 1:] == v[:, :-1]
    assert bool((ix[:, 1:][tie] > ix[:, :-1][tie]).all())
    mism = (ix.long() != oidx)
    # any disagreement must sit on a near-tie (|score gap| < 4 ulp-ish)
    if mism.any():
        gap = (v - ovals).abs()[mism]
        assert float(gap.max()) < 1e-4
    assert mism.float().mean() < 0.01
    valid = tok.cpu() != 0
    assert torch.equal(valid, torch.gather(seq, 1, ix.long()) != 0)
    assert bool((sel.cpu()[~valid] == 0).all())
    assert bool((v[~valid] == -1e9).all())
    # number of valid slots = min(K, l)
    assert torch.equal(valid.sum(1), torch.minimum(lens, torch.tensor(K)))


def test_retrieve_bwd_vs_oracle(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(99)
    B, V, D, L, K = 17, 120, 64, 50, 14
    att, rep = _rand(g, V, D), _rand(g, V, D)
    att[0] = 0; rep[0] = 0
    q = _rand(g, B, D)
    seq = torch.randint(1, 12, (B, L), generator=g)             # heavy token duplication -> atomics accumulate
    seq[:, :20] = 0
    seq[0] = 0
    rec = O.recency_table(L, 8.0)
    sel, vals, idx, tok = ops.retrieve_fwd(seq.to(dev), q.to(dev), att.to(dev), rep.to(dev), rec.to(dev), K, 0)
    dsel, dvals = _rand(g, B, K, D), _rand(g, B, K)
    datt = torch.zeros(V, D, device=dev); drep = torch.zeros(V, D, device=dev)
    dq = ops.retrieve_bwd(tok, q.to(dev), att.to(dev), dsel.to(dev), dvals.to(dev), 0, datt, drep)
    odq, odatt, odrep = O.retrieve_bwd(seq, q, att, rep, idx.cpu().long(), dsel, dvals, 0)
    assert max_err_scaled(dq.cpu(), odq) < TOL
    assert max_err_scaled(datt.cpu(), odatt) < TOL
    assert max_err_scaled(drep.cpu(), odrep) < TOL
    assert float(datt[0].abs().max()) == 0 and float(drep[0].abs().max()) == 0


def test_bce_wll_vs_oracle(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(4)
    for B, mode in [(1, "pos"), (5, "neg"), (777, "mix"), (8192, "mix")]:
        z, a = _rand(g, B) * 3, _rand(g, B)
        y = {"pos": torch.ones(B), "neg": torch.zeros(B), "mix": (torch.rand(B, generator=g) < 0.1).float()}[mode]
        loss, dz, da, prob, counts = ops.bce_wll(z.to(dev), a.to(dev), y.to(dev), 0.1)
        l0, d0 = O.bce_wll(z.double(), y.double())
        l1, d1 = O.bce_wll(a.double(), y.double())
        assert abs(float(loss.cpu()) - float(l0 + 0.1 * l1)) < 1e-5
        assert max_err_scaled(dz.cpu(), d0) < 1e-5 and max_err_scaled(da.cpu(), 0.1 * d1) < 1e-5
        assert torch.allclose(prob.cpu(), torch.sigmoid(z), atol=1e-6)
        assert counts.cpu().tolist() == [int(y.sum()), B - int(y.sum())]


def test_adamw_ema_clip_vs_oracle(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(8)
    n = 10007
    p, gr = _rand(g, n), _rand(g, n)
    m, v, ema = torch.zeros(n), torch.zeros(n), p.clone()
    pd, gd, md, vd, ed = (t.clone().to(dev) for t in (p, gr, m, v, ema))
    acc = torch.zeros(1, device=dev, dtype=torch.float64)
    coef = torch.ones(1, device=dev); tn = torch.zeros(1, device=dev)
    for step in (1, 2, 3):
        acc.zero_()
        ops.sqnorm_(gd, acc)
        ops.clip_coef(acc, 0.5, coef, tn)
        c, tot = O.clip_coef({"g": gr}, 0.5)
        assert abs(float(tn.cpu()) - tot) < 1e-4 * tot and abs(float(coef.cpu()) - c) < 1e-6
        ops.adamw_ema_(pd, gd, md, vd, ed, 3e-3, 1e-2, step, coef, 0.9)
        O.adamw_step(p, gr * c, m, v, step, 3e-3, 1e-2)
        O.ema_update(ema, p, 0.9)
        assert max_err_scaled(pd.cpu(), p) < 1e-5 and max_err_scaled(ed.cpu(), ema) < 1e-5
        assert max_err_scaled(md.cpu(), m) < 1e-5 and max_err_scaled(vd.cpu(), v) < 1e-4   # v carries coef^2 (fp32 coef)


def test_logit_mean_vs_reference_formula(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(2)
    P = torch.rand(5, 1000, generator=g)
    P[0, 0] = 0.0; P[1, 1] = 1.0
    Pc = torch.clamp(P, 1e-7, 1 - 1e-7)                                   # fp32, exactly as utils/metrics.py:31-33
    ref = torch.sigmoid((torch.log(Pc) - torch.log1p(-Pc)).mean(0))      # utils/metrics.py:68-71
    assert torch.allclose(ops.logit_mean(P.to(dev)).cpu(), ref, atol=1e-5)


def _bf16_round(t):
    return t.to(torch.bfloat16).double()


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 70, 129), (1000, 96, 64), (257, 192, 100), (4096, 384, 64),
                                   (512, 512, 1024), (130, 1152, 96), (64, 200, 4096)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
def test_gemm_bf16_tensor_core_all_layouts(dev, M, N, K, ta, tb):
    """tcgen05 path: operands rounded to bf16, fp32 accumulation => equals the fp64 product of the bf16-rounded
    operands to fp32 accumulation accuracy (a layout / descriptor error would show up as O(1) differences)."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    a = _rand(g, K, M) if ta else _rand(g, M, K)
    b = _rand(g, N, K) if tb else _rand(g, K, N)
    ref = (_bf16_round(a).t() if ta else _bf16_round(a)) @ (_bf16_round(b).t() if tb else _bf16_round(b))
    with ops.precision("bf16"):
        out = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, tc=True)
        out_sk = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, split_k=3, tc=True)
    assert max_err_scaled(out.cpu(), ref) < 2e-5
    assert max_err_scaled(out_sk.cpu(), ref) < 2e-5
    # and stays within bf16 rounding of the exact fp32 product
    exact = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    assert max_err_scaled(out.cpu(), exact) < 2e-2


def test_gemm_bf16_epilogue_and_strided_views(dev):
    from b200ctr import ops
    g = torch.Generator().manual_seed(21)
    M, N, K = 700, 192, 128
    big = _rand(g, M, 3 * K).to(dev)
    a = big[:, K:2 * K]
    w, bias, res = _rand(g, N, K).to(dev), _rand(g, N).to(dev), _rand(g, M, N).to(dev)
    lin = _bf16_round(a.cpu()) @ _bf16_round(w.cpu()).t()
    with ops.precision("bf16"):
        pre = torch.empty(M, N, device=dev)
        out = ops.gemm(a, w, trans_b=True, bias=bias, act=ops.ACT_GELU, pre_out=pre, residual=res, tc=True)
        assert max_err_scaled(pre.cpu(), lin + bias.cpu().double()) < 2e-5
        assert max_err_scaled(out.cpu(), O.gelu_fwd(lin + bias.cpu().double()) + res.cpu().double()) < 2e-5
        d = ops.drop(0.3, 99, 5)
        o1 = ops.gemm(a, w, trans_b=True, dropout=d, tc=True)
    o2 = ops.gemm(a, w, trans_b=True, dropout=d)             # fp32 path: same mask (same element ids)
    assert torch.equal(o1 == 0, o2 == 0)


# ---- fused tcgen05 FFN block (csrc/ffn_fused.cu) vs the unfused bf16 composition and vs torch --------------------
def _ffn_inputs(T, hidden, seed):
    g = torch.Generator().manual_seed(seed)
    x1 = _rand(g, T, 64)
    w1 = _rand(g, hidden, 64) * 64 ** -0.5
    b1 = _rand(g, hidden) * 0.1
    w2 = _rand(g, 64, hidden) * hidden ** -0.5
    b2 = _rand(g, 64) * 0.1
    n2 = 1.0 + 0.1 * _rand(g, 64)
    return x1, w1, b1, w2, b2, n2


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


# hidden = 192 / 320 / 512 run the 64-wide-chunk instantiation of the forward kernel, the multiples of 128 up to 384 the 128-wide one
@pytest.mark.parametrize("T,hidden,p", [(128, 128, 0.0), (1000, 384, 0.0), (1000, 384, 0.1), (128 * 300 + 17, 384, 0.1), (5, 256, 0.0),
                                        (700, 192, 0.1), (300, 320, 0.0), (300, 512, 0.1)])
def test_ffn_fwd_fused_matches_unfused(dev, T, hidden, p):
    from b200ctr import ops
    x1, w1, b1, w2, b2, n2 = (t.to(dev) for t in _ffn_inputs(T, hidden, 3 + T))
    dr = ops.drop(p, 1234, 77)
    x2, s2, r2 = ops.ffn_fwd(x1, w1, b1, w2, b2, n2, dr)
    with ops.precision("bf16"):
        fd = ops.gemm(x1, w1, trans_b=True, bias=b1, act=ops.ACT_GELU, dropout=dr, tc=True)
        h2 = ops.gemm(fd, w2, trans_b=True, bias=b2, tc=True)
    s2u, x2u, r2u = ops.add_rmsnorm_fwd(x1, h2, n2)
    assert max_err_scaled(s2.cpu(), s2u.cpu()) < 2e-3
    assert max_err_scaled(x2.cpu(), x2u.cpu()) < 2e-3
    assert max_err_scaled(r2.cpu(), r2u.cpu()) < 2e-3
    if p == 0.0:
        pre = _bf(x1) @ _bf(w1).t() + b1
        fdr = torch.nn.functional.gelu(pre)
        s2r = x1 + _bf(fdr) @ _bf(w2).t() + b2
        rr = torch.rsqrt((s2r * s2r).mean(1) + 1e-6)
        assert max_err_scaled(s2.cpu(), s2r.cpu()) < 2e-3
        assert max_err_scaled(x2.cpu(), (n2 * s2r * rr[:, None]).cpu()) < 2e-3


@pytest.mark.parametrize("T,hidden,p", [(128, 128, 0.0), (1000, 384, 0.0), (1000, 384, 0.1), (128 * 300 + 17, 384, 0.1), (5, 256, 0.0)])
def test_ffn_bwd_fused_matches_unfused(dev, T, hidden, p):
    from b200ctr import ops
    x1, w1, b1, w2, b2, n2 = (t.to(dev) for t in _ffn_inputs(T, hidden, 9 + T))
    dy = torch.randn(T, 64, generator=torch.Generator().manual_seed(T)).to(dev)
    dr = ops.drop(p, 4321, 55)
    x2, s2, r2, fkeep = ops.ffn_fwd(x1, w1, b1, w2, b2, n2, dr, save_keep=True)
    assert (fkeep is not None) == (p > 0)
    z = lambda *s: torch.zeros(*s, device=dev)
    dw1, db1, dw2, db2, dn2 = z(hidden, 64), z(hidden), z(64, hidden), z(64), z(64)
    dx1 = ops.ffn_bwd(dy, s2, r2, n2, x1, w1, b1, w2, dw1, db1, dw2, db2, dn2, dr)
    if fkeep is not None:      # the stored keep flags reproduce the re-hashed mask: bit-identical gradients
        k1, kb1, k2, kb2, kn2 = z(hidden, 64), z(hidden), z(64, hidden), z(64), z(64)
        dx1k = ops.ffn_bwd(dy, s2, r2, n2, x1, w1, b1, w2, k1, kb1, k2, kb2, kn2, dr, keep=fkeep)
        for a, b in ((dx1k, dx1), (k1, dw1), (kb1, db1), (k2, dw2), (kb2, db2), (kn2, dn2)):
            assert torch.equal(a, b)
    # unfused composition (hotpath.encoder_layer_bwd before fusion), same dropout stream
    with ops.precision("bf16"):
        pre = torch.empty(T, hidden, device=dev)
        fd = ops.gemm(x1, w1, trans_b=True, bias=b1, act=ops.ACT_GELU, pre_out=pre, dropout=dr, tc=True)
        dn2u = z(64)
        d_s2 = ops.rmsnorm_bwd(dy, s2, n2, r2, dn2u)
        dw2u = ops.gemm(d_s2, fd, trans_a=True, tc=True)
        db2u = ops.colsum(d_s2)
        dpre = ops.gemm(d_s2, w2, dropout=dr, mask_mode=ops.MASK_GELU_GRAD, mask_src=pre, tc=True)
        dw1u = ops.gemm(dpre, x1, trans_a=True, tc=True)
        db1u = ops.colsum(dpre)
        dx1u = ops.gemm(dpre, w1, residual=d_s2, tc=True)
    # tolerance: one bf16 ulp (2^-7) of the largest element.  The fused kernel computes GELU / GELU' on half2 pairs and rounds the
    # activation fp16 -> bf16 where the unfused composition rounds fp32 -> bf16, so single operand elements land on neighbouring
    # bf16 values; with T = 5 tokens nothing averages (measured 5.3e-3 there, <= 3e-3 at T >= 128).
    for name, a, b in (("dx1", dx1, dx1u), ("dw1", dw1, dw1u), ("db1", db1, db1u), ("dw2", dw2, dw2u), ("db2", db2, db2u),
                       ("dn2", dn2, dn2u)):
        assert max_err_scaled(a.cpu(), b.cpu()) < (7.8e-3 if T < 128 else 5e-3), name


# ---- tensor-core attention (packed-half csrc/attn_h16.cu = the default, TF32 csrc/attn_tc.cu) vs the exact fp32 CUDA-core
# kernels (same dropout masks).  do_scale: gradients of a mean-reduced loss are ~1e-7 -- far below fp16's range; the packed-half
# backward rescales dO per sample by a power of two, and rows of very different size share that scale.
@pytest.mark.parametrize("impl,do_scale", [("h16", 1.0), ("h16", 3e-8), ("tf32", 1.0)])
@pytest.mark.parametrize("B,K,p", [(3, 60, 0.0), (5, 60, 0.1), (2, 100, 0.1), (2, 148, 0.1), (4, 5, 0.0), (3, 16, 0.1), (2, 80, 0.0), (1, 1, 0.0),
                                   (2, 33, 0.2), (2, 112, 0.1), (2, 160, 0.0)])
def test_attn_tc_matches_fp32_kernels(dev, B, K, p, impl, do_scale):
    from b200ctr import ops
    g = torch.Generator().manual_seed(100 * B + K)
    qkv = (_rand(g, B * K, 192) * 0.8).to(dev)
    relmean = (_rand(g, 2 * K + 1) * 0.5).to(dev)
    d_o = _rand(g, B * K, 64) * do_scale
    if do_scale != 1.0:
        d_o = d_o * torch.logspace(0, -3, B * K)[torch.randperm(B * K, generator=g)][:, None]   # token rows spanning 3 decades
    d_o = d_o.to(dev)
    dr = ops.drop(p, 99, 3)
    o_ref, lse_ref = ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=False)
    o, lse, keep = ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl=impl)
    assert (keep is not None) == (p > 0)
    assert max_err_scaled(o.cpu(), o_ref.cpu()) < 3e-3
    assert max_err_scaled(lse.cpu(), lse_ref.cpu()) < 3e-3
    drel_ref = torch.zeros_like(relmean)
    drel = torch.zeros_like(relmean)
    dq_ref = ops.attn_bwd(qkv, relmean, K, o_ref, lse_ref, d_o, B, K, 8, 64, drel_ref, dr, tc=False)
    dq = ops.attn_bwd(qkv, relmean, K, o_ref, lse_ref, d_o, B, K, 8, 64, drel, dr, tc=True, impl=impl)
    # dq: TF32 operands; dk / dv: bf16 P^T / dS^T operands (the amp: bf16 matmul policy).  Scale by the largest gradient
    # of the call so that exactly-zero reference blocks (K = 1: softmax over one key) compare in absolute terms.
    gmax = float(dq_ref.abs().max()) + 1e-30
    for sl, name, tol in ((slice(0, 64), "dq", 5e-3), (slice(64, 128), "dk", 1.5e-2), (slice(128, 192), "dv", 1.5e-2)):
        assert float((dq[:, sl] - dq_ref[:, sl]).abs().max()) / gmax < tol, name
    assert float((drel - drel_ref).abs().max()) / (float(drel_ref.abs().max()) + gmax) < 5e-3
    if keep is not None:
        # the stored keep flags must reproduce the re-hashed mask exactly: same kernel arithmetic, same result
        drel_k = torch.zeros_like(relmean)
        dq_k = ops.attn_bwd(qkv, relmean, K, o_ref, lse_ref, d_o, B, K, 8, 64, drel_k, dr, tc=True, keep=keep, impl=impl)
        assert torch.equal(dq_k, dq)                                                  # no atomics anywhere: bit-identical
        assert float((drel_k - drel).abs().max()) / (float(drel.abs().max()) + gmax) < 1e-5
    # bf16 output for the fused in-projection backward: the same values, rounded once
    dq16 = ops.attn_bwd(qkv, relmean, K, o_ref, lse_ref, d_o, B, K, 8, 64, torch.zeros_like(relmean), dr, tc=True, keep=keep, out_bf16=True,
                        impl=impl)
    assert dq16.dtype == torch.bfloat16 and torch.equal(dq16, dq.to(torch.bfloat16))


# fp16 qkv hand-off: the in-projection writes fp16 with q pre-scaled, the packed-half attention stages it with copies.  Both
# halves of the hand-off are exact restatements of what the fp32 route computes, so every result must be BIT-identical.
H2_QSCALE = np.float32(0.51006972789764404297)


@pytest.mark.parametrize("B,K,p", [(3, 60, 0.1), (2, 100, 0.0), (2, 148, 0.1), (5, 7, 0.1), (1, 1, 0.0), (2, 33, 0.2)])
def test_attn_h16_fp16_qkv_is_bit_identical_to_the_fp32_route(dev, B, K, p):
    from b200ctr import ops
    g = torch.Generator().manual_seed(7 * B + K)
    x = _rand(g, B * K, 64).to(dev)
    w_in, b_in = (_rand(g, 192, 64) * 0.2).to(dev), (_rand(g, 192) * 0.1).to(dev)
    relmean = (_rand(g, 2 * K + 1) * 0.5).to(dev)
    d_o = (_rand(g, B * K, 64) * 1e-5).to(dev)
    with ops.precision("bf16"):
        qkv32 = ops.inproj_fwd(x, w_in, b_in)
        qkv16 = ops.inproj_fwd(x, w_in, b_in, h16=True)
    assert qkv16.dtype == torch.float16 and qkv16.shape == qkv32.shape
    want = torch.cat([(qkv32[:, :64] * float(H2_QSCALE)).half(), qkv32[:, 64:].half()], dim=1)
    assert torch.equal(qkv16, want)
    dr = ops.drop(p, 11, 5)
    o32, lse32, keep32 = ops.attn_fwd(qkv32, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl="h16")
    o16, lse16, keep16 = ops.attn_fwd(qkv16, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl="h16")
    assert torch.equal(o16, o32) and torch.equal(lse16, lse32)
    assert (keep16 is None and keep32 is None) or torch.equal(keep16, keep32)
    r32, r16 = torch.zeros_like(relmean), torch.zeros_like(relmean)
    dq32 = ops.attn_bwd(qkv32, relmean, K, o32, lse32, d_o, B, K, 8, 64, r32, dr, tc=True, keep=keep32, out_bf16=True, impl="h16")
    dq16 = ops.attn_bwd(qkv16, relmean, K, o32, lse32, d_o, B, K, 8, 64, r16, dr, tc=True, keep=keep16, out_bf16=True, impl="h16")
    assert torch.equal(dq16, dq32) and torch.equal(r16, r32)
    with pytest.raises(Exception):
        ops.attn_fwd(qkv16, relmean, K, B, K, 8, 64, dr, tc=True, impl="tf32")        # the TF32 kernels take the fp32 tensor only


# bf16 o / d_o hand-offs between the packed-half attention and the out-projection: every kernel with a bf16 port must give,
# bit for bit, what its fp32 port gives when fed the bf16-rounded values (inputs), resp. its fp32 output rounded once (outputs).
@pytest.mark.parametrize("B,K,p", [(3, 60, 0.1), (2, 100, 0.0), (3, 17, 0.1)])
def test_bf16_o_and_do_handoffs_are_exact_restatements(dev, B, K, p):
    from b200ctr import ops
    g = torch.Generator().manual_seed(3 * B + K)
    T = B * K
    qkv = (_rand(g, T, 192) * 0.8).to(dev)
    relmean = (_rand(g, 2 * K + 1) * 0.5).to(dev)
    dr = ops.drop(p, 21, 9)
    bf = lambda t: t.to(torch.bfloat16)
    # attention forward: bf16 o == fp32 o rounded once
    o32, lse, keep = ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl="h16")
    o16, lse2, keep2 = ops.attn_fwd(qkv, relmean, K, B, K, 8, 64, dr, tc=True, save_keep=True, impl="h16", out_bf16=True)
    assert o16.dtype == torch.bfloat16 and torch.equal(o16, bf(o32)) and torch.equal(lse, lse2)
    # attention backward: bf16 o / d_o ports == fp32 ports fed the rounded values
    d_o = (_rand(g, T, 64) * 1e-4).to(dev)
    ra, rb = torch.zeros_like(relmean), torch.zeros_like(relmean)
    dq_a = ops.attn_bwd(qkv, relmean, K, o16.float(), lse, bf(d_o).float(), B, K, 8, 64, ra, dr, tc=True, keep=keep, out_bf16=True, impl="h16")
    dq_b = ops.attn_bwd(qkv, relmean, K, o16, lse, bf(d_o), B, K, 8, 64, rb, dr, tc=True, keep=keep, out_bf16=True, impl="h16")
    assert torch.equal(dq_a, dq_b) and torch.equal(ra, rb)
    # out-projection forward / backward
    w_out, b_out, n1 = (_rand(g, 64, 64) * 0.2).to(dev), (_rand(g, 64) * 0.1).to(dev), (1 + 0.1 * _rand(g, 64)).to(dev)
    x = _rand(g, T, 64).to(dev)
    with ops.precision("bf16"):
        s1a, x1a, r1a = ops.outproj_norm_fwd(o16.float(), w_out, b_out, x, n1)
        s1b, x1b, r1b = ops.outproj_norm_fwd(o16, w_out, b_out, x, n1)
        assert torch.equal(s1a, s1b) and torch.equal(x1a, x1b) and torch.equal(r1a, r1b)
        dx1 = (_rand(g, T, 64) * 1e-3).to(dev)
        z = lambda *sh: torch.zeros(*sh, device=dev)
        ga, gb = (z(64, 64), z(64), z(64)), (z(64, 64), z(64), z(64))
        ds_a, do_a = ops.outproj_norm_bwd(dx1, s1a, r1a, n1, o16.float(), w_out, *ga)
        ds_b, do_b = ops.outproj_norm_bwd(dx1, s1a, r1a, n1, o16, w_out, *gb, out_bf16=True)
        assert do_b.dtype == torch.bfloat16 and torch.equal(do_b, bf(do_a)) and torch.equal(ds_a, ds_b)
        for a_, b_ in zip(ga, gb):
            assert torch.equal(a_, b_)


@pytest.mark.parametrize("B,H,R,P", [(1000, 6, 16, 192), (7, 6, 16, 192), (300, 8, 32, 128), (64, 2, 4, 24)])
def test_headproj_matches_einsum(dev, B, H, R, P):
    from b200ctr import ops
    assert ops.headproj_ok(H, R, P)
    g = torch.Generator().manual_seed(B + H)
    quad, V, dinter = _rand(g, B, H * R), _rand(g, H, R, P) * 0.1, _rand(g, B, H * P)
    inter = ops.headproj_fwd(quad.to(dev), V.to(dev))
    ref = torch.einsum("bhr,hrp->bhp", quad.view(B, H, R).double(), V.double()).reshape(B, H * P)
    assert max_err_scaled(inter.cpu(), ref) < 1e-5
    dV = torch.zeros(H, R, P, device=dev)
    dquad = ops.headproj_bwd(quad.to(dev), dinter.to(dev), V.to(dev), dV)
    assert max_err_scaled(dquad.cpu(), torch.einsum("bhp,hrp->bhr", dinter.view(B, H, P).double(), V.double()).reshape(B, H * R)) < 1e-5
    assert max_err_scaled(dV.cpu(), torch.einsum("bhr,bhp->hrp", quad.view(B, H, R).double(), dinter.view(B, H, P).double())) < 1e-5


@pytest.mark.parametrize("M,N,K,ta,tb", [(256, 512, 1280, False, True), (512, 1280, 256, True, False), (300, 1288, 512, False, False),
                                         (128, 64, 64, False, True)])
def test_gemm_bf16_operands(dev, M, N, K, ta, tb):
    """Operands already in bf16 (pure-copy staging) == the same contraction with fp32 operands rounded on load."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = _rand(g, K, M) if ta else _rand(g, M, K)
    b = _rand(g, N, K) if tb else _rand(g, K, N)
    ad, bd = a.to(dev), b.to(dev)
    ref = (_bf(a).t() if ta else _bf(a)).double() @ (_bf(b).t() if tb else _bf(b)).double()
    out = ops.gemm(ops.to_bf16(ad), ops.to_bf16(bd), trans_a=ta, trans_b=tb)
    assert max_err_scaled(out.cpu(), ref) < 1e-5
    # strided bf16 views (the W0[:, :N] / W0[:, N:] split of the first MLP layer)
    if not ta and tb:
        big = ops.to_bf16(torch.cat([bd, bd], dim=1).contiguous())
        out2 = ops.gemm(ops.to_bf16(ad), big[:, K:], trans_b=True)
        assert max_err_scaled(out2.cpu(), ref) < 1e-5


# ---- fused attention-side projections (csrc/proj_fused.cu) vs their unfused bf16 compositions --------------------------
@pytest.mark.parametrize("T", [128, 1000, 128 * 300 + 17, 5])
def test_fused_projections_match_unfused(dev, T):
    from b200ctr import ops
    g = torch.Generator().manual_seed(77 + T)
    x, o = _rand(g, T, 64).to(dev), _rand(g, T, 64).to(dev)
    w_in, b_in = (_rand(g, 192, 64) * 0.125).to(dev), (_rand(g, 192) * 0.1).to(dev)
    w_out, b_out = (_rand(g, 64, 64) * 0.125).to(dev), (_rand(g, 64) * 0.1).to(dev)
    n1 = (1.0 + 0.1 * _rand(g, 64)).to(dev)
    dqkv, dx1 = _rand(g, T, 192).to(dev), _rand(g, T, 64).to(dev)
    z = lambda *s: torch.zeros(*s, device=dev)
    with ops.precision("bf16"):
        qkv_u = ops.gemm(x, w_in, trans_b=True, bias=b_in, tc=True)
        h_u = ops.gemm(o, w_out, trans_b=True, bias=b_out, tc=True)
        s1_u, x1_u, r1_u = ops.add_rmsnorm_fwd(x, h_u, n1)
        dn1_u = z(64)
        ds1_u = ops.rmsnorm_bwd(dx1, s1_u, n1, r1_u, dn1_u)
        dwo_u = ops.gemm(ds1_u, o, trans_a=True, tc=True)
        dbo_u = ops.colsum(ds1_u)
        do_u = ops.gemm(ds1_u, w_out, tc=True)
        dwi_u = ops.gemm(dqkv, x, trans_a=True, tc=True)
        dbi_u = ops.colsum(dqkv)
        dx_u = ops.gemm(dqkv, w_in, residual=ds1_u, tc=True)
    qkv = ops.inproj_fwd(x, w_in, b_in)
    s1, x1, r1 = ops.outproj_norm_fwd(o, w_out, b_out, x, n1)
    dwo, dbo, dn1, dwi, dbi = z(64, 64), z(64), z(64), z(192, 64), z(192)
    ds1, d_o = ops.outproj_norm_bwd(dx1, s1_u, r1_u, n1, o, w_out, dwo, dbo, dn1)
    dx = ops.inproj_bwd(dqkv, x, w_in, ds1_u, dwi, dbi)
    dwi16, dbi16 = torch.empty_like(dwi), torch.empty_like(dbi)
    dx16 = ops.inproj_bwd(dqkv.to(torch.bfloat16), x, w_in, ds1_u, dwi16, dbi16)      # bf16 dqkv: the kernel rounds fp32 the same way
    assert torch.equal(dx16, dx) and torch.equal(dwi16, dwi) and torch.equal(dbi16, dbi)
    for name, a, b, tol in (("qkv", qkv, qkv_u, 2e-3), ("s1", s1, s1_u, 2e-3), ("x1", x1, x1_u, 2e-3), ("r1", r1, r1_u, 2e-3),
                            ("ds1", ds1, ds1_u, 1e-4), ("d_o", d_o, do_u, 3e-3), ("dw_out", dwo, dwo_u, 5e-3),
                            ("db_out", dbo, dbo_u, 5e-3), ("dn1", dn1, dn1_u, 1e-3), ("dx", dx, dx_u, 3e-3),
                            ("dw_in", dwi, dwi_u, 5e-3), ("db_in", dbi, dbi_u, 5e-3)):
        assert max_err_scaled(a.cpu(), b.cpu()) < tol, name


@pytest.mark.parametrize("B,F,C", [(300, 200, 96), (7, 200, 96), (200, 128, 64), (150, 256, 128), (33, 37, 32)])
def test_fm_fused_matches_unfused(dev, B, F, C):
    """Fused tcgen05 quadratic interaction == gemm + fm_reduce / fm_grad + gemms on the same bf16-rounded operands."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(B + F + C)
    z = (_rand(g, B, F * 64) * 0.5).to(dev)
    U = (_rand(g, 64, C) * 0.2).to(dev)
    dquad = _rand(g, B, C).to(dev)
    dzm = _rand(g, B, F * 64).to(dev)
    z16 = ops.to_bf16(z)
    zb = z16.float()
    s, quad = ops.fm_fused_fwd(z16, U, B, F)
    A = (zb.view(B * F, 64).double() @ _bf(U).double()).view(B, F, C)
    s_ref = A.sum(1)
    quad_ref = s_ref * s_ref - (A * A).sum(1)
    assert max_err_scaled(s.cpu(), s_ref.cpu()) < 1e-4
    assert max_err_scaled(quad.cpu(), quad_ref.cpu()) < 1e-4
    dz, dU = ops.fm_fused_bwd(z16, U, s, dquad, dzm, B, F)
    dA = (dquad.double().unsqueeze(1) * 2.0 * (s.double().unsqueeze(1) - A)).float()          # (B,F,C)
    dAb = _bf(dA).double().view(B * F, C)
    dz_ref = dAb @ _bf(U).double().t() + dzm.double().view(B * F, 64)
    dU_ref = zb.view(B * F, 64).double().t() @ dAb
    # dA is rounded to bf16 as a tensor-core operand: fp32-vs-fp64 noise in (s - A) occasionally flips one rounding
    # (1 bf16 ulp of a |dA| ~ 100 element), hence 3e-3 rather than accumulation-order tolerance
    assert max_err_scaled(dz.view(B * F, 64).cpu(), dz_ref.cpu()) < 3e-3
    assert max_err_scaled(dU.cpu(), dU_ref.cpu()) < 2e-3


@pytest.mark.parametrize("M,N,K", [(256, 512, 1280), (8192, 512, 1024), (512, 1280, 4096), (300, 96, 264), (128, 32, 256)])
def test_gemm_tma_kmajor(dev, M, N, K):
    """TMA-fed tcgen05 contraction (both operands bf16, K contiguous) incl. ragged M / N / K tiles and strided views."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(M * 3 + N + K)
    a, b = _rand(g, M, K).to(dev), _rand(g, N, K + 64).to(dev)
    a16, b16 = ops.to_bf16(a), ops.to_bf16(b)[:, 32:32 + K]            # row-strided bf16 view (like W0[:, :N])
    ref = _bf(a).double() @ _bf(b)[:, 32:32 + K].double().t()
    out = ops.gemm_kk(a16, b16)
    assert max_err_scaled(out.cpu(), ref.cpu()) < 1e-5
    # the transposing staging helpers
    t1 = ops.to_bf16_t(b[:, 32:32 + K])
    assert torch.equal(t1.float().cpu(), _bf(b)[:, 32:32 + K].t().contiguous().cpu())
    if M % 2 == 0 and K % 2 == 0:
        t2 = ops.transpose_bf16(a16)
        assert torch.equal(t2.cpu(), a16.t().contiguous().cpu())


@pytest.mark.parametrize("R,C", [(200, 136), (64, 64), (8, 8), (1000, 72), (130, 66), (8192, 1280)])
def test_transpose_bf16_both_kernels(dev, R, C):
    """bf16 (R, C) -> (C, R): the 16-byte kernel (R, C multiples of 8; ragged 64 x 64 tiles) and the 4-byte fallback."""
    from b200ctr import ops
    g = torch.Generator().manual_seed(R * 7 + C)
    a16 = ops.to_bf16(_rand(g, R, C).to(dev))
    assert torch.equal(ops.transpose_bf16(a16).cpu(), a16.t().contiguous().cpu())


# ---- Linear(K -> 1) as row dot products (csrc/qnn.cu lin1_*) vs torch ---------------------------------------------
@pytest.mark.parametrize("B,K", [(1000, 256), (7, 128), (4096, 512)])
def test_lin1_matches_torch(dev, B, K):
    from b200ctr import ops
    g = torch.Generator().manual_seed(B + K)
    h = torch.relu(_rand(g, B, K))                       # ReLU output: exact zeros carry the mask
    w, bias, dl = _rand(g, 1, K) * 0.1, _rand(g, 1), _rand(g, B, 1)
    out = ops.lin1_fwd(h.to(dev), w.to(dev), bias.to(dev))
    ref = h.double() @ w.double().t() + bias.double()
    assert max_err_scaled(out.cpu(), ref) < 1e-5
    dw, db = torch.empty(1, K, device=dev), torch.empty(1, device=dev)
    dh = ops.lin1_bwd(dl.to(dev), h.to(dev), w.to(dev), 1.25, dw, db)
    dh_ref = (dl.double() @ w.double()) * (h != 0).double() * 1.25
    assert max_err_scaled(dh.cpu(), dh_ref) < 1e-5
    assert max_err_scaled(dw.cpu(), dl.double().t() @ h.double()) < 1e-5
    assert abs(float(db.cpu()) - float(dl.double().sum())) < 1e-4 * (1 + float(dl.abs().sum()))


@pytest.mark.parametrize("mode,tol", [("x3", 3e-5), ("x6", 3e-5)])
@pytest.mark.parametrize("M,N,K,ta,tb", [(300, 192, 64, False, True), (1000, 96, 64, False, False), (257, 70, 136, True, False),
                                          (512, 384, 2048, False, True), (130, 64, 5000, True, False), (4096, 512, 1088, False, True),
                                          (2000, 64, 384, False, True), (384, 64, 9000, True, False)])
def test_gemm_split_precision_is_fp32_grade(M, N, K, ta, tb, mode, tol, monkeypatch):
    """b200ctr_gemm_split (the `amp: none` tensor-core contraction: bf16 parts, 3 or 6 MMAs per product) against fp64 math:
    error <= 3e-5 of max|C| for both (measured: 3 parts 2-8e-6 at K <= 2048 and 1-2e-5 at K = 5000-9000 per accumulator -- what
    is left is the tensor core's own fp32 accumulation, which truncates; the CUDA-core FMA chain sits at 0.3-1.4e-6) -- the plain bf16 kernel sits at ~3e-3 -- and the fused epilogue matches
    the CUDA-core kernel's."""
    from b200ctr import ops
    monkeypatch.setattr(ops, "FP32_GEMM", mode)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    with ops.precision("fp32"):
        ops.profile_start()
        out = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb)
        fams = {t.split("[")[0] for t in ops.profile_stop()}
        assert fams == {"gemm_" + mode}, fams
        bias = torch.randn(N, generator=g).to(dev)
        res = torch.randn(M, N, generator=g).to(dev)
        o_sp = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, bias=bias, act=ops.ACT_RELU, residual=res)
        o_fma = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, bias=bias, act=ops.ACT_RELU, residual=res, tc=False)
        o_sk = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, split_k=3)
        fma = ops.gemm(a.to(dev), b.to(dev), trans_a=ta, trans_b=tb, tc=False)
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    err_fma = float((fma.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < max(tol, 2 * err_fma), (err, err_fma)
    assert float((o_sk.double().cpu() - ref).abs().max() / ref.abs().max()) < max(tol, 2 * err_fma)
    assert float((o_sp - o_fma).abs().max() / o_fma.abs().max()) < max(tol, 2 * err_fma) * 2
