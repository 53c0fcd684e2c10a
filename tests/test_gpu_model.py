"""GPU parity of the whole hot path (drop-in modules and the fused trainer) against the reference-generated golden
fixtures and the oracle (run with -m gpu)."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import CASES, batch_of, group, load_case, max_err_scaled, spec_of

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _build(cfg, z, dev):
    import b200ctr
    g = cfg["golden"]
    cat_cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cat_cols}
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cat_cols)
    m.load_state_dict(group(z, "init/"), strict=True)
    if hasattr(m, "fc"):
        m.fc[2].p = 0.0          # same override oracle/make_golden.py applies to the reference (hard-coded Dropout(0.1))
    return m.to(dev)


@pytest.mark.parametrize("kind", CASES)
def test_module_forward_backward_matches_reference(kind):
    """The drop-in CTRModel on the reference's own inputs: logits/prob/aux and every autograd gradient."""
    from b200ctr import ops
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    m = _build(cfg, z, dev).train()
    batch, y = batch_of(z, 0)
    logits, prob, aux = m(batch)
    assert max_err_scaled(logits.detach().cpu(), torch.from_numpy(z["step0/logits"])) < TOL
    assert max_err_scaled(prob.cpu(), torch.from_numpy(z["step0/prob"])) < TOL
    assert max_err_scaled(aux.detach().cpu(), torch.from_numpy(z["step0/aux"])) < TOL
    loss, dz, da, _, _ = ops.bce_wll(logits.detach(), aux.detach(), y.to(dev), m.aux_weight)
    assert abs(float(loss.cpu()) - float(z["step0/loss"])) < 1e-5
    outs, grads = [logits], [dz]
    if m.aux_weight > 0:
        outs.append(aux); grads.append(da)
    torch.autograd.backward(outs, grads)
    ref = group(z, "grad0/")
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ref), set(got) ^ set(ref)
    for n, g in ref.items():
        assert max_err_scaled(got[n].cpu(), g) < TOL, (n, max_err_scaled(got[n].cpu(), g))


@pytest.mark.parametrize("kind", CASES)
def test_unmodified_step_body_with_torch_optimizer(kind):
    """Tier A: the reference's step body (torch AdamW + clip_grad_norm_) driving the drop-in module reproduces the
    reference's parameters after 3 steps."""
    from b200ctr import ops
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    m = _build(cfg, z, dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
    for s in range(3):
        batch, y = batch_of(z, s)
        opt.param_groups[0]["lr"] = float(z[f"step{s}/lr"])
        opt.zero_grad(set_to_none=True)
        logits, prob, aux = m(batch)
        loss, dzz, da, _, _ = ops.bce_wll(logits.detach(), aux.detach(), y.to(dev), m.aux_weight)
        outs, grads = [logits], [dzz]
        if m.aux_weight > 0:
            outs.append(aux); grads.append(da)
        torch.autograd.backward(outs, grads)
        if cfg["train"]["grad_clip_norm"] > 0:
            tn = torch.nn.utils.clip_grad_norm_(m.parameters(), cfg["train"]["grad_clip_norm"])
            assert abs(float(tn) - float(z[f"step{s}/total_norm"])) < 2e-4 * float(tn)
        opt.step()
        assert abs(float(loss.cpu()) - float(z[f"step{s}/loss"])) < 2e-4
    for k, v in group(z, "final/").items():
        assert max_err_scaled(m.state_dict()[k].cpu(), v) < 1e-3, k


@pytest.mark.parametrize("kind", CASES)
def test_fused_trainer_matches_reference(kind):
    """Tier B: FusedTrainer (flat buffers, fused clip/AdamW/EMA, row-sparse table grads) == reference after 3 steps."""
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    m = _build(cfg, z, dev).train()
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
    for s in range(3):
        batch, y = batch_of(z, s)
        assert abs(tr.lr_at(s + 1) - float(z[f"step{s}/lr"])) < 1e-15
        loss = tr.step(batch, y)
        assert abs(float(loss.cpu()) - float(z[f"step{s}/loss"])) < 2e-4
        if cfg["train"]["grad_clip_norm"] > 0:
            assert abs(float(tr.total_norm.cpu()) - float(z[f"step{s}/total_norm"])) < 2e-4 * float(z[f"step{s}/total_norm"])
    sd = m.state_dict()
    for k, v in group(z, "final/").items():
        assert max_err_scaled(sd[k].cpu(), v) < 1e-3, k
    ema = tr.ema_state()
    for k, v in group(z, "ema/").items():
        assert max_err_scaled(ema[k].cpu(), v) < 1e-3, k
    # the persistent dense gradient tables are back to all-zero
    for n in ("dare.emb_att.weight", "dare.emb_rep.weight"):
        assert float(tr.tables[n]["g"].abs().max()) == 0.0


def test_submodules_standalone_vs_oracle():
    """DARE / DAREEncoderLayer / QNNAlphaDetailed / SEBlock / feature embeddings used on their own (as in the
    reference's module API), checked against the oracle including input gradients."""
    import b200ctr
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    B, K, D = 9, 11, 64
    # encoder layer
    layer = b200ctr.DAREEncoderLayer(D, 8, 0.0, 96, 0.0, "rms", True, max_len=K).to(dev)
    x = torch.randn(B, K, D, generator=g)
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    dy = torch.randn(B, K, D, generator=g)
    y.backward(dy.to(dev))
    sd = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
    p = dict(in_w=sd["mha.in_proj_weight"], in_b=sd["mha.in_proj_bias"], out_w=sd["mha.out_proj.weight"],
             out_b=sd["mha.out_proj.bias"], n1=sd["norm1.w"], f1_w=sd["ffn.0.weight"], f1_b=sd["ffn.0.bias"],
             f2_w=sd["ffn.3.weight"], f2_b=sd["ffn.3.bias"], n2=sd["norm2.w"], rel=sd["pbias.rel.weight"])
    oy, c = O.encoder_layer_fwd(x, p, 8, K)
    odx, og = O.encoder_layer_bwd(dy, c, p, 8)
    assert max_err_scaled(y.detach().cpu(), oy) < TOL
    assert max_err_scaled(xd.grad.cpu(), odx) < TOL
    assert max_err_scaled(layer.pbias.rel.weight.grad.cpu(), og["rel"]) < TOL
    assert max_err_scaled(layer.mha.in_proj_weight.grad.cpu(), og["in_w"]) < TOL
    # numeric embedding
    ne = b200ctr.NumericFeatureEmbedding(13, 12, D).to(dev)
    with torch.no_grad():
        ne.bias.add_(0.1 * torch.randn(13, 12, generator=g).to(dev))
    xn = torch.randn(B, 13, generator=g)
    out = ne(xn)
    do = torch.randn(B, 13, D, generator=g)
    out.backward(do.to(dev))
    w, b, wo = ne.weight.detach().cpu(), ne.bias.detach().cpu(), ne.out_proj.weight.detach().cpu()
    assert max_err_scaled(out.detach().cpu(), O.numbin_fwd(xn, w, b, wo)) < TOL
    dw, db, dwo = O.numbin_bwd(xn, w, b, wo, do)
    assert max_err_scaled(ne.weight.grad.cpu(), dw) < TOL and max_err_scaled(ne.bias.grad.cpu(), db) < TOL
    assert max_err_scaled(ne.out_proj.weight.grad.cpu(), dwo) < TOL
    # SE block
    se = b200ctr.SEBlock(72, 8).to(dev)
    xs = torch.randn(B, 72, generator=g)
    xsd = xs.to(dev).requires_grad_(True)
    ys = se(xsd)
    ys.backward(torch.ones_like(ys))
    xr = xs.clone().requires_grad_(True)
    w1, b1, w2, b2 = (t.detach().cpu() for t in (se.fc[0].weight, se.fc[0].bias, se.fc[2].weight, se.fc[2].bias))
    gate = torch.sigmoid(torch.relu(xr.mean(0, keepdim=True) @ w1.t() + b1) @ w2.t() + b2)
    (xr * gate).sum().backward()
    assert max_err_scaled(ys.detach().cpu(), (xr * gate).detach()) < TOL
    assert max_err_scaled(xsd.grad.cpu(), xr.grad) < TOL


def test_state_dict_contract_and_ema_interop():
    """state_dict keys/shapes/dtypes follow SURVEY App. B; load_state_dict(strict=True) round-trips."""
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    m = _build(cfg, z, dev)
    sd = m.state_dict()
    init = group(z, "init/")
    assert list(sd.keys()) == list(init.keys())
    for k in sd:
        assert sd[k].dtype == torch.float32 and tuple(sd[k].shape) == tuple(init[k].shape)
    assert len(list(m.named_buffers())) == 0
    assert float(sd["dare.emb_att.weight"][0].abs().max()) == 0.0


def test_no_cpu_fallback():
    import b200ctr
    z, cfg = load_case("base_s1")
    g = cfg["golden"]
    cat_cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cat_cols}
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cat_cols)   # stays on CPU
    batch, _ = batch_of(z, 0)
    with pytest.raises(b200ctr.B200CtrError):
        m(batch)


def test_bf16_amp_mode_tracks_fp32():
    """amp: bf16 (tcgen05 contractions) vs exact fp32 on the same model/batch: logits and gradients agree to bf16
    rounding (stated tolerance: 3e-2 / 5e-2 of the tensor's max-abs; the top-k selection is fp32 in both)."""
    from b200ctr import ops
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    outs = {}
    for mode in ("fp32", "bf16"):
        m = _build(cfg, z, dev).train()
        batch, y = batch_of(z, 0)
        with ops.precision(mode):
            logits, prob, aux = m(batch)
            loss, dz, da, _, _ = ops.bce_wll(logits.detach(), aux.detach(), y.to(dev), m.aux_weight)
            torch.autograd.backward([logits, aux], [dz, da])
        outs[mode] = (logits.detach().cpu(), {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None})
    assert max_err_scaled(outs["bf16"][0], outs["fp32"][0]) < 3e-2
    for n, g in outs["fp32"][1].items():
        assert max_err_scaled(outs["bf16"][1][n], g) < 5e-2, n
    # and autocast selects the tensor-core path exactly like the reference's train loop would
    m = _build(cfg, z, dev).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l2, _, _ = m(batch)
    assert l2.dtype == torch.float32 and max_err_scaled(l2.detach().cpu(), outs["bf16"][0]) < 1e-6


@pytest.mark.parametrize("kind", ["qnn_concat", "s2_relu"])
def test_row_lazy_optimizer_equals_dense(kind):
    """Row-lazy AdamW+EMA on the sequence tables (catch-up before read, replay of skipped steps, closed-form decay of
    never-touched rows) vs the dense fused kernel over 8 steps with batches that touch different rows each step."""
    import b200ctr
    from b200ctr import synth
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    g = cfg["golden"]
    finals = {}
    for lazy in (False, True):
        m = _build(cfg, z, dev).train()
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=3, lazy_tables=lazy)
        gen = torch.Generator().manual_seed(5)
        for s in range(8):
            batch, y = batch_of(z, s % 3)
            seq = batch["seq"].clone()
            # rotate the token ids so that rows go untouched for several steps and then come back
            nz = seq != 0
            seq[nz] = (seq[nz] + 37 * (s % 4)) % (g["seq_vocab"] - 1) + 1
            if s == 5:
                seq[:, :] = 0                       # a step in which no table row is touched at all
            batch = dict(batch, seq=seq)
            tr.step(batch, y)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}      # pre-hook flushes the lazy tables
        finals[lazy] = (sd, {k: v.clone() for k, v in tr.ema_state().items()},
                        {n: (tr.tables[n]["m"].clone(), tr.tables[n]["v"].clone()) for n in tr.tables})
    for k in finals[False][0]:
        # ulp-level differences (closed-form vs sequential decay, and the order of the fp32 gradient atomics, which differs
        # between the two runs) fed back through 8 Adam steps: the observed spread between runs is 1e-4 .. 3.4e-4
        assert max_err_scaled(finals[True][0][k].cpu(), finals[False][0][k].cpu()) < 1e-3, k
    for k in finals[False][1]:
        assert max_err_scaled(finals[True][1][k].cpu(), finals[False][1][k].cpu()) < 1e-3, k
    for n in finals[False][2]:
        for a, b in zip(finals[True][2][n], finals[False][2][n]):
            assert max_err_scaled(a.cpu(), b.cpu()) < 2e-3, n


def _lazy_vs_dense(cfg, z, dev, n_steps, seq_of_step, small_hist=False):
    import b200ctr
    from b200ctr import ops
    finals = {}
    for lazy in (False, True):
        m = _build(cfg, z, dev).train()
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=3, lazy_tables=lazy)
        if lazy and small_hist:
            tr.hist = ops.RowHistory(dev, 4)          # far below the number of steps: push() must grow it
        for s in range(n_steps):
            batch, y = batch_of(z, s % 3)
            batch = dict(batch, seq=seq_of_step(s, batch["seq"].clone()))
            tr.step(batch, y)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        finals[lazy] = (sd, {k: v.clone() for k, v in tr.ema_state().items()})
        if lazy and small_hist:
            assert tr.hist.cap >= n_steps + 1
    return finals


def test_row_history_grows_past_its_initial_capacity():
    """the row-lazy optimizer's per-step history had a fixed capacity and aborted mid-training when the run was
    longer than estimated; it now grows geometrically.  12 steps through a history created with capacity 4 == dense."""
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    V = cfg["golden"]["seq_vocab"]

    def seq_of(s, seq):
        nz = seq != 0
        seq[nz] = (seq[nz] + 37 * (s % 4)) % (V - 1) + 1
        return seq
    f = _lazy_vs_dense(cfg, z, dev, 12, seq_of, small_hist=True)
    for k in f[False][0]:
        assert max_err_scaled(f[True][0][k].cpu(), f[False][0][k].cpu()) < 1e-3, k
    for k in f[False][1]:
        assert max_err_scaled(f[True][1][k].cpu(), f[False][1][k].cpu()) < 1e-3, k


def test_row_lazy_long_gap_closed_form_tail_equals_dense():
    """Rows with Adam moments that are not read for 300 steps (> the 256 skipped steps replayed exactly) and then come back:
    the closed-form tail (weight-decay product, EMA prefix, geometric moment decay) against the dense optimizer that really
    ran all 300 steps."""
    dev = torch.device("cuda:0")
    z, cfg = load_case("s2_relu")
    V = cfg["golden"]["seq_vocab"]
    half = (V - 1) // 2

    def seq_of(s, seq):
        nz = seq != 0
        first = s == 0 or s >= 300
        seq[nz] = (seq[nz] % half) + 1 + (0 if first else half)       # disjoint row sets: A at steps 0 and 300.., B in between
        return seq
    f = _lazy_vs_dense(cfg, z, dev, 302, seq_of)
    for k in f[False][0]:
        assert max_err_scaled(f[True][0][k].cpu(), f[False][0][k].cpu()) < 1e-3, k
    for k in f[False][1]:
        assert max_err_scaled(f[True][1][k].cpu(), f[False][1][k].cpu()) < 1e-3, k


def test_plain_module_eval_sees_up_to_date_tables_under_a_lazy_trainer():
    """the reference's validation pattern -- model.eval(); model(batch) (train.py:212-222) -- must not read rows whose
    pending weight decay / momentum steps have not been applied: the trainer's forward pre-hook flushes them."""
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    V = cfg["golden"]["seq_vocab"]
    m = _build(cfg, z, dev).train()
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=3)
    for s in range(4):
        batch, y = batch_of(z, s % 3)
        seq = batch["seq"].clone()
        nz = seq != 0
        seq[nz] = (seq[nz] + 53 * s) % (V - 1) + 1
        tr.step(dict(batch, seq=seq), y)
    batch, _ = batch_of(z, 0)                      # rows last touched at step 1: three steps of decay are pending
    m.eval()
    with torch.no_grad():
        logits, prob, aux = m(batch)
    ref_logits, _, _ = tr.predict(batch)
    assert max_err_scaled(logits.cpu(), ref_logits.cpu()) < 1e-6
    # and the dense (non-lazy) trainer agrees on what those weights should be
    m2 = _build(cfg, z, dev).train()
    tr2 = b200ctr.FusedTrainer(m2, cfg, steps_per_epoch=3, lazy_tables=False)
    for s in range(4):
        b2, y = batch_of(z, s % 3)
        seq = b2["seq"].clone()
        nz = seq != 0
        seq[nz] = (seq[nz] + 53 * s) % (V - 1) + 1
        tr2.step(dict(b2, seq=seq), y)
    m2.eval()
    with torch.no_grad():
        l2, _, _ = m2(batch)
    assert max_err_scaled(logits.cpu(), l2.cpu()) < 1e-3


def test_fc_head_sees_undropped_categorical_embeddings():
    """embedding_dropout applies to cat_stack (QNN input) only; the FC head concatenates the un-dropped cat_embs
    (wrapper.py:149-150 vs 167-171).  With p_emb = 0.5 in train mode the FC-head logits must equal the p_emb = 0 logits."""
    import copy
    dev = torch.device("cuda:0")
    z, cfg = load_case("base_s1")
    batch, _ = batch_of(z, 0)
    outs = []
    for p in (0.0, 0.5):
        c = copy.deepcopy(cfg)
        c["model"]["embedding_dropout"] = p
        m = _build(c, z, dev).train()
        with torch.no_grad():
            outs.append(m(batch)[0].cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("kind,prec", [("qnn_concat", "bf16"), ("s2_relu", "bf16"), ("base_s1", "none")])
def test_two_runs_are_bit_identical(kind, prec):
    """No floating-point atomics on the training path: table-gradient rows are reduced after a stable sort in slot order
    (csrc/scatter.cu), small accumulators through per-CTA / per-sample partial rows summed in order, the gradient norm as an
    integer sum.  Two runs of 4 steps from the same state must agree BITWISE in every parameter, Adam moment and EMA shadow.
    (bf16 mode = the benchmarked path; in fp32 mode one accumulator, the relative-position bias gradient of the CUDA-core
    attention backward, still uses shared-memory atomics inside a CTA, so fp32 is checked on the encoder-less case.)"""
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    cfg = dict(cfg, amp=prec)
    V = cfg["golden"]["seq_vocab"]
    runs = []
    for _ in range(2):
        m = _build(cfg, z, dev).train()
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
        for s in range(4):
            batch, y = batch_of(z, s % 3)
            seq = batch["seq"].clone()
            nz = seq != 0
            seq[nz] = (seq[nz] % 7) + 1 + (s % 2)            # a handful of distinct tokens: every gradient row collects many slots
            tr.step(dict(batch, seq=seq), y)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        mom = {n: (tr.tables[n]["m"].clone(), tr.tables[n]["v"].clone()) for n in tr.tables}
        runs.append((sd, {k: v.clone() for k, v in tr.ema_state().items()}, mom, tr.flat_m.clone(), tr.flat_v.clone(),
                     float(tr.total_norm.cpu())))
    a, b = runs
    assert a[5] == b[5]
    for k in a[0]:
        assert torch.equal(a[0][k], b[0][k]), k
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), ("ema", k)
    for n in a[2]:
        assert torch.equal(a[2][n][0], b[2][n][0]) and torch.equal(a[2][n][1], b[2][n][1]), n
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])


def test_staged_batches_give_the_same_steps_as_batches_handed_to_step():
    """FusedTrainer.stage (copy stream + two preallocated staging sets, the next batch staged while the current step runs)
    feeds step() the same bytes as step(batch, y): parameters, EMA and the losses agree BITWISE over 5 steps, including a
    batch of another shape in the middle (the slot is re-allocated) and the two-ahead limit."""
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    cfg = dict(cfg, amp="bf16")
    batches = []
    for s in range(5):
        batch, y = batch_of(z, s % 3)
        if s == 2:                                        # a ragged batch: half the rows
            n = y.shape[0] // 2
            batch, y = {k: v[:n].contiguous() for k, v in batch.items()}, y[:n].contiguous()
        batches.append(({k: v.pin_memory() for k, v in batch.items()}, y.pin_memory()))
    runs = []
    for staged in (False, True):
        m = _build(cfg, z, dev).train()
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
        losses = []
        if staged:
            with pytest.raises(RuntimeError):
                tr.step()                                  # nothing staged yet
            tr.stage(*batches[0])
            for i in range(5):
                loss = tr.step()
                if i + 1 < 5:
                    tr.stage(*batches[i + 1])
                losses.append(float(loss))
            tr.stage(*batches[0]); tr.stage(*batches[1])
            with pytest.raises(RuntimeError):
                tr.stage(*batches[2])                      # two batches are already ahead
        else:
            for b, y in batches:
                losses.append(float(tr.step(b, y)))
        runs.append((losses, {k: v.detach().clone() for k, v in m.state_dict().items()},
                     {k: v.clone() for k, v in tr.ema_state().items()}))
    a, b = runs
    assert a[0] == b[0]
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), k
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), ("ema", k)


def test_sorted_scatter_equals_atomic_scatter_and_the_oracle():
    """b200ctr_scatter_rows (sort + segmented reduction) against the fp64 index_add of the same contributions, with heavy
    duplication (rows that collect thousands of slots and span many 32-slot chunks) and pad slots."""
    from b200ctr import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    B, K, D, V = 300, 37, 64, 5000
    tok = torch.randint(1, 40, (B, K), generator=g)                        # 39 hot rows
    tok[torch.rand(B, K, generator=g) < 0.3] = torch.randint(40, V, (1,), generator=g).item()
    tok[:, -3:] = torch.randint(1, V, (B, 3), generator=g)                 # + rows hit once
    tok[:, :15] = 7                                                        # + one row with 4,500 slots: ~140 chunks, solid groups
    tok[torch.rand(B, K, generator=g) < 0.1] = 0                           # pad slots
    q = torch.randn(B, D, generator=g)
    dvals = torch.randn(B, K, generator=g)
    drep = torch.randn(B, K, D, generator=g)
    g_att = torch.zeros(V, D, device=dev)
    g_rep = torch.zeros(V, D, device=dev)
    ops.scatter_rows(tok.int().to(dev), 1, B * K, 0, V, K, q.to(dev), dvals.to(dev), drep.to(dev), None, None, 0, 1, 0, g_att, g_rep)
    flat = tok.reshape(-1)
    keep = flat != 0
    ref_att = torch.zeros(V, D, dtype=torch.float64).index_add_(0, flat[keep], (dvals.reshape(-1, 1).double() * q.double().repeat_interleave(K, 0))[keep])
    ref_rep = torch.zeros(V, D, dtype=torch.float64).index_add_(0, flat[keep], drep.reshape(-1, D).double()[keep])
    assert max_err_scaled(g_att.cpu(), ref_att) < 1e-6 and max_err_scaled(g_rep.cpu(), ref_rep) < 1e-6
    assert float(g_att[0].abs().max()) == 0.0                               # the padding row gets no gradient
    g2a, g2r = torch.zeros(V, D, device=dev), torch.zeros(V, D, device=dev)
    ops.scatter_rows(tok.int().to(dev), 1, B * K, 0, V, K, q.to(dev), dvals.to(dev), drep.to(dev), None, None, 0, 1, 0, g2a, g2r)
    assert torch.equal(g_att, g2a) and torch.equal(g_rep, g2r)


@pytest.mark.parametrize("kind", ["qnn_concat", "base_s1"])
def test_row_lazy_categorical_tables_match_reference(kind):
    """Enlarged hashed-categorical tables (cfgs/v4_k120_s1.yaml: 2,000,503 x 32 and 4,000,513 x 57 rows) take the sequence
    tables' route in FusedTrainer: stored 64 wide outside the flat buffer, gradient rows from the sort + segmented reduction,
    row-lazy AdamW + EMA.  Forced here on the small golden tables (threshold 100 rows: d_c = 33, 57 and 64 wide tables qualify)
    and held to the reference's 3-step parameters and EMA shadows, and to the dense path over 8 steps with wandering ids."""
    import copy
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case(kind)
    cfg = copy.deepcopy(cfg)
    cfg["b200ctr"] = {"lazy_table_min_rows": 100}
    m = _build(cfg, z, dev).train()
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
    assert len(tr.cat_tables) >= (2 if kind == "qnn_concat" else 1) and not any(n in tr.offsets for n in tr.cat_tables)
    for s in range(3):
        batch, y = batch_of(z, s)
        loss = tr.step(batch, y)
        assert abs(float(loss.cpu()) - float(z[f"step{s}/loss"])) < 2e-4
        if cfg["train"]["grad_clip_norm"] > 0:
            assert abs(float(tr.total_norm.cpu()) - float(z[f"step{s}/total_norm"])) < 2e-4 * float(z[f"step{s}/total_norm"])
    sd = m.state_dict()
    for k, v in group(z, "final/").items():
        assert max_err_scaled(sd[k].cpu(), v) < 1e-3, k
    ema = tr.ema_state()
    for k, v in group(z, "ema/").items():
        assert max_err_scaled(ema[k].cpu(), v) < 1e-3, k
    assert float(tr.cat_rows_acc.abs().max()) == 0.0                       # the row accumulator is back to all-zero
    # lazy vs dense over 8 steps, ids wandering so that rows go stale and come back
    finals = {}
    cards = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cfg["data"]["cat_cols"]}
    cvec = torch.tensor([cards[c] for c in cfg["data"]["cat_cols"]])
    for thr in (100, 10 ** 9):
        c2 = copy.deepcopy(cfg)
        c2["b200ctr"] = {"lazy_table_min_rows": thr}
        m2 = _build(c2, z, dev).train()
        t2 = b200ctr.FusedTrainer(m2, c2, steps_per_epoch=3)
        for s in range(8):
            batch, y = batch_of(z, s % 3)
            xc = (batch["X_cat"] + 17 * (s % 4)) % cvec
            t2.step(dict(batch, X_cat=xc), y)
        finals[thr] = ({k: v.detach().clone() for k, v in m2.state_dict().items()}, {k: v.clone() for k, v in t2.ema_state().items()})
    for k in finals[100][0]:
        assert max_err_scaled(finals[100][0][k].cpu(), finals[10 ** 9][0][k].cpu()) < 1e-3, k
    for k in finals[100][1]:
        assert max_err_scaled(finals[100][1][k].cpu(), finals[10 ** 9][1][k].cpu()) < 1e-3, k
