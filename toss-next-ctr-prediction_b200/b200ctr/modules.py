"""Drop-in mirrors of the reference's nn.Module surfaces, backed by libb200ctr.

Same class names, constructor signatures, attribute / sub-module names (hence identical `state_dict` keys, shapes
and dtypes: SURVEY App. B) and the same construction ORDER as the reference, so that under the same torch seed the
initial parameters are bit-identical:
    src/models/feature_embed.py  -> NumericFeatureEmbedding, BinaryFeatureEmbedding
    src/models/dare.py           -> RMSNorm, make_norm, PositionalBias, DAREEncoderLayer, DARE
    src/models/qnn_alpha.py      -> RMSNorm, make_norm, SEBlock, QNNAlphaDetailed
    src/models/wrapper.py        -> CTRModel
torch.nn modules (nn.Embedding, nn.Linear, nn.MultiheadAttention ...) appear only as PARAMETER CONTAINERS for
initialisation and state_dict layout; their forward methods are never called.  All compute goes through
b200ctr.hotpath -> the C ABI.  CPU tensors as *inputs* are copied to the parameters' device (the reference does the
same inside forward, wrapper.py:139-143); parameters on a CPU device raise: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import hotpath as hp
from . import ops
from ._lib import B200CtrError


def _require_cuda(p: torch.Tensor, what: str):
    if not p.is_cuda:
        raise B200CtrError(f"{what}: parameters live on {p.device}; b200ctr runs on a B200 only (no CPU fallback). "
                           "Call .to('cuda') first.")


class _GradStore:
    """G(name) -> zero-initialised gradient tensor shaped like the parameter; remembers which names were touched."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params, self.grads = params, {}

    def __call__(self, name: str) -> torch.Tensor:
        g = self.grads.get(name)
        if g is None:
            g = torch.zeros_like(self.params[name])
            self.grads[name] = g
        return g


# ---------------------------------------------------------------------------------------------------------------
# feature_embed.py
# ---------------------------------------------------------------------------------------------------------------
class _NumBinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, w_out):
        B, F = x.shape
        D = w_out.shape[0]
        x = x.contiguous()
        Pm, Qm = ops.numbin_fold(w, b, w_out)
        out = torch.empty((B, F, D), device=w.device, dtype=torch.float32)
        ops.numbin_fwd(x, Pm, Qm, out, F * D, None, 0)
        ctx.save_for_backward(x, w, b if b is not None else torch.empty(0, device=w.device), w_out)
        ctx.has_b = b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, b, w_out = ctx.saved_tensors
        b = b if ctx.has_b else None
        dout = dout.contiguous()
        F, D = x.shape[1], w_out.shape[0]
        dP, dQ = ops.numbin_bwd(x, dout, F * D, None, 0, D)
        dw, db, dwo = ops.numbin_unfold(dP, dQ, w, b, w_out)
        return None, dw, db, dwo


class NumericFeatureEmbedding(nn.Module):
    """(B,F) -> (B,F,D): (x*slope + bias) @ out_proj^T.  Mirrors feature_embed.py:5-27."""

    def __init__(self, n_features: int, f_embed: int, out_dim: int, bias: bool = True):
        super().__init__()
        self.n_features, self.f_embed = n_features, f_embed
        self.weight = nn.Parameter(torch.randn(n_features, f_embed) * 0.02)
        self.bias = nn.Parameter(torch.zeros(n_features, f_embed)) if bias else None
        self.out_proj = nn.Linear(f_embed, out_dim, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, F = x.shape
        assert F == self.n_features
        _require_cuda(self.weight, type(self).__name__)
        x = x.to(self.weight.device, non_blocking=True).float()
        return _NumBinFn.apply(x, self.weight, self.bias, self.out_proj.weight)


class BinaryFeatureEmbedding(NumericFeatureEmbedding):
    """Mirrors feature_embed.py:29-48 (bias off by default)."""

    def __init__(self, n_features: int, f_embed: int, out_dim: int, bias: bool = False):
        super().__init__(n_features, f_embed, out_dim, bias=bias)


# ---------------------------------------------------------------------------------------------------------------
# dare.py
# ---------------------------------------------------------------------------------------------------------------
class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        if shp[-1] == 64:
            s, y, r = ops.add_rmsnorm_fwd(x2, None, w, eps, save_sum=False)
            ctx.row = False
        else:
            y, r = ops.rowrms_fwd(x2, w, eps)
            ctx.row = True
        ctx.save_for_backward(x2, w, r)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, r = ctx.saved_tensors
        d2 = dy.reshape(x2.shape).contiguous()
        if ctx.row:
            dx, dw = ops.rowrms_bwd(d2, x2, w, r)
        else:
            dw = torch.zeros_like(w)
            dx = ops.rmsnorm_bwd(d2, x2, w, r, dw)
        return dx.view(dy.shape), dw, None


class RMSNorm(nn.Module):
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.w = nn.Parameter(torch.ones(d))

    def forward(self, x):
        _require_cuda(self.w, "RMSNorm")
        if self.w.numel() % 4:
            raise B200CtrError("RMSNorm width must be a multiple of 4")
        return _RMSNormFn.apply(x.float(), self.w, self.eps)


def make_norm(name: str, d: int):
    if name.lower() == "rms":
        return RMSNorm(d)
    raise B200CtrError("only norm='rms' is built in b200ctr (every shipped cfg uses it)")


class PositionalBias(nn.Module):
    """Parameter container for the relative-position table (dare.py:20-37); forward kept for API parity."""

    def __init__(self, max_len=512, n_heads=4, dim=32):
        super().__init__()
        self.rel = nn.Embedding(2 * max_len + 1, n_heads)
        self.max_len, self.n_heads = max_len, n_heads

    def forward(self, qlen, klen):
        # (n_heads, qlen, klen) gather of a tiny table: index plumbing, not on the hot path (the fused attention
        # kernel reads the head-mean table directly)
        dev = self.rel.weight.device
        i = torch.arange(qlen, device=dev).unsqueeze(1)
        j = torch.arange(klen, device=dev).unsqueeze(0)
        d = (j - i).clamp(-self.max_len, self.max_len) + self.max_len
        return self.rel.weight[d].permute(2, 0, 1)


def _spec_from_layer(d_model, n_heads, mha_do, ffn_hidden, ffn_do, add_pos, max_len) -> hp.ModelSpec:
    cfg = {"model": {"emb_dim": d_model, "dare_dropout": 0.0, "qnn_alpha": {"enabled": False}},
           "sequence": {"max_len": max_len, "top_k": max_len, "recency_tau": 1, "pad_id": 0, "query_mode": "S2",
                        "query_key": "", "transformer_block": True,
                        "tfm": {"n_layers": 1, "n_heads": n_heads, "mha_dropout": mha_do, "ffn_hidden": ffn_hidden,
                                "ffn_dropout": ffn_do, "norm": "rms", "add_positional_bias": add_pos}}}
    return hp.ModelSpec(cfg, 0, 0, [])


_LAYER_KEYS = ["mha.in_proj_weight", "mha.in_proj_bias", "mha.out_proj.weight", "mha.out_proj.bias", "norm1.w",
               "ffn.0.weight", "ffn.0.bias", "ffn.3.weight", "ffn.3.bias", "norm2.w", "pbias.rel.weight"]


class _StageFn(torch.autograd.Function):
    """Generic bridge: runs a (fwd, bwd) pair of hotpath stage functions over a name->parameter dict."""

    @staticmethod
    def forward(ctx, holder, *tensors):
        ctx.set_materialize_grads(False)
        ctx.holder = holder
        # the reference's train loop wraps forward+loss in torch autocast when cfg.amp is bf16/fp16 (train.py:158-161):
        # under autocast the contractions run on the tensor cores with bf16 operands; validation/inference run
        # outside autocast and stay exact fp32.  The backward pass reuses the forward's mode.
        ctx.prec = "bf16" if torch.is_autocast_enabled() else ops.get_precision()
        names = holder["names"]
        n_in = holder["n_inputs"]
        inputs = tensors[:n_in]
        P = dict(zip(names, tensors[n_in:]))
        with ops.precision(ctx.prec):
            outs, sv = holder["fwd"](P, inputs)
        ctx.P, ctx.sv = P, sv
        ctx.n_in = n_in
        nd = holder.get("non_diff", ())
        if nd:
            ctx.mark_non_differentiable(*[outs[i] for i in nd])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        holder = ctx.holder
        G = _GradStore(ctx.P)
        with ops.precision(ctx.prec):
            dins = holder["bwd"](ctx.P, ctx.sv, [d.float() if d is not None else None for d in douts], G)
        grads = [G.grads.get(n) for n in holder["names"]]
        dins = list(dins) + [None] * (ctx.n_in - len(dins))
        ctx.sv = None
        return (None, *dins, *grads)


class DAREEncoderLayer(nn.Module):
    """Post-norm encoder layer (dare.py:39-70): MHA with head-mean Toeplitz bias, RMSNorm, GELU FFN."""

    def __init__(self, d_model, n_heads, mha_dropout, ffn_hidden, ffn_dropout, norm="rms", add_pos_bias=True, max_len=512):
        super().__init__()
        self.n_heads = n_heads
        self.mha = nn.MultiheadAttention(d_model, n_heads, dropout=mha_dropout, batch_first=True)   # parameter container
        self.norm1 = make_norm(norm, d_model)
        self.ffn = nn.Sequential(nn.Linear(d_model, ffn_hidden), nn.GELU(), nn.Dropout(ffn_dropout),
                                 nn.Linear(ffn_hidden, d_model))
        self.norm2 = make_norm(norm, d_model)
        self.add_pos_bias = add_pos_bias
        self.pbias = PositionalBias(max_len=max_len, n_heads=n_heads) if add_pos_bias else None
        self._spec = _spec_from_layer(d_model, n_heads, mha_dropout, ffn_hidden, ffn_dropout, add_pos_bias, max_len)
        self._step = 0

    def forward(self, x, key_padding_mask=None):
        if key_padding_mask is not None:
            raise B200CtrError("key_padding_mask is not supported (the reference never passes one, dare.py:143-146)")
        _require_cuda(self.norm1.w, "DAREEncoderLayer")
        B, K, D = x.shape
        spec = self._spec
        sd = {k: v for k, v in self.named_parameters()}
        names = [k for k in _LAYER_KEYS if k in sd]
        dc = hp.DropCtl(self.training, seed=torch.initial_seed(), step=self._step)
        self._step += int(self.training)

        def fwd(P, inputs):
            P2 = {"dare.layers.0." + k: v for k, v in P.items()}
            y, sv = hp.encoder_layer_fwd(inputs[0].reshape(B * K, D).contiguous().float(), P2, 0, spec, B, K, dc)
            return [y.view(B, K, D)], sv

        def bwd(P, sv, douts, G):
            P2 = {"dare.layers.0." + k: v for k, v in P.items()}
            dx = hp.encoder_layer_bwd(douts[0].reshape(B * K, D).contiguous(), sv, P2, 0, spec, B, K, dc,
                                      lambda n: G(n[len("dare.layers.0."):]))
            return [dx.view(B, K, D)]

        holder = dict(names=names, n_inputs=1, fwd=fwd, bwd=bwd)
        return _StageFn.apply(holder, x, *[sd[k] for k in names])[0]


class DARE(nn.Module):
    """Target-aware retrieval encoder (dare.py:72-162)."""

    def __init__(self, seq_vocab: int, emb_dim: int, dropout: float, query_mode: str = "S1",
                 transformer_block: bool = True, top_k: int = 80, recency_tau: int = 256, pad_id: int = 0,
                 tfm_cfg: Optional[dict] = None):
        super().__init__()
        self.emb_att = nn.Embedding(seq_vocab, emb_dim, padding_idx=pad_id)
        self.emb_rep = nn.Embedding(seq_vocab, emb_dim, padding_idx=pad_id)
        self.query_mode = query_mode
        self.top_k = top_k
        self.tau = recency_tau
        self.pad_id = pad_id
        self.dropout = nn.Dropout(dropout)
        self.transformer_block = transformer_block
        if transformer_block:
            n_layers = tfm_cfg.get("n_layers", 2)
            n_heads = tfm_cfg.get("n_heads", 4)
            mha_dropout = tfm_cfg.get("mha_dropout", 0.1)
            ffn_hidden = tfm_cfg.get("ffn_hidden", 256)
            ffn_dropout = tfm_cfg.get("ffn_dropout", 0.1)
            norm = tfm_cfg.get("norm", "rms")
            add_pos = tfm_cfg.get("add_positional_bias", True)
            self.layers = nn.ModuleList([
                DAREEncoderLayer(emb_dim, n_heads, mha_dropout, ffn_hidden, ffn_dropout, norm, add_pos, max_len=top_k)
                for _ in range(n_layers)])
        self.gating = tfm_cfg.get("gating", "softmax") if tfm_cfg else "softmax"
        self.aux_head = nn.Linear(emb_dim, 1)
        cfg = {"model": {"emb_dim": emb_dim, "dare_dropout": dropout, "qnn_alpha": {"enabled": False}},
               "sequence": {"max_len": 0, "top_k": top_k, "recency_tau": recency_tau, "pad_id": pad_id,
                            "query_mode": "S2", "query_key": "", "transformer_block": transformer_block,
                            "tfm": tfm_cfg}}
        self._spec = hp.ModelSpec(cfg, 0, 0, [])
        self._step = 0

    def _named(self) -> Dict[str, torch.Tensor]:
        return {"dare." + k: v for k, v in self.named_parameters()}

    def _inputs(self, seq_ids, query_vec):
        dev = self.emb_att.weight.device
        _require_cuda(self.emb_att.weight, "DARE")
        seq = seq_ids.to(dev, non_blocking=True)
        if seq.dtype not in (torch.int64, torch.int32):
            seq = seq.long()
        return seq.contiguous(), query_vec.to(dev).float().contiguous()

    def topk_select(self, seq_ids, query_vec):
        """(B,L) ids, (B,D) query -> (sel_rep (B,K,D), vals (B,K)), sorted by (score desc, position asc)."""
        seq, q = self._inputs(seq_ids, query_vec)
        spec = self._spec
        P = self._named()
        names = ["dare.emb_att.weight", "dare.emb_rep.weight"]

        def fwd(Pd, inputs):
            B, L = seq.shape
            K = min(spec.top_k, L)
            rec = hp.recency_table(L, spec.tau, q.device)
            sel, vals, idx, tok = ops.retrieve_fwd(seq, inputs[0], Pd[names[0]], Pd[names[1]], rec, K, spec.pad_id)
            self.last_idx = idx
            return [sel, vals], dict(tok=tok, q=inputs[0])

        def bwd(Pd, sv, douts, G):
            dsel = douts[0].contiguous() if douts[0] is not None else torch.zeros_like(sv["q"]).new_zeros(
                (*sv["tok"].shape, spec.D))
            dv = douts[1].contiguous() if douts[1] is not None else torch.zeros(sv["tok"].shape, device=dsel.device)
            dq = ops.retrieve_bwd(sv["tok"], sv["q"], Pd[names[0]], dsel, dv, spec.pad_id, G(names[0]), G(names[1]))
            return [dq]

        holder = dict(names=names, n_inputs=1, fwd=fwd, bwd=bwd)
        sel, vals = _StageFn.apply(holder, q, *[P[n] for n in names])
        return sel, vals

    def forward(self, seq_ids, query_vec):
        seq, q = self._inputs(seq_ids, query_vec)
        spec = self._spec
        P = self._named()
        names = list(P.keys())
        B = seq.shape[0]
        dc = hp.DropCtl(self.training, seed=torch.initial_seed(), step=self._step)
        self._step += int(self.training)

        def fwd(Pd, inputs):
            u = torch.empty((B, spec.D), device=q.device, dtype=torch.float32)
            aux, sv = hp.dare_fwd(Pd, spec, seq, inputs[0], u, spec.D, dc)
            return [u, aux], sv

        def bwd(Pd, sv, douts, G):
            du = douts[0].contiguous() if douts[0] is not None else None
            daux = douts[1].contiguous() if douts[1] is not None else None
            dq = hp.dare_bwd(Pd, spec, sv, du, spec.D, daux, dc, G, want_aux_grads=daux is not None)
            return [dq]

        holder = dict(names=names, n_inputs=1, fwd=fwd, bwd=bwd)
        u, aux = _StageFn.apply(holder, q, *[P[n] for n in names])
        return u, aux


# ---------------------------------------------------------------------------------------------------------------
# qnn_alpha.py
# ---------------------------------------------------------------------------------------------------------------
class SEBlock(nn.Module):
    """Batch-statistic squeeze-excite gate (qnn_alpha.py:17-26): x * sigmoid(W2 relu(W1 mean_b(x)))."""

    def __init__(self, c: int, r: int = 8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c, c // r), nn.ReLU(inplace=True), nn.Linear(c // r, c), nn.Sigmoid())

    def forward(self, x):
        _require_cuda(self.fc[0].weight, "SEBlock")
        names = ["w1", "b1", "w2", "b2"]
        params = [self.fc[0].weight, self.fc[0].bias, self.fc[2].weight, self.fc[2].bias]
        B = x.shape[0]

        def fwd(P, inputs):
            xi = inputs[0].contiguous().float()
            mean, h1, gate = ops.se_gate_fwd(ops.colsum(xi), B, P["w1"], P["b1"], P["w2"], P["b2"])
            return [ops.se_apply_fwd(xi, gate)], dict(x=xi, mean=mean, h1=h1, gate=gate)

        def bwd(P, sv, douts, G):
            d = douts[0].contiguous()
            _, gprod = ops.se_apply_bwd(d, sv["x"], sv["gate"], None, B, False, True)
            dw1, db1, dw2, db2, dmean = ops.se_gate_bwd(ops.colsum(gprod), sv["mean"], sv["h1"], sv["gate"], P["w1"], P["w2"])
            G("w1").copy_(dw1); G("b1").copy_(db1); G("w2").copy_(dw2); G("b2").copy_(db2)
            dx, _ = ops.se_apply_bwd(d, sv["x"], sv["gate"], dmean, B, True, False)
            return [dx]

        return _StageFn.apply(dict(names=names, n_inputs=1, fwd=fwd, bwd=bwd), x, *params)[0]


class QNNAlphaDetailed(nn.Module):
    """QNN-alpha head (qnn_alpha.py:28-130): RMSNorm over F*D, multi-head low-rank FM interaction, SE gate,
    residual concat, MLP."""

    def __init__(self, *, in_feat: int, emb_dim: int, heads: int = 8, rank: int = 32, proj_dim: int = 256,
                 mlp_hidden=(512, 256), dropout: float = 0.1, use_se: bool = True, se_reduction: int = 8,
                 use_residual: bool = True, norm: str = "rms", pair_grouping: str = "block",
                 block_slices: Optional[dict] = None):
        super().__init__()
        self.in_feat, self.emb_dim, self.heads, self.rank, self.proj_dim = in_feat, emb_dim, heads, rank, proj_dim
        self.use_se, self.use_residual, self.pair_grouping = use_se, use_residual, pair_grouping
        self.block_slices = block_slices or {}
        self.pre_norm = make_norm(norm, in_feat * emb_dim)
        self.U = nn.Parameter(torch.randn(heads, emb_dim, rank) * 0.02)
        self.V = nn.Parameter(torch.randn(heads, rank, proj_dim) * 0.02)
        self.dropout = nn.Dropout(dropout)
        out_dim = heads * proj_dim
        self.se = SEBlock(out_dim, r=se_reduction) if use_se else nn.Identity()
        layers = []
        in_dim = out_dim + in_feat * emb_dim
        for h in mlp_hidden:
            layers += [nn.Linear(in_dim, h), nn.ReLU(inplace=True), nn.Dropout(dropout)]
            in_dim = h
        layers.append(nn.Linear(in_dim, 1))
        self.mlp = nn.Sequential(*layers)
        if pair_grouping == "block" and self.block_slices:
            raise B200CtrError("qnn_alpha.pair_grouping='block' is not built (every shipped cfg uses 'all')")
        cfg = {"model": {"emb_dim": emb_dim, "dare_dropout": 0.0,
                         "qnn_alpha": {"enabled": True, "heads": heads, "rank": rank, "proj_dim": proj_dim,
                                       "mlp_hidden": list(mlp_hidden), "dropout": dropout, "use_se": use_se,
                                       "se_reduction": se_reduction, "use_residual": use_residual, "norm": norm,
                                       "pair_grouping": "all"}},
               "sequence": {"max_len": 0, "top_k": 1, "recency_tau": 1, "pad_id": 0, "query_mode": "S2", "query_key": "",
                            "transformer_block": False, "tfm": None}}
        self._spec = hp.ModelSpec(cfg, 0, 0, [])
        self._step = 0

    def forward(self, feats):
        _require_cuda(self.U, "QNNAlphaDetailed")
        spec = self._spec
        P = {"qnn." + k: v for k, v in self.named_parameters()}
        names = list(P.keys())
        dc = hp.DropCtl(self.training, seed=torch.initial_seed(), step=self._step)
        self._step += int(self.training)

        def fwd(Pd, inputs):
            logits, sv = hp.qnn_fwd(Pd, spec, inputs[0].contiguous().float(), dc)
            return [logits], sv

        def bwd(Pd, sv, douts, G):
            return [hp.qnn_bwd(Pd, spec, sv, douts[0].contiguous(), dc, G)]

        return _StageFn.apply(dict(names=names, n_inputs=1, fwd=fwd, bwd=bwd), feats, *[P[n] for n in names])[0]


# ---------------------------------------------------------------------------------------------------------------
# wrapper.py
# ---------------------------------------------------------------------------------------------------------------
class CTRModel(nn.Module):
    """Drop-in for src/models/wrapper.py:CTRModel: same ctor, same state_dict, forward(batch) -> (logits, prob, aux)."""

    def __init__(self, cfg, seq_vocab: int, num_feat_dim: int, mask_feat_dim: int, cat_cardinals: dict,
                 cat_cols_order: list):
        super().__init__()
        D = cfg["model"]["emb_dim"]
        f_embed = int(cfg["model"]["qnn_alpha"].get("feature_embed_dim", max(8, D // 4)))
        self.cat_cols_order = cat_cols_order
        self.num_dim, self.mask_dim, self.D = num_feat_dim, mask_feat_dim, D
        p = float(cfg["model"].get("embedding_dropout", 0.0))
        self.emb_dropout = nn.Dropout(p)
        self.num_embed = NumericFeatureEmbedding(num_feat_dim, f_embed, D) if num_feat_dim > 0 else None
        self.mask_embed = BinaryFeatureEmbedding(mask_feat_dim, f_embed, D) if mask_feat_dim > 0 else None
        self.cat_embs = nn.ModuleDict()
        self.cat_proj = nn.ModuleDict()
        cat_dims_map = cfg["model"].get("cat_embedding_dims", {}) if "model" in cfg else {}
        for c, card in cat_cardinals.items():
            d_in = int(cat_dims_map.get(c, D))
            self.cat_embs[c] = nn.Embedding(card, d_in)
            self.cat_proj[c] = nn.Linear(d_in, D, bias=False)
        ctx_in = D * ((num_feat_dim > 0) + (mask_feat_dim > 0) + 1)
        self.ctx_mlp = nn.Sequential(nn.Linear(ctx_in, D), nn.ReLU(inplace=True))
        tfm_cfg = cfg["sequence"].get("tfm", {})
        self.dare = DARE(seq_vocab=seq_vocab, emb_dim=D, dropout=cfg["model"]["dare_dropout"],
                         query_mode=cfg["sequence"]["query_mode"], transformer_block=cfg["sequence"]["transformer_block"],
                         top_k=cfg["sequence"]["top_k"], recency_tau=cfg["sequence"]["recency_tau"],
                         pad_id=cfg["sequence"]["pad_id"], tfm_cfg=tfm_cfg)
        self.use_qnn = cfg["model"]["qnn_alpha"]["enabled"]
        if self.use_qnn:
            F_all = 1 + max(num_feat_dim, 0) + max(mask_feat_dim, 0) + len(cat_cardinals)
            ofs = 0
            block_slices = {"seq": (0, 1)}
            ofs = 1
            if num_feat_dim > 0:
                block_slices["num"] = (ofs, ofs + num_feat_dim); ofs += num_feat_dim
            if mask_feat_dim > 0:
                block_slices["mask"] = (ofs, ofs + mask_feat_dim); ofs += mask_feat_dim
            block_slices["cat"] = (ofs, ofs + len(cat_cardinals)); ofs += len(cat_cardinals)
            assert ofs == F_all
            qa = cfg["model"]["qnn_alpha"]
            if qa["pair_grouping"] != "all":
                raise B200CtrError("qnn_alpha.pair_grouping='block' is not built (every shipped cfg uses 'all')")
            self.qnn = QNNAlphaDetailed(in_feat=F_all, emb_dim=D, heads=qa["heads"], rank=qa["rank"],
                                        proj_dim=qa["proj_dim"], mlp_hidden=qa["mlp_hidden"], dropout=qa["dropout"],
                                        use_se=qa["use_se"], se_reduction=qa["se_reduction"],
                                        use_residual=qa["use_residual"], norm=qa["norm"], pair_grouping="all",
                                        block_slices=block_slices)
        else:
            self.fc = nn.Sequential(
                nn.Linear(D * (1 + (num_feat_dim > 0) + (mask_feat_dim > 0) + len(cat_cardinals)), 512),
                nn.ReLU(inplace=True), nn.Dropout(0.1), nn.Linear(512, 1))
        self.sigmoid = nn.Sigmoid()
        self.query_key = cfg["sequence"]["query_key"]
        self.aux_weight = float(cfg["model"]["qnn_alpha"].get("aux_head_weight", 0.0))
        if cfg["model"]["qnn_alpha"]["enabled"] is False and "tfm" not in cfg["sequence"]:
            pass
        self._spec = hp.ModelSpec(cfg, num_feat_dim, mask_feat_dim, list(cat_cols_order))
        self._cache: dict = {}
        self._step = 0

    # -- the hot path --------------------------------------------------------------------------------------
    def _device_batch(self, batch) -> Dict[str, torch.Tensor]:
        dev = next(self.parameters()).device
        _require_cuda(next(self.parameters()), "CTRModel")
        out = {"X_num": batch["X_num"].to(dev, non_blocking=True).float().contiguous(),
               "X_mask": batch["X_mask"].to(dev, non_blocking=True).float().contiguous()}
        for k in ("X_cat", "seq"):
            t = batch[k].to(dev, non_blocking=True)
            if t.dtype not in (torch.int64, torch.int32):
                t = t.long()
            out[k] = t.contiguous()
        return out

    def forward(self, batch):
        spec = self._spec
        if not self.use_qnn:
            spec.p_fc = float(self.fc[2].p)          # honour the module attribute like the reference's nn.Dropout does
        dbatch = self._device_batch(batch)
        P = dict(self.named_parameters())
        names = list(P.keys())
        dc = hp.DropCtl(self.training, seed=torch.initial_seed(), step=self._step)
        self._step += int(self.training)
        cache = self._cache

        def fwd(Pd, inputs):
            logits, aux, sv = hp.model_fwd(Pd, spec, dbatch, dc, cache)
            prob = ops.sigmoid(logits)
            return [logits, prob, aux], sv

        def bwd(Pd, sv, douts, G):
            dl = douts[0].contiguous() if douts[0] is not None else torch.zeros(sv["B"], device=sv["feat"].device)
            daux = douts[2].contiguous() if douts[2] is not None else None
            hp.model_bwd(Pd, spec, sv, dl, daux, dc, G)
            return []

        holder = dict(names=names, n_inputs=0, fwd=fwd, bwd=bwd, non_diff=(1,))
        logits, prob, aux = _StageFn.apply(holder, *[P[n] for n in names])
        return logits, prob, aux
