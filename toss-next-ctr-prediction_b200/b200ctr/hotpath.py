"""The DARE + QNN-alpha hot path as a sequence of libb200ctr kernel calls (forward and backward).

This module is the host-side orchestration shared by
  * the drop-in nn.Module mirrors (b200ctr.modules: autograd.Functions call model_fwd / model_bwd), and
  * the fused trainer (b200ctr.trainer: no autograd, flat buffers, row-sparse table gradients).
Parameters are addressed by the reference's state_dict names, so the same dict works for
`CTRModel.state_dict()` tensors and for views into the trainer's flat buffers.

Reference call graph being replaced: src/models/wrapper.py:138-176 -> feature_embed.py:19-48, wrapper.py:106-136,
dare.py:116-162 (+53-70 per encoder layer), qnn_alpha.py:109-130.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

from . import ops
from ._lib import NO_DROP

Tensor = torch.Tensor

# dropout sites (stream ids); one RNG stream per site per step
SITE_EMB, SITE_DARE_U, SITE_QNN_INTER, SITE_MLP0, SITE_LAYER0 = 1, 2, 3, 4, 16


class ModelSpec:
    """What CTRModel.__init__ (wrapper.py:8-104) derives from cfg, as plain attributes."""

    def __init__(self, cfg: dict, num_dim: int, mask_dim: int, cat_cols: List[str]):
        m, s = cfg["model"], cfg["sequence"]
        self.D = int(m["emb_dim"])
        self.L_cfg = int(s["max_len"])
        self.top_k = int(s["top_k"])
        self.tau = float(s["recency_tau"])
        self.pad_id = int(s["pad_id"])
        self.query_mode = s["query_mode"]
        self.query_key = s["query_key"]
        self.transformer = bool(s["transformer_block"])
        tfm = s.get("tfm") or {}
        self.n_layers = int(tfm.get("n_layers", 2)) if self.transformer else 0
        self.n_heads = int(tfm.get("n_heads", 4))
        self.ffn_hidden = int(tfm.get("ffn_hidden", 256))
        self.p_mha = float(tfm.get("mha_dropout", 0.1))
        self.p_ffn = float(tfm.get("ffn_dropout", 0.1))
        self.norm = str(tfm.get("norm", "rms")).lower()
        self.gating = tfm.get("gating", "softmax") if tfm else "softmax"
        self.pos_bias = bool(tfm.get("add_positional_bias", True))
        self.p_dare = float(m["dare_dropout"])
        self.p_emb = float(m.get("embedding_dropout", 0.0))
        self.num_dim, self.mask_dim = int(num_dim), int(mask_dim)
        self.cat_cols = list(cat_cols)
        qa = m["qnn_alpha"]
        self.use_qnn = bool(qa["enabled"])
        self.f_embed = int(qa.get("feature_embed_dim", max(8, self.D // 4)))
        self.aux_weight = float(qa.get("aux_head_weight", 0.0))
        if self.use_qnn:
            self.heads, self.rank, self.proj = int(qa["heads"]), int(qa["rank"]), int(qa["proj_dim"])
            self.mlp_hidden = [int(h) for h in qa["mlp_hidden"]]
            self.p_qnn = float(qa["dropout"])
            self.use_se = bool(qa["use_se"])
            self.se_reduction = int(qa["se_reduction"])
            self.use_residual = bool(qa["use_residual"])
            self.qnn_norm = str(qa["norm"]).lower()
            self.pair_grouping = qa["pair_grouping"]
        self.p_fc = 0.1          # Dropout(0.1) of the FC head is hard-coded in the reference (wrapper.py:98)
        self.F_all = 1 + self.num_dim + self.mask_dim + len(self.cat_cols)
        self.q_idx = self.cat_cols.index(self.query_key) if self.query_mode != "S2" else -1
        self.validate()

    def validate(self):
        from ._lib import B200CtrError
        if self.D != 64:
            raise B200CtrError(f"emb_dim={self.D}: the sm_100a kernels are specialised for emb_dim=64")
        if self.transformer:
            if self.D // self.n_heads != 8 or self.D % self.n_heads:
                raise B200CtrError("encoder kernels are built for head_dim=8 (n_heads=8 at emb_dim=64)")
            if self.norm != "rms":
                raise B200CtrError("only tfm.norm='rms' is built (every shipped cfg uses it)")
        if self.use_qnn:
            if self.qnn_norm != "rms":
                raise B200CtrError("only qnn_alpha.norm='rms' is built")
            if self.pair_grouping != "all":
                raise B200CtrError("only qnn_alpha.pair_grouping='all' is built (every shipped cfg uses it)")


class DropCtl:
    """Per-step dropout descriptors: p = 0 (identity) unless training."""

    def __init__(self, train: bool, seed: int = 0, step: int = 0):
        self.train, self.seed, self.step = train, int(seed), int(step)

    def __call__(self, p: float, site: int):
        if not self.train or p <= 0.0:
            return NO_DROP
        return ops.drop(p, self.seed, self.step * 64 + site)

    def scale(self, p: float) -> float:
        return 1.0 / (1.0 - p) if (self.train and p > 0.0) else 1.0


_REC_CACHE: Dict[tuple, Tensor] = {}


def recency_table(L: int, tau: float, device) -> Tensor:
    """r[p] = log(exp(-(L-1-p)/max(1,tau)) + 1e-8) evaluated with the reference's own fp32 torch ops on the host
    (dare.py:126-127,130), so the table is bit-identical to the reference's; cached per (L, tau, device)."""
    key = (L, float(tau), str(device))
    t = _REC_CACHE.get(key)
    if t is None:
        pos = torch.arange(L)
        decay = torch.exp(-(L - 1 - pos).float() / max(1.0, float(tau)))
        t = torch.log(decay + 1e-8).to(device)
        _REC_CACHE[key] = t
    return t


def _layer_names(i: int) -> Dict[str, str]:
    p = f"dare.layers.{i}."
    return dict(in_w=p + "mha.in_proj_weight", in_b=p + "mha.in_proj_bias", out_w=p + "mha.out_proj.weight",
                out_b=p + "mha.out_proj.bias", n1=p + "norm1.w", f1_w=p + "ffn.0.weight", f1_b=p + "ffn.0.bias",
                f2_w=p + "ffn.3.weight", f2_b=p + "ffn.3.bias", n2=p + "norm2.w", rel=p + "pbias.rel.weight")


# =============================================================================================================
# DARE stage: retrieval -> encoder layers -> gated pooling (+aux head)
# =============================================================================================================
def encoder_layer_fwd(x: Tensor, P: Dict[str, Tensor], i: int, spec: ModelSpec, B: int, K: int, dc: DropCtl) -> tuple:
    """x: (B*K, D).  Returns (y, saved)."""
    n = _layer_names(i)
    D, H = spec.D, spec.n_heads
    relmean = ops.relmean_fwd(P[n["rel"]]) if spec.pos_bias else None
    fused_proj = ops.get_precision() == "bf16" and D == 64 and H == 8
    if fused_proj:        # persistent tcgen05 kernels with resident weights (csrc/proj_fused.cu)
        # packed-half attention: qkv leaves the projection as fp16 (q pre-scaled), the attention kernels stage it with copies
        # and hand o to the out-projection as bf16 (its tensor-core products round it to bf16 in any case)
        h16 = ops.attn_takes_h16(K, H, D)
        qkv = ops.inproj_fwd(x, P[n["in_w"]], P[n["in_b"]], h16=h16)
        o, lse, keep = ops.attn_fwd(qkv, relmean, spec.top_k, B, K, H, D, dc(spec.p_mha, SITE_LAYER0 + 2 * i), save_keep=True,
                                    out_bf16=h16)
        s1, x1, r1 = ops.outproj_norm_fwd(o, P[n["out_w"]], P[n["out_b"]], x, P[n["n1"]])
    else:
        qkv = ops.gemm(x, P[n["in_w"]], trans_b=True, bias=P[n["in_b"]])
        o, lse, keep = ops.attn_fwd(qkv, relmean, spec.top_k, B, K, H, D, dc(spec.p_mha, SITE_LAYER0 + 2 * i), save_keep=True)
        h = ops.gemm(o, P[n["out_w"]], trans_b=True, bias=P[n["out_b"]])
        s1, x1, r1 = ops.add_rmsnorm_fwd(x, h, P[n["n1"]])
    if ops.get_precision() == "bf16" and ops.ffn_fused_ok(spec.ffn_hidden, D):
        # one tcgen05 kernel: the (tokens, ffn_hidden) activation stays on the SM; the backward recomputes it from x1
        x2, s2, r2, fkeep = ops.ffn_fwd(x1, P[n["f1_w"]], P[n["f1_b"]], P[n["f2_w"]], P[n["f2_b"]], P[n["n2"]],
                                        dc(spec.p_ffn, SITE_LAYER0 + 2 * i + 1), save_keep=True)
        pre = fd = None
    else:
        fkeep = None
        pre = torch.empty((x.shape[0], spec.ffn_hidden), device=x.device, dtype=torch.float32)
        fd = ops.gemm(x1, P[n["f1_w"]], trans_b=True, bias=P[n["f1_b"]], act=ops.ACT_GELU, pre_out=pre,
                      dropout=dc(spec.p_ffn, SITE_LAYER0 + 2 * i + 1))
        h2 = ops.gemm(fd, P[n["f2_w"]], trans_b=True, bias=P[n["f2_b"]])
        s2, x2, r2 = ops.add_rmsnorm_fwd(x1, h2, P[n["n2"]])
    return x2, dict(x=x, relmean=relmean, qkv=qkv, o=o, lse=lse, keep=keep, fkeep=fkeep, s1=s1, x1=x1, r1=r1, pre=pre, fd=fd, s2=s2, r2=r2,
                    fused_proj=fused_proj)


def encoder_layer_bwd(dy: Tensor, sv: dict, P: Dict[str, Tensor], i: int, spec: ModelSpec, B: int, K: int, dc: DropCtl,
                      G: Callable[[str], Tensor]) -> Tensor:
    n = _layer_names(i)
    D, H = spec.D, spec.n_heads
    if sv["pre"] is None:         # fused tcgen05 FFN (forward saved only x1, s2, r2)
        dx1 = ops.ffn_bwd(dy, sv["s2"], sv["r2"], P[n["n2"]], sv["x1"], P[n["f1_w"]], P[n["f1_b"]], P[n["f2_w"]],
                          G(n["f1_w"]), G(n["f1_b"]), G(n["f2_w"]), G(n["f2_b"]), G(n["n2"]),
                          dc(spec.p_ffn, SITE_LAYER0 + 2 * i + 1), keep=sv.get("fkeep"))
    else:
        d_s2 = ops.rmsnorm_bwd(dy, sv["s2"], P[n["n2"]], sv["r2"], G(n["n2"]))
        # FFN second linear
        ops.gemm(d_s2, sv["fd"], trans_a=True, out=G(n["f2_w"]))
        ops.colsum(d_s2, out=G(n["f2_b"]))
        # dpre = (d_s2 @ W2) * dropout_mask * gelu'(pre): dropout regenerated from the same stream + element ids
        dpre = ops.gemm(d_s2, P[n["f2_w"]], dropout=dc(spec.p_ffn, SITE_LAYER0 + 2 * i + 1), mask_mode=ops.MASK_GELU_GRAD,
                        mask_src=sv["pre"])
        ops.gemm(dpre, sv["x1"], trans_a=True, out=G(n["f1_w"]))
        ops.colsum(dpre, out=G(n["f1_b"]))
        dx1 = ops.gemm(dpre, P[n["f1_w"]], residual=d_s2)
    fused_proj = sv.get("fused_proj", False)
    if fused_proj:
        # d_o goes to the packed-half attention backward as bf16 (recognised by the dtype of the saved qkv)
        d_s1, d_o = ops.outproj_norm_bwd(dx1, sv["s1"], sv["r1"], P[n["n1"]], sv["o"], P[n["out_w"]], G(n["out_w"]),
                                         G(n["out_b"]), G(n["n1"]), out_bf16=sv["qkv"].dtype == torch.float16)
    else:
        d_s1 = ops.rmsnorm_bwd(dx1, sv["s1"], P[n["n1"]], sv["r1"], G(n["n1"]))
        ops.gemm(d_s1, sv["o"], trans_a=True, out=G(n["out_w"]))
        ops.colsum(d_s1, out=G(n["out_b"]))
        d_o = ops.gemm(d_s1, P[n["out_w"]])
    drelmean = None
    if spec.pos_bias:
        drelmean = torch.zeros_like(sv["relmean"])
    dqkv = ops.attn_bwd(sv["qkv"], sv["relmean"], spec.top_k, sv["o"], sv["lse"], d_o, B, K, H, D, drelmean,
                        dc(spec.p_mha, SITE_LAYER0 + 2 * i), keep=sv.get("keep"), out_bf16=bool(sv.get("fused_proj")))
    if spec.pos_bias:
        G(n["rel"]).copy_(ops.relmean_bwd(drelmean, H))
    if fused_proj:
        return ops.inproj_bwd(dqkv, sv["x"], P[n["in_w"]], d_s1, G(n["in_w"]), G(n["in_b"]))
    ops.gemm(dqkv, sv["x"], trans_a=True, out=G(n["in_w"]))
    ops.colsum(dqkv, out=G(n["in_b"]))
    dx = ops.gemm(dqkv, P[n["in_w"]], residual=d_s1)
    return dx


def dare_fwd(P: Dict[str, Tensor], spec: ModelSpec, seq: Tensor, q: Tensor, u_out: Tensor, ld_u: int, dc: DropCtl):
    """DARE.forward (dare.py:140-162).  u_seq is written at u_out (+b*ld_u).  Returns (aux, saved)."""
    B, L = seq.shape
    K = min(spec.top_k, L)
    rec = recency_table(L, spec.tau, q.device)
    sel, vals, idx, tok = ops.retrieve_fwd(seq, q, P["dare.emb_att.weight"], P["dare.emb_rep.weight"], rec, K, spec.pad_id)
    x = sel.view(B * K, spec.D)
    layers = []
    for i in range(spec.n_layers):
        x, lsv = encoder_layer_fwd(x, P, i, spec, B, K, dc)
        layers.append(lsv)
    x3 = x.view(B, K, spec.D)
    aux, w = ops.pool_fwd(x3, vals, P["dare.aux_head.weight"], P["dare.aux_head.bias"], spec.gating == "relu", u_out, ld_u,
                          dc(spec.p_dare, SITE_DARE_U))
    return aux, dict(q=q, vals=vals, idx=idx, tok=tok, x_enc=x3, layers=layers, w=w, B=B, K=K, u=u_out, ld_u=ld_u)


def dare_bwd(P, spec: ModelSpec, sv: dict, du: Optional[Tensor], ld_du: int, daux: Optional[Tensor], dc: DropCtl,
             G: Callable[[str], Tensor], want_aux_grads: bool = True) -> Tensor:
    """Returns dq (B,D); accumulates table gradients into G('dare.emb_att.weight') / G('dare.emb_rep.weight')."""
    B, K = sv["B"], sv["K"]
    daw = G("dare.aux_head.weight") if (daux is not None and want_aux_grads) else None
    dab = G("dare.aux_head.bias") if (daux is not None and want_aux_grads) else None
    dx3, dvals = ops.pool_bwd(sv["x_enc"], sv["vals"], sv["w"], P["dare.aux_head.weight"], sv["u"], sv["ld_u"], du, ld_du,
                              daux, spec.gating == "relu", daw, dab, dc(spec.p_dare, SITE_DARE_U))
    dx = dx3.view(B * K, spec.D)
    for i in range(spec.n_layers - 1, -1, -1):
        dx = encoder_layer_bwd(dx, sv["layers"][i], P, i, spec, B, K, dc, G)
    dq = ops.retrieve_bwd(sv["tok"], sv["q"], P["dare.emb_att.weight"], dx, dvals, spec.pad_id,
                          G("dare.emb_att.weight"), G("dare.emb_rep.weight"))
    return dq


# =============================================================================================================
# QNN-alpha stage
# =============================================================================================================
def _qnn_mlp_names(P) -> List[int]:
    idx = []
    i = 0
    while f"qnn.mlp.{i}.weight" in P:
        idx.append(i)
        i += 3
    return idx


def qnn_fwd(P: Dict[str, Tensor], spec: ModelSpec, xF: Tensor, dc: DropCtl, B_total: Optional[int] = None,
            colsum_hook: Optional[Callable[[Tensor], Tensor]] = None):
    """QNNAlphaDetailed.forward (qnn_alpha.py:109-130).  xF: (B, F, D) contiguous.  Returns (logits (B,), saved).
    colsum_hook: data-parallel hook that all-reduces the SE squeeze sum over ranks (C4); B_total = global batch."""
    B, F, D = xF.shape
    N = F * D
    H, r, Pd = spec.heads, spec.rank, spec.proj
    C, HP = H * r, H * Pd
    flat = xF.view(B, N)
    # amp: bf16 -- the deep (K = F*D) first-layer contractions read bf16 copies of z (emitted by the norm kernel itself)
    # and of W0 (converted once per step) instead of converting fp32 operands inside the GEMM's load path
    bf16_ops = ops.get_precision() == "bf16" and N % 8 == 0 and B >= 128
    z16 = W0_16 = None
    fm_fused = bf16_ops and ops.fm_fused_ok(F, C, D)
    if bf16_ops:
        # every consumer of z reads the bf16 copy when the fused interaction and the bf16 first-layer products run: the
        # fp32 z (B x F*D) is then never written
        W0_ = P[f"qnn.mlp.{_qnn_mlp_names(P)[0]}.weight"]
        all_bf16 = fm_fused and W0_.shape[1] % 8 == 0 and W0_.is_contiguous() and W0_.shape[0] % 8 == 0
        z, rz, z16 = ops.rowrms_fwd(flat, P["qnn.pre_norm.w"], want_bf16=True, want_f32=not all_bf16)
    else:
        z, rz = ops.rowrms_fwd(flat, P["qnn.pre_norm.w"])
    Ucat = P["qnn.U"].permute(1, 0, 2).reshape(D, C).contiguous()            # (D, H*r): layout plumbing only
    if fm_fused:          # one tcgen05 kernel; the (B*F, C) product z.Ucat stays in TMEM
        A = None
        s, quad = ops.fm_fused_fwd(z16, Ucat, B, F)
    else:
        A = ops.gemm(z.view(B * F, D), Ucat)                                   # (B*F, C)
        s, quad = ops.fm_reduce(A, B, F, C)
    V = P["qnn.V"]
    if ops.headproj_ok(H, r, Pd):
        inter = ops.headproj_fwd(quad, V)
    else:
        inter = torch.empty((B, HP), device=xF.device, dtype=torch.float32)
        for h in range(H):
            ops.gemm(quad[:, h * r:(h + 1) * r], V[h], out=inter[:, h * Pd:(h + 1) * Pd])
    sv = dict(xF=xF, rz=rz, z=z, A=A, s=s, quad=quad, inter=inter, Ucat=Ucat)
    Bt = B_total or B
    if spec.use_se:
        cs = ops.colsum(inter)
        if colsum_hook is not None:
            cs = colsum_hook(cs)
        mean, h1, gate = ops.se_gate_fwd(cs, Bt, P["qnn.se.fc.0.weight"], P["qnn.se.fc.0.bias"], P["qnn.se.fc.2.weight"],
                                         P["qnn.se.fc.2.bias"])
        sv.update(mean=mean, h1=h1, gate=gate)
    else:
        gate = None
    igd = ops.se_apply_fwd(inter, gate, dc(spec.p_qnn, SITE_QNN_INTER))
    mlp = _qnn_mlp_names(P)
    acts = []
    W0 = P[f"qnn.mlp.{mlp[0]}.weight"]
    if bf16_ops and W0.shape[1] % 8 == 0 and W0.is_contiguous():
        W0_16 = ops.to_bf16(W0)
    zin, W0z = (z16, W0_16[:, :N]) if W0_16 is not None else (z, W0[:, :N])
    # the deep (K = F*D) contraction: TMA-fed tcgen05 kernel when both operands are bf16 with K contiguous
    use_kk = W0_16 is not None and ops.gemm_kk_ok(B, W0.shape[0], N)
    # the interaction part of the first layer: bf16 operands (16-byte async copies) when the bf16 weight copy exists --
    # converting fp32 operands inside the GEMM's load path made this 1,152-deep product as slow as the 12,800-deep one
    ig_op, Wi_op = igd, W0[:, N:]
    if W0_16 is not None and igd.is_contiguous() and (W0.shape[1] - N) % 8 == 0 and N % 8 == 0:
        ig_op, Wi_op = ops.to_bf16(igd), W0_16[:, N:]
    if len(mlp) == 1:
        hcur = ops.gemm_kk(zin, W0z) if use_kk else ops.gemm(zin, W0z, trans_b=True)
        logit = ops.gemm(ig_op, Wi_op, trans_b=True, accum=hcur, bias=P[f"qnn.mlp.{mlp[0]}.bias"])
    else:
        hcur = ops.gemm_kk(zin, W0z) if use_kk else ops.gemm(zin, W0z, trans_b=True)
        hcur = ops.gemm(ig_op, Wi_op, trans_b=True, accum=hcur, bias=P[f"qnn.mlp.{mlp[0]}.bias"], act=ops.ACT_RELU,
                        dropout=dc(spec.p_qnn, SITE_MLP0), out=hcur)
        acts.append(hcur)
        for li in range(1, len(mlp) - 1):
            hcur = ops.gemm(hcur, P[f"qnn.mlp.{mlp[li]}.weight"], trans_b=True, bias=P[f"qnn.mlp.{mlp[li]}.bias"],
                            act=ops.ACT_RELU, dropout=dc(spec.p_qnn, SITE_MLP0 + li))
            acts.append(hcur)
        Wl = P[f"qnn.mlp.{mlp[-1]}.weight"]
        if Wl.shape[0] == 1 and ops.lin1_ok(Wl.shape[1]) and hcur.is_contiguous():      # Linear(K -> 1): row dot products
            logit = ops.lin1_fwd(hcur, Wl, P[f"qnn.mlp.{mlp[-1]}.bias"])
        else:
            logit = ops.gemm(hcur, Wl, trans_b=True, bias=P[f"qnn.mlp.{mlp[-1]}.bias"])
    sv.update(igd=igd, acts=acts, mlp=mlp, B_total=Bt, z16=z16 if W0_16 is not None else None, W0_16=W0_16,
              z16_fm=z16 if fm_fused else None)
    return logit.view(B), sv


def qnn_bwd(P, spec: ModelSpec, sv: dict, dlogit: Tensor, dc: DropCtl, G: Callable[[str], Tensor],
            colsum_hook: Optional[Callable[[Tensor], Tensor]] = None) -> Tensor:
    """Returns dxF (B,F,D)."""
    xF = sv["xF"]
    B, F, D = xF.shape
    N = F * D
    H, r, Pd = spec.heads, spec.rank, spec.proj
    C, HP = H * r, H * Pd
    mlp, acts = sv["mlp"], sv["acts"]
    z, igd = sv["z"], sv["igd"]
    ms = dc.scale(spec.p_qnn)
    dl = dlogit.view(B, 1)
    last = mlp[-1]
    if len(mlp) == 1:
        dh = dl
    else:
        Wl = P[f"qnn.mlp.{last}.weight"]
        if Wl.shape[0] == 1 and ops.lin1_ok(Wl.shape[1]) and acts[-1].is_contiguous() and dl.is_contiguous():
            dh = ops.lin1_bwd(dl, acts[-1], Wl, ms, G(f"qnn.mlp.{last}.weight"), G(f"qnn.mlp.{last}.bias"))
        else:
            ops.gemm(dl, acts[-1], trans_a=True, out=G(f"qnn.mlp.{last}.weight"))
            ops.colsum(dl, out=G(f"qnn.mlp.{last}.bias"))
            dh = ops.gemm(dl, Wl, mask_mode=ops.MASK_NONZERO, mask_src=acts[-1], mask_scale=ms)
        for li in range(len(mlp) - 2, 0, -1):
            ops.gemm(dh, acts[li - 1], trans_a=True, out=G(f"qnn.mlp.{mlp[li]}.weight"))
            ops.colsum(dh, out=G(f"qnn.mlp.{mlp[li]}.bias"))
            dh = ops.gemm(dh, P[f"qnn.mlp.{mlp[li]}.weight"], mask_mode=ops.MASK_NONZERO, mask_src=acts[li - 1], mask_scale=ms)
    # first layer: input = cat([z, igd])
    W0 = P[f"qnn.mlp.{mlp[0]}.weight"]
    gW0 = G(f"qnn.mlp.{mlp[0]}.weight")
    z16, W0_16 = sv.get("z16"), sv.get("W0_16")
    dh16 = ops.to_bf16(dh) if (z16 is not None and dh.is_contiguous() and dh.shape[1] % 8 == 0) else None
    H0 = W0.shape[0]
    use_kk = dh16 is not None and ops.gemm_kk_ok(H0, N, B) and ops.gemm_kk_ok(B, N, H0) and B % 2 == 0 and N % 2 == 0
    HPc = W0.shape[1] - N
    kk_inter = use_kk and ops.gemm_kk_ok(H0, HPc, B) and ops.gemm_kk_ok(B, HPc, H0)
    if use_kk:
        # K-major forms for the TMA kernel: dW0z = dh^T . z as (dh^T)(z^T)^T, dz = dh . W0z as dh (W0z^T)^T
        dh16t = ops.to_bf16_t(dh)
        ops.gemm_kk(dh16t, ops.transpose_bf16(z16), out=gW0[:, :N])
    elif dh16 is not None:
        ops.gemm(dh16, z16, trans_a=True, out=gW0[:, :N])
    else:
        ops.gemm(dh, z, trans_a=True, out=gW0[:, :N])
    if kk_inter:
        ops.gemm_kk(dh16t, ops.to_bf16_t(igd), out=gW0[:, N:])
    else:
        ops.gemm(dh, igd, trans_a=True, out=gW0[:, N:])
    ops.colsum(dh, out=G(f"qnn.mlp.{mlp[0]}.bias"))
    digd = ops.gemm_kk(dh16, ops.to_bf16_t(W0[:, N:])) if kk_inter else ops.gemm(dh, W0[:, N:])
    inter = sv["inter"]
    dcq = dc(spec.p_qnn, SITE_QNN_INTER)
    if spec.use_se:
        _, gprod = ops.se_apply_bwd(digd, inter, sv["gate"], None, sv["B_total"], False, True, dcq)
        dgate = ops.colsum(gprod)
        if colsum_hook is not None:
            dgate = colsum_hook(dgate)
        dw1, db1, dw2, db2, dmean = ops.se_gate_bwd(dgate, sv["mean"], sv["h1"], sv["gate"], P["qnn.se.fc.0.weight"],
                                                    P["qnn.se.fc.2.weight"])
        G("qnn.se.fc.0.weight").copy_(dw1); G("qnn.se.fc.0.bias").copy_(db1)
        G("qnn.se.fc.2.weight").copy_(dw2); G("qnn.se.fc.2.bias").copy_(db2)
        dinter, _ = ops.se_apply_bwd(digd, inter, sv["gate"], dmean, sv["B_total"], True, False, dcq)
    else:
        dinter, _ = ops.se_apply_bwd(digd, inter, None, None, sv["B_total"], True, False, dcq)
    quad, V = sv["quad"], P["qnn.V"]
    gV = G("qnn.V")
    if ops.headproj_ok(H, r, Pd):
        dquad = ops.headproj_bwd(quad, dinter, V, gV)
    else:
        dquad = torch.empty((B, C), device=xF.device, dtype=torch.float32)
        for h in range(H):
            ops.gemm(quad[:, h * r:(h + 1) * r], dinter[:, h * Pd:(h + 1) * Pd], trans_a=True, out=gV[h])
            ops.gemm(dinter[:, h * Pd:(h + 1) * Pd], V[h], trans_b=True, out=dquad[:, h * r:(h + 1) * r])
    dz_mlp = None
    if spec.use_residual:
        if use_kk:
            dz_mlp = ops.gemm_kk(dh16, ops.to_bf16_t(W0[:, :N]))                               # (B, N)
        else:
            dz_mlp = ops.gemm(dh16, W0_16[:, :N]) if dh16 is not None else ops.gemm(dh, W0[:, :N])
    if sv.get("z16_fm") is not None:
        # recompute z.Ucat in TMEM, dA as a bf16 operand tile, dz (+ the MLP branch) and dUcat in one kernel
        dz, dUcat = ops.fm_fused_bwd(sv["z16_fm"], sv["Ucat"], sv["s"], dquad, dz_mlp, B, F, dz_out=dz_mlp)
        G("qnn.U").copy_(dUcat.view(D, H, r).permute(1, 0, 2))
    else:
        dA = ops.fm_grad_(sv["A"], sv["s"], dquad, B, F, C)                   # in place over the saved A
        z3 = z.view(B * F, D)
        dUcat = ops.gemm(z3, dA, trans_a=True)                                # (D, C)
        G("qnn.U").copy_(dUcat.view(D, H, r).permute(1, 0, 2))
        if spec.use_residual:
            dz = dz_mlp
            ops.gemm(dA, sv["Ucat"], trans_b=True, residual=dz.view(B * F, D), out=dz.view(B * F, D))
        else:
            dz = ops.gemm(dA, sv["Ucat"], trans_b=True).view(B, N)
    dflat, dpw = ops.rowrms_bwd(dz, xF.view(B, N), P["qnn.pre_norm.w"], sv["rz"])
    G("qnn.pre_norm.w").copy_(dpw)
    return dflat.view(B, F, D)


# =============================================================================================================
# whole model  (wrapper.py:138-176)
# =============================================================================================================
def _cat_tables(P, spec: ModelSpec, cache: dict) -> ops.CatTables:
    tabs = [P[f"cat_embs.{c}.weight"] for c in spec.cat_cols]
    projs = [P[f"cat_proj.{c}.weight"] for c in spec.cat_cols]
    ct = cache.get("ct")
    if ct is None or not ct.matches(tabs, projs):
        ct = ops.CatTables(tabs, projs)
        cache["ct"] = ct
    return ct


def model_fwd(P: Dict[str, Tensor], spec: ModelSpec, batch: Dict[str, Tensor], dc: DropCtl, cache: dict,
              B_total: Optional[int] = None, colsum_hook=None):
    """Returns (logits, aux, saved).  batch tensors must already be on the device: X_num/X_mask float32 (B,Fn/Fm),
    X_cat / seq int64 or int32."""
    X_num, X_mask, X_cat, seq = batch["X_num"], batch["X_mask"], batch["X_cat"], batch["seq"]
    dev = seq.device
    B = seq.shape[0]
    D, Fn, Fm, Fc = spec.D, spec.num_dim, spec.mask_dim, len(spec.cat_cols)
    nparts = (Fn > 0) + (Fm > 0) + 1
    ctx = torch.empty((B, nparts * D), device=dev, dtype=torch.float32)
    sv: dict = dict(B=B, X_num=X_num, X_mask=X_mask, X_cat=X_cat, ctx=ctx)
    if spec.use_qnn:
        feat = torch.empty((B, spec.F_all, D), device=dev, dtype=torch.float32)     # xF
        ld = spec.F_all * D
        u_out, ld_u = feat, ld                                                        # feature 0
        off_num, off_mask, off_cat = 1, 1 + Fn, 1 + Fn + Fm
        num_dst = feat[:, off_num:] if Fn > 0 else None
        mask_dst = feat[:, off_mask:] if Fm > 0 else None
        cat_dst = feat[:, off_cat:]
    else:
        width = D * (1 + (Fn > 0) + (Fm > 0) + Fc)
        feat = torch.empty((B, width), device=dev, dtype=torch.float32)              # [u | num_mean | mask_mean | cats]
        ld = width
        u_out, ld_u = feat, ld
        num_dst = mask_dst = None
        cat_dst = feat[:, D * (1 + (Fn > 0) + (Fm > 0)):]
    o = 0
    if Fn > 0:
        Pn, Qn = ops.numbin_fold(P["num_embed.weight"], P["num_embed.bias"], P["num_embed.out_proj.weight"])
        ops.numbin_fwd(X_num, Pn, Qn, num_dst, ld, ctx[:, o:], nparts * D)
        o += D
    if Fm > 0:
        Pm, Qm = ops.numbin_fold(P["mask_embed.weight"], None, P["mask_embed.out_proj.weight"])
        ops.numbin_fwd(X_mask, Pm, Qm, mask_dst, ld, ctx[:, o:], nparts * D)
        o += D
    ct = _cat_tables(P, spec, cache)
    s1 = torch.empty((B, D), device=dev, dtype=torch.float32) if spec.q_idx >= 0 else None
    # embedding_dropout acts on cat_stack, which only the QNN head consumes; the FC head concatenates the un-dropped
    # cat_embs (wrapper.py:149-150 vs 167-171)
    emb_drop = dc(spec.p_emb, SITE_EMB) if spec.use_qnn else NO_DROP
    ops.cat_fwd(X_cat, ct, cat_dst, ld, ctx[:, o:], nparts * D, spec.q_idx, s1, D, emb_drop)
    if not spec.use_qnn:
        # the FC head consumes the feature means: copy them next to u (device-to-device plumbing)
        k = 1
        if Fn > 0:
            feat[:, k * D:(k + 1) * D].copy_(ctx[:, 0:D]); k += 1
        if Fm > 0:
            feat[:, k * D:(k + 1) * D].copy_(ctx[:, (k - 1) * D:k * D]); k += 1
    if spec.query_mode == "S1":
        q = s1
        s2 = None
    else:
        s2 = ops.gemm(ctx, P["ctx_mlp.0.weight"], trans_b=True, bias=P["ctx_mlp.0.bias"], act=ops.ACT_RELU)
        q = s2 if spec.query_mode == "S2" else ops.axpby(s1, s2, 0.5, 0.5)
    sv.update(s2=s2, ct=ct, cat_rows_acc=cache.get("cat_rows_acc"), cat_keep_rows=cache.get("cat_keep_rows"))
    if cache.get("cat_defer_slot") is not None:        # data parallel: the row-lazy tables' slots are reduced by the trainer
        nd = cache["cat_defer_n"]
        sv["cat_defer"] = dict(slot=cache["cat_defer_slot"], n=nd, rows=torch.empty((B, nd, D), device=dev, dtype=torch.float32),
                               keys=torch.empty((B, nd), device=dev, dtype=torch.int32))
    aux, dsv = dare_fwd(P, spec, seq, q, u_out, ld_u, dc)
    sv["dare"] = dsv
    if spec.use_qnn:
        logits, qsv = qnn_fwd(P, spec, feat, dc, B_total, colsum_hook)
        sv["qnn"] = qsv
    else:
        fh = ops.gemm(feat, P["fc.0.weight"], trans_b=True, bias=P["fc.0.bias"], act=ops.ACT_RELU,
                      dropout=dc(spec.p_fc, SITE_MLP0))
        logits = ops.gemm(fh, P["fc.3.weight"], trans_b=True, bias=P["fc.3.bias"]).view(B)
        sv.update(fh=fh)
    sv.update(feat=feat, ld=ld)
    return logits, aux, sv


def model_bwd(P: Dict[str, Tensor], spec: ModelSpec, sv: dict, dlogits: Tensor, daux: Optional[Tensor], dc: DropCtl,
              G: Callable[[str], Tensor], colsum_hook=None, after_head: Optional[Callable[[], None]] = None):
    """Backward of model_fwd.  G(name) returns the (zero-initialised) gradient tensor of parameter `name`.
    after_head: called once the head's (QNN / FC) parameter gradients are final -- the data-parallel trainer starts their
    all-reduce there, under the encoder / embedding backward."""
    B, D = sv["B"], spec.D
    Fn, Fm, Fc = spec.num_dim, spec.mask_dim, len(spec.cat_cols)
    nparts = (Fn > 0) + (Fm > 0) + 1
    feat, ld = sv["feat"], sv["ld"]
    dev = feat.device
    if spec.use_qnn:
        dfeat = qnn_bwd(P, spec, sv["qnn"], dlogits, dc, G, colsum_hook)          # (B,F,D)
        du, ld_d = dfeat, ld
        dnum = dfeat[:, 1:] if Fn > 0 else None
        dmask = dfeat[:, 1 + Fn:] if Fm > 0 else None
        dcat = dfeat[:, 1 + Fn + Fm:]
        dmean_fc = None
    else:
        dl = dlogits.view(B, 1)
        fh = sv["fh"]
        ops.gemm(dl, fh, trans_a=True, out=G("fc.3.weight"))
        ops.colsum(dl, out=G("fc.3.bias"))
        dfh = ops.gemm(dl, P["fc.3.weight"], mask_mode=ops.MASK_NONZERO, mask_src=fh, mask_scale=dc.scale(spec.p_fc))
        ops.gemm(dfh, feat, trans_a=True, out=G("fc.0.weight"))
        ops.colsum(dfh, out=G("fc.0.bias"))
        dfeat = ops.gemm(dfh, P["fc.0.weight"])                                     # (B, width)
        du, ld_d = dfeat, ld
        dnum = dmask = None
        dcat = dfeat[:, D * (1 + (Fn > 0) + (Fm > 0)):]
        dmean_fc = dfeat[:, D:]                                                     # [dnum_mean | dmask_mean] slices
    if after_head is not None:
        after_head()
    dq = dare_bwd(P, spec, sv["dare"], du, ld_d, daux, dc, G, want_aux_grads=spec.aux_weight > 0)
    # query path
    dctx = None
    dq_cat = None
    if spec.query_mode == "S1":
        dq_cat = dq
    else:
        scale = 1.0 if spec.query_mode == "S2" else 0.5
        ds2 = ops.relu_bwd(dq, sv["s2"], scale)                                   # ReLU of ctx_mlp
        ops.gemm(ds2, sv["ctx"], trans_a=True, out=G("ctx_mlp.0.weight"))
        ops.colsum(ds2, out=G("ctx_mlp.0.bias"))
        dctx = ops.gemm(ds2, P["ctx_mlp.0.weight"])                                  # (B, nparts*D)
        if spec.query_mode != "S2":
            dq_cat = ops.axpby(dq, None, 0.5, 0.0)
    if not spec.use_qnn:
        # FC head also back-propagates into the feature means; merge with the ctx gradient
        k = (Fn > 0) + (Fm > 0)
        if k > 0:
            if dctx is None:
                dctx = torch.zeros((B, nparts * D), device=dev, dtype=torch.float32)
            ops.add2d_(dctx[:, :k * D], dmean_fc[:, :k * D])
    o = 0
    ldc = nparts * D
    if Fn > 0:
        dmean = dctx[:, o:] if dctx is not None else None
        dP, dQ = ops.numbin_bwd(sv["X_num"], dnum, ld_d if dnum is not None else 0, dmean, ldc, D)
        dw, db, dwo = ops.numbin_unfold(dP, dQ, P["num_embed.weight"], P["num_embed.bias"], P["num_embed.out_proj.weight"])
        G("num_embed.weight").copy_(dw); G("num_embed.bias").copy_(db); G("num_embed.out_proj.weight").copy_(dwo)
        o += D
    if Fm > 0:
        dmean = dctx[:, o:] if dctx is not None else None
        dP, dQ = ops.numbin_bwd(sv["X_mask"], dmask, ld_d if dmask is not None else 0, dmean, ldc, D)
        dw, _, dwo = ops.numbin_unfold(dP, dQ, P["mask_embed.weight"], None, P["mask_embed.out_proj.weight"])
        G("mask_embed.weight").copy_(dw); G("mask_embed.out_proj.weight").copy_(dwo)
        o += D
    dmean_cat = dctx[:, o:] if dctx is not None else None
    ops.cat_bwd(sv["X_cat"], sv["ct"], dcat, ld_d, dmean_cat, ldc, spec.q_idx if dq_cat is not None else -1, dq_cat,
                [G(f"cat_embs.{c}.weight") for c in spec.cat_cols], [G(f"cat_proj.{c}.weight") for c in spec.cat_cols],
                D, dc(spec.p_emb, SITE_EMB) if spec.use_qnn else NO_DROP, rows_acc=sv.get("cat_rows_acc"),
                keep_rows=sv.get("cat_keep_rows"), defer=sv.get("cat_defer"))
