"""CPU oracle for the DARE + QNN-alpha hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* of the reference's algorithm for the
training / inference hot path (forward, hand-derived backward, loss, clipping,
AdamW, EMA).  It is written with torch CPU tensor ops purely as a threaded
array library (no nn.Module, no autograd): every gradient below is an explicit
formula, which makes this file the executable spec the CUDA kernels are built
to.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it; the product path (b200ctr) never does.

Parity pinning: `oracle/make_golden.py` imports the *real* reference modules
from /root/reference (possible only in the build container), runs them on
seeded inputs and stores inputs + outputs + autograd gradients + post-step
parameters under `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks
every function here against those fixtures, so the oracle is pinned to the
reference implementation itself (there are no golden vectors upstream:
SURVEY.md section 8c).

Reference anchors (file:line relative to /root/reference):
  numbin_*            src/models/feature_embed.py:19-27, 42-48
  cat_*               src/models/wrapper.py:106-112
  context/query       src/models/wrapper.py:114-136
  recency_table,
  retrieve_*          src/models/dare.py:116-138
  encoder_layer_*     src/models/dare.py:29-37, 53-70 (+ torch MHA semantics)
  pool_*              src/models/dare.py:151-161
  qnn_*               src/models/qnn_alpha.py:24-26, 86-97, 109-130
  fc head             src/models/wrapper.py:95-100, 167-173
  bce_wll             src/train.py:71-90
  cosine_warmup_lr    src/utils/sched.py:3-11
  clip / adamw        src/train.py:138, 193-195 (torch.optim.AdamW, clip_grad_norm_)
  ema_*               src/utils/ema.py:72-131
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
PAD_SCORE = -1e9


# --------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------
def rmsnorm_fwd(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tuple[Tensor, Tensor]:
    """y = w * x * rsqrt(mean(x^2) + eps)   (dare.py:6-13, qnn_alpha.py:6-12)."""
    r = torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps)
    return w * x * r, r


def rmsnorm_bwd(dy: Tensor, x: Tensor, w: Tensor, r: Tensor) -> Tuple[Tensor, Tensor]:
    g = dy * w
    dx = r * g - x * (r ** 3) * (g * x).mean(dim=-1, keepdim=True)
    dw = (dy * x * r).reshape(-1, x.shape[-1]).sum(0)
    return dx, dw


def gelu_fwd(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def gelu_grad(x: Tensor) -> Tensor:
    cdf = 0.5 * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))
    pdf = torch.exp(-0.5 * x * x) * (1.0 / math.sqrt(2.0 * math.pi))
    return cdf + x * pdf


def softplus(x: Tensor) -> Tensor:
    # torch.nn.functional.softplus(beta=1, threshold=20)
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


# --------------------------------------------------------------------------
# numeric / binary feature embedding   (feature_embed.py:19-27, 42-48)
# --------------------------------------------------------------------------
def numbin_fwd(x: Tensor, w: Tensor, b: Optional[Tensor], w_out: Tensor) -> Tensor:
    """out[b,f,:] = (x[b,f]*w[f,:] + b[f,:]) @ w_out^T ; w_out is (D, fe)."""
    e = x.unsqueeze(-1) * w
    if b is not None:
        e = e + b
    return e @ w_out.t()


def numbin_bwd(x: Tensor, w: Tensor, b: Optional[Tensor], w_out: Tensor, dout: Tensor):
    """Returns (dw, db|None, dw_out).  x is an input: no dx."""
    e = x.unsqueeze(-1) * w
    if b is not None:
        e = e + b
    de = dout @ w_out                                   # (B,F,fe)
    dw_out = torch.einsum("bfd,bfe->de", dout, e)       # (D,fe)
    dw = (x.unsqueeze(-1) * de).sum(0)
    db = de.sum(0) if b is not None else None
    return dw, db, dw_out


# --------------------------------------------------------------------------
# hashed-categorical gather + projection   (wrapper.py:106-112)
# --------------------------------------------------------------------------
def cat_fwd(x_cat: Tensor, tables: List[Tensor], projs: List[Tensor]) -> Tensor:
    """(B,Fc) int -> (B,Fc,D): E_c[x_cat[:,c]] @ W_c^T, W_c is (D, d_c)."""
    outs = []
    for c, (tab, pw) in enumerate(zip(tables, projs)):
        outs.append(tab[x_cat[:, c]] @ pw.t())
    return torch.stack(outs, dim=1)


def cat_bwd(x_cat: Tensor, tables: List[Tensor], projs: List[Tensor], dout: Tensor):
    dtabs, dprojs = [], []
    for c, (tab, pw) in enumerate(zip(tables, projs)):
        rows = tab[x_cat[:, c]]
        d = dout[:, c, :]
        dprojs.append(d.t() @ rows)                      # (D, d_c)
        drows = d @ pw                                   # (B, d_c)
        dt = torch.zeros_like(tab)
        dt.index_add_(0, x_cat[:, c], drows)
        dtabs.append(dt)
    return dtabs, dprojs


# --------------------------------------------------------------------------
# DARE retrieval   (dare.py:116-138)
# --------------------------------------------------------------------------
def recency_table(L: int, tau: float, dtype=torch.float32) -> Tensor:
    """r[p] = log(exp(-(L-1-p)/max(1,tau)) + 1e-8), evaluated exactly as the
    reference does (fp32 torch ops), so the table is bit-identical."""
    pos = torch.arange(L)
    decay = torch.exp(-(L - 1 - pos).float() / max(1.0, float(tau)))
    return torch.log(decay + 1e-8).to(dtype)


def canonical_topk(scores: Tensor, K: int) -> Tuple[Tensor, Tensor]:
    """Top-K with the canonical tie rule (score desc, position asc).
    torch.topk's tie order is unspecified (SURVEY hard part 1); a stable
    descending sort gives the canonical one."""
    vals, idx = torch.sort(scores, dim=1, descending=True, stable=True)
    return vals[:, :K].contiguous(), idx[:, :K].contiguous()


def retrieve_fwd(seq: Tensor, q: Tensor, emb_att: Tensor, emb_rep: Tensor,
                 rec: Tensor, top_k: int, pad_id: int = 0):
    """seq (B,L) int64, q (B,D) -> sel_rep (B,K,D), vals (B,K), idx (B,K) int64 positions."""
    B, L = seq.shape
    att = emb_att[seq]                                    # (B,L,D)
    scores = (att * q.unsqueeze(1)).sum(-1) + rec.to(att.dtype).unsqueeze(0)
    scores = scores.masked_fill(seq == pad_id, PAD_SCORE)
    K = min(top_k, L)
    vals, idx = canonical_topk(scores, K)
    tok = torch.gather(seq, 1, idx)
    sel_rep = emb_rep[tok]                                # pad rows are zero rows
    return sel_rep, vals, idx


def retrieve_bwd(seq: Tensor, q: Tensor, emb_att: Tensor, emb_rep: Tensor, idx: Tensor,
                 dsel_rep: Tensor, dvals: Tensor, pad_id: int = 0):
    """Returns dq (B,D), demb_att, demb_rep (dense, pad row zero)."""
    tok = torch.gather(seq, 1, idx)                       # (B,K)
    valid = (tok != pad_id).to(q.dtype)
    dv = dvals * valid                                    # masked_fill blocks grad at pads
    att_sel = emb_att[tok]                                # (B,K,D)
    dq = (dv.unsqueeze(-1) * att_sel).sum(1)
    datt = torch.zeros_like(emb_att)
    datt.index_add_(0, tok.reshape(-1), (dv.unsqueeze(-1) * q.unsqueeze(1)).reshape(-1, q.shape[1]))
    drep = torch.zeros_like(emb_rep)
    drep.index_add_(0, tok.reshape(-1), (dsel_rep * valid.unsqueeze(-1)).reshape(-1, q.shape[1]))
    datt[pad_id].zero_()
    drep[pad_id].zero_()
    return dq, datt, drep


# --------------------------------------------------------------------------
# DARE encoder layer  (dare.py:29-37, 53-70; torch.nn.MultiheadAttention math)
# --------------------------------------------------------------------------
def toeplitz_bias(rel_w: Tensor, K: int, max_len: int) -> Tensor:
    """bias[i,j] = mean_h rel_w[clamp(j-i,-max_len,max_len)+max_len, h]   (dare.py:29-37,57-60)."""
    i = torch.arange(K).unsqueeze(1)
    j = torch.arange(K).unsqueeze(0)
    d = (j - i).clamp(-max_len, max_len) + max_len
    return rel_w.mean(dim=1)[d]


def encoder_layer_fwd(x: Tensor, p: Dict[str, Tensor], n_heads: int, max_len: int,
                      attn_mask_drop: Optional[Tensor] = None, ffn_mask_drop: Optional[Tensor] = None):
    """x (B,K,D).  p keys: in_w (3D,D) in_b (3D) out_w (D,D) out_b (D) n1 (D)
    f1_w (Hf,D) f1_b f2_w (D,Hf) f2_b n2 (D) rel (2*max_len+1, H).
    *_mask_drop: optional pre-scaled dropout keep masks (mask/(1-p)), None = no dropout."""
    B, K, D = x.shape
    H = n_heads
    hd = D // H
    bias = toeplitz_bias(p["rel"], K, max_len) if "rel" in p and p["rel"] is not None else None
    qkv = x @ p["in_w"].t() + p["in_b"]
    q, k, v = qkv.split(D, dim=-1)
    qh = q.reshape(B, K, H, hd).permute(0, 2, 1, 3) * (1.0 / math.sqrt(hd))
    kh = k.reshape(B, K, H, hd).permute(0, 2, 1, 3)
    vh = v.reshape(B, K, H, hd).permute(0, 2, 1, 3)
    S = qh @ kh.transpose(-1, -2)
    if bias is not None:
        S = S + bias
    P = torch.softmax(S, dim=-1)
    Pd = P * attn_mask_drop if attn_mask_drop is not None else P
    o = (Pd @ vh).permute(0, 2, 1, 3).reshape(B, K, D)
    h = o @ p["out_w"].t() + p["out_b"]
    r1_in = x + h
    x1, r1 = rmsnorm_fwd(r1_in, p["n1"])
    pre = x1 @ p["f1_w"].t() + p["f1_b"]
    f = gelu_fwd(pre)
    fd = f * ffn_mask_drop if ffn_mask_drop is not None else f
    h2 = fd @ p["f2_w"].t() + p["f2_b"]
    r2_in = x1 + h2
    x2, r2 = rmsnorm_fwd(r2_in, p["n2"])
    cache = dict(x=x, qh=qh, kh=kh, vh=vh, P=P, Pd=Pd, o=o, r1_in=r1_in, r1=r1, x1=x1, pre=pre,
                 fd=fd, r2_in=r2_in, r2=r2, amd=attn_mask_drop, fmd=ffn_mask_drop, K=K, max_len=max_len)
    return x2, cache


def encoder_layer_bwd(dy: Tensor, c: dict, p: Dict[str, Tensor], n_heads: int):
    B, K, D = c["x"].shape
    H = n_heads
    hd = D // H
    g: Dict[str, Tensor] = {}
    d_r2in, g["n2"] = rmsnorm_bwd(dy, c["r2_in"], p["n2"], c["r2"])
    dx1 = d_r2in.clone()
    dh2 = d_r2in
    g["f2_w"] = dh2.reshape(-1, D).t() @ c["fd"].reshape(-1, c["fd"].shape[-1])
    g["f2_b"] = dh2.reshape(-1, D).sum(0)
    dfd = dh2 @ p["f2_w"]
    df = dfd * c["fmd"] if c["fmd"] is not None else dfd
    dpre = df * gelu_grad(c["pre"])
    g["f1_w"] = dpre.reshape(-1, dpre.shape[-1]).t() @ c["x1"].reshape(-1, D)
    g["f1_b"] = dpre.reshape(-1, dpre.shape[-1]).sum(0)
    dx1 = dx1 + dpre @ p["f1_w"]
    d_r1in, g["n1"] = rmsnorm_bwd(dx1, c["r1_in"], p["n1"], c["r1"])
    dx = d_r1in.clone()
    dh = d_r1in
    g["out_w"] = dh.reshape(-1, D).t() @ c["o"].reshape(-1, D)
    g["out_b"] = dh.reshape(-1, D).sum(0)
    do = (dh @ p["out_w"]).reshape(B, K, H, hd).permute(0, 2, 1, 3)      # (B,H,K,hd)
    dPd = do @ c["vh"].transpose(-1, -2)                                  # (B,H,K,K)
    dvh = c["Pd"].transpose(-1, -2) @ do
    dP = dPd * c["amd"] if c["amd"] is not None else dPd
    P = c["P"]
    dS = P * (dP - (dP * P).sum(-1, keepdim=True))
    dqh = (dS @ c["kh"]) * (1.0 / math.sqrt(hd))
    dkh = dS.transpose(-1, -2) @ c["qh"]          # qh already carries the 1/sqrt(hd) scale
    if "rel" in p and p["rel"] is not None:
        dbias = dS.sum(dim=(0, 1))                                        # (K,K)
        max_len = c["max_len"]
        i = torch.arange(K).unsqueeze(1)
        j = torch.arange(K).unsqueeze(0)
        d = ((j - i).clamp(-max_len, max_len) + max_len).reshape(-1)
        drm = torch.zeros(p["rel"].shape[0], dtype=dy.dtype)
        drm.index_add_(0, d, dbias.reshape(-1))
        g["rel"] = (drm / H).unsqueeze(1).expand(-1, H).contiguous()
    dq = dqh.permute(0, 2, 1, 3).reshape(B, K, D)
    dk = dkh.permute(0, 2, 1, 3).reshape(B, K, D)
    dv = dvh.permute(0, 2, 1, 3).reshape(B, K, D)
    dqkv = torch.cat([dq, dk, dv], dim=-1)
    g["in_w"] = dqkv.reshape(-1, 3 * D).t() @ c["x"].reshape(-1, D)
    g["in_b"] = dqkv.reshape(-1, 3 * D).sum(0)
    dx = dx + dqkv @ p["in_w"]
    return dx, g


# --------------------------------------------------------------------------
# gated pooling + aux head   (dare.py:151-161)
# --------------------------------------------------------------------------
def pool_fwd(x: Tensor, vals: Tensor, aux_w: Tensor, aux_b: Tensor, gating: str = "softmax",
             drop_mask: Optional[Tensor] = None):
    if gating == "relu":
        wr = torch.relu(vals)
        w = wr / (wr.sum(dim=1, keepdim=True) + 1e-12)
    else:
        w = torch.softmax(vals, dim=1)
    u0 = (x * w.unsqueeze(-1)).sum(1)
    u = u0 * drop_mask if drop_mask is not None else u0
    aux = (u @ aux_w.t()).squeeze(1) + aux_b
    return u, aux, w


def pool_bwd(x: Tensor, vals: Tensor, w: Tensor, aux_w: Tensor, du_ext: Tensor, daux: Tensor, u: Tensor,
             gating: str = "softmax", drop_mask: Optional[Tensor] = None):
    """du_ext: gradient reaching u_seq from the head; daux (B,).  Returns dx, dvals, daux_w, daux_b."""
    du = du_ext + daux.unsqueeze(1) * aux_w                     # (B,D)
    daux_w = (daux.unsqueeze(1) * u).sum(0, keepdim=True)
    daux_b = daux.sum().reshape(1)
    du0 = du * drop_mask if drop_mask is not None else du
    dx = w.unsqueeze(-1) * du0.unsqueeze(1)
    dw = (x * du0.unsqueeze(1)).sum(-1)                         # (B,K)
    if gating == "relu":
        wr = torch.relu(vals)
        s = wr.sum(dim=1, keepdim=True) + 1e-12
        dwr = dw / s - (dw * wr).sum(1, keepdim=True) / (s * s)
        dvals = dwr * (vals > 0).to(vals.dtype)
    else:
        dvals = w * (dw - (dw * w).sum(1, keepdim=True))
    return dx, dvals, daux_w, daux_b


# --------------------------------------------------------------------------
# QNN-alpha head   (qnn_alpha.py:24-26, 86-97, 109-130)
# --------------------------------------------------------------------------
def qnn_fwd(xF: Tensor, p: Dict[str, Tensor], use_se: bool = True, use_residual: bool = True,
            drop_inter: Optional[Tensor] = None, drop_mlp: Optional[List[Tensor]] = None):
    """xF (B,F,D).  p: pre_w (F*D) U (H,D,r) V (H,r,P) se1_w se1_b se2_w se2_b,
    mlp_w: list, mlp_b: list (last one -> 1)."""
    B, F, D = xF.shape
    H, _, r = p["U"].shape
    flat = xF.reshape(B, F * D)
    z, rz = rmsnorm_fwd(flat, p["pre_w"])
    z3 = z.reshape(B, F, D)
    A = torch.einsum("bfd,hdr->bhfr", z3, p["U"])          # (B,H,F,r)
    s = A.sum(2)                                           # (B,H,r)
    quad = s * s - (A * A).sum(2)
    inter = torch.einsum("bhr,hrp->bhp", quad, p["V"]).reshape(B, -1)    # (B,H*P)
    cache = dict(xF=xF, flat=flat, rz=rz, z=z, s=s, quad=quad, inter=inter)
    if use_se:
        m = inter.mean(0, keepdim=True)
        a1 = m @ p["se1_w"].t() + p["se1_b"]
        h1 = torch.relu(a1)
        gate = torch.sigmoid(h1 @ p["se2_w"].t() + p["se2_b"])           # (1,HP)
        ig = inter * gate
        cache.update(m=m, h1=h1, gate=gate)
    else:
        ig = inter
    igd = ig * drop_inter if drop_inter is not None else ig
    h = torch.cat([z, igd], dim=1)
    acts = [h]
    relu_pos = []
    n = len(p["mlp_w"])
    for li in range(n - 1):
        h = torch.relu(h @ p["mlp_w"][li].t() + p["mlp_b"][li])
        relu_pos.append(h > 0)
        if drop_mlp is not None and drop_mlp[li] is not None:
            h = h * drop_mlp[li]
        acts.append(h)
    logit = (h @ p["mlp_w"][n - 1].t() + p["mlp_b"][n - 1]).squeeze(1)
    cache.update(acts=acts, relu_pos=relu_pos, drop_inter=drop_inter, drop_mlp=drop_mlp, use_se=use_se, use_residual=use_residual)
    return logit, cache


def qnn_bwd(dlogit: Tensor, c: dict, p: Dict[str, Tensor]):
    g: Dict[str, object] = {}
    xF = c["xF"]
    B, F, D = xF.shape
    H, _, r = p["U"].shape
    n = len(p["mlp_w"])
    gw: List[Optional[Tensor]] = [None] * n
    gb: List[Optional[Tensor]] = [None] * n
    acts = c["acts"]
    dh = dlogit.unsqueeze(1) @ p["mlp_w"][n - 1]                      # (B, last_hidden)
    gw[n - 1] = dlogit.unsqueeze(0) @ acts[n - 1]
    gb[n - 1] = dlogit.sum().reshape(1)
    for li in range(n - 2, -1, -1):
        if c["drop_mlp"] is not None and c["drop_mlp"][li] is not None:
            dh = dh * c["drop_mlp"][li]
        dh = dh * c["relu_pos"][li].to(dh.dtype)
        gw[li] = dh.t() @ acts[li]
        gb[li] = dh.sum(0)
        dh = dh @ p["mlp_w"][li]
    g["mlp_w"], g["mlp_b"] = gw, gb
    dz = dh[:, :F * D]
    if not c["use_residual"]:
        dz = torch.zeros_like(dz)
    digd = dh[:, F * D:]
    dig = digd * c["drop_inter"] if c["drop_inter"] is not None else digd
    inter = c["inter"]
    if c["use_se"]:
        gate = c["gate"]
        dinter = dig * gate
        dgate = (dig * inter).sum(0, keepdim=True)
        da2 = dgate * gate * (1.0 - gate)
        g["se2_w"] = da2.t() @ c["h1"]
        g["se2_b"] = da2.reshape(-1)
        dh1 = (da2 @ p["se2_w"]) * (c["h1"] > 0).to(dig.dtype)
        g["se1_w"] = dh1.t() @ c["m"]
        g["se1_b"] = dh1.reshape(-1)
        dm = dh1 @ p["se1_w"]
        dinter = dinter + dm / B
    else:
        dinter = dig
    dinter3 = dinter.reshape(B, H, -1)
    g["V"] = torch.einsum("bhr,bhp->hrp", c["quad"], dinter3)
    dquad = torch.einsum("bhp,hrp->bhr", dinter3, p["V"])
    z3 = c["z"].reshape(B, F, D)
    A = torch.einsum("bfd,hdr->bhfr", z3, p["U"])
    dA = dquad.unsqueeze(2) * 2.0 * (c["s"].unsqueeze(2) - A)        # (B,H,F,r)
    g["U"] = torch.einsum("bfd,bhfr->hdr", z3, dA)
    dz = dz + torch.einsum("bhfr,hdr->bfd", dA, p["U"]).reshape(B, F * D)
    dflat, g["pre_w"] = rmsnorm_bwd(dz, c["flat"], p["pre_w"], c["rz"])
    return dflat.reshape(B, F, D), g


# --------------------------------------------------------------------------
# loss   (train.py:71-90)
# --------------------------------------------------------------------------
def bce_wll(logits: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """Returns (loss scalar, dloss/dlogits)."""
    pos = y > 0.5
    n_pos = int(pos.sum())
    n_neg = int((~pos).sum())
    loss = logits.new_zeros(())
    dz = torch.zeros_like(logits)
    if n_pos > 0:
        loss = loss + 0.5 * softplus(-logits[pos]).mean()
        dz[pos] = -0.5 * torch.sigmoid(-logits[pos]) / n_pos
    if n_neg > 0:
        loss = loss + 0.5 * softplus(logits[~pos]).mean()
        dz[~pos] = 0.5 * torch.sigmoid(logits[~pos]) / n_neg
    return loss, dz


# --------------------------------------------------------------------------
# schedule / clip / AdamW / EMA
# --------------------------------------------------------------------------
def cosine_warmup_lr(epoch: int, step: int, steps_per_epoch: int, base_lr: float,
                     warmup_epochs: int = 1, total_epochs: int = 10) -> float:
    g = epoch * steps_per_epoch + step
    warm = warmup_epochs * steps_per_epoch
    total = total_epochs * steps_per_epoch
    if g < warm:
        return base_lr * (g + 1) / max(1, warm)
    prog = (g - warm) / max(1, total - warm)
    return 0.5 * base_lr * (1.0 + math.cos(math.pi * prog))


def clip_coef(grads: Dict[str, Tensor], max_norm: float) -> Tuple[float, float]:
    """torch clip_grad_norm_: coef = min(1, max_norm / (||g||_2 + 1e-6))."""
    tot = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    return min(1.0, max_norm / (tot + 1e-6)), tot


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float,
               b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """In-place torch.optim.AdamW single-tensor math (train.py:138,195)."""
    p.mul_(1.0 - lr * wd)
    m.lerp_(g, 1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def ema_decay_at(n_updates: int, base: float, warmup_steps: int = 0, warmup_type: str = "linear") -> float:
    """utils/ema.py:72-88 (0-based update index)."""
    d = base
    if warmup_steps <= 0 or warmup_type == "none":
        return d
    t = min(1.0, (n_updates + 1) / warmup_steps)
    if warmup_type == "linear":
        dec = 1.0 - (1.0 - d) * t
    elif warmup_type == "cosine":
        c = 0.5 * (1 + math.cos(math.pi * (1 - t)))
        dec = 1.0 - (1.0 - d) * c
    else:
        dec = d
    return float(max(0.0, min(1.0, dec)))


def ema_update(shadow: Tensor, p: Tensor, decay: float):
    shadow.mul_(decay).add_(p, alpha=1.0 - decay)


# --------------------------------------------------------------------------
# whole model, keyed by the reference state_dict names
# --------------------------------------------------------------------------
class ModelSpec:
    """Static description of one CTRModel instance (what CTRModel.__init__ derives from cfg)."""

    def __init__(self, cfg: dict, num_dim: int, mask_dim: int, cat_cols: List[str]):
        self.D = int(cfg["model"]["emb_dim"])
        seqc = cfg["sequence"]
        self.L = int(seqc["max_len"])
        self.top_k = int(seqc["top_k"])
        self.tau = float(seqc["recency_tau"])
        self.pad_id = int(seqc["pad_id"])
        self.query_mode = seqc["query_mode"]
        self.query_key = seqc["query_key"]
        self.transformer = bool(seqc["transformer_block"])
        tfm = seqc.get("tfm") or {}
        self.n_layers = int(tfm.get("n_layers", 2)) if self.transformer else 0
        self.n_heads = int(tfm.get("n_heads", 4))
        self.gating = tfm.get("gating", "softmax") if tfm else "softmax"
        self.pos_bias = bool(tfm.get("add_positional_bias", True))
        self.num_dim, self.mask_dim = num_dim, mask_dim
        self.cat_cols = list(cat_cols)
        qa = cfg["model"]["qnn_alpha"]
        self.use_qnn = bool(qa["enabled"])
        self.use_se = bool(qa.get("use_se", True))
        self.use_residual = bool(qa.get("use_residual", True))
        self.aux_weight = float(qa.get("aux_head_weight", 0.0))
        self.q_idx = self.cat_cols.index(self.query_key) if self.query_mode != "S2" else -1


def _layer_params(sd: Dict[str, Tensor], i: int, pos_bias: bool) -> Dict[str, Tensor]:
    pre = f"dare.layers.{i}."
    return dict(in_w=sd[pre + "mha.in_proj_weight"], in_b=sd[pre + "mha.in_proj_bias"],
                out_w=sd[pre + "mha.out_proj.weight"], out_b=sd[pre + "mha.out_proj.bias"],
                n1=sd[pre + "norm1.w"], f1_w=sd[pre + "ffn.0.weight"], f1_b=sd[pre + "ffn.0.bias"],
                f2_w=sd[pre + "ffn.3.weight"], f2_b=sd[pre + "ffn.3.bias"], n2=sd[pre + "norm2.w"],
                rel=sd.get(pre + "pbias.rel.weight") if pos_bias else None)


_LAYER_KEYMAP = dict(in_w="mha.in_proj_weight", in_b="mha.in_proj_bias", out_w="mha.out_proj.weight",
                     out_b="mha.out_proj.bias", n1="norm1.w", f1_w="ffn.0.weight", f1_b="ffn.0.bias",
                     f2_w="ffn.3.weight", f2_b="ffn.3.bias", n2="norm2.w", rel="pbias.rel.weight")


def _qnn_params(sd: Dict[str, Tensor]) -> Dict[str, object]:
    ws, bs = [], []
    i = 0
    while f"qnn.mlp.{i}.weight" in sd:
        ws.append(sd[f"qnn.mlp.{i}.weight"])
        bs.append(sd[f"qnn.mlp.{i}.bias"])
        i += 3
    p = dict(pre_w=sd["qnn.pre_norm.w"], U=sd["qnn.U"], V=sd["qnn.V"], mlp_w=ws, mlp_b=bs)
    if "qnn.se.fc.0.weight" in sd:
        p.update(se1_w=sd["qnn.se.fc.0.weight"], se1_b=sd["qnn.se.fc.0.bias"],
                 se2_w=sd["qnn.se.fc.2.weight"], se2_b=sd["qnn.se.fc.2.bias"])
    return p


def model_forward(sd: Dict[str, Tensor], spec: ModelSpec, batch: Dict[str, Tensor], keep: bool = False):
    """CTRModel.forward (wrapper.py:138-176) with dropout off.  Returns (logits, prob, aux, cache)."""
    dt = sd["dare.emb_att.weight"].dtype
    X_num, X_mask = batch["X_num"].to(dt), batch["X_mask"].to(dt)
    X_cat, seq = batch["X_cat"].long(), batch["seq"].long()
    D = spec.D
    c: dict = {}
    num_e = numbin_fwd(X_num, sd["num_embed.weight"], sd["num_embed.bias"], sd["num_embed.out_proj.weight"]) \
        if spec.num_dim > 0 else None
    mask_e = numbin_fwd(X_mask, sd["mask_embed.weight"], None, sd["mask_embed.out_proj.weight"]) \
        if spec.mask_dim > 0 else None
    tabs = [sd[f"cat_embs.{n}.weight"] for n in spec.cat_cols]
    projs = [sd[f"cat_proj.{n}.weight"] for n in spec.cat_cols]
    cat_e = cat_fwd(X_cat, tabs, projs)                                  # (B,Fc,D)
    parts = []
    if num_e is not None:
        parts.append(num_e.mean(1))
    if mask_e is not None:
        parts.append(mask_e.mean(1))
    parts.append(cat_e.mean(1))
    ctx = torch.cat(parts, dim=1)
    s2_pre = None
    if spec.query_mode == "S1":
        qv = cat_e[:, spec.q_idx]
    else:
        s2_pre = ctx @ sd["ctx_mlp.0.weight"].t() + sd["ctx_mlp.0.bias"]
        s2 = torch.relu(s2_pre)
        qv = s2 if spec.query_mode == "S2" else 0.5 * (cat_e[:, spec.q_idx] + s2)
    rec = recency_table(seq.shape[1], spec.tau, dt)
    sel_rep, vals, idx = retrieve_fwd(seq, qv, sd["dare.emb_att.weight"], sd["dare.emb_rep.weight"],
                                      rec, spec.top_k, spec.pad_id)
    x = sel_rep
    lcaches = []
    for i in range(spec.n_layers):
        x, lc = encoder_layer_fwd(x, _layer_params(sd, i, spec.pos_bias), spec.n_heads, spec.top_k)
        lcaches.append(lc)
    u, aux, w = pool_fwd(x, vals, sd["dare.aux_head.weight"], sd["dare.aux_head.bias"], spec.gating)
    if spec.use_qnn:
        feats = [u.unsqueeze(1)]
        if num_e is not None:
            feats.append(num_e)
        if mask_e is not None:
            feats.append(mask_e)
        feats.append(cat_e)
        xF = torch.cat(feats, dim=1)
        logits, qc = qnn_fwd(xF, _qnn_params(sd), spec.use_se, spec.use_residual)
        c["qc"] = qc
    else:
        ff = [u]
        if num_e is not None:
            ff.append(num_e.mean(1))
        if mask_e is not None:
            ff.append(mask_e.mean(1))
        ff.append(cat_e.reshape(cat_e.shape[0], -1))
        allf = torch.cat(ff, dim=1)
        fpre = allf @ sd["fc.0.weight"].t() + sd["fc.0.bias"]
        fh = torch.relu(fpre)
        logits = (fh @ sd["fc.3.weight"].t() + sd["fc.3.bias"]).squeeze(1)
        c.update(allf=allf, fh=fh)
    prob = torch.sigmoid(logits)
    if keep:
        c.update(X_num=X_num, X_mask=X_mask, X_cat=X_cat, seq=seq, num_e=num_e, mask_e=mask_e, cat_e=cat_e,
                 ctx=ctx, s2_pre=s2_pre, qv=qv, sel_rep=sel_rep, vals=vals, idx=idx, x_enc=x, lcaches=lcaches,
                 u=u, w=w, tabs=tabs, projs=projs)
    return logits, prob, aux, c


def model_backward(sd: Dict[str, Tensor], spec: ModelSpec, c: dict, dlogits: Tensor, daux: Tensor
                   ) -> Dict[str, Tensor]:
    """Hand-derived backward of model_forward.  Returns dense grads keyed like state_dict."""
    G: Dict[str, Tensor] = {}
    D = spec.D
    B = dlogits.shape[0]
    Fn, Fm, Fc = spec.num_dim, spec.mask_dim, len(spec.cat_cols)
    dnum_e = torch.zeros_like(c["num_e"]) if c["num_e"] is not None else None
    dmask_e = torch.zeros_like(c["mask_e"]) if c["mask_e"] is not None else None
    dcat_e = torch.zeros_like(c["cat_e"])
    if spec.use_qnn:
        dxF, gq = qnn_bwd(dlogits, c["qc"], _qnn_params(sd))
        G["qnn.pre_norm.w"], G["qnn.U"], G["qnn.V"] = gq["pre_w"], gq["U"], gq["V"]
        if spec.use_se:
            G["qnn.se.fc.0.weight"], G["qnn.se.fc.0.bias"] = gq["se1_w"], gq["se1_b"]
            G["qnn.se.fc.2.weight"], G["qnn.se.fc.2.bias"] = gq["se2_w"], gq["se2_b"]
        for li, (gw, gb) in enumerate(zip(gq["mlp_w"], gq["mlp_b"])):
            G[f"qnn.mlp.{3 * li}.weight"], G[f"qnn.mlp.{3 * li}.bias"] = gw, gb
        du = dxF[:, 0]
        o = 1
        if dnum_e is not None:
            dnum_e += dxF[:, o:o + Fn]; o += Fn
        if dmask_e is not None:
            dmask_e += dxF[:, o:o + Fm]; o += Fm
        dcat_e += dxF[:, o:o + Fc]
    else:
        fh, allf = c["fh"], c["allf"]
        G["fc.3.weight"] = dlogits.unsqueeze(0) @ fh
        G["fc.3.bias"] = dlogits.sum().reshape(1)
        dfh = (dlogits.unsqueeze(1) @ sd["fc.3.weight"]) * (fh > 0).to(fh.dtype)
        G["fc.0.weight"] = dfh.t() @ allf
        G["fc.0.bias"] = dfh.sum(0)
        dall = dfh @ sd["fc.0.weight"]
        du = dall[:, :D]
        o = D
        if dnum_e is not None:
            dnum_e += (dall[:, o:o + D] / Fn).unsqueeze(1); o += D
        if dmask_e is not None:
            dmask_e += (dall[:, o:o + D] / Fm).unsqueeze(1); o += D
        dcat_e += dall[:, o:].reshape(B, Fc, D)
    dx, dvals, G["dare.aux_head.weight"], G["dare.aux_head.bias"] = pool_bwd(
        c["x_enc"], c["vals"], c["w"], sd["dare.aux_head.weight"], du, daux, c["u"], spec.gating)
    for i in range(spec.n_layers - 1, -1, -1):
        dx, gl = encoder_layer_bwd(dx, c["lcaches"][i], _layer_params(sd, i, spec.pos_bias), spec.n_heads)
        for k, v in gl.items():
            G[f"dare.layers.{i}." + _LAYER_KEYMAP[k]] = v
    dq, G["dare.emb_att.weight"], G["dare.emb_rep.weight"] = retrieve_bwd(
        c["seq"], c["qv"], sd["dare.emb_att.weight"], sd["dare.emb_rep.weight"], c["idx"], dx, dvals, spec.pad_id)
    dctx = None
    if spec.query_mode == "S1":
        dcat_e[:, spec.q_idx] += dq
    else:
        scale = 1.0 if spec.query_mode == "S2" else 0.5
        ds2 = scale * dq * (c["s2_pre"] > 0).to(dq.dtype)
        G["ctx_mlp.0.weight"] = ds2.t() @ c["ctx"]
        G["ctx_mlp.0.bias"] = ds2.sum(0)
        dctx = ds2 @ sd["ctx_mlp.0.weight"]
        if spec.query_mode != "S2":
            dcat_e[:, spec.q_idx] += 0.5 * dq
    if dctx is not None:
        o = 0
        if dnum_e is not None:
            dnum_e += (dctx[:, o:o + D] / Fn).unsqueeze(1); o += D
        if dmask_e is not None:
            dmask_e += (dctx[:, o:o + D] / Fm).unsqueeze(1); o += D
        dcat_e += (dctx[:, o:o + D] / Fc).unsqueeze(1)
    else:
        # ctx_mlp is unused in S1 mode: the reference leaves its .grad = None (treated as zero)
        pass
    dtabs, dprojs = cat_bwd(c["X_cat"], c["tabs"], c["projs"], dcat_e)
    for n, dt_, dp_ in zip(spec.cat_cols, dtabs, dprojs):
        G[f"cat_embs.{n}.weight"], G[f"cat_proj.{n}.weight"] = dt_, dp_
    if dnum_e is not None:
        G["num_embed.weight"], G["num_embed.bias"], G["num_embed.out_proj.weight"] = numbin_bwd(
            c["X_num"], sd["num_embed.weight"], sd["num_embed.bias"], sd["num_embed.out_proj.weight"], dnum_e)
    if dmask_e is not None:
        G["mask_embed.weight"], _, G["mask_embed.out_proj.weight"] = numbin_bwd(
            c["X_mask"], sd["mask_embed.weight"], None, sd["mask_embed.out_proj.weight"], dmask_e)
    return G


def loss_and_grads(sd, spec: ModelSpec, batch, y: Tensor):
    """One forward+backward of the train.py:164-185 step body (dropout off)."""
    logits, prob, aux, c = model_forward(sd, spec, batch, keep=True)
    loss, dz = bce_wll(logits, y.to(logits.dtype))
    daux = torch.zeros_like(aux)
    if spec.aux_weight > 0:
        la, da = bce_wll(aux, y.to(aux.dtype))
        loss = loss + spec.aux_weight * la
        daux = spec.aux_weight * da
    G = model_backward(sd, spec, c, dz, daux)
    if spec.aux_weight <= 0:
        # aux_logit never enters the loss: autograd leaves aux_head.grad = None, so AdamW skips it entirely
        G.pop("dare.aux_head.weight"); G.pop("dare.aux_head.bias")
    return loss, logits, prob, aux, G, c


class OracleTrainer:
    """Literal restatement of the train.py:152-203 step (fp32, dropout off): lr schedule ->
    fwd -> loss -> bwd -> global-norm clip -> AdamW (wd on every tensor) -> EMA."""

    def __init__(self, sd: Dict[str, Tensor], spec: ModelSpec, lr: float, wd: float, clip: float,
                 ema_decay: Optional[float] = None, ema_warmup_steps: int = 0, ema_warmup_type: str = "linear"):
        self.sd, self.spec = sd, spec
        self.lr, self.wd, self.clip = lr, wd, clip
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.t = 0
        self.ema_decay = ema_decay
        self.ema_warm, self.ema_type = ema_warmup_steps, ema_warmup_type
        self.shadow = {k: v.clone() for k, v in sd.items()} if ema_decay is not None else None
        self.n_ema = 0

    def step(self, batch, y, lr: Optional[float] = None):
        lr = self.lr if lr is None else lr
        loss, logits, prob, aux, G, _ = loss_and_grads(self.sd, self.spec, batch, y)
        if self.clip > 0:
            coef, _ = clip_coef(G, self.clip)
            for g in G.values():
                g.mul_(coef)
        self.t += 1
        for k, g in G.items():
            adamw_step(self.sd[k], g, self.m[k], self.v[k], self.t, lr, self.wd)
        if self.shadow is not None:
            d = ema_decay_at(self.n_ema, self.ema_decay, self.ema_warm, self.ema_type)
            for k in self.sd.keys():          # EMA walks every requires-grad param (ema.py:103-113)
                ema_update(self.shadow[k], self.sd[k], d)
            self.n_ema += 1
        return float(loss), logits, G
