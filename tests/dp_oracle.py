"""Data-parallel restatement of the oracle step, mirroring FusedTrainer's collectives (C2 dense-grad all-reduce,
C3 global class counts, C4 global SE squeeze, C1 summed table-row gradients).  Test infrastructure (gloo, CPU)."""
import torch

from oracle import ctr_oracle as O


def dp_loss_and_grads(sd, spec, batch, y, B_total, dist):
    """Each rank holds a slice of the global batch; returns the all-reduced gradients (== global-batch gradients)."""
    # --- forward with a global SE mean: run the local forward up to `inter`, all-reduce its column sum ---
    logits, prob, aux, c = O.model_forward(sd, spec, batch, keep=True)
    qc = c["qc"]
    p = O._qnn_params(sd)
    inter = qc["inter"]
    s = inter.sum(0, keepdim=True)
    dist.all_reduce(s)
    m = s / B_total
    a1 = m @ p["se1_w"].t() + p["se1_b"]
    h1 = torch.relu(a1)
    gate = torch.sigmoid(h1 @ p["se2_w"].t() + p["se2_b"])
    ig = inter * gate
    h = torch.cat([qc["z"], ig], dim=1)
    acts, relu_pos = [h], []
    n = len(p["mlp_w"])
    for li in range(n - 1):
        h = torch.relu(h @ p["mlp_w"][li].t() + p["mlp_b"][li])
        relu_pos.append(h > 0)
        acts.append(h)
    logits = (h @ p["mlp_w"][n - 1].t() + p["mlp_b"][n - 1]).squeeze(1)
    qc.update(m=m, h1=h1, gate=gate, acts=acts, relu_pos=relu_pos)
    # --- loss with global class counts ---
    pos = y > 0.5
    cnt = torch.tensor([float(pos.sum()), float((~pos).sum())])
    dist.all_reduce(cnt)
    dz = torch.where(pos, -0.5 * torch.sigmoid(-logits) / max(cnt[0], 1), 0.5 * torch.sigmoid(logits) / max(cnt[1], 1))
    daux = spec.aux_weight * torch.where(pos, -0.5 * torch.sigmoid(-aux) / max(cnt[0], 1), 0.5 * torch.sigmoid(aux) / max(cnt[1], 1))
    # --- backward; the SE gate gradient needs the global sum of dgate and dm is divided by the global batch ---
    orig_bwd = O.qnn_bwd

    def qnn_bwd_dp(dlogit, cc, pp):
        # replicate O.qnn_bwd with the two cross-rank reductions
        import types
        g = {}
        xF = cc["xF"]
        B, F, D = xF.shape
        H = pp["U"].shape[0]
        nn_ = len(pp["mlp_w"])
        gw, gb = [None] * nn_, [None] * nn_
        dh = dlogit.unsqueeze(1) @ pp["mlp_w"][nn_ - 1]
        gw[nn_ - 1] = dlogit.unsqueeze(0) @ cc["acts"][nn_ - 1]
        gb[nn_ - 1] = dlogit.sum().reshape(1)
        for li in range(nn_ - 2, -1, -1):
            dh = dh * cc["relu_pos"][li].to(dh.dtype)
            gw[li] = dh.t() @ cc["acts"][li]
            gb[li] = dh.sum(0)
            dh = dh @ pp["mlp_w"][li]
        g["mlp_w"], g["mlp_b"] = gw, gb
        dzz = dh[:, :F * D]
        dig = dh[:, F * D:]
        gate_ = cc["gate"]
        dinter = dig * gate_
        dgate = (dig * cc["inter"]).sum(0, keepdim=True)
        dist.all_reduce(dgate)                                                     # C4 backward
        da2 = dgate * gate_ * (1.0 - gate_)
        g["se2_w"] = da2.t() @ cc["h1"]; g["se2_b"] = da2.reshape(-1)
        dh1 = (da2 @ pp["se2_w"]) * (cc["h1"] > 0).to(dig.dtype)
        g["se1_w"] = dh1.t() @ cc["m"]; g["se1_b"] = dh1.reshape(-1)
        dinter = dinter + (dh1 @ pp["se1_w"]) / B_total
        # SE weight grads are identical on every rank (computed from global sums): keep them on rank 0 only so the
        # final all-reduce(sum) does not multiply them by the world size
        if dist.get_rank() != 0:
            for k in ("se1_w", "se1_b", "se2_w", "se2_b"):
                g[k] = torch.zeros_like(g[k])
        dinter3 = dinter.reshape(B, H, -1)
        g["V"] = torch.einsum("bhr,bhp->hrp", cc["quad"], dinter3)
        dquad = torch.einsum("bhp,hrp->bhr", dinter3, pp["V"])
        z3 = cc["z"].reshape(B, F, D)
        A = torch.einsum("bfd,hdr->bhfr", z3, pp["U"])
        dA = dquad.unsqueeze(2) * 2.0 * (cc["s"].unsqueeze(2) - A)
        g["U"] = torch.einsum("bfd,bhfr->hdr", z3, dA)
        dzz = dzz + torch.einsum("bhfr,hdr->bfd", dA, pp["U"]).reshape(B, F * D)
        dflat, g["pre_w"] = O.rmsnorm_bwd(dzz, cc["flat"], pp["pre_w"], cc["rz"])
        return dflat.reshape(B, F, D), g

    O.qnn_bwd = qnn_bwd_dp
    try:
        G = O.model_backward(sd, spec, c, dz, daux)
    finally:
        O.qnn_bwd = orig_bwd
    for k in G:
        dist.all_reduce(G[k])                                                      # C2 (+ C1 for the tables)
    return G
