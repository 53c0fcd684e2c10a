"""GPU: the batched shard reader's device form and the fused fold-ensemble inference (run with -m gpu)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import batch_of, group, load_case, max_err_scaled

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reader_device_form_equals_contract_form(tmp_path):
    from b200ctr import BatchedShardReader, configs, synth
    cfg = configs.dare_qnn_next()
    cfg["sequence"]["max_len"] = 64
    mtr, mte = synth.write_cache(str(tmp_path), cfg, n_train=3000, n_test=500, n_num=13, shard_rows=1100, seed=5)
    index = np.random.default_rng(0).permutation(3000)
    dev = torch.device("cuda:0")
    host = BatchedShardReader(mtr, index, train=True, batch_size=512, shuffle=True, seed=4)
    gpu = BatchedShardReader(mtr, index, train=True, batch_size=512, shuffle=True, seed=4, device=dev)
    n = 0
    for hb, gb in zip(host, gpu):
        assert gb["X_num"].is_cuda and gb["X_mask"].dtype == torch.float32 and gb["seq"].dtype == torch.int32 and gb["y"].dtype == torch.float32
        for k in ("X_num", "X_mask", "y"):
            assert torch.equal(gb[k].cpu(), hb[k]), k
        for k in ("X_cat", "seq"):
            assert torch.equal(gb[k].cpu().long(), hb[k]), k
        assert torch.equal(gb["groups"], hb["groups"])
        n += 1
    assert n == len(host) == 6


def test_fused_trainer_consumes_reader_batches(tmp_path):
    """One epoch of FusedTrainer fed by the device-form reader == the same epoch fed by the reference-contract CPU batches."""
    import b200ctr
    from b200ctr import BatchedShardReader, synth
    z, cfg = load_case("qnn_concat")
    g = cfg["golden"]
    cfg = copy.deepcopy(cfg)
    mtr, _ = synth.write_cache(str(tmp_path), cfg, n_train=256, n_test=8, n_num=g["num_dim"], shard_rows=100, seed=2, planted=False)
    dev = torch.device("cuda:0")
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    finals = []
    for device in (None, dev):
        m = b200ctr.CTRModel(cfg, 1_000_001, g["num_dim"], g["mask_dim"], card, cols)
        sd = group(z, "init/")
        sd = {k: v for k, v in sd.items() if not k.startswith("dare.emb_")}
        torch.manual_seed(0)
        for k in ("dare.emb_att.weight", "dare.emb_rep.weight"):
            sd[k] = torch.randn(1_000_001, 64) * 0.1
            sd[k][0] = 0
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).train()
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=4)
        rd = BatchedShardReader(mtr, np.arange(256), train=True, batch_size=64, shuffle=False, device=device)
        for b in rd:
            y = b.pop("y")
            tr.step(b, y)
        finals.append({k: v.detach().cpu().clone() for k, v in m.state_dict().items() if not k.startswith("dare.emb_")})
    for k in finals[0]:
        assert max_err_scaled(finals[1][k], finals[0][k]) < 1e-4, k          # only the table-gradient atomics' order differs (fed through 4 Adam steps)


@pytest.mark.parametrize("with_T", [False, True])
def test_fold_ensemble_matches_the_reference_inference_loop(with_T):
    """FoldEnsemble.predict == infer.py:104-158 restated with torch on the drop-in models: per-fold forward, temperature
    calibration in numpy on the CPU (Calibrator.predict_proba), clamp, ensemble_probs('logit_mean')."""
    import b200ctr
    dev = torch.device("cuda:0")
    z, cfg = load_case("qnn_concat")
    g = cfg["golden"]
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    models = []
    for f in range(3):
        m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cols)
        sd = group(z, "init/")
        gen = torch.Generator().manual_seed(100 + f)
        sd = {k: v + 0.05 * torch.randn(v.shape, generator=gen) * (0 if k.startswith("dare.emb_") else 1) for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
        models.append(m.to(dev).eval())
    batch, _ = batch_of(z, 0)

    class Cal:                                              # the attributes FoldEnsemble reads off src.utils.calibration.Calibrator
        def __init__(self, T):
            self.temp_scaler = type("S", (), {"log_temp": torch.tensor(np.log(T), dtype=torch.float32)})()
            self.iso, self.clamp_T = None, (0.2, 5.0)
    Ts = [1.37, 0.61, 7.5] if with_T else None               # 7.5 is clamped to 5.0 like TemperatureScaler.forward does
    cals = [Cal(t) for t in Ts] if with_T else None
    ens = b200ctr.FoldEnsemble(models, cals)
    p = ens.predict(batch)
    # reference loop
    p_list = []
    with torch.no_grad():
        for i, m in enumerate(models):
            zl, pl, _ = m(batch)
            if with_T:
                T = min(max(Ts[i], 0.2), 5.0)
                zz = np.clip((zl.cpu().double().numpy() / np.float32(T)).astype(np.float32).astype(np.float64), -50, 50)
                pl = torch.from_numpy(np.clip(1.0 / (1.0 + np.exp(-zz)), 1e-7, 1 - 1e-7)).float().to(dev)
            p_list.append(torch.clamp(pl, 1e-7, 1 - 1e-7))
        P = torch.stack(p_list, 0)
        L = torch.log(P) - torch.log1p(-P)
        ref = torch.sigmoid(L.mean(0))
    assert float((p - ref).abs().max()) < 2e-6
    # one model, no calibrator: the clamped probability itself (infer.py:126-127)
    one = b200ctr.FoldEnsemble(models[:1]).predict(batch)
    assert float((one - p_list[0]).abs().max()) < 1e-6 or with_T
