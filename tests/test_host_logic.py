"""CPU-side tests of the host logic: spec derivation, schedules, state_dict contract, configs, synthetic data, the
no-fallback rule, and the data-parallel plumbing under a 2-rank gloo group."""
import json
import os

import numpy as np
import pytest
import torch

import b200ctr
from b200ctr import configs, synth
from b200ctr.hotpath import DropCtl, ModelSpec, recency_table
from oracle import ctr_oracle as O
from tests.helpers import CASES, group, load_case


@pytest.mark.parametrize("kind", CASES)
def test_state_dict_contract_matches_reference(kind):
    z, cfg = load_case(kind)
    g = cfg["golden"]
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    torch.manual_seed(cfg["seed"])
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cols)
    sd = m.state_dict()
    init = group(z, "init/")
    assert list(sd.keys()) == list(init.keys())                    # same names, same order
    for k in sd:
        assert tuple(sd[k].shape) == tuple(init[k].shape) and sd[k].dtype == torch.float32
    # same construction order => same RNG consumption => identical initial values for untouched tensors
    same = sum(bool(torch.equal(sd[k], init[k])) for k in sd)
    assert same >= len(sd) // 2
    m.load_state_dict(init, strict=True)
    assert float(m.dare.emb_att.weight[0].abs().max()) == 0.0       # padding row


def test_full_size_contract_key_count():
    """SURVEY App. B key structure at the shipped cfgs (tables shrunk to keep the CPU test light): 127 tensors with
    3 encoder layers (dare_qnn_next), 138 with 4 (v3_k148_s1), 27 for dare_base."""
    cfg = configs.dare_qnn_next()
    m = b200ctr.CTRModel(cfg, 1000, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cfg["data"]["cat_cols"])
    assert len(m.state_dict()) == 127
    assert m.qnn.mlp[0].weight.shape == (512, 6 * 192 + 200 * 64)
    cfg = configs.v3_k148_s1()
    m = b200ctr.CTRModel(cfg, 1000, configs.N_NUM, configs.N_NUM, configs.cardinals(cfg), cfg["data"]["cat_cols"])
    assert len(m.state_dict()) == 138
    assert m.dare.layers[3].pbias.rel.weight.shape == (2 * 148 + 1, 8)
    cfg = configs.dare_base()
    m = b200ctr.CTRModel(cfg, 1000, 5, 5, {c: 50 for c in cfg["data"]["cat_cols"]}, cfg["data"]["cat_cols"])
    assert len(m.state_dict()) == 27
    assert m.fc[0].weight.shape == (512, 64 * (3 + 6))


def test_spec_rejects_unbuilt_variants():
    cfg = configs.dare_qnn_next()
    bad = configs.shrink(cfg, **{"model.emb_dim": 32})
    with pytest.raises(b200ctr.B200CtrError):
        ModelSpec(bad, 4, 4, cfg["data"]["cat_cols"])
    bad = configs.shrink(cfg, **{"model.qnn_alpha.pair_grouping": "block"})
    with pytest.raises(b200ctr.B200CtrError):
        ModelSpec(bad, 4, 4, cfg["data"]["cat_cols"])
    bad = configs.shrink(cfg, **{"sequence.tfm.n_heads": 4})
    with pytest.raises(b200ctr.B200CtrError):
        ModelSpec(bad, 4, 4, cfg["data"]["cat_cols"])


def test_cpu_parameters_raise_instead_of_falling_back():
    z, cfg = load_case("s2_relu")
    g = cfg["golden"]
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cols)
    batch = {k: torch.from_numpy(np.array(z[f"step0/batch/{k}"])) for k in ("X_num", "X_mask", "X_cat", "seq")}
    with pytest.raises(b200ctr.B200CtrError):
        m(batch)
    with pytest.raises(b200ctr.B200CtrError):
        m.dare(batch["seq"], torch.zeros(batch["seq"].shape[0], 64))
    with pytest.raises(RuntimeError):
        b200ctr.FusedTrainer(m, cfg)
    with pytest.raises(RuntimeError):
        b200ctr.ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2))


def test_schedules_match_oracle():
    for spe, warm, tot in [(7, 2, 8), (1, 1, 10), (100, 0, 3)]:
        for g in range(spe * tot):
            e, s = divmod(g, spe)
            assert b200ctr.cosine_warmup_lr(e, s, spe, 3e-4, warm, tot) == O.cosine_warmup_lr(e, s, spe, 3e-4, warm, tot)
    for n in range(0, 50, 7):
        for kind in ("linear", "cosine", "none"):
            assert b200ctr.ema_decay_at(n, 0.999, 20, kind) == O.ema_decay_at(n, 0.999, 20, kind)
    assert b200ctr.ema_decay_at(5, 0.999) == 0.999


def test_recency_table_is_bit_identical_to_oracle():
    for L, tau in [(400, 512), (600, 256), (40, 16), (1, 0.5)]:
        assert torch.equal(recency_table(L, tau, "cpu"), O.recency_table(L, tau))


def test_dropctl():
    assert DropCtl(False)(0.5, 3).p == 0.0
    d = DropCtl(True, seed=(7 << 32) | 5, step=3)(0.25, 4)
    assert abs(d.p - 0.25) < 1e-7 and d.seed_lo == 5 and d.seed_hi == 7 and d.stream_id == 3 * 64 + 4
    assert DropCtl(True).scale(0.2) == 1.25 and DropCtl(False).scale(0.2) == 1.0


def test_restated_configs_and_synthetic_batch_contract():
    cfg = configs.dare_qnn_next()
    assert len(cfg["data"]["cat_cols"]) == 35 and cfg["sequence"]["top_k"] == 60 and cfg["train"]["batch_size"] == 8192
    assert sum(cfg["model"]["cat_embedding_dims"].values()) == 597          # sum_c d_c -> 2,388 B / sample
    assert configs.dare_qnn_next_k100_s1()["sequence"]["top_k"] == 100
    assert configs.v3_k148_s1()["sequence"]["tfm"]["n_layers"] == 4 and configs.cardinals(configs.v3_k148_s1())["gender"] == 1505
    s = synth.ToySchema(cfg, configs.N_NUM, configs.N_NUM)
    b, y = s.batch(257, seed=3)
    assert b["X_num"].dtype == np.float32 and b["X_mask"].dtype == np.float32
    assert b["X_cat"].dtype == np.int64 and b["seq"].dtype == np.int64 and y.dtype == np.float32
    assert b["seq"].shape == (257, 400) and b["X_cat"].shape == (257, 35)
    card = configs.cardinals(cfg)
    for i, c in enumerate(cfg["data"]["cat_cols"]):
        assert 0 <= b["X_cat"][:, i].min() and b["X_cat"][:, i].max() < card[c]
    nz = b["seq"] != 0
    assert bool((nz[:, 1:] >= nz[:, :-1]).all())                            # left-padded: pads only on the left
    assert b["seq"].max() <= 1_000_000
    b2, _ = s.batch(257, seed=3)
    assert all(np.array_equal(b[k], b2[k]) for k in b)                       # seeded


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, cfg = load_case("qnn_concat")
        sd = group(z, "init/")
        spec = O.ModelSpec(cfg, cfg["golden"]["num_dim"], cfg["golden"]["mask_dim"], cfg["data"]["cat_cols"])
        batch = {k: torch.from_numpy(np.array(z[f"step0/batch/{k}"])) for k in ("X_num", "X_mask", "X_cat", "seq")}
        y = torch.from_numpy(np.array(z["step0/y"]))
        B = y.shape[0]
        lo, hi = rank * B // world, (rank + 1) * B // world
        # the data-parallel recipe of FusedTrainer, on the oracle: local slice, global SE mean (C4), global class
        # counts (C3), summed gradients (C2, C1)
        from tests.dp_oracle import dp_loss_and_grads
        G = dp_loss_and_grads(sd, spec, {k: v[lo:hi] for k, v in batch.items()}, y[lo:hi], B, dist)
        if rank == 0:
            _, _, _, _, Gref, _ = O.loss_and_grads(sd, spec, batch, y)
            worst = max(float((G[k] - Gref[k]).abs().max() / (Gref[k].abs().max() + 1e-30)) for k in Gref)
            q.put(worst)
    finally:
        dist.destroy_process_group()


def test_data_parallel_recipe_reproduces_global_batch_gloo():
    """2 ranks x B/2 samples with the C2/C3/C4 reductions == 1 rank x B samples (host-side logic, gloo, CPU)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-4


def test_shape_predicates_and_shard_layout_helpers():
    """Host-side dispatch predicates of the fused tensor-core kernels and the row-shard index math (no GPU needed)."""
    from b200ctr import ops
    assert ops.ffn_fused_ok(384) and ops.ffn_fused_ok(256) and not ops.ffn_fused_ok(512) and not ops.ffn_fused_ok(100)
    assert ops.fm_fused_ok(200, 96, 64) and not ops.fm_fused_ok(300, 96, 64) and not ops.fm_fused_ok(200, 100, 64)
    assert ops.headproj_ok(6, 16, 192) and not ops.headproj_ok(6, 64, 192)
    assert not ops.headproj_ok(6, 16, 191)          # rows of H*P floats are copied in 16-byte pieces
    assert ops.lin1_ok(256) and ops.lin1_ok(512) and not ops.lin1_ok(200) and not ops.lin1_ok(640)
    assert ops.bitmap_words(10_000_000) == 312_500 and ops.bitmap_words(33) == 2
    # row r lives on rank r % N at local row r // N: the shards of a table partition its rows exactly once
    V, N = 1003, 4
    rows_local = (V + N - 1) // N
    seen = np.zeros(V, dtype=np.int64)
    for rank in range(N):
        mine = np.arange(V)[rank::N]
        assert mine.shape[0] <= rows_local
        assert np.array_equal(mine // N, np.arange(mine.shape[0])) and np.all(mine % N == rank)
        seen[mine] += 1
    assert np.all(seen == 1)
    with pytest.raises(Exception):
        ops.ShardedTable([0, 0, 0], 10)          # 3 shards: not a power of two


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE line, the JSON (library banners must go to stderr).  The reference arm is the
    unmodified reference from baseline/_ref on the host CPU (the oracle port when that copy is absent), so this runs without a
    GPU; a tiny per-step sample keeps it short."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-batch", "64"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    have_ref = os.path.isdir(os.path.join(root, "baseline", "_ref", "src", "models"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
