"""Data-parallel FusedTrainer on 2 GPUs (NCCL) == the same global batch on 1 GPU.  Needs >= 2 CUDA devices
(skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import batch_of, group, load_case, max_err_scaled

pytestmark = pytest.mark.gpu


def _build(cfg, z, dev):
    import b200ctr
    g = cfg["golden"]
    cols = cfg["data"]["cat_cols"]
    card = {c: cfg["data"]["hash_buckets"][c] + cfg["data"]["hash_buckets_margin"] for c in cols}
    m = b200ctr.CTRModel(cfg, g["seq_vocab"], g["num_dim"], g["mask_dim"], card, cols)
    m.load_state_dict(group(z, "init/"), strict=True)
    return m.to(dev).train()


def _worker(rank, world, port, out_path, shard, lazy_cat=False):
    import torch.distributed as dist
    import b200ctr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        z, cfg = load_case("qnn_concat")
        if lazy_cat:
            cfg = dict(cfg, b200ctr={"lazy_table_min_rows": 100})
        m = _build(cfg, z, dev)
        tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"], process_group=dist.group.WORLD,
                                  shard_tables=shard)
        assert tr.shard == shard and (len(tr.cat_tables) >= 2) == lazy_cat
        if shard:       # the full replicas are gone: each rank holds every other row
            assert m.dare.emb_att.weight.numel() == 0 and tr.tables["dare.emb_att.weight"]["p"].shape[0] == (g_vocab(cfg) + world - 1) // world
        for s in range(3):
            batch, y = batch_of(z, s)
            B = y.shape[0]
            lo, hi = rank * B // world, (rank + 1) * B // world
            tr.step({k: v[lo:hi] for k, v in batch.items()}, y[lo:hi])
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}       # sharded: re-assembles the tables (collective)
        if rank == 0:
            torch.save(sd, out_path)
    finally:
        dist.destroy_process_group()


def g_vocab(cfg):
    return cfg["golden"]["seq_vocab"]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("shard,lazy_cat", [(False, False), (True, False), (True, True)],
                         ids=["replicated_tables", "row_sharded_tables_p2p", "row_sharded_plus_row_lazy_categorical_tables"])
def test_two_rank_data_parallel_matches_single_gpu_global_batch(tmp_path, shard, lazy_cat):
    import torch.multiprocessing as mp
    import b200ctr
    z, cfg = load_case("qnn_concat")
    out = str(tmp_path / "dp.pt")
    port = 29600 + (os.getpid() % 1000) + (500 if shard else 0) + (250 if lazy_cat else 0)
    mp.spawn(_worker, args=(2, port, out, shard, lazy_cat), nprocs=2, join=True)
    dp = torch.load(out)
    dev = torch.device("cuda:0")
    m = _build(cfg, z, dev)
    tr = b200ctr.FusedTrainer(m, cfg, steps_per_epoch=cfg["golden"]["steps_per_epoch"])
    for s in range(3):
        batch, y = batch_of(z, s)
        tr.step(batch, y)
    ref = m.state_dict()
    for k, v in ref.items():
        assert max_err_scaled(dp[k], v.cpu()) < 1e-3, k
    # and both equal the reference's own 3-step result
    for k, v in group(z, "final/").items():
        assert max_err_scaled(dp[k], v) < 1e-3, k
