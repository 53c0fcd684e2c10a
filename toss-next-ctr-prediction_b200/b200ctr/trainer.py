"""Fused training step (the fast tier): mirrors the step body of the reference's train loop
(src/train.py:152-203) without autograd and without per-parameter optimizer launches.

    lr <- cosine_warmup_lr        (utils/sched.py:3-11)
    logits, aux <- model(batch)   (hotpath.model_fwd, dropout on)
    loss <- bce_wll(logits) + aux_w * bce_wll(aux)            (train.py:71-90,165-168; one kernel, no host sync)
    backward                       (hotpath.model_bwd into ONE flat gradient buffer + row-sparse table grads)
    clip_grad_norm_                (global L2 over dense grads + the distinct touched table rows; coef stays on device)
    AdamW                          (one fused kernel over the flat buffer, one per big table)
    EMA                            (fused into the AdamW kernels; utils/ema.py:91-131)

Layout: every "small" parameter (everything except the two seq_vocab x D tables) lives in one flat fp32 buffer
with 16-byte aligned segments; the nn.Module's parameters are re-pointed at views of it, so state_dict(), eval and
checkpoints keep working on the same storage.  Parameters that never receive a gradient under the cfg (ctx_mlp in
S1 mode, aux_head when aux_head_weight == 0) sit in a tail segment the optimizer skips: torch.optim.AdamW skips
grad-less parameters entirely (no weight decay either), and this reproduces that.

The two big tables keep separate p/m/v/ema buffers, a persistent all-zero gradient table and a per-row lazy state: the
backward kernel accumulates the K touched rows per sample into the gradient table, the rows a batch reads / touches are
kept as bitmaps (one bit per row), and the row-lazy optimizer (csrc/rowopt.cu) catches rows up before they are read, applies
step t to the touched rows and re-zeroes their gradient rows.  Everything that reads the tables outside FusedTrainer.step
(state_dict, plain-module forward for validation, EMA swap-in) flushes them first through module hooks.

Data parallel (one process per GPU): dense grads are all-reduced (NCCL), class counts (C3), the SE squeeze (C4)
and the table-row gradients are exchanged so that N ranks with B/N samples each reproduce the 1-GPU global batch.

Row-sharded sequence tables (shard_tables=True, the default for world > 1): row r of emb_att / emb_rep -- and its
Adam moments, EMA shadow, gradient row and lazy-optimizer state -- lives only on rank r % N at local row r // N.
Every rank maps the other ranks' parameter and gradient shards through CUDA IPC (peer memory over NVLink), so the
retrieval kernel gathers remote rows with plain loads; the backward kernel reduces locally-owned gradient rows in place and
writes rows owned by other ranks into the owner's peer-mapped STAGING buffer (stage[owner][src_rank][slot]), which the owner
adds into its shard after the step's second all-gather (remote fp32 atomics over NVLink were 7x slower).  Per step the owners
need to know which of their rows are read / touched: one all-gather of the row bitmaps (before the forward) and one of the K
selected tokens (after the backward); those two collectives double as the device-side barriers that order "owner caught the
rows up" -> "peers read them" and "peers stored gradients" -> "owner applies AdamW".
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import torch

from . import hotpath as hp
from . import ops
from .modules import CTRModel


def cosine_warmup_lr(epoch, step, steps_per_epoch, base_lr, warmup_epochs=1, total_epochs=10):
    """utils/sched.py:3-11."""
    g = epoch * steps_per_epoch + step
    warm = warmup_epochs * steps_per_epoch
    total = total_epochs * steps_per_epoch
    if g < warm:
        return base_lr * (g + 1) / max(1, warm)
    prog = (g - warm) / max(1, total - warm)
    return 0.5 * base_lr * (1.0 + math.cos(math.pi * prog))


def ema_decay_at(n_updates: int, base: float, warmup_steps: int = 0, warmup_type: str = "linear") -> float:
    """utils/ema.py:72-88."""
    d = base
    if warmup_steps <= 0 or warmup_type == "none":
        return d
    t = min(1.0, (n_updates + 1) / warmup_steps)
    if warmup_type == "linear":
        dec = 1.0 - (1.0 - d) * t
    elif warmup_type == "cosine":
        dec = 1.0 - (1.0 - d) * 0.5 * (1 + math.cos(math.pi * (1 - t)))
    else:
        dec = d
    return float(max(0.0, min(1.0, dec)))


BIG_TABLES = ("dare.emb_att.weight", "dare.emb_rep.weight")
import os as _os
C5_MODE = _os.environ.get("B200CTR_C5", "gather")        # "allreduce": NCCL int64 sum (A/B timing)


class FusedTrainer:
    def __init__(self, model: CTRModel, cfg: dict, steps_per_epoch: int = 1, seed: Optional[int] = None,
                 process_group=None, lazy_tables: bool = True, shard_tables: Optional[bool] = None):
        self.model, self.cfg = model, cfg
        self.spec: hp.ModelSpec = model._spec
        spec = self.spec
        tr = cfg["train"]
        self.base_lr, self.wd = float(tr["lr"]), float(tr["weight_decay"])
        self.clip = float(tr.get("grad_clip_norm", 0.0))
        self.warmup_epochs, self.epochs = tr.get("warmup_epochs", 1), int(tr.get("epochs", 10))      # fractional warm-up allowed, as in sched.py
        self.steps_per_epoch = int(steps_per_epoch)
        ema = cfg.get("ema") or {}
        self.ema_on = bool(ema.get("enabled", False))
        self.ema_decay = float(ema.get("decay", 0.999))
        self.ema_warm, self.ema_type = int(ema.get("warmup_steps", 0)), str(ema.get("warmup_type", "linear"))
        self.seed = int(cfg.get("seed", 0) if seed is None else seed)
        amp = str(cfg.get("amp", "none")).lower()
        if amp not in ("none", "bf16"):
            raise ValueError("FusedTrainer supports amp: none (fp32) and amp: bf16; fp16 + GradScaler is not built")
        self.prec = "bf16" if amp == "bf16" else "fp32"
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        params = dict(model.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer needs the model on a CUDA device (no CPU fallback)")
        self.dev = dev
        # parameters that never get a gradient under this cfg (torch AdamW skips them)
        frozen = set()
        if spec.query_mode == "S1":
            frozen |= {n for n in params if n.startswith("ctx_mlp.")}
        if spec.aux_weight <= 0:
            frozen |= {n for n in params if n.startswith("dare.aux_head.")}
        # Hashed-categorical tables with an HBM-sized bucket count (cfgs/v4_k120_s1.yaml: 2,000,503 x 32 and 4,000,513 x 57;
        # dare_base: 2,000,000 x 64 and 1,000,000 x 64) are kept OUT of the flat dense buffer and updated by the row-lazy
        # optimizer like the sequence tables (a dense AdamW + EMA pass over them would cost 36 B per parameter per step, and a
        # dense all-reduce under data parallelism).  Under DP they stay REPLICATED (1.2 GB at the v4 sizes) and exchange sparsely:
        # every rank all-gathers its B slots per table (row id + 64-wide contribution row) and reduces all ranks' slots with
        # the same sort + segmented reduction, so every replica applies bit-identical row updates.
        thr = int(_os.environ.get("B200CTR_LAZY_THR", (cfg.get("b200ctr") or {}).get("lazy_table_min_rows", 500_000)))
        self.cat_lazy = [c for c in spec.cat_cols if lazy_tables and params[f"cat_embs.{c}.weight"].shape[0] >= thr]
        lazy_names = {f"cat_embs.{c}.weight" for c in self.cat_lazy}
        small = [n for n in params if n not in BIG_TABLES and n not in lazy_names]
        active = [n for n in small if n not in frozen] + [n for n in small if n in frozen]
        self.offsets: Dict[str, int] = {}
        off = 0
        n_active = 0
        for n in active:
            if n in frozen:
                continue
            self.offsets[n] = off
            off += (params[n].numel() + 3) // 4 * 4
        n_active = off
        self.slot0 = off          # 4 reserved floats between the active and the frozen segment: the step's loss rides on the
        off += 4                  # dense-gradient all-reduce there (one collective less per step)
        for n in active:
            if n in frozen:
                self.offsets[n] = off
                off += (params[n].numel() + 3) // 4 * 4
        self.n_flat, self.n_active = off, n_active
        # head (qnn.* / fc.*) gradients are final first in the backward pass: if they form the tail of the active segment their
        # all-reduce is launched asynchronously under the encoder / embedding backward
        head = [n for n in active if n not in frozen and (n.startswith("qnn.") or n.startswith("fc."))]
        self.head_off = min((self.offsets[n] for n in head), default=n_active)
        if any(self.offsets[n] > self.head_off for n in active if n not in frozen and n not in head):
            self.head_off = n_active                       # not a contiguous tail: single all-reduce
        self.flat_p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n_active, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n_active, device=dev, dtype=torch.float32)
        self.P: Dict[str, torch.Tensor] = {}
        self.Gv: Dict[str, torch.Tensor] = {}
        for n in active:
            p = params[n]
            o = self.offsets[n]
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view                       # the module now aliases the flat buffer
            self.P[n] = view
            self.Gv[n] = self.flat_g[o:o + p.numel()].view(p.shape)
        self.flat_ema = self.flat_p.clone() if self.ema_on else None
        self.tables: Dict[str, dict] = {}
        self.vocab_global = params[BIG_TABLES[0]].shape[0]
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self.shard = (self.world > 1) if shard_tables is None else (bool(shard_tables) and self.world > 1)
        self.lazy = bool(lazy_tables)
        if self.shard:
            if not self.lazy:
                raise ValueError("row-sharded tables use the row-lazy optimizer (lazy_tables=True)")
            self._init_sharded_tables(params)
        else:
            for n in BIG_TABLES:
                p = params[n].data
                self.P[n] = p
                t = dict(p=p, m=torch.zeros_like(p), v=torch.zeros_like(p), g=torch.zeros_like(p),
                         ema=p.clone() if self.ema_on else None)
                self.tables[n] = t
                self.Gv[n] = t["g"]
        self.vocab = self.tables[BIG_TABLES[0]]["p"].shape[0]          # rows held by this rank
        self.cat_tables: Dict[str, dict] = {}
        if self.cat_lazy:
            cards = [params[f"cat_embs.{c}.weight"].shape[0] for c in spec.cat_cols]
            # ONE persistent all-zero (sum of cards, 64) accumulator in global-row order: cat_bwd leaves the reduced 64-wide
            # gradient rows of the row-lazy tables here (their `g`), the optimizer consumes and re-zeroes them
            self.cat_rows_acc = torch.zeros((sum(cards), 64), device=dev, dtype=torch.float32)
            keep = torch.zeros(len(cards), dtype=torch.int32)
            dslot = torch.full((len(cards),), -1, dtype=torch.int32)
            base = 0
            for ci, c in enumerate(spec.cat_cols):
                if c in self.cat_lazy:
                    n = f"cat_embs.{c}.weight"
                    full = params[n].data
                    card, dcw = full.shape
                    p = torch.zeros((card, 64), device=dev, dtype=torch.float32)       # stored 64 wide: the row kernels' width
                    p[:, :dcw].copy_(full)
                    params[n].data = p[:, :dcw]                                        # the module sees the same storage
                    self.P[n] = p[:, :dcw]
                    t = dict(p=p, m=torch.zeros_like(p), v=torch.zeros_like(p), g=self.cat_rows_acc[base:base + card],
                             ema=p.clone() if self.ema_on else None, state=torch.zeros(card, device=dev, dtype=torch.int32),
                             words=ops.bitmap_words(card), col=ci, dc=dcw, card=card, base=base)
                    t["bm"] = torch.zeros(t["words"], device=dev, dtype=torch.int32)
                    self.cat_tables[n] = t
                    self.Gv[n] = t["g"]
                    keep[ci] = 1
                    dslot[ci] = len(self.cat_tables) - 1
                base += cards[ci]
            self.cat_keep_rows = keep.to(dev)
            self.cat_keep_rows._host = [int(x) for x in keep]          # host copy for the launch geometry (ops.cat_bwd)
            self.cat_defer_slot = dslot.to(dev) if self.world > 1 else None
        if self.lazy:
            # initial capacity only: RowHistory grows geometrically when the run outlives the estimate
            cap = min(max(self.epochs * self.steps_per_epoch + 64, 4096), 1 << 22)
            self.hist = ops.RowHistory(dev, cap)
            for n in BIG_TABLES:
                self.tables[n]["state"] = torch.zeros(self.vocab, device=dev, dtype=torch.int32)
            # state_dict()/checkpointing through the nn.Module sees fully-updated (and, if sharded, re-assembled) tables
            model.register_state_dict_pre_hook(lambda *a, **k: self.materialize_tables())
            # the reference's validation pattern -- model.eval(); model(batch) (train.py:212-222), ema.store()/copy_to() --
            # goes through the nn.Module: the lazily-updated rows must be brought to the current step first (sharded mode:
            # the full tables are re-assembled, a collective every rank reaches because every rank validates)
            model.register_forward_pre_hook(lambda *a, **k: self._before_module_forward())
        self.owner = torch.empty(max(self.vocab, 1), device=dev, dtype=torch.int32)
        # row sets as bitmaps (one bit per global row) unless the tables are replicated under DP (legacy all-gather path)
        self.bitmaps = self.lazy and (self.world == 1 or self.shard)
        self.n_shards = self.world if self.shard else 1
        self.shard_rank = self.rank if self.shard else 0
        self.bm_words = ops.bitmap_words(self.vocab_global)
        if self.bitmaps:
            self.bm_mine = torch.zeros(self.bm_words, device=dev, dtype=torch.int32)
            self.bm_all = torch.zeros(self.n_shards * self.bm_words, device=dev, dtype=torch.int32) if self.shard else None
        self.acc = torch.zeros(1, device=dev, dtype=torch.float64)
        self.acc_all = torch.zeros(max(1, self.world), device=dev, dtype=torch.float64)
        self.coef = torch.ones(1, device=dev, dtype=torch.float32)
        self._bar = torch.zeros(4, device=dev, dtype=torch.float32)      # payload of the barrier all-reduces (+ class counts)
        self.total_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.global_step = 0
        self._equal_batch = -1
        self.comm_prof = None            # bench.py: list of (tag, start_event, end_event) per collective
        self._stg = None                 # stage(): copy stream + two preallocated device staging sets
        self._flushed_at = 0             # global_step at which every table row was last brought up to date (sync_tables)
        self.n_ema = 0
        self.cache: dict = {}
        if self.cat_lazy:
            self.cache["cat_rows_acc"], self.cache["cat_keep_rows"] = self.cat_rows_acc, self.cat_keep_rows
            if self.world > 1:
                self.cache["cat_defer_slot"], self.cache["cat_defer_n"] = self.cat_defer_slot, len(self.cat_tables)
        defaults = {"update_after_step": 0, "update_interval": 1, "param_filter": [], "offload_to_cpu": False}
        for k, dv in defaults.items():
            v = ema.get(k, dv)
            if self.ema_on and (list(v) if k == "param_filter" and v is not None else v) not in (dv, None):
                raise ValueError(f"FusedTrainer: ema.{k}={v!r} is not supported by the fused AdamW+EMA kernels (only the default "
                                 f"{dv!r}: update every step from step 0, all parameters, shadows on the device)")

    # ---- row-sharded tables ----------------------------------------------------------------------------------
    def _init_sharded_tables(self, params):
        """Keep rows rank::world of each big table (+ optimizer state) and map every peer's parameter / gradient
        shard (cudaIpcOpenMemHandle with lazy peer access => NVLink loads / atomics from inside the kernels)."""
        import torch.distributed as dist
        N, rank, dev = self.world, self.rank, self.dev
        if N & (N - 1) or N > 8:
            raise ValueError(f"row-sharded tables need a power-of-two world size <= 8, got {N}")
        rows_local = (self.vocab_global + N - 1) // N
        for n in BIG_TABLES:
            full = params[n].data
            D = full.shape[1]
            p = torch.zeros((rows_local, D), device=dev, dtype=torch.float32)
            mine = full[rank::N]
            p[:mine.shape[0]].copy_(mine)
            g = torch.zeros_like(p)
            t = dict(p=p, m=torch.zeros_like(p), v=torch.zeros_like(p), g=g, ema=p.clone() if self.ema_on else None)
            self.tables[n] = t
            params[n].data = torch.empty((0, D), device=dev, dtype=torch.float32)     # the full replica is released
            del full, mine
            shards = {}
            for kind, tens in (("p", p), ("g", g)):
                handles = [None] * N
                dist.all_gather_object(handles, ops.ipc_export(tens), group=self.pg)
                lst = [tens if r == rank else ops.ipc_import(*handles[r]) for r in range(N)]
                shards[kind] = ops.ShardedTable(lst, self.vocab_global, D)
            self.P[n] = shards["p"]
            self.Gv[n] = shards["g"]
        torch.cuda.empty_cache()
        dist.barrier(group=self.pg)

    def _ensure_staging(self, n_slots: int):
        """Receive buffers of the staged gradient scatter: stage[n][src_rank, slot, :] on every rank, peer-mapped.  COLLECTIVE
        on first use (and if a later batch is larger than any seen before)."""
        import torch.distributed as dist
        if getattr(self, "stage_cap", 0) >= n_slots:
            return
        N, D = self.world, self.spec.D
        self.stage_cap = int(n_slots)
        for n in BIG_TABLES:
            buf = torch.empty((N, self.stage_cap, D), device=self.dev, dtype=torch.float32)
            self.tables[n]["stage"] = buf
            handles = [None] * N
            dist.all_gather_object(handles, ops.ipc_export(buf), group=self.pg)
            addrs = [buf.data_ptr() if r == self.rank else ops.ipc_import(*handles[r]) for r in range(N)]
            gsh = self.Gv[n]
            gsh.stage_ptrs = (ctypes.c_void_p * N)(*addrs)
            gsh.stage_rank, gsh.stage_cap = self.rank, self.stage_cap
        dist.barrier(group=self.pg)

    def _row_set(self, ids: torch.Tensor):  # noqa: D401
        """Bitmap(s) of the global rows listed in `ids`: (maps, n_maps).  Sharded mode: all ranks' bitmaps, all-gathered
        (1.25 MB per rank at 10 M rows; the collective doubles as a device-side barrier)."""
        self.bm_mine.zero_()
        ops.bitmap_mark(ids, self.spec.pad_id, self.vocab_global, self.bm_mine)
        if not self.shard:
            return self.bm_mine, 1
        self._allgather(self.bm_all, self.bm_mine, "C1_bitmaps_allgather")
        return self.bm_all, self.world

    def materialize_tables(self):
        """Fully-updated tables behind the nn.Module's parameters (state_dict / checkpoint / plain-module eval).
        Sharded mode: COLLECTIVE -- every rank re-assembles the full (vocab, D) tables from the shards."""
        self.sync_tables()
        if not self.shard:
            return
        params = dict(self.model.named_parameters())
        N = self.world
        for n in BIG_TABLES:
            p = self.tables[n]["p"]
            buf = torch.empty((N,) + tuple(p.shape), device=self.dev, dtype=torch.float32)
            torch.distributed.all_gather_into_tensor(buf, p, group=self.pg)
            params[n].data = buf.permute(1, 0, 2).reshape(N * p.shape[0], p.shape[1])[:self.vocab_global].contiguous()

    def _before_module_forward(self):
        """Plain-module forward while this trainer owns the tables lazily: flush pending decay / momentum / EMA steps (a no-op
        pass when nothing is pending), and in sharded mode materialise the replicas the module parameters point at."""
        if self._flushed_at != self.global_step or (self.shard and dict(self.model.named_parameters())[BIG_TABLES[0]].shape[0] == 0):
            self.materialize_tables()

    def release_tables(self):
        """Drop the re-assembled replicas again (sharded mode)."""
        if self.shard:
            params = dict(self.model.named_parameters())
            for n in BIG_TABLES:
                params[n].data = torch.empty((0, self.spec.D), device=self.dev, dtype=torch.float32)

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _G(self, name: str) -> torch.Tensor:
        return self.Gv[name]

    def to_device(self, batch: Dict[str, torch.Tensor], y: Optional[torch.Tensor]):
        """Host (pinned) -> device copies of one collate_sharded batch (dataset.py:98-124)."""
        dev = self.dev
        out = {"X_num": batch["X_num"].to(dev, non_blocking=True), "X_mask": batch["X_mask"].to(dev, non_blocking=True),
               "X_cat": batch["X_cat"].to(dev, non_blocking=True), "seq": batch["seq"].to(dev, non_blocking=True)}
        if out["X_num"].dtype != torch.float32:
            out["X_num"] = out["X_num"].float()
        if out["X_mask"].dtype != torch.float32:
            out["X_mask"] = out["X_mask"].float()
        yd = y.to(dev, non_blocking=True).float() if y is not None else None
        return out, yd

    def lr_at(self, global_step_1based: int) -> float:
        g = global_step_1based - 1
        epoch, step = divmod(g, max(1, self.steps_per_epoch))
        return cosine_warmup_lr(epoch, step, self.steps_per_epoch, self.base_lr, self.warmup_epochs, self.epochs)

    def _comm(self, tag: str, fn):
        """Run one collective; when bench.py switched `comm_prof` on, bracket it with CUDA events on the launch stream."""
        if self.comm_prof is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.comm_prof.append((tag, a, b))
        return out

    def _allreduce(self, t: torch.Tensor, tag: str = "allreduce"):
        if self.pg is not None and self.world > 1:
            self._comm(tag, lambda: torch.distributed.all_reduce(t, group=self.pg))
        return t

    def _allgather(self, out: torch.Tensor, inp: torch.Tensor, tag: str):
        self._comm(tag, lambda: torch.distributed.all_gather_into_tensor(out, inp, group=self.pg))

    def _check_equal_batch(self, B_local: int):
        """Data parallel: staging slots, all_gather_into_tensor sizes and B_total = B_local * world all assume one batch size
        per rank.  Verified collectively at first use and again whenever the local batch size changes (a ragged last batch
        must change on every rank alike: pad or drop it otherwise)."""
        if self.world == 1 or B_local == self._equal_batch:
            return
        mm = torch.tensor([B_local, -B_local], device=self.dev, dtype=torch.int64)
        torch.distributed.all_reduce(mm, op=torch.distributed.ReduceOp.MAX, group=self.pg)
        lo, hi = -int(mm[1]), int(mm[0])
        if lo != hi:
            raise ValueError(f"FusedTrainer: per-rank batch sizes differ ({lo}..{hi}); every rank must step with the same B "
                             f"(pad or drop the ragged last batch)")
        self._equal_batch = B_local

    # ---- one optimisation step ---------------------------------------------------------------------------------
    def stage(self, batch: Dict[str, torch.Tensor], y: torch.Tensor) -> None:
        """Start the host-to-device copy of one collate_sharded batch (CPU, ideally pinned) into one of the trainer's two
        PREALLOCATED device staging sets, on the trainer's copy stream; the next `step()` called without a batch consumes it.
        Called right after `step()` returned (the step is enqueued, the GPU busy) the next batch crosses PCIe under the
        running step -- the overlap the reference gets from DataLoader workers + `non_blocking=True` (train.py:106-110, 152-156):

            tr.stage(*batches[0])
            for i in range(n):
                loss = tr.step()                       # consumes the staged batch
                if i + 1 < n: tr.stage(*batches[i + 1])
                loss_value = float(loss)               # train.py:202

        The staging sets are allocated once per (shape, dtype) signature and reused: copies into FRESH allocations on a side
        stream made the overlapped form 2x SLOWER than copy-then-step (the caching allocator has to cudaMalloc while the GPU is
        busy and defers the reuse of blocks shared between streams; tools/e2e_ab.py: 12-19 ms against 9.35 ms with static sets,
        9.86 ms copy-then-step).  At most two batches can be staged ahead."""
        dev = self.dev
        if self._stg is None:
            self._stg = {"cs": torch.cuda.Stream(dev), "b": [None, None], "y": [None, None], "sig": [None, None],
                         "ready": [torch.cuda.Event(), torch.cuda.Event()], "used": [torch.cuda.Event(), torch.cuda.Event()],
                         "w": 0, "pending": []}
            cur = torch.cuda.current_stream(dev)
            for e in self._stg["used"]:
                e.record(cur)
        st = self._stg
        if len(st["pending"]) >= 2:
            raise RuntimeError("FusedTrainer.stage: two batches are already staged; call step() first")
        slot = st["w"]
        keys = ("X_num", "X_mask", "X_cat", "seq")
        sig = tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in keys) + (("y", tuple(y.shape), y.dtype),)
        if st["sig"][slot] != sig:       # first use of the slot, or another batch shape (a ragged last batch)
            st["b"][slot] = {k: torch.empty(batch[k].shape, device=dev, dtype=batch[k].dtype) for k in keys}
            st["y"][slot] = torch.empty(y.shape, device=dev, dtype=y.dtype)
            st["sig"][slot] = sig
            # the allocator may hand out memory whose previous tenant is still in use by kernels queued on this stream
            st["cs"].wait_stream(torch.cuda.current_stream(dev))
        cs = st["cs"]
        cs.wait_event(st["used"][slot])                 # the step that read this slot has finished
        with torch.cuda.stream(cs):
            for k in keys:
                st["b"][slot][k].copy_(batch[k], non_blocking=True)
            st["y"][slot].copy_(y, non_blocking=True)
            st["ready"][slot].record(cs)
        st["pending"].append(slot)
        st["w"] = slot ^ 1

    def step(self, batch: Optional[Dict[str, torch.Tensor]] = None, y: Optional[torch.Tensor] = None,
             lr: Optional[float] = None) -> torch.Tensor:
        """batch: collate_sharded dict (CPU pinned or device tensors); None = the batch handed to `stage()` earlier.  Returns
        the loss as a 1-element device tensor (no host sync; the caller decides when to read it)."""
        slot = None
        if batch is None:
            if self._stg is None or not self._stg["pending"]:
                raise RuntimeError("FusedTrainer.step() without a batch needs a batch staged with stage()")
            slot = self._stg["pending"].pop(0)
            torch.cuda.current_stream(self.dev).wait_event(self._stg["ready"][slot])
            batch, y = self._stg["b"][slot], self._stg["y"][slot]
        with ops.precision(self.prec):
            loss = self._step(batch, y, lr)
        if slot is not None:
            self._stg["used"][slot].record(torch.cuda.current_stream(self.dev))
        return loss

    def _step(self, batch, y, lr):
        spec = self.spec
        if not spec.use_qnn:
            spec.p_fc = float(self.model.fc[2].p)
        self._check_equal_batch(int(y.shape[0]))
        self.global_step += 1
        lr = self.lr_at(self.global_step) if lr is None else lr
        dbatch, yd = self.to_device(batch, y)
        B_local = yd.shape[0]
        # every rank draws its own masks (same seed on all ranks would give local row i identical masks on every rank)
        dc = hp.DropCtl(True, seed=self.seed + 0x9E3779B1 * self.rank, step=self.global_step)
        counts_ready = None
        self.flat_g.zero_()
        d = ema_decay_at(self.n_ema, self.ema_decay, self.ema_warm, self.ema_type) if self.ema_on else 0.0
        if self.lazy:
            # rows this batch is about to read are brought to step t-1 (exactly what dense AdamW/EMA would hold)
            self.hist.push(self.global_step, lr, self.wd, d if self.ema_on else 1.0)
            if self.bitmaps:
                # the set of rows this step reads, one bit per global row; sharded: the all-gather of the ranks' bitmaps
                # tells every owner which of ITS rows any rank will read (and is the barrier that orders the previous
                # step's owner-side AdamW before this step's peer reads)
                maps, n_maps = self._row_set(dbatch["seq"])
                if self.global_step > 1:
                    for n in BIG_TABLES:
                        t = self.tables[n]
                        ops.rowopt_catchup_bm(maps, n_maps, self.bm_words, self.n_shards, self.shard_rank, t["state"], t["p"],
                                              t["m"], t["v"], t["ema"], self.global_step - 1, self.hist)
                if self.shard:
                    # barrier: every shard is caught up before any peer gathers from it.  Its payload carries the class counts
                    # (C3) so they need no collective of their own (exact in fp32: counts <= 2^24)
                    self._bar[0] = 0.0
                    self._bar[1] = (yd > 0.5).sum()
                    self._bar[2] = (yd <= 0.5).sum()
                    self._allreduce(self._bar, "barrier+C3_counts")
                    counts_ready = self._bar[1:3].to(torch.int32)
            elif self.global_step > 1:
                seq = dbatch["seq"]
                ops.rows_claim_any(seq, spec.pad_id, self.vocab, self.owner)
                for n in BIG_TABLES:
                    t = self.tables[n]
                    ops.rowopt_catchup(seq, spec.pad_id, self.owner, t["state"], t["p"], t["m"], t["v"], t["ema"],
                                       self.global_step - 1, self.hist)
        for n, t in self.cat_tables.items():
            # rows of the row-lazy categorical tables this batch reads: mark, then bring them to step t-1
            ids = dbatch["X_cat"][:, t["col"]].contiguous()
            t["bm"].zero_()
            ops.bitmap_mark(ids, -1, t["card"], t["bm"])
            if self.global_step > 1:
                ops.rowopt_catchup_bm(t["bm"], 1, t["words"], 1, 0, t["state"], t["p"], t["m"], t["v"], t["ema"], self.global_step - 1,
                                      self.hist)
        if self.shard:
            self._ensure_staging(B_local * min(spec.top_k, dbatch["seq"].shape[1]))
        hook = (lambda t: self._allreduce(t, "C4_se_squeeze")) if self.world > 1 else None
        logits, aux, sv = hp.model_fwd(self.P, spec, dbatch, dc, self.cache, B_total=B_local * self.world, colsum_hook=hook)
        counts_override = None
        if self.world > 1:
            if counts_ready is not None:
                counts_override = counts_ready
            else:
                counts_override = torch.stack([(yd > 0.5).sum(), (yd <= 0.5).sum()]).to(torch.int32)   # 2 ints (C3)
                self._allreduce(counts_override, "C3_counts")
        loss, dz, daux, _, _ = ops.bce_wll(logits, aux, yd, spec.aux_weight, counts_override, want_prob=False)
        pending = []

        def head_grads_final():
            """C2, first part: the head's gradients (28 of the 37.6 MB: qnn.mlp.0.weight) + the loss slot, all-reduced on
            NCCL's stream while the encoder / embedding backward keeps this stream busy."""
            # the SE gate's weight gradients are functions of batch-global sums (already all-reduced inside
            # qnn_bwd), hence identical on every rank: pre-divide so the dense all-reduce(sum) leaves them unchanged
            if spec.use_qnn and spec.use_se:
                for n in ("qnn.se.fc.0.weight", "qnn.se.fc.0.bias", "qnn.se.fc.2.weight", "qnn.se.fc.2.bias"):
                    g = self.Gv[n].view(-1)
                    ops.axpby(g, None, 1.0 / self.world, 0.0, out=g)
            self.flat_g[self.slot0:self.slot0 + 1].copy_(loss.view(1))
            seg = self.flat_g[self.head_off:self.slot0 + 4]
            pending.append(self._comm("C2_head_async", lambda: torch.distributed.all_reduce(seg, group=self.pg, async_op=True)))

        overlap = self.world > 1 and self.head_off < self.n_active
        hp.model_bwd(self.P, spec, sv, dz, daux if spec.aux_weight > 0 else None, dc, self._G, colsum_hook=hook,
                     after_head=head_grads_final if overlap else None)
        tok = sv["dare"]["tok"]
        if self.world > 1 and self.cat_tables:
            # replicated row-lazy categorical tables: gather every rank's slots, reduce them once -- in (rank, sample, table)
            # order, the same on every replica -- into the row accumulator, and mark the union of the touched rows
            df = sv["cat_defer"]
            rows_all = torch.empty((self.world,) + tuple(df["rows"].shape), device=self.dev, dtype=torch.float32)
            keys_all = torch.empty((self.world,) + tuple(df["keys"].shape), device=self.dev, dtype=torch.int32)
            self._allgather(rows_all.view(-1), df["rows"].view(-1), "C1_cat_rows_allgather")
            self._allgather(keys_all.view(-1), df["keys"].view(-1), "C1_cat_keys_allgather")
            n_sl = keys_all.numel()
            ops.scatter_rows(keys_all.view(-1), 1, n_sl, -1, self.cat_rows_acc.shape[0], 1, None, None, rows_all.view(n_sl, spec.D),
                             None, None, 0, 1, 0, None, self.cat_rows_acc)
            for k, t in enumerate(self.cat_tables.values()):
                ids = (keys_all[:, :, k] - t["base"]).contiguous()
                t["bm"].zero_()
                ops.bitmap_mark(ids, -1, t["card"], t["bm"])
        if self.world > 1:
            if overlap:
                self._allreduce(self.flat_g[:self.head_off], "C2_rest")                             # C2, second part
                self._comm("C2_head_wait", lambda: pending[0].wait())
            else:
                if spec.use_qnn and spec.use_se:
                    for n in ("qnn.se.fc.0.weight", "qnn.se.fc.0.bias", "qnn.se.fc.2.weight", "qnn.se.fc.2.bias"):
                        g = self.Gv[n].view(-1)
                        ops.axpby(g, None, 1.0 / self.world, 0.0, out=g)
                self.flat_g[self.slot0:self.slot0 + 1].copy_(loss.view(1))
                self._allreduce(self.flat_g[:self.slot0 + 4], "C2")                                 # C2 (+ loss slot)
            loss = self.flat_g[self.slot0:self.slot0 + 1].clone()
            if not self.shard:
                tok = self._exchange_table_grads(tok)                                               # C1 (replicated tables)
        # ---- global-norm clip (coef stays on the device) ----
        self.acc.zero_()
        if self.lazy and self.bitmaps:
            if self.shard:
                # C1 happened inside retrieve_bwd: rows owned by other ranks were stored into the owners' staging buffers over
                # NVLink.  The all-gather of the selected token ids tells every owner which staged rows are its own (and is the
                # barrier that orders every peer's stores before the owner-side reduction).
                n_slots = tok.numel()
                gtok = torch.empty((self.world * n_slots,), device=self.dev, dtype=torch.int32)
                self._allgather(gtok, tok.reshape(-1), "C1_selected_tokens_allgather")
                if ops.SCATTER == "sorted":
                    # every rank staged ALL its slots' rows at their owners: reduce each owned row once, in (rank, slot) order
                    ops.scatter_rows(gtok, self.world, n_slots, spec.pad_id, self.vocab_global, tok.shape[1], None, None, None,
                                     self.tables[BIG_TABLES[0]]["stage"], self.tables[BIG_TABLES[1]]["stage"], self.stage_cap,
                                     self.world, self.rank, self.tables[BIG_TABLES[0]]["g"], self.tables[BIG_TABLES[1]]["g"])
                else:
                    ops.stage_apply(gtok, self.world, n_slots, self.stage_cap, spec.pad_id, self.world, self.rank,
                                    self.tables[BIG_TABLES[0]]["stage"], self.tables[BIG_TABLES[1]]["stage"],
                                    self.tables[BIG_TABLES[0]]["g"], self.tables[BIG_TABLES[1]]["g"])
                self.bm_mine.zero_()
                ops.bitmap_mark(gtok, spec.pad_id, self.vocab_global, self.bm_mine)
                maps, n_maps = self.bm_mine, 1
            else:
                maps, n_maps = self._row_set(tok)
            for n in BIG_TABLES:
                ops.rows_sqnorm_bm(maps, n_maps, self.bm_words, self.n_shards, self.shard_rank, self.tables[n]["g"], self.acc)
            if self.shard:
                # C5: each owner contributes the squared norm of its rows.  The accumulator is a fixed-point 64-bit integer
                # (order-independent sum); the ranks' 8-byte slots are all-gathered and added as integers
                if C5_MODE == "gather":
                    self._allgather(self.acc_all, self.acc, "C5_sqnorm")
                    self.acc.view(torch.int64).copy_(self.acc_all.view(torch.int64).sum(0, keepdim=True))
                else:
                    self._allreduce(self.acc.view(torch.int64), "C5_sqnorm")
        else:
            ops.rows_claim(tok, spec.pad_id, self.owner)
            for n in BIG_TABLES:
                ops.rows_sqnorm(tok, spec.pad_id, self.owner, self.tables[n]["g"], self.acc)
        for n, t in self.cat_tables.items():      # the rows a batch reads are the rows that received a gradient: same bitmap
            ops.rows_sqnorm_bm(t["bm"], 1, t["words"], 1, 0, t["g"], self.acc)
        ops.sqnorm_(self.flat_g[:self.n_active], self.acc)
        ops.clip_coef(self.acc, self.clip, self.coef, self.total_norm)
        # ---- AdamW (+EMA) ----
        ops.adamw_ema_(self.flat_p[:self.n_active], self.flat_g[:self.n_active], self.flat_m, self.flat_v,
                       self.flat_ema[:self.n_active] if self.ema_on else None, lr, self.wd, self.global_step, self.coef, d)
        for n in BIG_TABLES:
            t = self.tables[n]
            if self.lazy and self.bitmaps:
                ops.rowopt_update_bm(maps, n_maps, self.bm_words, self.n_shards, self.shard_rank, t["state"], t["p"], t["g"],
                                     t["m"], t["v"], t["ema"], self.global_step, self.coef, self.hist)
            elif self.lazy:
                ops.rowopt_update(tok, spec.pad_id, self.owner, t["state"], t["p"], t["g"], t["m"], t["v"], t["ema"],
                                  self.global_step, self.coef, self.hist)
            else:
                ops.adamw_ema_(t["p"], t["g"], t["m"], t["v"], t["ema"], lr, self.wd, self.global_step, self.coef, d)
                ops.rows_zero(tok, spec.pad_id, self.owner, t["g"])
        for n, t in self.cat_tables.items():
            ops.rowopt_update_bm(t["bm"], 1, t["words"], 1, 0, t["state"], t["p"], t["g"], t["m"], t["v"], t["ema"], self.global_step,
                                 self.coef, self.hist)
        if self.ema_on:
            self.n_ema += 1
        self.last_logits = logits
        return loss

    def _exchange_table_grads(self, tok: torch.Tensor) -> torch.Tensor:
        """Replicated tables under data parallelism: gather every rank's (token, gradient-row) pairs and apply the
        remote ones locally, so each replica sees the full global-batch table gradient (C1, all-gather form)."""
        D = self.spec.D
        flat_tok = tok.reshape(-1)
        rows = {n: self.tables[n]["g"][flat_tok.long()] for n in BIG_TABLES}     # local contributions (dedup'ed sums)
        toks = [torch.empty_like(flat_tok) for _ in range(self.world)]
        torch.distributed.all_gather(toks, flat_tok, group=self.pg)
        rank = torch.distributed.get_rank(self.pg)
        # a row summed locally appears once per local slot: only the owner slot carries it, the others send zeros
        ops.rows_claim(flat_tok, self.spec.pad_id, self.owner)
        own = (self.owner[flat_tok.long()] == torch.arange(flat_tok.numel(), device=self.dev, dtype=torch.int32)) & \
              (flat_tok != self.spec.pad_id)
        for n in BIG_TABLES:
            send = rows[n] * own.unsqueeze(1)
            recv = [torch.empty_like(send) for _ in range(self.world)]
            torch.distributed.all_gather(recv, send, group=self.pg)
            for r in range(self.world):
                if r != rank:
                    self.tables[n]["g"].index_add_(0, toks[r].long(), recv[r])
        return torch.cat(toks)

    # ---- evaluation / state ----------------------------------------------------------------------------------
    def sync_tables(self):
        """Bring every row of the big tables (p, m, v, ema) to the current step.  Needed before anything other than
        FusedTrainer.step reads them (evaluation, state_dict, EMA swap-in); a dense pass, once per epoch in practice."""
        if self.lazy and self.global_step > 0 and self._flushed_at != self.global_step:
            for t in list(self.tables.values()) + list(self.cat_tables.values()):
                ops.rowopt_flush(t["state"], t["p"], t["m"], t["v"], t["ema"], self.global_step, self.hist)
        self._flushed_at = self.global_step

    @torch.no_grad()
    def predict(self, batch: Dict[str, torch.Tensor]):
        """Eval-mode forward on the trainer's own storage (sharded mode: collective, every rank calls it)."""
        self.sync_tables()
        if self.shard:
            self._allreduce(self._bar)                   # barrier: all shards flushed before peers read them
        dbatch, _ = self.to_device(batch, None)
        dc = hp.DropCtl(False)
        logits, aux, _ = hp.model_fwd(self.P, self.spec, dbatch, dc, self.cache)
        return logits, ops.sigmoid(logits), aux

    def ema_state(self) -> Dict[str, torch.Tensor]:
        """EMA shadows keyed like ModelEMA.shadow_params (utils/ema.py:51-58)."""
        if not self.ema_on:
            return {}
        self.sync_tables()
        out = {}
        for n, o in self.offsets.items():
            p = self.P[n]
            out[n] = self.flat_ema[o:o + p.numel()].view(p.shape)
        for n in BIG_TABLES:
            out[n] = self.tables[n]["ema"]
        for n, t in self.cat_tables.items():
            out[n] = t["ema"][:, :t["dc"]]
        return out
