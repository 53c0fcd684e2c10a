"""Thin tensor-level wrappers over the C ABI (one Python function per entry point).

PyTorch is used for device memory and streams only: every function takes CUDA tensors, passes raw
pointers + the current stream to libb200ctr, and returns the output tensor(s).  No compute happens in torch.
All wrappers raise if a tensor is not on a CUDA device (there is no fallback path).
"""
from __future__ import annotations

import ctypes
import os as _os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import Dropout, Epilogue, NO_DROP, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MASK_NONE, MASK_NONZERO, MASK_GELU_GRAD = 0, 1, 2

LAUNCHES = 0          # number of C-ABI kernel-launching calls made (bench.py reports it)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


# ---- optional per-call device timing (bench.py: roofline of the dominant kernel, measured live) ---------------
_PROF = None          # None = off; else {tag: [ (start_event, end_event), ... ]}
_TAG = None           # extra tag for the next call (gemm shapes)


def profile_start():
    global _PROF
    _PROF = {}


def profile_stop():
    """Returns {tag: (n_calls, total_ms)}; synchronises the device."""
    global _PROF
    prof, _PROF = _PROF, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (prof or {}).items()}


# NVTX range per C-ABI call (B200CTR_NVTX=1): the kernels of one op show up under the op's name in Nsight timelines
_NVTX = _os.environ.get("B200CTR_NVTX", "0") not in ("", "0")


def _run(what, fn, *args):
    global _TAG
    if _NVTX:
        torch.cuda.nvtx.range_push(what if _TAG is None else f"{what}{_TAG}")
        try:
            rc = fn(*args)
        finally:
            torch.cuda.nvtx.range_pop()
        _TAG = None
        check(rc, what)
        return
    if _PROF is None:
        rc = fn(*args)
    else:
        tag = what if _TAG is None else f"{what}{_TAG}"
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        _PROF.setdefault(tag, []).append((a, b))
    _TAG = None
    check(rc, what)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("b200ctr ops need CUDA tensors (no CPU fallback exists)")
    return t.data_ptr()


def _f32(t: torch.Tensor, name="tensor"):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


def _ld(t: torch.Tensor) -> int:
    """Leading dimension of a 2-D (possibly row-strided) view with unit inner stride."""
    assert t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1), (t.shape, t.stride())
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def drop(p: float, seed: int, stream_id: int) -> Dropout:
    if p <= 0.0:
        return NO_DROP
    return Dropout(float(p), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, stream_id & 0xFFFFFFFF)


# ---- matmul precision -------------------------------------------------------------------------------------------
# "fp32": CUDA-core contraction, exact fp32 (the reference's default `amp: none`, TF32 off).
# "bf16": tcgen05 tensor-core contraction with bf16 operands / fp32 accumulation (the reference's `amp: bf16` autocast).
_PRECISION = "fp32"
# contraction kernel of the fp32 mode (b200ctr_gemm_split): "x6" = tcgen05 split precision with 3 bf16 parts per operand (6 MMAs
# per product, fp32-grade), "x3" = 2 parts (3 MMAs, ~1e-5), "fma" = the CUDA-core fp32 kernel (b200ctr_gemm)
import os as _os
FP32_GEMM = _os.environ.get("B200CTR_FP32_GEMM", "x6")


def get_precision() -> str:
    return _PRECISION


class precision:
    """Context manager: `with ops.precision("bf16"): ...`"""

    def __init__(self, mode: str):
        if mode not in ("fp32", "bf16"):
            raise ValueError(f"unknown matmul precision {mode!r}")
        self.mode = mode

    def __enter__(self):
        global _PRECISION
        self.prev, _PRECISION = _PRECISION, self.mode
        return self

    def __exit__(self, *exc):
        global _PRECISION
        _PRECISION = self.prev
        return False


def _split_k_for(M: int, N: int, K: int) -> int:
    bn = 128 if N > 64 else 64
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    if tiles >= 148 or K < 2048:
        return 1
    return int(max(1, min(296 // tiles, K // 256, 256)))


def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, out: Optional[torch.Tensor] = None,
         bias=None, act=ACT_NONE, pre_out=None, dropout: Dropout = NO_DROP, mask_mode=MASK_NONE, mask_src=None,
         mask_scale=1.0, residual=None, accum=None, alpha=1.0, split_k: Optional[int] = None,
         tc: Optional[bool] = None) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * op(a) @ op(b)); see include/b200ctr.h:b200ctr_gemm."""
    lib = _lib.load()
    bf16_in = a.dtype == torch.bfloat16
    if bf16_in:
        if b.dtype != torch.bfloat16 or (trans_a and trans_b):
            raise TypeError("gemm: bf16 operands must both be bf16 (and not transA && transB)")
    else:
        _f32(a, "a"); _f32(b, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    Kb, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, trans_a, trans_b)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    assert out.shape == (M, N)
    e = Epilogue()
    e.alpha = alpha
    e.bias = _p(bias)
    e.act = act
    e.pre_out = _p(pre_out); e.ld_pre = _ld(pre_out) if pre_out is not None else 0
    e.drop = dropout
    e.mask_mode = mask_mode
    e.mask_src = _p(mask_src); e.ld_mask = _ld(mask_src) if mask_src is not None else 0
    e.mask_scale = mask_scale
    e.residual = _p(residual); e.ld_res = _ld(residual) if residual is not None else 0
    e.accum = _p(accum); e.ld_accum = _ld(accum) if accum is not None else 0
    if split_k is None:
        split_k = _split_k_for(M, N, K)
    ws = None
    if split_k > 1:
        ws = torch.empty((split_k * M * N,), device=a.device, dtype=torch.float32)
    global _TAG
    if tc is None:      # tiny contractions are not worth a TMEM allocation; transA && transB is not built
        use_tc = (_PRECISION == "bf16") and not (trans_a and trans_b) and (M * N * K >= (1 << 20)) and K >= 32
    else:
        use_tc = bool(tc)
    # `amp: none`: the same tcgen05 kernels in split-precision mode (3 x bf16 per product, fp32-grade accuracy) instead of
    # the CUDA-core fp32 contraction; tc=False (tests) and tiny shapes keep the exact FMA kernel
    use_split = (tc is None and not use_tc and not bf16_in and _PRECISION == "fp32" and FP32_GEMM in ("x3", "x6")
                 and not (trans_a and trans_b) and (M * N * K >= (1 << 20)) and K >= 32)
    if _PROF is not None:
        _TAG = f"[{M}x{N}x{K}{'T' if trans_a else 'N'}{'T' if trans_b else 'N'}]"
    if use_split:
        _run("gemm_" + FP32_GEMM, lib.b200ctr_gemm_split, 2 if FP32_GEMM == "x3" else 3, int(trans_a), int(trans_b), M, N, K, _p(a),
             _ld(a), _p(b), _ld(b), _p(out), _ld(out), ctypes.byref(e), split_k, _p(ws), _stream())
    else:
        fn = lib.b200ctr_gemm_bf16in if bf16_in else (lib.b200ctr_gemm_bf16 if use_tc else lib.b200ctr_gemm)
        _run("gemm_bf16" if (use_tc or bf16_in) else "gemm", fn, int(trans_a), int(trans_b), M, N, K, _p(a), _ld(a), _p(b), _ld(b),
             _p(out), _ld(out), ctypes.byref(e), split_k, _p(ws), _stream())
    _count(2 if split_k > 1 else 1)
    return out


def gemm_kk(a16: torch.Tensor, b16: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (M,N) fp32 = a16 (M,K) @ b16 (N,K)^T, both bf16 with unit inner stride (row-strided views allowed): TMA-fed tcgen05."""
    lib = _lib.load()
    assert a16.dtype == torch.bfloat16 and b16.dtype == torch.bfloat16
    M, K = a16.shape
    N, Kb = b16.shape
    assert K == Kb
    if out is None:
        out = torch.empty((M, N), device=a16.device, dtype=torch.float32)
    assert out.shape == (M, N)
    global _TAG
    if _PROF is not None:
        _TAG = f"[{M}x{N}x{K}NT]"
    _run("gemm_tma", lib.b200ctr_gemm_bf16_kk, _p(a16), _ld(a16), _p(b16), _ld(b16), _p(out), _ld(out), M, N, K, _stream())
    _count()
    return out


def gemm_kk_ok(M: int, N: int, K: int) -> bool:
    return N % 32 == 0 and K % 8 == 0 and M >= 128 and K >= 256


def to_bf16_t(x: torch.Tensor) -> torch.Tensor:
    """bf16 transpose (C, R) of a row-strided fp32 (R, C) view."""
    lib = _lib.load()
    R, C = x.shape
    out = torch.empty((C, R), device=x.device, dtype=torch.bfloat16)
    _run("to_bf16_t", lib.b200ctr_to_bf16_transposed, _p(_f32(x)), R, C, _ld(x), _p(out), _stream())
    _count()
    return out


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    R, C = x.shape
    out = torch.empty((C, R), device=x.device, dtype=torch.bfloat16)
    _run("transpose_bf16", lib.b200ctr_transpose_bf16, _p(x), R, C, _p(out), _stream())
    _count()
    return out


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    M, N = x.shape
    if out is None:
        out = torch.empty((N,), device=x.device, dtype=torch.float32)
    ws = torch.empty((1184 * N,), device=x.device, dtype=torch.float32)
    _run("colsum", lib.b200ctr_colsum, _p(x), _ld(x), M, N, _p(out), _p(ws), _stream())
    _count(2)
    return out


# ---- numeric / binary embedding ---------------------------------------------------------------------------
def numbin_fold(w, b, w_out):
    lib = _lib.load()
    F, fe = w.shape
    D = w_out.shape[0]
    P = torch.empty((F, D), device=w.device, dtype=torch.float32)
    Q = torch.empty((F, D), device=w.device, dtype=torch.float32)
    _run("numbin_fold", lib.b200ctr_numbin_fold, _p(w), _p(b), _p(w_out), _p(P), _p(Q), F, fe, D, _stream())
    _count()
    return P, Q


def numbin_fwd(x, P, Q, out, ld_b, mean_out, ld_mean):
    """out: tensor whose data_ptr is the address of element [0,0,0] of the (B,F,D) destination (or None)."""
    lib = _lib.load()
    B, F = x.shape
    D = P.shape[1]
    _run("numbin_fwd", lib.b200ctr_numbin_fwd, _p(x), _p(P), _p(Q), _p(out), ld_b, _p(mean_out), ld_mean, B, F, D, _stream())
    _count()


def numbin_bwd(x, dout, ld_b, dmean, ld_dmean, D):
    lib = _lib.load()
    B, F = x.shape
    rows = int(lib.b200ctr_numbin_bwd_rows(B))
    part = torch.empty((2, rows, F * D), device=x.device, dtype=torch.float32)       # per-chunk partial sums, no atomics
    _run("numbin_bwd", lib.b200ctr_numbin_bwd, _p(x), _p(dout), ld_b, _p(dmean), ld_dmean, _p(part[0]), _p(part[1]), B, F, D, _stream())
    _count()
    dP, dQ = colsum(part[0]).view(F, D), colsum(part[1]).view(F, D)                  # summed in chunk order
    return dP, dQ


def numbin_unfold(dP, dQ, w, b, w_out):
    lib = _lib.load()
    F, fe = w.shape
    D = w_out.shape[0]
    dw = torch.empty_like(w)
    db = torch.empty_like(b) if b is not None else None
    dwo = torch.empty_like(w_out)
    _run("numbin_unfold", lib.b200ctr_numbin_unfold, _p(dP), _p(dQ), _p(w), _p(b), _p(w_out), _p(dw), _p(db), _p(dwo), F, fe, D,
                                    _stream())
    _count()
    return dw, db, dwo


# ---- categorical ----------------------------------------------------------------------------------------------
class CatTables:
    """Device-side descriptor arrays (pointer tables, dims, cards) for the Fc categorical tables."""

    def __init__(self, tables: Sequence[torch.Tensor], projs: Sequence[torch.Tensor]):
        dev = tables[0].device
        self.tables, self.projs = list(tables), list(projs)
        self.Fc = len(tables)
        for t, p in zip(tables, projs):
            # a table is either dense (row stride d_c) or a [:, :d_c] view of a 64-wide store (row-lazy tables)
            assert t.stride(1) == 1 and t.stride(0) in (t.shape[1], 64) and p.is_contiguous() and t.dtype == torch.float32
            if t.shape[1] > 64:
                raise _lib.B200CtrError(f"categorical embedding dim {t.shape[1]} > 64 is not built")
        self.t_ptrs = torch.tensor([t.data_ptr() for t in tables], dtype=torch.int64, device=dev)
        self.p_ptrs = torch.tensor([p.data_ptr() for p in projs], dtype=torch.int64, device=dev)
        self.dims = torch.tensor([t.shape[1] for t in tables], dtype=torch.int32, device=dev)
        self.cards = torch.tensor([t.shape[0] for t in tables], dtype=torch.int32, device=dev)
        self.lds = torch.tensor([t.stride(0) if t.shape[0] > 1 else t.shape[1] for t in tables], dtype=torch.int32, device=dev)
        self._key = tuple(t.data_ptr() for t in tables) + tuple(p.data_ptr() for p in projs)

    def matches(self, tables, projs) -> bool:
        return self._key == tuple(t.data_ptr() for t in tables) + tuple(p.data_ptr() for p in projs)


_PTR_TABLES: dict = {}


def _ptr_table(ts: Sequence[torch.Tensor]) -> torch.Tensor:
    """Device array of the tensors' addresses, cached by the addresses themselves: `torch.tensor(list, device=cuda)` is a
    pageable host-to-device copy, i.e. a host wait for everything queued on the stream -- made once per step (cat_bwd's two
    gradient pointer tables) it drained the launch queue near the end of every backward and left the GPU waiting for the host
    through the launch-bound start of the next step.  The trainer's gradient views live in one persistent buffer, so the
    tables are built once."""
    key = (ts[0].device, tuple(t.data_ptr() for t in ts))
    tab = _PTR_TABLES.get(key)
    if tab is None:
        if len(_PTR_TABLES) >= 256:          # callers with ever-changing gradient tensors (autograd): do not grow without bound
            _PTR_TABLES.clear()
        tab = torch.tensor(key[1], dtype=torch.int64, device=ts[0].device)
        _PTR_TABLES[key] = tab
    return tab


def _idx64(t: torch.Tensor) -> int:
    if t.dtype == torch.int64:
        return 1
    if t.dtype == torch.int32:
        return 0
    raise TypeError(f"index tensor must be int64 or int32, got {t.dtype}")


def cat_fwd(x_cat, ct: CatTables, out, ld_b, mean_out, ld_mean, q_idx, q_out, D, dropout=NO_DROP):
    lib = _lib.load()
    B, Fc = x_cat.shape
    assert Fc == ct.Fc and x_cat.is_contiguous()
    _run("cat_fwd", lib.b200ctr_cat_fwd, _p(x_cat), _idx64(x_cat), _p(ct.t_ptrs), _p(ct.p_ptrs), _p(ct.dims), _p(ct.cards),
                              _p(ct.lds), _p(out), ld_b, _p(mean_out), ld_mean, q_idx, _p(q_out), B, Fc, D, dropout, _stream())
    _count()


def cat_bwd(x_cat, ct: CatTables, dout, ld_b, dmean, ld_dmean, q_idx, dq, dtables, dprojs, D, dropout=NO_DROP, rows_acc=None,
            keep_rows=None, defer=None):
    """dtables/dprojs: lists of ZEROED gradient tensors (accumulated into).  rows_acc / keep_rows: the trainer's persistent
    all-zero (sum_cards, 64) row accumulator and the int32 per-table flags of its row-lazy tables (their reduced rows stay in
    rows_acc; their dtables entry is ignored and may be any tensor of the right row count)."""
    lib = _lib.load()
    B, Fc = x_cat.shape
    dt_ptrs, dp_ptrs = _ptr_table(dtables), _ptr_table(dprojs)
    sum_dims = int(sum(t.shape[1] for t in ct.tables))
    # rows of the widest table that is UNPADDED into its dense gradient (the row-lazy tables keep their rows in rows_acc): the
    # unpad grid is sized by it, so HBM-sized row-lazy tables do not cost millions of idle blocks
    kept = getattr(keep_rows, "_host", None) if keep_rows is not None else None
    max_card = int(max([t.shape[0] for i, t in enumerate(ct.tables) if not (kept and kept[i])] or [1]))
    sum_cards = int(sum(t.shape[0] for t in ct.tables))
    ws_bytes = int(lib.b200ctr_cat_bwd_workspace_bytes(B, Fc, sum_dims, sum_cards))
    ws = torch.empty((ws_bytes,), device=x_cat.device, dtype=torch.uint8)
    _run("cat_bwd", lib.b200ctr_cat_bwd, _p(x_cat), _idx64(x_cat), _p(ct.t_ptrs), _p(ct.p_ptrs), _p(ct.dims), _p(ct.cards),
                              _p(ct.lds), _p(dout), ld_b, _p(dmean), ld_dmean, q_idx, _p(dq), _p(dt_ptrs), _p(dp_ptrs),
                              B, Fc, D, sum_dims, sum_cards, max_card, dropout, _p(rows_acc), _p(keep_rows),
                              _p(defer["slot"]) if defer else None, defer["n"] if defer else 0, _p(defer["rows"]) if defer else None,
                              _p(defer["keys"]) if defer else None, _p(ws), ws_bytes, _stream())
    _count(7)


def axpby(x, y, a, b, out=None):
    lib = _lib.load()
    ref = x if x is not None else y
    if out is None:
        out = torch.empty_like(ref)
    _run("axpby", lib.b200ctr_axpby, _p(x), _p(y), a, b, _p(out), ref.numel(), _stream())
    _count()
    return out


def relu_bwd(dy, y, scale=1.0):
    lib = _lib.load()
    out = torch.empty_like(dy)
    _run("relu_bwd", lib.b200ctr_relu_bwd, _p(dy), _p(y), float(scale), _p(out), dy.numel(), _stream())
    _count()
    return out


def sigmoid(x):
    lib = _lib.load()
    out = torch.empty_like(x)
    _run("sigmoid", lib.b200ctr_sigmoid, _p(x), _p(out), x.numel(), _stream())
    _count()
    return out


def add2d_(dst, src):
    """dst += src for 2-D row-strided views of equal shape."""
    lib = _lib.load()
    assert dst.shape == src.shape
    _run("add2d", lib.b200ctr_add2d, _p(dst), _ld(dst), _p(src), _ld(src), dst.shape[0], dst.shape[1], _stream())
    _count()
    return dst


# ---- DARE ------------------------------------------------------------------------------------------------------
def enable_peer_access(peer_device: int):
    """Let kernels on the current device dereference memory of `peer_device` (NVLink P2P)."""
    check(_lib.load().b200ctr_enable_peer_access(int(peer_device)), "enable_peer_access")


def ipc_export(t: torch.Tensor):
    """(handle bytes, offset) of a CUDA tensor's memory for another process on this node."""
    lib = _lib.load()
    h = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    check(lib.b200ctr_ipc_export(_p(t), h, ctypes.byref(off)), "ipc_export")
    return bytes(h.raw), int(off.value)


_IPC_BASES: dict = {}


def ipc_import(handle: bytes, offset: int) -> int:
    """Device address (valid on the current device, peer access enabled) of memory exported by another process."""
    base = _IPC_BASES.get(handle)
    if base is None:
        lib = _lib.load()
        out = ctypes.c_void_p(0)
        check(lib.b200ctr_ipc_import(ctypes.create_string_buffer(handle, 64), ctypes.byref(out)), "ipc_import")
        base = int(out.value)
        _IPC_BASES[handle] = base            # an allocation may be opened only once per process
    return base + offset


class ShardedTable:
    """A (vocab, D) table stored as n row-interleaved shards: row r lives in shard r % n at local row r // n.  `shards`
    holds this rank's tensor and, for the other ranks, raw device addresses of their shards mapped through CUDA IPC
    (peer memory over NVLink); the kernels dereference them directly."""

    def __init__(self, shards: Sequence, vocab: int, D: int = 64):
        n = len(shards)
        if n < 1 or n > 8 or (n & (n - 1)):
            raise _lib.B200CtrError(f"row-sharded tables need a power-of-two shard count <= 8, got {n}")
        addrs = []
        for t in shards:
            if isinstance(t, torch.Tensor):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 2
                addrs.append(t.data_ptr())
            else:
                addrs.append(int(t))
        self.shards, self.n, self.vocab, self.D = list(shards), n, int(vocab), D
        self.ptrs = (ctypes.c_void_p * n)(*addrs)


def retrieve_fwd(seq, q, emb_att, emb_rep, rec, K, pad_id):
    lib = _lib.load()
    B, L = seq.shape
    D = q.shape[1]
    dev = q.device
    assert seq.is_contiguous() and q.is_contiguous()
    sel = torch.empty((B, K, D), device=dev, dtype=torch.float32)
    vals = torch.empty((B, K), device=dev, dtype=torch.float32)
    idx = torch.empty((B, K), device=dev, dtype=torch.int32)
    tok = torch.empty((B, K), device=dev, dtype=torch.int32)
    if isinstance(emb_att, ShardedTable):
        assert isinstance(emb_rep, ShardedTable) and emb_rep.n == emb_att.n
        _run("retrieve_fwd", lib.b200ctr_retrieve_fwd_sharded, _p(seq), _idx64(seq), _p(q), emb_att.ptrs, emb_rep.ptrs, emb_att.n,
             _p(rec), B, L, K, D, pad_id, emb_att.vocab, _p(sel), _p(vals), _p(idx), _p(tok), _stream())
    else:
        assert emb_att.is_contiguous() and emb_rep.is_contiguous()
        _run("retrieve_fwd", lib.b200ctr_retrieve_fwd, _p(seq), _idx64(seq), _p(q), _p(emb_att), _p(emb_rep), _p(rec), B, L, K, D,
             pad_id, emb_att.shape[0], _p(sel), _p(vals), _p(idx), _p(tok), _stream())
    _count()
    return sel, vals, idx, tok


# table-gradient scatter: "sorted" = radix sort by token + segmented reduction in slot order (deterministic, csrc/scatter.cu);
# "atomic" = the first version's fp32 vector atomics (kept for A/B timing)
SCATTER = _os.environ.get("B200CTR_SCATTER", "sorted")


def scatter_rows(tok, n_src, n_slots, pad_id, vocab, K, q, dvals, rows_rep, stage_att, stage_rep, stage_cap, n_shards, rank, g_att, g_rep):
    lib = _lib.load()
    n = int(n_src) * int(n_slots)
    ws_bytes = int(lib.b200ctr_scatter_rows_workspace_bytes(n, int(vocab)))
    ws = torch.empty((ws_bytes,), device=tok.device, dtype=torch.uint8)
    _run("scatter_rows", lib.b200ctr_scatter_rows, _p(tok), int(n_src), int(n_slots), pad_id, int(vocab), int(K), _p(q), _p(dvals),
         _p(rows_rep), _p(stage_att), _p(stage_rep), int(stage_cap), int(n_shards), int(rank), _p(g_att), _p(g_rep), _p(ws), ws_bytes,
         _stream())
    _count(4)


def retrieve_bwd(tok, q, emb_att, dsel_rep, dvals, pad_id, datt_dense=None, drep_dense=None):
    lib = _lib.load()
    B, K = tok.shape
    D = q.shape[1]
    dq = torch.empty_like(q)
    if SCATTER == "sorted" and not isinstance(emb_att, ShardedTable):
        # dq only; then every distinct gradient row is reduced once, in slot order
        _run("retrieve_bwd", lib.b200ctr_retrieve_bwd, _p(tok), _p(q), _p(emb_att), _p(dsel_rep), _p(dvals), B, K, D, pad_id, _p(dq),
             None, None, _stream())
        _count()
        if datt_dense is not None or drep_dense is not None:
            scatter_rows(tok, 1, B * K, pad_id, emb_att.shape[0], K, q, dvals, dsel_rep.contiguous() if dsel_rep is not None else None,
                         None, None, 0, 1, 0, datt_dense, drep_dense)
        return dq
    if isinstance(emb_att, ShardedTable):
        sa = getattr(datt_dense, "stage_ptrs", None) if datt_dense is not None else None
        sr = getattr(drep_dense, "stage_ptrs", None) if drep_dense is not None else None
        ref = datt_dense if datt_dense is not None else drep_dense
        _run("retrieve_bwd", lib.b200ctr_retrieve_bwd_staged, _p(tok), _p(q), emb_att.ptrs, emb_att.n, _p(dsel_rep), _p(dvals), B, K,
             D, pad_id, _p(dq), datt_dense.ptrs if datt_dense is not None else None,
             drep_dense.ptrs if drep_dense is not None else None, sa, sr,
             getattr(ref, "stage_rank", 0) if ref is not None else 0, getattr(ref, "stage_cap", 0) if ref is not None else 0,
             1 if (SCATTER == "sorted" and (sa is not None or sr is not None)) else 0, _stream())
    else:
        _run("retrieve_bwd", lib.b200ctr_retrieve_bwd, _p(tok), _p(q), _p(emb_att), _p(dsel_rep), _p(dvals), B, K, D, pad_id, _p(dq),
             _p(datt_dense), _p(drep_dense), _stream())
    _count()
    return dq


def stage_apply(gtok, n_src, n_slots, stage_cap, pad_id, n_shards, rank, stage_att, stage_rep, g_att, g_rep):
    """Owner side of the staged scatter: add the rows other ranks staged for tokens this rank owns."""
    lib = _lib.load()
    _run("stage_apply", lib.b200ctr_stage_apply, _p(gtok), n_src, n_slots, stage_cap, pad_id, n_shards, rank, _p(stage_att),
         _p(stage_rep), _p(g_att), _p(g_rep), g_att.shape[1] if g_att is not None else g_rep.shape[1], _stream())
    _count()


def rows_localize(tok, pad_id, n_shards, rank):
    lib = _lib.load()
    out = torch.empty(tok.shape, device=tok.device, dtype=torch.int32)
    _run("rows_localize", lib.b200ctr_rows_localize, _p(tok), _idx64(tok), tok.numel(), pad_id, n_shards, rank, _p(out), _stream())
    _count()
    return out


def _attn_tc(tc: Optional[bool], K: int, H: int, D: int) -> bool:
    """bf16 precision mode runs attention on the tensor cores (TF32 mma.sync); fp32 mode keeps the exact CUDA-core kernels."""
    if tc is None:
        tc = _PRECISION == "bf16"
    return bool(tc) and D == 64 and H == 8 and K <= 160


def _qkv_f32(qkv):
    if qkv.dtype != torch.float32:
        raise _lib.B200CtrError("an fp16 qkv tensor (inproj_fwd(h16=True)) feeds the packed-half attention kernels only")


def _attn_impl(impl: Optional[str]) -> str:
    """Tensor-core attention flavour: "h16" (packed-half kernels, csrc/attn_h16.cu; default) or "tf32" (csrc/attn_tc.cu).
    B200CTR_ATTN overrides the default for A/B runs."""
    impl = impl or _os.environ.get("B200CTR_ATTN", "h16")
    if impl not in ("h16", "tf32"):
        raise _lib.B200CtrError(f"attention implementation {impl!r}: expected 'h16' or 'tf32'")
    return impl


def attn_fwd(qkv, relmean, max_len, B, K, H, D, dropout=NO_DROP, tc: Optional[bool] = None, save_keep: bool = False,
             impl: Optional[str] = None, out_bf16: bool = False):
    """Returns (o, lse), or (o, lse, keep) with save_keep: keep = the dropout keep-flag words the tensor-core backward reads
    instead of re-hashing (None when dropout is off or the fp32 kernels run)."""
    lib = _lib.load()
    use_tc = _attn_tc(tc, K, H, D)
    out_bf16 = bool(out_bf16 and use_tc and _attn_impl(impl) == "h16")      # packed-half kernels only: o as bf16 for the out-projection
    o = torch.empty((B * K, D), device=qkv.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    lse = torch.empty((B, H, K), device=qkv.device, dtype=torch.float32)
    keep = None
    if use_tc:
        if save_keep and dropout.p > 0.0:
            keep = torch.empty((int(lib.b200ctr_attn_tc_keep_words(B, K)),), device=qkv.device, dtype=torch.int32)
        if _attn_impl(impl) == "h16":
            _run("attn_tc_fwd", lib.b200ctr_attn_h16_fwd, _p(qkv), int(qkv.dtype == torch.float16), _p(relmean), max_len, B, K, H, D,
                 _p(o), int(out_bf16), _p(lse), _p(keep), dropout, _stream())
        else:
            _qkv_f32(qkv)
            _run("attn_tc_fwd", lib.b200ctr_attn_tc_fwd, _p(qkv), _p(relmean), max_len, B, K, H, D, _p(o), _p(lse), _p(keep), dropout,
                 _stream())
    else:
        _run("attn_fwd", lib.b200ctr_attn_fwd, _p(qkv), _p(relmean), max_len, B, K, H, D, _p(o), _p(lse), dropout, _stream())
    _count()
    return (o, lse, keep) if save_keep else (o, lse)


def attn_bwd(qkv, relmean, max_len, o, lse, d_o, B, K, H, D, drelmean, dropout=NO_DROP, tc: Optional[bool] = None, keep=None,
             out_bf16: bool = False, impl: Optional[str] = None):
    """out_bf16 (tensor-core path only): dqkv comes back as bf16 for inproj_bwd, which consumes it as bf16 in any case."""
    lib = _lib.load()
    use_tc = _attn_tc(tc, K, H, D)
    out_bf16 = bool(out_bf16 and use_tc)
    dqkv = torch.empty(qkv.shape, device=qkv.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    drel_out = drelmean
    h16 = use_tc and _attn_impl(impl) == "h16"
    if drelmean is not None:      # the kernels write one partial row per CTA (no float atomics); summed in row order below
        rows = int(lib.b200ctr_attn_h16_drel_rows(B)) if h16 else B
        drelmean = torch.zeros((rows, 2 * max_len + 1), device=qkv.device, dtype=torch.float32)
    if use_tc:
        if h16:
            _run("attn_tc_bwd", lib.b200ctr_attn_h16_bwd, _p(qkv), int(qkv.dtype == torch.float16), _p(relmean), max_len, _p(o),
                 int(o.dtype == torch.bfloat16), _p(lse), _p(d_o), int(d_o.dtype == torch.bfloat16), B, K, H, D, _p(dqkv), int(out_bf16),
                 _p(drelmean), _p(keep), dropout, _stream())
        else:
            _qkv_f32(qkv)
            _run("attn_tc_bwd", lib.b200ctr_attn_tc_bwd, _p(qkv), _p(relmean), max_len, _p(o), _p(lse), _p(d_o), B, K, H, D, _p(dqkv),
                 int(out_bf16), _p(drelmean), _p(keep), dropout, _stream())
    else:
        _run("attn_bwd", lib.b200ctr_attn_bwd, _p(qkv), _p(relmean), max_len, _p(o), _p(lse), _p(d_o), B, K, H, D, _p(dqkv),
             _p(drelmean), dropout, _stream())
    _count()
    if drel_out is not None:
        colsum(drelmean, out=drel_out)
    return dqkv


def add_rmsnorm_fwd(a, b, w, eps=1e-6, save_sum=True):
    lib = _lib.load()
    M, D = a.shape
    s = torch.empty_like(a) if save_sum else None
    y = torch.empty_like(a)
    r = torch.empty((M,), device=a.device, dtype=torch.float32)
    _run("add_rmsnorm_fwd", lib.b200ctr_add_rmsnorm_fwd, _p(a), _p(b), _p(w), eps, M, D, _p(s), _p(y), _p(r), _stream())
    _count()
    return s, y, r


# ---- fused attention-side projections (csrc/proj_fused.cu) ---------------------------------------------------------
def inproj_fwd(x, w_in, b_in, h16: bool = False):
    """qkv = x . w_in^T + b_in.  h16: as fp16 with the q third pre-multiplied by log2(e)/sqrt(8) -- the operand format of the
    packed-half attention kernels (attn_fwd / attn_bwd recognise it by its dtype)."""
    lib = _lib.load()
    T, D = x.shape
    if h16:
        qkv = torch.empty((T, 3 * D), device=x.device, dtype=torch.float16)
        _run("inproj_fwd", lib.b200ctr_inproj_fwd_h16, _p(_f32(x)), _p(w_in), _p(b_in), T, D, _p(qkv), _stream())
    else:
        qkv = torch.empty((T, 3 * D), device=x.device, dtype=torch.float32)
        _run("inproj_fwd", lib.b200ctr_inproj_fwd, _p(_f32(x)), _p(w_in), _p(b_in), T, D, _p(qkv), _stream())
    _count()
    return qkv


def attn_takes_h16(K: int, H: int, D: int) -> bool:
    """True when the attention of the current precision mode is the packed-half implementation (so the in-projection may hand
    it fp16 operands)."""
    return _attn_tc(None, K, H, D) and _attn_impl(None) == "h16"


def outproj_norm_fwd(o, w_out, b_out, x_resid, n1w, eps=1e-6):
    """Returns (s1, x1, r1) like add_rmsnorm_fwd(x_resid, o @ w_out^T + b_out, n1w)."""
    lib = _lib.load()
    T, D = o.shape
    s1, x1 = torch.empty_like(x_resid), torch.empty_like(x_resid)
    r1 = torch.empty((T,), device=o.device, dtype=torch.float32)
    o16 = o.dtype == torch.bfloat16                 # bf16 o from attn_fwd(out_bf16=True): staged with plain copies
    _run("outproj_norm_fwd", lib.b200ctr_outproj_norm_fwd, _p(o if o16 else _f32(o)), int(o16), _p(w_out), _p(b_out), _p(x_resid), _p(n1w),
         eps, T, D, _p(s1), _p(r1), _p(x1), _stream())
    _count()
    return s1, x1, r1


def _proj_ws(dev):
    return torch.empty((_lib.load().b200ctr_proj_bwd_workspace_bytes(),), device=dev, dtype=torch.uint8)


def outproj_norm_bwd(dx1, s1, r1, n1w, o, w_out, dw_out, db_out, dn1, out_bf16: bool = False):
    """Returns (d_s1, d_o); overwrites dw_out, db_out, dn1.  o may be bf16 (attn_fwd(out_bf16=True)); out_bf16: d_o comes back as
    bf16 for the packed-half attention backward."""
    lib = _lib.load()
    T, D = o.shape
    ds1 = torch.empty_like(s1)
    d_o = torch.empty((T, D), device=o.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ws = _proj_ws(o.device)
    _run("outproj_norm_bwd", lib.b200ctr_outproj_norm_bwd, _p(_f32(dx1)), _p(s1), _p(r1), _p(n1w), _p(o), int(o.dtype == torch.bfloat16),
         _p(w_out), T, D, _p(ds1), _p(d_o), int(out_bf16), _p(dw_out), _p(db_out), _p(dn1), _p(ws), _stream())
    _count(2)
    return ds1, d_o


def inproj_bwd(dqkv, x, w_in, d_resid, dw_in, db_in):
    """Returns dx = dqkv @ w_in + d_resid; overwrites dw_in, db_in."""
    lib = _lib.load()
    T, D = x.shape
    dx = torch.empty_like(x)
    ws = _proj_ws(x.device)
    if dqkv.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"inproj_bwd: dqkv must be float32 or bfloat16, got {dqkv.dtype}")
    _run("inproj_bwd", lib.b200ctr_inproj_bwd, _p(dqkv), int(dqkv.dtype == torch.bfloat16), _p(x), _p(w_in), _p(d_resid), T, D, _p(dx),
         _p(dw_in), _p(db_in), _p(ws), _stream())
    _count(2)
    return dx


def ffn_fused_ok(hidden: int, D: int = 64) -> bool:
    """Shapes the fused tcgen05 FFN kernels are built for (else: the unfused gemm composition)."""
    return D == 64 and hidden % 128 == 0 and 0 < hidden <= 384


def ffn_fwd(x1, w1, b1, w2, b2, n2w, dropout: Dropout = NO_DROP, eps=1e-6, save=True, save_keep=False):
    """Fused FFN block (bf16 tensor cores): returns (x2, s2, r2) -- s2/r2 are None when save=False -- or, with save_keep,
    (x2, s2, r2, keep): the dropout keep-flag bits ffn_bwd reads instead of re-hashing (None when dropout is off)."""
    lib = _lib.load()
    T, D = x1.shape
    hidden = w1.shape[0]
    x2 = torch.empty_like(x1)
    s2 = torch.empty_like(x1) if save else None
    r2 = torch.empty((T,), device=x1.device, dtype=torch.float32) if save else None
    keep = None
    if save_keep and dropout.p > 0.0:
        keep = torch.empty((int(lib.b200ctr_ffn_keep_bytes(T, hidden)),), device=x1.device, dtype=torch.uint8)
    _run("ffn_fwd", lib.b200ctr_ffn_fwd, _p(_f32(x1)), _p(w1), _p(b1), _p(w2), _p(b2), _p(n2w), eps, T, D, hidden,
         _p(s2), _p(r2), _p(x2), _p(keep), dropout, _stream())
    _count()
    return (x2, s2, r2, keep) if save_keep else (x2, s2, r2)


def ffn_bwd(dy, s2, r2, n2w, x1, w1, b1, w2, dw1, db1, dw2, db2, dn2, dropout: Dropout = NO_DROP, keep=None):
    """Backward of ffn_fwd; returns dx1 and overwrites the five parameter-gradient tensors."""
    lib = _lib.load()
    T, D = x1.shape
    hidden = w1.shape[0]
    nbytes = lib.b200ctr_ffn_bwd_workspace_bytes(T, hidden)
    ws = torch.empty((nbytes,), device=x1.device, dtype=torch.uint8)
    dx1 = torch.empty_like(x1)
    for t in (dw1, db1, dw2, db2, dn2):
        assert t.is_contiguous()
    _run("ffn_bwd", lib.b200ctr_ffn_bwd, _p(_f32(dy)), _p(s2), _p(r2), _p(n2w), _p(x1), _p(w1), _p(b1), _p(w2), T, D, hidden,
         _p(dx1), _p(dw1), _p(db1), _p(dw2), _p(db2), _p(dn2), _p(ws), _p(keep), dropout, _stream())
    _count(3)
    return dx1


def rmsnorm_bwd(dy, s, w, r, dw, dx_add=None):
    """dw: (D,) tensor, overwritten with sum_m dy*s*r."""
    lib = _lib.load()
    M, D = dy.shape
    dx = torch.empty_like(dy)
    part = torch.empty((int(lib.b200ctr_rmsnorm_bwd_rows(M)), D), device=dy.device, dtype=torch.float32)   # one row per CTA
    _run("rmsnorm_bwd", lib.b200ctr_rmsnorm_bwd, _p(dy), _p(s), _p(w), _p(r), M, D, _p(dx_add), _p(dx), _p(part), _stream())
    _count()
    colsum(part, out=dw)                                                                  # summed in CTA order: no float atomics
    return dx


def relmean_fwd(rel):
    lib = _lib.load()
    T, H = rel.shape
    out = torch.empty((T,), device=rel.device, dtype=torch.float32)
    _run("relmean_fwd", lib.b200ctr_relmean_fwd, _p(rel), T, H, _p(out), _stream())
    _count()
    return out


def relmean_bwd(drelmean, H):
    lib = _lib.load()
    T = drelmean.shape[0]
    out = torch.empty((T, H), device=drelmean.device, dtype=torch.float32)
    _run("relmean_bwd", lib.b200ctr_relmean_bwd, _p(drelmean), T, H, _p(out), _stream())
    _count()
    return out


def pool_fwd(x, vals, aux_w, aux_b, gating_relu, u, ld_u, dropout=NO_DROP):
    lib = _lib.load()
    B, K, D = x.shape
    aux = torch.empty((B,), device=x.device, dtype=torch.float32)
    w = torch.empty((B, K), device=x.device, dtype=torch.float32)
    _run("pool_fwd", lib.b200ctr_pool_fwd, _p(x), _p(vals), _p(aux_w), _p(aux_b), int(gating_relu), B, K, D, _p(u), ld_u, _p(aux), _p(w),
                               dropout, _stream())
    _count()
    return aux, w


def pool_bwd(x, vals, w, aux_w, u, ld_u, du, ld_du, daux, gating_relu, daux_w, daux_b, dropout=NO_DROP):
    lib = _lib.load()
    B, K, D = x.shape
    dx = torch.empty_like(x)
    dvals = torch.empty_like(vals)
    want = daux is not None and daux_w is not None
    pw = torch.empty((B, D), device=x.device, dtype=torch.float32) if want else None      # per-sample partial rows, no atomics
    pb = torch.empty((B, 1), device=x.device, dtype=torch.float32) if (want and daux_b is not None) else None
    _run("pool_bwd", lib.b200ctr_pool_bwd, _p(x), _p(vals), _p(w), _p(aux_w), _p(u), ld_u, _p(du), ld_du, _p(daux), int(gating_relu),
                               B, K, D, _p(dx), _p(dvals), _p(pw), _p(pb), dropout, _stream())
    _count()
    if want:                                                                              # summed in sample order
        colsum(pw, out=daux_w.view(-1))
        if pb is not None:
            colsum(pb, out=daux_b.view(-1))
    return dx, dvals


# ---- QNN-alpha -------------------------------------------------------------------------------------------------
def rowrms_fwd(x, w, eps=1e-6, want_bf16=False, want_f32=True):
    """z = w * x * rsqrt(mean(x^2) + eps) per row; want_bf16 also returns a bf16 copy of z for the tensor-core GEMMs
    (want_f32=False with want_bf16: only the bf16 copy is written, z is None)."""
    lib = _lib.load()
    B, N = x.shape
    assert want_f32 or want_bf16
    z = torch.empty((B, N), device=x.device, dtype=torch.float32) if want_f32 else None
    r = torch.empty((B,), device=x.device, dtype=torch.float32)
    z16 = torch.empty((B, N), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _run("rowrms_fwd", lib.b200ctr_rowrms_fwd_bf16, _p(x), _ld(x), _p(w), eps, B, N, _p(z), N, _p(r), _p(z16), _stream())
    _count()
    return (z, r, z16) if want_bf16 else (z, r)


def lin1_ok(K: int) -> bool:
    return K % 128 == 0 and K <= 512


def lin1_fwd(h: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Linear(K -> 1): out (B, 1) = h @ w^T + bias, one warp per row."""
    lib = _lib.load()
    B, K = h.shape
    out = torch.empty((B, 1), device=h.device, dtype=torch.float32)
    _run("lin1_fwd", lib.b200ctr_lin1_fwd, _p(_f32(h)), _ld(h), _p(_f32(w)), _p(bias), B, K, _p(out), _stream())
    _count()
    return out


def lin1_bwd(dl: torch.Tensor, act: torch.Tensor, w: torch.Tensor, mask_scale: float, dw: torch.Tensor, db: Optional[torch.Tensor]):
    """Backward of Linear(K -> 1) behind ReLU + dropout: returns dh (B, K); writes dw (1, K) and db (1,)."""
    lib = _lib.load()
    B, K = act.shape
    dh = torch.empty((B, K), device=act.device, dtype=torch.float32)
    ws = torch.empty((int(lib.b200ctr_lin1_bwd_workspace_bytes(K)) // 4,), device=act.device, dtype=torch.float32)
    _run("lin1_bwd", lib.b200ctr_lin1_bwd, _p(_f32(dl)), _p(_f32(act)), _ld(act), _p(_f32(w)), B, K, float(mask_scale), _p(dh), K,
         _p(dw), _p(db), _p(ws), _stream())
    _count(2)
    return dh


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a contiguous fp32 tensor (GEMM operand staging: weights once per step, small activations)."""
    lib = _lib.load()
    assert x.is_contiguous() and x.dtype == torch.float32
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _run("to_bf16", lib.b200ctr_to_bf16, _p(x), x.numel(), _p(out), _stream())
    _count()
    return out


def rowrms_bwd(dz, x, w, r, n_partials=296):
    lib = _lib.load()
    B, N = x.shape
    n_partials = max(1, min(n_partials, B))
    dx = torch.empty((B, N), device=x.device, dtype=torch.float32)
    part = torch.empty((n_partials, N), device=x.device, dtype=torch.float32)
    _run("rowrms_bwd", lib.b200ctr_rowrms_bwd, _p(dz), _ld(dz), _p(x), _ld(x), _p(w), _p(r), B, N, _p(dx), N, _p(part), n_partials,
                                 _stream())
    _count()
    return dx, colsum(part)


def fm_fused_ok(F: int, C: int, D: int) -> bool:
    return D == 64 and 1 <= F <= 256 and 16 <= C <= 128 and C % 32 == 0


def fm_fused_fwd(z16, Ucat, B, F):
    """(s, quad), each (B, C), from the bf16 normed features z16 (B, F*D) and Ucat (D, C)."""
    lib = _lib.load()
    D, C = Ucat.shape
    assert z16.dtype == torch.bfloat16 and z16.is_contiguous() and Ucat.is_contiguous()
    s = torch.empty((B, C), device=z16.device, dtype=torch.float32)
    quad = torch.empty((B, C), device=z16.device, dtype=torch.float32)
    _run("fm_fused_fwd", lib.b200ctr_fm_fused_fwd, _p(z16), _p(Ucat), B, F, C, D, _p(s), _p(quad), _stream())
    _count()
    return s, quad


def fm_fused_bwd(z16, Ucat, s, dquad, dz_resid, B, F, dz_out=None):
    """Returns (dz (B, F*D), dUcat (D, C)); dz_resid (B, F*D) or None is added; dz_out may alias dz_resid."""
    lib = _lib.load()
    D, C = Ucat.shape
    dz = dz_out if dz_out is not None else torch.empty((B, F * D), device=z16.device, dtype=torch.float32)
    dU = torch.empty((D, C), device=z16.device, dtype=torch.float32)
    ws = torch.empty((lib.b200ctr_fm_fused_bwd_workspace_bytes(),), device=z16.device, dtype=torch.uint8)
    assert dz.is_contiguous() and (dz_resid is None or dz_resid.is_contiguous())
    _run("fm_fused_bwd", lib.b200ctr_fm_fused_bwd, _p(z16), _p(Ucat), _p(s), _p(dquad), _p(dz_resid), B, F, C, D, _p(dz), _p(dU),
         _p(ws), _stream())
    _count(2)
    return dz, dU


def headproj_ok(H: int, R: int, P: int) -> bool:
    C = H * R
    if R > 32 or H * P > 1536 or C > 384 or C <= 0 or (H * P) % 4 or C % 4:
        return False
    return (C * (P + 1) + 4 + 2 * 4 * (H * P + C) + 4 * C) * 4 <= 200 * 1024


def headproj_fwd(quad, V):
    """inter (B, H*P) = per-head quad_h @ V_h, all heads in one launch."""
    lib = _lib.load()
    H, R, P = V.shape
    B = quad.shape[0]
    inter = torch.empty((B, H * P), device=quad.device, dtype=torch.float32)
    _run("headproj_fwd", lib.b200ctr_headproj_fwd, _p(_f32(quad)), _p(V), B, H, R, P, _p(inter), _stream())
    _count()
    return inter


def headproj_bwd(quad, dinter, V, dV):
    """Returns dquad (B, H*R); overwrites dV (H, R, P)."""
    lib = _lib.load()
    H, R, P = V.shape
    B = quad.shape[0]
    assert dinter.is_contiguous() and quad.is_contiguous() and dV.is_contiguous()
    dquad = torch.empty((B, H * R), device=quad.device, dtype=torch.float32)
    ws = torch.empty((lib.b200ctr_headproj_bwd_workspace_bytes(H, R, P),), device=quad.device, dtype=torch.uint8)
    _run("headproj_bwd", lib.b200ctr_headproj_bwd, _p(quad), _p(dinter), _p(V), B, H, R, P, _p(dquad), _p(dV), _p(ws), _stream())
    _count(2)
    return dquad


def fm_reduce(A, B, F, C):
    lib = _lib.load()
    s = torch.empty((B, C), device=A.device, dtype=torch.float32)
    quad = torch.empty((B, C), device=A.device, dtype=torch.float32)
    _run("fm_reduce", lib.b200ctr_fm_reduce, _p(A), B, F, C, _p(s), _p(quad), _stream())
    _count()
    return s, quad


def fm_grad_(A, s, dquad, B, F, C):
    lib = _lib.load()
    _run("fm_grad", lib.b200ctr_fm_grad, _p(A), _p(s), _p(dquad), B, F, C, _stream())
    _count()
    return A


def se_gate_fwd(csum, B, w1, b1, w2, b2):
    lib = _lib.load()
    C, Cr = w2.shape
    mean = torch.empty((C,), device=csum.device, dtype=torch.float32)
    h1 = torch.empty((Cr,), device=csum.device, dtype=torch.float32)
    gate = torch.empty((C,), device=csum.device, dtype=torch.float32)
    _run("se_gate_fwd", lib.b200ctr_se_gate_fwd, _p(csum), 1.0 / B, _p(w1), _p(b1), _p(w2), _p(b2), C, Cr, _p(mean), _p(h1), _p(gate),
                                  _stream())
    _count(2)
    return mean, h1, gate


def se_gate_bwd(dgate, mean, h1, gate, w1, w2):
    lib = _lib.load()
    C, Cr = w2.shape
    dw1, dw2 = torch.empty_like(w1), torch.empty_like(w2)
    db1 = torch.empty((Cr,), device=w1.device, dtype=torch.float32)
    db2 = torch.empty((C,), device=w1.device, dtype=torch.float32)
    dmean = torch.empty((C,), device=w1.device, dtype=torch.float32)
    _run("se_gate_bwd", lib.b200ctr_se_gate_bwd, _p(dgate), _p(mean), _p(h1), _p(gate), _p(w1), _p(w2), C, Cr, _p(dw1), _p(db1), _p(dw2),
                                  _p(db2), _p(dmean), _stream())
    _count()
    return dw1, db1, dw2, db2, dmean


def se_apply_fwd(inter, gate, dropout=NO_DROP):
    lib = _lib.load()
    B, C = inter.shape
    out = torch.empty_like(inter)
    _run("se_apply_fwd", lib.b200ctr_se_apply_fwd, _p(inter), _p(gate), B, C, _p(out), C, dropout, _stream())
    _count()
    return out


def se_apply_bwd(dout, inter, gate, dmean, B_total, want_dinter, want_gprod, dropout=NO_DROP):
    lib = _lib.load()
    B, C = inter.shape
    dinter = torch.empty_like(inter) if want_dinter else None
    gprod = torch.empty_like(inter) if want_gprod else None
    _run("se_apply_bwd", lib.b200ctr_se_apply_bwd, _p(dout), _ld(dout), _p(inter), _p(gate), _p(dmean), 1.0 / B_total, B, C, _p(dinter),
                                   _p(gprod), dropout, _stream())
    _count()
    return dinter, gprod


# ---- loss / optimiser -----------------------------------------------------------------------------------------
def bce_wll(logits, aux, y, aux_w, counts_override=None, want_prob=True):
    lib = _lib.load()
    B = logits.shape[0]
    dev = logits.device
    loss = torch.empty((1,), device=dev, dtype=torch.float32)
    counts = torch.empty((2,), device=dev, dtype=torch.int32)
    dz = torch.empty_like(logits)
    da = torch.empty_like(logits) if aux is not None else None
    prob = torch.empty_like(logits) if want_prob else None
    _run("bce_wll", lib.b200ctr_bce_wll, _p(logits), _p(aux), _p(y), float(aux_w), B, _p(counts_override), _p(loss), _p(counts),
                              _p(dz), _p(da), _p(prob), _stream())
    _count()
    return loss, dz, da, prob, counts


def sqnorm_(x, acc):
    lib = _lib.load()
    _run("sqnorm", lib.b200ctr_sqnorm, _p(x), x.numel(), _p(acc), _stream())
    _count()


def clip_coef(acc, max_norm, coef, total_norm=None):
    lib = _lib.load()
    _run("clip_coef", lib.b200ctr_clip_coef, _p(acc), float(max_norm), _p(coef), _p(total_norm), _stream())
    _count()


def adamw_ema_(p, g, m, v, ema, lr, wd, step, coef, ema_decay, beta1=0.9, beta2=0.999, eps=1e-8):
    lib = _lib.load()
    _run("adamw_ema", lib.b200ctr_adamw_ema, _p(p), _p(g), _p(m), _p(v), _p(ema), p.numel(), lr, wd, beta1, beta2, eps, step, _p(coef),
                                ema_decay, _stream())
    _count()


def rows_claim(tok, pad_id, owner):
    lib = _lib.load()
    _run("rows_claim", lib.b200ctr_rows_claim, _p(tok), tok.numel(), pad_id, _p(owner), _stream())
    _count()


def rows_sqnorm(tok, pad_id, owner, g_dense, acc):
    lib = _lib.load()
    _run("rows_sqnorm", lib.b200ctr_rows_sqnorm, _p(tok), tok.numel(), pad_id, _p(owner), _p(g_dense), g_dense.shape[1], _p(acc),
                                  _stream())
    _count()


def rows_zero(tok, pad_id, owner, g_dense):
    lib = _lib.load()
    _run("rows_zero", lib.b200ctr_rows_zero, _p(tok), tok.numel(), pad_id, _p(owner), _p(g_dense), g_dense.shape[1], _stream())
    _count()


def logit_mean(probs, eps=1e-7):
    lib = _lib.load()
    M, B = probs.shape
    out = torch.empty((B,), device=probs.device, dtype=torch.float32)
    _run("logit_mean", lib.b200ctr_logit_mean, _p(probs), M, B, eps, _p(out), _stream())
    _count()
    return out


def ensemble_logit_mean(logits, inv_T=None, eps=1e-7):
    """(M, B) fold logits [+ per-fold 1/T] -> sigmoid(mean_m logit(clip(sigmoid(z_m / T_m)))): infer.py:108-158 in one launch."""
    lib = _lib.load()
    M, B = logits.shape
    assert logits.is_contiguous() and (inv_T is None or (inv_T.numel() == M and inv_T.dtype == torch.float32))
    out = torch.empty((B,), device=logits.device, dtype=torch.float32)
    _run("ensemble_logit_mean", lib.b200ctr_ensemble_logit_mean, _p(_f32(logits)), _p(inv_T), M, B, eps, _p(out), _stream())
    _count()
    return out


# ---- row-lazy table optimiser -------------------------------------------------------------------------------------
class RowHistory:
    """Device-side per-step constants for the row-lazy AdamW/EMA (see csrc/rowopt.cu) + their host recurrences."""

    def __init__(self, device, capacity: int, beta1=0.9, beta2=0.999, eps=1e-8):
        import numpy as np
        self.np = np
        self.cap = int(capacity)
        self.h4 = torch.zeros((self.cap, 4), device=device, dtype=torch.float32)
        self.C = torch.ones(self.cap, device=device, dtype=torch.float32)
        self.S = torch.ones(self.cap, device=device, dtype=torch.float32)
        self.logD = torch.zeros(self.cap, device=device, dtype=torch.float32)
        self.b1, self.b2, self.eps = float(np.float32(beta1)), float(np.float32(beta2)), float(eps)
        self.c, self.s, self.ld = 1.0, 1.0, 0.0          # double recurrences on the host
        self.t = 0

    def push(self, t: int, lr: float, wd: float, ema_d: float):
        """Register step t (must be called once per step, in order)."""
        import math
        np = self.np
        assert t == self.t + 1, (t, self.t)
        if t >= self.cap:
            self._grow(max(2 * self.cap, t + 1))
        lib = _lib.load()
        lr32 = np.float32(lr)
        decay_mul = np.float32(1.0) - lr32 * np.float32(wd)                 # as adamw_ema_kernel: 1.f - lr*wd
        bc1 = np.float32(1.0 - self.b1 ** t)
        step_size = np.float32(lr32 / bc1)
        bc2s = np.float32(math.sqrt(1.0 - self.b2 ** t))
        d32 = np.float32(ema_d)
        self.c *= float(decay_mul)
        self.s = float(d32) * self.s + (1.0 - float(d32)) * self.c
        self.ld += math.log(float(d32)) if d32 > 0 else -1e30
        _run("rowopt_push", lib.b200ctr_rowopt_push, _p(self.h4), _p(self.C), _p(self.S), _p(self.logD), self.cap, t,
             float(decay_mul), float(step_size), float(bc2s), float(d32), self.c, self.s, self.ld, _stream())
        self.t = t
        _count()

    def _grow(self, cap: int):
        """Geometric growth of the device-side history (the recurrences live on the host, the arrays are plain prefixes):
        a run longer than the initial estimate -- more epochs, an underestimated steps_per_epoch, millions of steps -- keeps
        going like the reference's loop does.  Stream-ordered copies; the kernels receive the pointers per call."""
        def grown(old, fill):
            new = torch.full((cap,) + tuple(old.shape[1:]), fill, device=old.device, dtype=old.dtype)
            new[:old.shape[0]].copy_(old)
            return new
        self.h4, self.C, self.S, self.logD = grown(self.h4, 0.0), grown(self.C, 1.0), grown(self.S, 1.0), grown(self.logD, 0.0)
        self.cap = int(cap)

    def args(self):
        return (_p(self.h4), _p(self.C), _p(self.S), _p(self.logD), self.b1, self.b2, self.eps)


# ---- bitmap-driven row sets (see csrc/rowopt.cu) -------------------------------------------------------------------
def bitmap_words(vocab: int) -> int:
    return (int(vocab) + 31) // 32


def bitmap_mark(tok, pad_id, vocab, bitmap):
    """bitmap (int32 tensor of bitmap_words(vocab) words, zeroed by the caller) |= {rows listed in tok}."""
    lib = _lib.load()
    _run("bitmap_mark", lib.b200ctr_bitmap_mark, _p(tok), _idx64(tok), tok.numel(), pad_id, vocab, _p(bitmap), _stream())
    _count()


def rowopt_catchup_bm(bitmaps, n_maps, words, n_shards, rank, state, p, m, v, ema, target, hist: "RowHistory"):
    lib = _lib.load()
    _run("rowopt_catchup", lib.b200ctr_rowopt_catchup_bm, _p(bitmaps), n_maps, words, n_shards, rank, _p(state), _p(p), _p(m), _p(v),
         _p(ema), p.shape[1], target, *hist.args(), _stream())
    _count()


def rows_sqnorm_bm(bitmaps, n_maps, words, n_shards, rank, g, acc):
    lib = _lib.load()
    _run("rows_sqnorm", lib.b200ctr_rows_sqnorm_bm, _p(bitmaps), n_maps, words, n_shards, rank, _p(g), g.shape[1], _p(acc), _stream())
    _count()


def rowopt_update_bm(bitmaps, n_maps, words, n_shards, rank, state, p, g, m, v, ema, step, coef, hist: "RowHistory"):
    lib = _lib.load()
    _run("rowopt_update", lib.b200ctr_rowopt_update_bm, _p(bitmaps), n_maps, words, n_shards, rank, _p(state), _p(p), _p(g), _p(m),
         _p(v), _p(ema), p.shape[1], step, _p(coef), *hist.args(), _stream())
    _count()


def rows_claim_any(tok, pad_id, vocab, owner):
    lib = _lib.load()
    _run("rows_claim_any", lib.b200ctr_rows_claim_any, _p(tok), _idx64(tok), tok.numel(), pad_id, vocab, _p(owner), _stream())
    _count()


def rowopt_catchup(tok, pad_id, owner, state, p, m, v, ema, target, hist: RowHistory):
    lib = _lib.load()
    _run("rowopt_catchup", lib.b200ctr_rowopt_catchup, _p(tok), _idx64(tok), tok.numel(), pad_id, p.shape[0], _p(owner), _p(state),
         _p(p), _p(m), _p(v), _p(ema), p.shape[1], target, *hist.args(), _stream())
    _count()


def rowopt_update(tok, pad_id, owner, state, p, g, m, v, ema, step, coef, hist: RowHistory):
    lib = _lib.load()
    _run("rowopt_update", lib.b200ctr_rowopt_update, _p(tok), tok.numel(), pad_id, _p(owner), _p(state), _p(p), _p(g), _p(m), _p(v),
         _p(ema), p.shape[1], step, _p(coef), *hist.args(), _stream())
    _count()


def rowopt_flush(state, p, m, v, ema, target, hist: RowHistory):
    lib = _lib.load()
    _run("rowopt_flush", lib.b200ctr_rowopt_flush, _p(state), _p(p), _p(m), _p(v), _p(ema), p.shape[0], p.shape[1], target,
         *hist.args(), _stream())
    _count()
