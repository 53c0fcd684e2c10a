/*
 * b200ctr.h — C ABI of libb200ctr.so: the sm_100a implementation of the DARE + QNN-alpha
 * training / inference hot path of biyotteu/toss-next-ctr-prediction.
 *
 * The reference has no FFI/plugin layer: its boundary is the Python surface of
 * src/models/{wrapper,dare,qnn_alpha,feature_embed}.py plus the step body of src/train.py:152-203.
 * Every entry point below replaces the ATen/cuBLAS op chain behind one of those call sites
 * (cited per function, file:line relative to the reference root).  INTEGRATION.md shows the
 * ctypes binding and the namespace-package overlay that makes the reference's own train.py/infer.py
 * call into this library unmodified.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch) owns all memory;
 *   - row-major tensors, explicit leading dimensions (in elements) where a slice of a wider buffer is legal;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all calls are
 *     asynchronous on it, never synchronise, never allocate;
 *   - return value: 0 = OK, negative = B200CTR_E_* ; b200ctr_last_error() gives a thread-local message;
 *   - index tensors (seq, x_cat) come as int64 (the collate contract, dataset.py:98-124) or int32 (the on-disk
 *     dtype) selected by `idx_is_64`;
 *   - dropout uses a counter-based hash RNG keyed by (seed, stream_id, element index): the backward pass
 *     regenerates masks instead of storing them.  p = 0 disables it (eval / parity mode).
 *   - there is no CPU fallback: a non-sm_100 device yields B200CTR_E_ARCH.
 */
#ifndef B200CTR_H
#define B200CTR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CTR_OK 0
#define B200CTR_E_INVALID_ARG (-1)
#define B200CTR_E_UNSUPPORTED_SHAPE (-2)
#define B200CTR_E_CUDA (-3)
#define B200CTR_E_ARCH (-4)

#define B200CTR_ABI_VERSION 1

int b200ctr_version(void);
const char* b200ctr_last_error(void);
/* 0 if the current device is sm_100 (B200); B200CTR_E_ARCH otherwise. */
int b200ctr_check_device(void);
/* Enable loads / stores / atomics from the current device into `peer_device`'s memory (row-sharded tables whose remote
 * shards are mapped through CUDA IPC).  Idempotent.                                                            */
int b200ctr_enable_peer_access(int peer_device);
/* CUDA IPC plumbing for the row-sharded tables: export = 64-byte handle of the allocation containing `ptr` + the byte
 * offset of `ptr` in it; import (on the consumer's current device) = base address of that allocation mapped into this
 * process with peer access enabled.  An allocation must be imported only once per process (cache by handle).       */
int b200ctr_ipc_export(const void* ptr, void* handle_out64, int64_t* offset_out);
int b200ctr_ipc_import(const void* handle64, void** base_out);

/* ---- dropout descriptor ------------------------------------------------------------------------- */
typedef struct {
  float p;            /* drop probability; 0 => identity */
  uint32_t seed_lo;   /* 64-bit seed */
  uint32_t seed_hi;
  uint32_t stream_id; /* distinct per dropout site and per step */
} b200ctr_dropout_t;

/* ---- dense contraction ---------------------------------------------------------------------------
 * C[M,N] = epilogue( alpha * opA(A)[M,K] . opB(B)[K,N] )
 *   transA = 0: A stored [M,K] (lda >= K)      transA = 1: A stored [K,M] (lda >= M)
 *   transB = 0: B stored [K,N] (ldb >= N)      transB = 1: B stored [N,K] (ldb >= K)   (nn.Linear weight)
 * epilogue order: v = alpha*acc; v += accum[m,n]; v += bias[n]; if pre_out: pre_out = v; v = act(v);
 *                 v *= dropout keep/(1-p); v *= mask(mask_src); v += residual; C = v.
 *                 (accum = pre-activation addend: a K-split contraction such as cat([z, inter]) @ W^T;
 *                  residual = post-activation addend; both may alias C.)
 * Replaces every F.linear / matmul on the path: dare.py:64,67 (in_proj, out_proj, FFN), qnn_alpha.py:92,96,129
 * (z.U, quad.V, MLP), wrapper.py:41,97 (ctx_mlp, fc head) and their autograd backward GEMMs.
 * `workspace` (>= b200ctr_gemm_workspace_bytes) is required only when split_k > 1.                       */
#define B200CTR_ACT_NONE 0
#define B200CTR_ACT_RELU 1
#define B200CTR_ACT_GELU 2
#define B200CTR_MASK_NONE 0
#define B200CTR_MASK_NONZERO 1  /* v *= (mask_src != 0) ? mask_scale : 0   (ReLU(+dropout) backward from the saved output) */
#define B200CTR_MASK_GELU_GRAD 2 /* v *= gelu'(mask_src)                    (GELU backward from the saved pre-activation)    */
typedef struct {
  float alpha;
  const float* bias;       /* [N] or NULL */
  int act;
  float* pre_out;          /* [M,N] (ld_pre) or NULL */
  int64_t ld_pre;
  b200ctr_dropout_t drop;  /* applied after act (row key = m, column = n) */
  int mask_mode;
  const float* mask_src;   /* [M,N] (ld_mask) */
  int64_t ld_mask;
  float mask_scale;
  const float* residual;   /* [M,N] (ld_res) or NULL */
  int64_t ld_res;
  const float* accum;      /* [M,N] (ld_accum) or NULL */
  int64_t ld_accum;
} b200ctr_epilogue_t;

int64_t b200ctr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int split_k);
int b200ctr_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K,
                 const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                 const b200ctr_epilogue_t* epi_host, int split_k, void* workspace, void* stream);
/* Same contract as b200ctr_gemm, computed on the tcgen05 tensor cores: operands are rounded to bf16 on their way
 * into shared memory (the reference's `amp: bf16` autocast semantics for matmul / linear), accumulation and the
 * epilogue are fp32, inputs and outputs stay fp32 in HBM.  transA && transB is not built.                        */
int b200ctr_gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K,
                      const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                      const b200ctr_epilogue_t* epi_host, int split_k, void* workspace, void* stream);
/* Same contract as b200ctr_gemm on the tcgen05 tensor cores at fp32-grade accuracy (the `amp: none` mode, SURVEY hard
 * part 5): each fp32 operand is split into `parts` bf16 pieces on its way into shared memory (a0 = bf16(a), a1 = bf16(a - a0),
 * a2 = bf16(a - a0 - a1)) and the product is accumulated as the sum of a_i.b_j over i + j < parts in the fp32 TMEM accumulator:
 * parts = 2 -> 3 MMAs per k-step, error ~2^-17 per product term; parts = 3 -> 6 MMAs, dropped terms ~2^-24 (fp32 rounding).
 * Replaces the reference's true-fp32 cuBLAS matmul (TF32 off): src/models/dare.py:64-69, qnn_alpha.py:89-97,123-129 under
 * `amp: none`.  transA && transB is not built.                                                                          */
int b200ctr_gemm_split(int parts, int transA, int transB, int64_t M, int64_t N, int64_t K,
                       const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                       const b200ctr_epilogue_t* epi_host, int split_k, void* workspace, void* stream);
/* Same contraction with operands ALREADY in bf16 (row-major, leading dimensions in elements): activations that their
 * producer kernel also emitted in bf16 (b200ctr_rowrms_fwd_bf16) and weights converted once per step (b200ctr_to_bf16);
 * staging is then a pure 16-byte copy into the UMMA core-matrix layout.  Persistent warp-specialised kernel.       */
int b200ctr_gemm_bf16in(int transA, int transB, int64_t M, int64_t N, int64_t K,
                        const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                        const b200ctr_epilogue_t* epi_host, int split_k, void* workspace, void* stream);
int b200ctr_to_bf16(const float* x, int64_t n, void* out_bf16, void* stream);
/* Deep-K contraction fed by TMA: C (M,N) fp32 = A (M,K) . B (N,K)^T with BOTH operands bf16 and K contiguous
 * (cp.async.bulk.tensor 128-byte-swizzled tiles -> tcgen05.mma, persistent, double-buffered TMEM accumulators).
 * N % 32 == 0; K, lda, ldb % 8 == 0; 16-byte aligned bases.  Helpers to put operands into that form:
 * to_bf16_transposed: fp32 (R,C) row stride ld_in -> bf16 (C,R) contiguous;  transpose_bf16: bf16 (R,C) -> bf16 (C,R).   */
int b200ctr_gemm_bf16_kk(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                         int64_t N, int64_t K, void* stream);
int b200ctr_to_bf16_transposed(const float* in, int64_t R, int64_t C, int64_t ld_in, void* out_bf16, void* stream);
int b200ctr_transpose_bf16(const void* in_bf16, int64_t R, int64_t C, void* out_bf16, void* stream);
int b200ctr_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* workspace, void* stream);

/* ---- numeric / binary feature embedding  (feature_embed.py:19-27, 42-48) ------------------------------
 * fold:  P = w @ w_out^T (F,D),  Q = b @ w_out^T (F,D) (Q = 0 when b == NULL)
 * fwd :  out[b,f,:] = x[b,f]*P[f,:] + Q[f,:]   written at out + b*ld_b + f*D (skipped when out == NULL);
 *        mean_out[b,:] (ld_mean) = mean_f out[b,f,:]
 * bwd :  dP[f,:] = sum_b x[b,f]*g[b,f,:], dQ[f,:] = sum_b g[b,f,:], g = dout[b,f,:] + dmean[b,:]/F
 * unfold: dw = dP @ w_out, db = dQ @ w_out, dw_out = dP^T @ w + dQ^T @ b                                   */
int b200ctr_numbin_fold(const float* w, const float* b, const float* w_out, float* P, float* Q,
                        int F, int fe, int D, void* stream);
int b200ctr_numbin_fwd(const float* x, const float* P, const float* Q, float* out, int64_t ld_b,
                       float* mean_out, int64_t ld_mean, int64_t B, int F, int D, void* stream);
/* numbin_bwd writes PARTIAL sums: dP / dQ are (b200ctr_numbin_bwd_rows(B), F*D) matrices, one row per 64-sample chunk, every
 * entry written (no atomics); the caller sums the rows in order (b200ctr_colsum) -> bit-identical runs.                 */
int64_t b200ctr_numbin_bwd_rows(int64_t B);
int b200ctr_numbin_bwd(const float* x, const float* dout, int64_t ld_b, const float* dmean, int64_t ld_dmean,
                       float* dP, float* dQ, int64_t B, int F, int D, void* stream);
int b200ctr_numbin_unfold(const float* dP, const float* dQ, const float* w, const float* b, const float* w_out,
                          float* dw, float* db, float* dw_out, int F, int fe, int D, void* stream);

/* ---- hashed-categorical gather + projection  (wrapper.py:106-112, 149-150, 119-121, 128-136) ---------
 * tables/projs: device arrays of Fc device pointers; dims/cards: device int32[Fc].
 * fwd: e = E_c[x_cat[b,c]] @ W_c^T ; out[b,c,:] = dropout(e) at out + b*ld_b + c*D ; mean_out = mean_c e (pre-dropout);
 *      q_out[b,:] = e[b,q_idx,:] (pre-dropout, S1 query) when q_idx >= 0.
 * bwd: g = dropout'(dout[b,c,:]) + dmean[b,:]/Fc + (c==q_idx)*dq[b,:]; dE_c[idx] += W_c^T g ; dW_c += g (x) row.   */
/* lds: device int32[Fc] row strides of the tables in elements (NULL = d_c, i.e. dense rows): tables kept under the row-lazy
 * optimizer are stored 64 floats wide.                                                                                */
int b200ctr_cat_fwd(const void* x_cat, int idx_is_64, const float* const* tables, const float* const* projs,
                    const int32_t* dims, const int32_t* cards, const int32_t* lds, float* out, int64_t ld_b,
                    float* mean_out, int64_t ld_mean, int q_idx, float* q_out,
                    int64_t B, int Fc, int D, b200ctr_dropout_t drop, void* stream);
/* cat_bwd is deterministic (no float atomics): kernel 1 writes every (sample, column) slot's contribution W_c^T g as a 64-wide
 * row plus the slot's global table row id (row_base_c + id), and per-CTA partials of dW_c; b200ctr_scatter_rows sorts the slots
 * by table row and reduces each distinct row once, in (sample, column) order; the reduced rows are added into dE_c and the
 * dW_c partials are added in CTA order.  sum_dims = sum_c d_c, sum_cards = sum_c card_c, max_card = max_c card_c; dtables /
 * dprojs are accumulated into (single writer per element).  The same path serves HBM-sized hashed tables (v4 cfgs):
 * rows_acc (optional) = a caller-owned ALL-ZERO (sum_cards, 64) buffer in global-row order that outlives the call, and
 * keep_rows[c] != 0 leaves table c's reduced 64-wide rows there (they are the gradient rows the row-lazy optimizer
 * consumes and re-zeroes) instead of adding them into dtables[c]; the other tables' rows are zeroed again.
 * defer_slot[c] >= 0 (data parallel, replicated row-lazy tables): column c's slots are NOT reduced here; sample b's 64-wide
 * contribution row and its global row id are written to defer_rows[b, defer_slot[c]] / defer_keys[b, defer_slot[c]]
 * (n_defer slots per sample) for the caller to all-gather and reduce with b200ctr_scatter_rows.                          */
int64_t b200ctr_cat_bwd_workspace_bytes(int64_t B, int Fc, int sum_dims, int64_t sum_cards);
int b200ctr_cat_bwd(const void* x_cat, int idx_is_64, const float* const* tables, const float* const* projs,
                    const int32_t* dims, const int32_t* cards, const int32_t* lds, const float* dout, int64_t ld_b,
                    const float* dmean, int64_t ld_dmean, int q_idx, const float* dq, float* const* dtables,
                    float* const* dprojs, int64_t B, int Fc, int D, int sum_dims, int64_t sum_cards, int max_card,
                    b200ctr_dropout_t drop, float* rows_acc, const int32_t* keep_rows, const int32_t* defer_slot,
                    int n_defer, float* defer_rows, int32_t* defer_keys, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* ---- query mix  (wrapper.py:128-136): q = a*s1 + b*s2 ; bwd: ds2pre = b*dq*(s2>0) --------------------- */
int b200ctr_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream);
/* out = scale * dy * (y > 0)   (ReLU backward from the saved output, ctx_mlp: wrapper.py:41-43) */
int b200ctr_relu_bwd(const float* dy, const float* y, float scale, float* out, int64_t n, void* stream);
/* out = sigmoid(x)   (prob = sigmoid(logits), wrapper.py:175) */
int b200ctr_sigmoid(const float* x, float* out, int64_t n, void* stream);
/* dst[r, c] += src[r, c] over a (rows, cols) window of two row-strided buffers */
int b200ctr_add2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* stream);

/* ---- DARE retrieval  (dare.py:116-138) -----------------------------------------------------------------
 * fwd: score[b,p] = dot(emb_att[seq[b,p]], q[b]) + rec[p]; pad -> -1e9; sorted top-K under the canonical
 *      order (score desc, position asc); sel_rep[b,k,:] = emb_rep[seq[b,idx[b,k]]] (zero row for pad slots).
 *      Outputs: sel_rep (B,K,D), vals (B,K), idx (B,K) int32 positions, tok (B,K) int32 tokens (pad_id at pad slots).
 * bwd: dq[b,:] = sum_k dvals[b,k]*emb_att[tok]; sparse table gradients, accumulated either
 *      into dense gradient tables (datt/drep, atomics) ...                                                  */
int b200ctr_retrieve_fwd(const void* seq, int idx_is_64, const float* q, const float* emb_att, const float* emb_rep,
                         const float* rec, int64_t B, int L, int K, int D, int pad_id, int64_t vocab,
                         float* sel_rep, float* vals, int32_t* idx, int32_t* tok, void* stream);
int b200ctr_retrieve_bwd(const int32_t* tok, const float* q, const float* emb_att,
                         const float* dsel_rep, const float* dvals, int64_t B, int K, int D, int pad_id,
                         float* dq, float* datt_dense, float* drep_dense, void* stream);

/* Row-sharded form (data parallel, one process per GPU): row r of a table lives on shard r % n_shards at local row
 * r / n_shards.  att_shards / rep_shards / d*_shards are HOST arrays of n_shards device pointers (peer-mapped over
 * NVLink for the remote shards), so the gather and the gradient scatter cross GPUs inside the kernel.
 * n_shards is a power of two <= 8; n_shards == 1 is the plain layout above.                                  */
int b200ctr_retrieve_fwd_sharded(const void* seq, int idx_is_64, const float* q, const float* const* att_shards,
                                 const float* const* rep_shards, int n_shards, const float* rec, int64_t B, int L, int K,
                                 int D, int pad_id, int64_t vocab, float* sel_rep, float* vals, int32_t* idx, int32_t* tok,
                                 void* stream);
int b200ctr_retrieve_bwd_sharded(const int32_t* tok, const float* q, const float* const* att_shards, int n_shards,
                                 const float* dsel_rep, const float* dvals, int64_t B, int K, int D, int pad_id,
                                 float* dq, float* const* datt_shards, float* const* drep_shards, void* stream);

/* Staged scatter for row-sharded tables: gradient rows of tokens owned by another rank are written with plain stores
 * into that owner's staging buffer stage[owner][my_rank][slot] (slot = b*K + k; stage_* = HOST arrays of n_shards device
 * pointers to each rank's (n_shards, stage_cap, D) buffer, peer-mapped), rows owned by this rank are reduced locally.
 * After a cross-rank barrier the owner runs b200ctr_stage_apply over the all-gathered token ids gtok (n_src, n_slots)
 * to add the staged rows into its local gradient shard.                                                        */
int b200ctr_retrieve_bwd_staged(const int32_t* tok, const float* q, const float* const* att_shards, int n_shards,
                                const float* dsel_rep, const float* dvals, int64_t B, int K, int D, int pad_id,
                                float* dq, float* const* datt_shards, float* const* drep_shards,
                                float* const* stage_att, float* const* stage_rep, int my_rank, int64_t stage_cap, int stage_all,
                                void* stream);
int b200ctr_stage_apply(const int32_t* gtok, int n_src, int64_t n_slots, int64_t stage_cap, int pad_id, int n_shards,
                        int rank, const float* stage_att, const float* stage_rep, float* g_att, float* g_rep, int D,
                        void* stream);

/* ---- DARE encoder layer pieces  (dare.py:29-37, 53-70) ---------------------------------------------------
 * attention: qkv (M=B*K, 3D) packed [q|k|v]; bias[i,j] = relmean[j-i+max_len] (relmean = mean over heads of
 *   pbias.rel.weight, length 2*max_len+1, may be NULL); softmax(q k^T/sqrt(hd) + bias); dropout on probabilities;
 *   o (M,D); lse (B,H,K) saved for the backward, which recomputes probabilities.
 * add_rmsnorm: s = a + b ; y = w * s * rsqrt(mean(s^2)+eps) ; saves s and r.                                 */
int b200ctr_attn_fwd(const float* qkv, const float* relmean, int max_len, int64_t B, int K, int H, int D,
                     float* o, float* lse, b200ctr_dropout_t drop, void* stream);
/* drelmean (both attention backward kernels): a ZEROED (B, 2*max_len+1) matrix of per-sample partial rows, written without
 * atomics; d relmean = its column sum in sample order (b200ctr_colsum) -> bit-identical runs.  NULL = not wanted.        */
int b200ctr_attn_bwd(const float* qkv, const float* relmean, int max_len, const float* o, const float* lse,
                     const float* d_o, int64_t B, int K, int H, int D, float* dqkv, float* drelmean,
                     b200ctr_dropout_t drop, void* stream);
/* Same contracts on the warp-level tensor cores (mma.sync m16n8k8, TF32 operands, fp32 accumulation): the
 * `amp: bf16` path.  head_dim 8 == the k extent of the instruction; dropout masks are identical to the fp32 kernels'.
 * keep (optional, b200ctr_attn_tc_keep_words(B, K) 32-bit words): the forward stores one keep flag per probability,
 * the backward reads them instead of re-hashing; NULL = the backward recomputes the mask (same result).            */
int64_t b200ctr_attn_tc_keep_words(int64_t B, int K);
int b200ctr_attn_tc_fwd(const float* qkv, const float* relmean, int max_len, int64_t B, int K, int H, int D,
                        float* o, float* lse, uint32_t* keep, b200ctr_dropout_t drop, void* stream);
/* dqkv_bf16 != 0: dqkv is written as bf16 (same (M, 3D) layout) for b200ctr_inproj_bwd(dqkv_bf16 = 1), which rounds
 * it to bf16 for its tensor-core products in any case (identical results, half the traffic).                       */
int b200ctr_attn_tc_bwd(const float* qkv, const float* relmean, int max_len, const float* o, const float* lse,
                        const float* d_o, int64_t B, int K, int H, int D, void* dqkv, int dqkv_bf16, float* drelmean,
                        const uint32_t* keep, b200ctr_dropout_t drop, void* stream);
/* Same contracts again in PACKED HALF arithmetic (csrc/attn_h16.cu; the default of the `amp: bf16` path): scores with
 * the bias as the accumulator input of mma.m16n8k8.f16, max / exp2 / dropout mask on half2 pairs, probabilities straight
 * into mma.m16n8k16 (fp32 accumulate), operand fragments by ldmatrix from transposed fp16 images; the backward rescales dO
 * per sample by a power of two.  Under torch autocast the reference rounds scores and probabilities to bf16 (8 bits);
 * fp16 keeps 11.  Same keep-word count (b200ctr_attn_tc_keep_words) but its own bit layout: pair the h16 forward with
 * the h16 backward.  Dropout masks are identical to the other attention kernels'.                                     */
/* the h16 backward runs two CTAs per sample (four heads each): its drelmean is a ZEROED (b200ctr_attn_h16_drel_rows(B),
 * 2*max_len+1) matrix of per-CTA partial rows (column-summed in row order, like the others').                           */
int64_t b200ctr_attn_h16_drel_rows(int64_t B);
/* qkv_f16 != 0: qkv is the (M, 3D) fp16 tensor written by b200ctr_inproj_fwd_h16 (q already times log2(e) / sqrt(hd)): the
 * kernels then stage their operands with plain copies -- same values, half the bytes; 0: the fp32 tensor of the other kernels. */
/* o_bf16 / do_bf16 != 0: o (written by the forward, read by the backward) and d_o (read by the backward) are bf16 (M, D): see
 * b200ctr_outproj_norm_fwd/bwd.                                                                                         */
int b200ctr_attn_h16_fwd(const void* qkv, int qkv_f16, const float* relmean, int max_len, int64_t B, int K, int H, int D,
                         void* o, int o_bf16, float* lse, uint32_t* keep, b200ctr_dropout_t drop, void* stream);
int b200ctr_attn_h16_bwd(const void* qkv, int qkv_f16, const float* relmean, int max_len, const void* o, int o_bf16,
                         const float* lse, const void* d_o, int do_bf16, int64_t B, int K, int H, int D, void* dqkv,
                         int dqkv_bf16, float* drelmean, const uint32_t* keep, b200ctr_dropout_t drop, void* stream);
int b200ctr_add_rmsnorm_fwd(const float* a, const float* b, const float* w, float eps, int64_t M, int D,
                            float* s_out, float* y_out, float* r_out, void* stream);
/* dx = rmsnorm'(dy) (+ dx_add) ; dw: (b200ctr_rmsnorm_bwd_rows(M), D) PARTIAL sums of dy*s*r, one row per CTA, every entry
 * written (no atomics); the caller sums the rows in order (b200ctr_colsum).                                            */
int64_t b200ctr_rmsnorm_bwd_rows(int64_t M);
int b200ctr_rmsnorm_bwd(const float* dy, const float* s, const float* w, const float* r, int64_t M, int D,
                        const float* dx_add, float* dx, float* dw, void* stream);
/* relmean[t] = mean_h rel[t,h] ; drel[t,h] = drelmean[t]/H */
int b200ctr_relmean_fwd(const float* rel, int T, int H, float* relmean, void* stream);
int b200ctr_relmean_bwd(const float* drelmean, int T, int H, float* drel, void* stream);

/* ---- fused encoder FFN block on tcgen05 (dare.py:67-69; the `amp: bf16` path) -----------------------------
 * x2 = RMSNorm_n2(s2), s2 = x1 + W2.dropout(GELU(W1.x1 + b1)) + b2 ; operands rounded to bf16, fp32 accumulation
 * in TMEM; the (T, hidden) activation never reaches HBM.  w1 (hidden, D), w2 (D, hidden) row-major as in the
 * reference's nn.Linear; dropout element = (row m, column n) (identical to the unfused gemm epilogue, so the
 * backward regenerates the same mask).  s2_out / r2_out (saved for the backward) may be NULL.
 * D == 64, hidden % 64 == 0, hidden <= 512.                                                                 */
/* keep_bits (optional, b200ctr_ffn_keep_bytes(T, hidden) bytes): the forward stores one dropout keep flag per hidden
 * activation, b200ctr_ffn_bwd reads them instead of re-hashing; NULL = the backward regenerates the same mask.        */
int64_t b200ctr_ffn_keep_bytes(int64_t T, int hidden);
int b200ctr_ffn_fwd(const float* x1, const float* w1, const float* b1, const float* w2, const float* b2,
                    const float* n2w, float eps, int64_t T, int D, int hidden, float* s2_out, float* r2_out,
                    float* x2_out, void* keep_bits, b200ctr_dropout_t drop, void* stream);

/* Backward of b200ctr_ffn_fwd: recomputes the hidden activation from x1 (dropout mask regenerated), accumulates
 * dW1 / dW2 in TMEM across a persistent loop (deterministic cross-CTA reduction afterwards) and returns
 * dx1 = dL/dx1 (including the residual branch).  dw1 (hidden, D), db1 (hidden), dw2 (D, hidden), db2 (D), dn2 (D)
 * are overwritten.  workspace >= b200ctr_ffn_bwd_workspace_bytes(T, hidden).  hidden % 128 == 0, hidden <= 384.  */
int64_t b200ctr_ffn_bwd_workspace_bytes(int64_t T, int hidden);
int b200ctr_ffn_bwd(const float* dy, const float* s2, const float* r2, const float* n2w, const float* x1,
                    const float* w1, const float* b1, const float* w2, int64_t T, int D, int hidden,
                    float* dx1, float* dw1, float* db1, float* dw2, float* db2, float* dn2,
                    void* workspace, const void* keep_bits, b200ctr_dropout_t drop, void* stream);

/* ---- fused attention-side projections on tcgen05 (dare.py:64-66; the `amp: bf16` path) ---------------------------
 * inproj_fwd:        qkv (T,192) = x . w_in^T + b_in                              (w_in (192,64) = mha.in_proj_weight)
 * outproj_norm_fwd:  s1 = x_resid + o . w_out^T + b_out ; r1 = rsqrt(mean(s1^2)+eps) ; x1 = n1w * s1 * r1
 * outproj_norm_bwd:  ds1 = RMSNorm'(dx1; s1, r1, n1w) ; d_o = ds1 . w_out ; dw_out, db_out, dn1 (overwritten)
 * inproj_bwd:        dx = dqkv . w_in + d_resid (may be NULL) ; dw_in, db_in (overwritten)
 * Persistent kernels, weights resident in shared memory, weight AND bias gradients accumulated in TMEM (the bias
 * gradient is the 65th accumulator column: the activation operand carries a ones column), deterministic cross-CTA
 * reduction.  workspace >= b200ctr_proj_bwd_workspace_bytes().  D == 64 (8 heads x head_dim 8).                  */
int b200ctr_inproj_fwd(const float* x, const float* w_in, const float* b_in, int64_t T, int D, float* qkv, void* stream);
/* the same projection with qkv written as fp16 (T, 192), its q third times log2(e) / sqrt(head_dim = 8): the operand format of
 * b200ctr_attn_h16_fwd/bwd(qkv_f16 = 1) -- bit-identical attention results, half the qkv traffic in three kernels            */
int b200ctr_inproj_fwd_h16(const float* x, const float* w_in, const float* b_in, int64_t T, int D, void* qkv16, void* stream);
/* o_bf16 / do_bf16: the attention output o and its gradient d_o travel as bf16 (T, 64) between the packed-half attention
 * kernels and the out-projection: both projections' tensor-core products round them to bf16 in any case, so only the attention
 * backward's D = dO . o sees the narrower type (the reference's autocast keeps both in bf16 as well).                        */
int b200ctr_outproj_norm_fwd(const void* o, int o_bf16, const float* w_out, const float* b_out, const float* x_resid,
                             const float* n1w, float eps, int64_t T, int D, float* s1_out, float* r1_out,
                             float* x1_out, void* stream);
int64_t b200ctr_proj_bwd_workspace_bytes(void);
int b200ctr_outproj_norm_bwd(const float* dx1, const float* s1, const float* r1, const float* n1w, const void* o, int o_bf16,
                             const float* w_out, int64_t T, int D, float* ds1_out, void* d_o, int do_bf16, float* dw_out,
                             float* db_out, float* dn1, void* workspace, void* stream);
int b200ctr_inproj_bwd(const void* dqkv, int dqkv_bf16, const float* x, const float* w_in, const float* d_resid, int64_t T,
                       int D, float* dx, float* dw_in, float* db_in, void* workspace, void* stream);

/* ---- gated pooling + aux head  (dare.py:151-161) ------------------------------------------------------- */
int b200ctr_pool_fwd(const float* x, const float* vals, const float* aux_w, const float* aux_b, int gating_relu,
                     int64_t B, int K, int D, float* u, int64_t ld_u, float* aux, float* w_out,
                     b200ctr_dropout_t drop, void* stream);
/* daux_w / daux_b of pool_bwd are per-sample PARTIAL rows, (B, D) and (B), every entry written (no atomics); the aux-head
 * gradients are their column sums in sample order (b200ctr_colsum).                                                      */
int b200ctr_pool_bwd(const float* x, const float* vals, const float* w, const float* aux_w, const float* u, int64_t ld_u,
                     const float* du, int64_t ld_du, const float* daux, int gating_relu, int64_t B, int K, int D,
                     float* dx, float* dvals, float* daux_w, float* daux_b, b200ctr_dropout_t drop, void* stream);

/* ---- QNN-alpha pieces  (qnn_alpha.py:24-26, 86-97, 109-130) ----------------------------------------------
 * rowrms: z[b,:] = w * x[b,:] * rsqrt(mean(x^2)+eps) over N = F*D columns (z written with ld_z), r[b].
 * fm_reduce: A (B,F,C) -> s[b,c] = sum_f A, quad[b,c] = s^2 - sum_f A^2.
 * fm_grad:   dA[b,f,c] = dquad[b,c] * 2 * (s[b,c] - A[b,f,c])  (in place over A).
 * se_gate:   m = colsum/B; gate = sigmoid(W2 relu(W1 m + b1) + b2) (single CTA); bwd gives dm and weight grads.
 * se_apply:  out = inter * gate * dropout ; bwd: dinter = dout*drop*gate + dm/B ; gprod = dout*drop*inter.     */
int b200ctr_rowrms_fwd_bf16(const float* x, int64_t ldx, const float* w, float eps, int64_t B, int64_t N,
                            float* z, int64_t ld_z, float* r, void* z_bf16, void* stream);   /* + bf16 copy of z (B, N) */
int b200ctr_rowrms_fwd(const float* x, int64_t ldx, const float* w, float eps, int64_t B, int64_t N,
                       float* z, int64_t ld_z, float* r, void* stream);
int b200ctr_rowrms_bwd(const float* dz, int64_t ld_dz, const float* x, int64_t ldx, const float* w, const float* r,
                       int64_t B, int64_t N, float* dx, int64_t ld_dx, float* dw_partial, int n_partials, void* stream);
/* Fused quadratic interaction on tcgen05 (qnn_alpha.py:86-97; `amp: bf16` path): z_bf16 = the bf16 copy of the normed
 * features (B*F, 64) from b200ctr_rowrms_fwd_bf16, Ucat (64, C) = qnn.U laid out [d][head*rank].
 *   fwd: s[b,c] = sum_f (z.Ucat), quad = s^2 - sum_f (z.Ucat)^2          -- the (B*F, C) product never reaches HBM
 *   bwd: dz (B*F, 64) = (dquad * 2 (s - z.Ucat)) . Ucat^T + dz_resid (may be NULL / may alias dz) ; dUcat (64, C) overwritten
 * F <= 256, C % 32 == 0, C <= 128, D == 64.  workspace >= b200ctr_fm_fused_bwd_workspace_bytes().                */
int b200ctr_fm_fused_fwd(const void* z_bf16, const float* Ucat, int64_t B, int F, int C, int D, float* s_out,
                         float* quad_out, void* stream);
int64_t b200ctr_fm_fused_bwd_workspace_bytes(void);
int b200ctr_fm_fused_bwd(const void* z_bf16, const float* Ucat, const float* s, const float* dquad,
                         const float* dz_resid, int64_t B, int F, int C, int D, float* dz, float* dUcat,
                         void* workspace, void* stream);
/* The MLP's final Linear(K -> 1) (qnn_alpha.py:85-93) as row dot products instead of 1-wide GEMMs:
 *   fwd: out[b] = h[b,:] . w + bias[0]
 *   bwd: dh[b,n] = dl[b] * w[n] * (act[b,n] != 0 ? mask_scale : 0) ; dw[n] = sum_b dl[b] act[b,n] ; db[0] = sum_b dl[b]
 *        (act = the layer's input after ReLU + dropout, whose mask it carries; K a multiple of 128, <= 512)          */
int b200ctr_lin1_fwd(const float* h, int64_t ldh, const float* w, const float* bias, int64_t B, int K, float* out, void* stream);
int64_t b200ctr_lin1_bwd_workspace_bytes(int K);
int b200ctr_lin1_bwd(const float* dl, const float* act, int64_t ldact, const float* w, int64_t B, int K, float mask_scale,
                     float* dh, int64_t lddh, float* dw, float* db, void* workspace, void* stream);
/* headproj: inter[b, h*P+p] = sum_r quad[b, h*R+r] * V[h][r][p] for all heads in one launch (qnn_alpha.py:96);
 * bwd: dquad and dV (overwritten; deterministic two-stage batch reduction through `workspace`).              */
int b200ctr_headproj_fwd(const float* quad, const float* V, int64_t B, int H, int R, int P, float* inter, void* stream);
int64_t b200ctr_headproj_bwd_workspace_bytes(int H, int R, int P);
int b200ctr_headproj_bwd(const float* quad, const float* dinter, const float* V, int64_t B, int H, int R, int P,
                         float* dquad, float* dV, void* workspace, void* stream);
int b200ctr_fm_reduce(const float* A, int64_t B, int F, int C, float* s, float* quad, void* stream);
int b200ctr_fm_grad(float* A, const float* s, const float* dquad, int64_t B, int F, int C, void* stream);
int b200ctr_se_gate_fwd(const float* colsum, float inv_B, const float* w1, const float* b1, const float* w2,
                        const float* b2, int C, int Cr, float* mean, float* h1, float* gate, void* stream);
int b200ctr_se_gate_bwd(const float* dgate, const float* mean, const float* h1, const float* gate,
                        const float* w1, const float* w2, int C, int Cr,
                        float* dw1, float* db1, float* dw2, float* db2, float* dmean, void* stream);
int b200ctr_se_apply_fwd(const float* inter, const float* gate, int64_t B, int C, float* out, int64_t ld_out,
                         b200ctr_dropout_t drop, void* stream);
int b200ctr_se_apply_bwd(const float* dout, int64_t ld_dout, const float* inter, const float* gate, const float* dmean,
                         float inv_B, int64_t B, int C, float* dinter, float* gprod, b200ctr_dropout_t drop, void* stream);

/* ---- loss  (train.py:71-90, 165-168) ---------------------------------------------------------------------
 * loss = wll(logits,y) + aux_w * wll(aux,y), wll = 0.5*(mean_{y=1} softplus(-z) + mean_{y=0} softplus(z)).
 * Writes loss[0], counts[0..1] = (n_pos, n_neg), dlogits, daux (daux = 0 when aux_w == 0), prob = sigmoid(logits).
 * n_pos_total / n_neg_total > 0 override the local counts (data-parallel: global class counts, C3).            */
int b200ctr_bce_wll(const float* logits, const float* aux, const float* y, float aux_w, int64_t B,
                    const int32_t* counts_override, float* loss, int32_t* counts, float* dlogits, float* daux,
                    float* prob, void* stream);

/* ---- clip + AdamW + EMA  (train.py:193-199, utils/ema.py:91-131) -----------------------------------------
 * sqnorm: acc += sum x^2.  `acc` is an opaque 8-byte accumulator the caller zeroes (all-zero bits): every CTA adds its double
 *   partial as a 2^-32 fixed-point 64-bit INTEGER, so the total does not depend on the order the CTAs retire (bit-identical
 *   runs); only clip_coef reads it back.  Across ranks it is summed as int64.
 * clip_coef: coef[0] = min(1, max_norm / (sqrt(acc[0]) + 1e-6)) ; max_norm <= 0 => 1.
 * adamw_ema: torch.optim.AdamW step order (decay, lerp m, v, bias-corrected update) with g scaled by coef[0],
 *            then ema = decay*ema + (1-decay)*p when ema != NULL.                                             */
int b200ctr_sqnorm(const float* x, int64_t n, double* acc, void* stream);
int b200ctr_clip_coef(const double* acc, float max_norm, float* coef, float* total_norm, void* stream);
int b200ctr_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                      float lr, float wd, float beta1, float beta2, float eps, int64_t step,
                      const float* coef, float ema_decay, void* stream);
/* multi-tensor form for the unmodified reference loop (separate parameter tensors): recs = device array of n_t records of six
 * 64-bit words {p, g, m, v, ema, n} (ema may be 0); chunk c covers elements [chunk_off[c], chunk_off[c] + chunk) of tensor
 * chunk_tensor[c].  mode 0 = AdamW as above (+ EMA where a record has a shadow), mode 1 = EMA only.  The betas are doubles:
 * torch forms 1 - beta and the bias corrections from Python doubles, and float(0.999) is 1.3e-5 away in 1 - beta.           */
int b200ctr_multi_adamw_ema(const void* recs, const int32_t* chunk_tensor, const int64_t* chunk_off, int64_t n_chunks,
                            int64_t chunk, float lr, float wd, double beta1, double beta2, float eps, int64_t step,
                            const float* coef, float ema_decay, int mode, void* stream);
/* sparse rows of a big table: rows listed in tok[n_slots] (duplicates + pad_id allowed).  owner[vocab] scratch. */
int b200ctr_rows_claim(const int32_t* tok, int64_t n_slots, int pad_id, int32_t* owner, void* stream);
int b200ctr_rows_sqnorm(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner,
                        const float* g_dense, int D, double* acc, void* stream);
int b200ctr_rows_zero(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner,
                      float* g_dense, int D, void* stream);

/* ---- row-lazy, replay-exact AdamW + EMA for the seq_vocab x 64 tables (csrc/rowopt.cu) ----------------------------
 * state[row] = (last_step << 1) | has_moments.  History (device arrays indexed by step, entry 0 = identity):
 *   hist4[s] = {1 - lr_s*wd, lr_s/(1-b1^s), sqrt(1-b2^s), ema_decay_s}; C[s] = prod_{u<=s}(1-lr_u*wd);
 *   S[s] = d_s*S[s-1] + (1-d_s)*C[s]; logD[s] = sum_{u<=s} log d_u.
 * catchup: rows listed in tok (owner slots only) -> step target_step with zero gradient (what dense AdamW/EMA
 *          would have done to an untouched row); update: catch up to step-1, then the AdamW step with the
 *          accumulated gradient row scaled by coef[0], EMA, and the gradient row is re-zeroed; flush: all rows.  */
int b200ctr_rowopt_push(void* hist4, float* C, float* S, float* logD, int64_t capacity, int64_t t,
                        float decay_mul, float step_size, float bc2_sqrt, float ema_d, float C_t, float S_t, float logD_t,
                        void* stream);
/* Bitmap-driven row sets (one bit per GLOBAL row; `bitmaps` = n_maps maps of `words` 32-bit words, e.g. one per
 * data-parallel rank after an all-gather).  The *_bm kernels visit every row owned by `rank` (row % n_shards == rank,
 * local row = row / n_shards) whose bit is set in any map, each exactly once: catch-up before the forward reads it,
 * squared gradient norm, AdamW/EMA step t (+ re-zeroing of the gradient row).  n_shards in {1,2,4,8}.          */
int b200ctr_bitmap_mark(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab, uint32_t* bitmap,
                        void* stream);
int b200ctr_rowopt_catchup_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, int32_t* state,
                              float* p, float* m, float* v, float* ema, int D, int64_t target_step, const void* hist4,
                              const float* C, const float* S, const float* logD, float beta1, float beta2, float eps,
                              void* stream);
int b200ctr_rows_sqnorm_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, const float* g,
                           int D, double* acc, void* stream);
int b200ctr_rowopt_update_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, int32_t* state,
                             float* p, float* g, float* m, float* v, float* ema, int D, int64_t step, const float* coef,
                             const void* hist4, const float* C, const float* S, const float* logD, float beta1,
                             float beta2, float eps, void* stream);
/* Row-sharded tables: out[i] = local row (tok / n_shards) if tok is owned by `rank` (tok % n_shards == rank) and is
 * not pad_id, else -1; the owner-side row kernels are then called with pad_id = -1 on the local shard.           */
int b200ctr_rows_localize(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int n_shards, int rank,
                          int32_t* out, void* stream);
int b200ctr_rows_claim_any(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab, int32_t* owner,
                           void* stream);
int b200ctr_rowopt_catchup(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab, const int32_t* owner,
                           int32_t* state, float* p, float* m, float* v, float* ema, int D, int64_t target_step,
                           const void* hist4, const float* C, const float* S, const float* logD,
                           float beta1, float beta2, float eps, void* stream);
int b200ctr_rowopt_update(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner, int32_t* state,
                          float* p, float* g, float* m, float* v, float* ema, int D, int64_t step, const float* coef,
                          const void* hist4, const float* C, const float* S, const float* logD,
                          float beta1, float beta2, float eps, void* stream);
int b200ctr_rowopt_flush(int32_t* state, float* p, float* m, float* v, float* ema, int64_t n_rows, int D,
                         int64_t target_step, const void* hist4, const float* C, const float* S, const float* logD,
                         float beta1, float beta2, float eps, void* stream);

/* ---- deterministic scatter-add of the sequence-table gradients (csrc/scatter.cu) ----
 * Replaces the dense `embedding_dense_backward` + top-k `scatter` of autograd (src/models/dare.py:122-123,135-137 under
 * train.py:192) without atomics: the (token, slot) pairs are radix-sorted by token and every distinct row is reduced once,
 * in slot order, so two runs are bit-identical.  tok: n_src * n_slots ids.  DIRECT form (n_src = 1): slot s adds
 * dvals[s] * q[s / K] to g_att[tok_s] and rows_rep[s] to g_rep[tok_s].  STAGED form (row-sharded tables): the rows are read
 * from the owner's staging buffers, stage[(src * stage_cap + slot)], written by b200ctr_retrieve_bwd_staged(stage_all = 1) on
 * every rank.  Only tokens with (t & (n_shards-1)) == rank are reduced, into local row t >> log2(n_shards).  The target rows
 * must be zero on entry: each touched row is written, not accumulated.                                                  */
int64_t b200ctr_scatter_rows_workspace_bytes(int64_t n_total_slots, int64_t vocab);
int b200ctr_scatter_rows(const int32_t* tok, int n_src, int64_t n_slots, int pad_id, int64_t vocab, int K,
                         const float* q, const float* dvals, const float* rows_rep, const float* stage_att,
                         const float* stage_rep, int64_t stage_cap, int n_shards, int rank, float* g_att,
                         float* g_rep, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fold-ensemble  (infer.py:122-158, utils/metrics.py:68-71): p = sigmoid(mean_m logit(clamp(p_m))) ---- */
int b200ctr_logit_mean(const float* probs, int n_models, int64_t B, float eps, float* out, void* stream);
/* The whole per-batch tail of infer.py:108-158 for `ensemble.method: logit_mean` in one launch: for every fold m,
 * p_m = sigmoid(z_m / T_m) (T_m = 1 when the fold has no temperature calibrator: utils/calibration.py:17-21,103-110),
 * clipped to [eps, 1-eps] exactly where the reference clips (Calibrator.predict_proba and infer.py:122), then
 * out = sigmoid(mean_m logit(p_m)) (utils/metrics.py:30-32,68-71).  logits (n_models, B) row-major; inv_T may be NULL. */
int b200ctr_ensemble_logit_mean(const float* logits, const float* inv_T, int n_models, int64_t B, float eps, float* out,
                                void* stream);

/* ---- batched shard reader  (data/dataset.py:28-43,98-124): host row gather + device-side widening ---- */
/* dst row (dst_idx ? dst_idx[i] : i) <- src row src_idx[i] for i < n, rows of row_bytes bytes; n_threads host threads.  Host
 * memory only (memory-mapped .npy payload -> pinned staging buffer); makes no CUDA call.                              */
int b200ctr_gather_rows(const void* src, int64_t row_bytes, const int64_t* src_idx, const int64_t* dst_idx, int64_t n,
                        void* dst, int n_threads);
/* On-disk dtypes -> the collate contract's, on the device: kind 0 uint8 -> float32 (X_mask), 1 int8 -> float32 (y),
 * 2 int32 -> int64 (ids; the kernels also take int32 directly).                                                      */
int b200ctr_widen(int kind, const void* in, int64_t n, void* out, void* stream);

/* ---- on-GPU cache builder  (data/build_cache_v1.py:28-75 medians, 78-173 `_process_arrow_batch`) ----------
 * Arrow column buffers in, App. D shard arrays out.  Strings are (bytes, offsets[rows+1]) with int32 (offsets64 = 0) or int64
 * offsets; validity = Arrow validity bitmap at bit offset 0 (LSB first) or NULL.  The reference's categorical / group hash is
 * polars.Series.hash (unpinned, absent here), so this builder pins its own (cache format v2):
 *     h(bytes) = splitmix64_finalize(fnv1a64(bytes, basis 0xcbf29ce484222325 ^ seed)), nulls hash as "NA", integers as their
 *     decimal string; out[r * out_stride] = h % mod as int32 (out_int32) or int64.
 * b200ctr_cache_hash_bytes_host: the same hash of one byte string on the HOST (no GPU; returns the 64 bits as int64).
 * parse_seq: "12,345,6789" -> last max_len tokens right-aligned in seq_out (rows, max_len), pad_id on the left; n_bad
 *     (device uint64, caller-zeroed) counts bytes other than digits / ',' and tokens > INT32_MAX.
 * col_to_f32: one numeric column buffer (dtype 0 f32, 1 f64, 2 i32, 3 i64, 4 u8) -> float32 with nulls as NaN.
 * numeric: x = (Fn, ld) column-major float32, rows r0..r0+rows -> X_num (rows, Fn) float32 with NaN -> med[c], then
 *     nan_to_num(nan 0, +-inf -> +-1e6), and X_mask (rows, Fn) uint8 = isnan.
 * select_hist: one 8-bit digit pass of an exact radix select over order-preserving float keys: query q looks at column
 *     q_col[q], elements whose key equals q_prefix[q] above bit shift + 8; hist (n_q, 256) uint64, caller-zeroed.            */
int64_t b200ctr_cache_hash_bytes_host(const void* bytes, int64_t n, uint64_t seed);
int b200ctr_cache_hash_strings(const void* bytes, const void* offsets, int offsets64, const void* validity, int64_t rows,
                               uint64_t mod, uint64_t seed, void* out, int64_t out_stride, int out_int32, void* stream);
int b200ctr_cache_hash_int64(const void* vals, const void* validity, int64_t rows, uint64_t mod, uint64_t seed, void* out,
                             int64_t out_stride, int out_int32, void* stream);
int b200ctr_cache_parse_seq(const void* bytes, const void* offsets, int offsets64, const void* validity, int64_t rows,
                            int max_len, int pad_id, int32_t* seq_out, void* n_bad, void* stream);
int b200ctr_cache_col_to_f32(const void* col, int dtype, const void* validity, int64_t rows, float* out, void* stream);
int b200ctr_cache_numeric(const float* x, int64_t ld, int64_t r0, int64_t rows, int Fn, const float* med, float* X_num,
                          void* X_mask, void* stream);
int b200ctr_cache_select_hist(const float* x, int64_t ld, int64_t rows, const int* q_col, const uint32_t* q_prefix, int n_q,
                              int shift, void* hist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CTR_H */
