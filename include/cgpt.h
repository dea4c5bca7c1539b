/* cgpt_b200 — C ABI of the B200-native codon-GPT step (libcgpt_b200.so).
 *
 * Drop-in boundary for the hot path of AvishaiBarnoy/genomics-lm:
 *   src/codonlm/model_tiny_gpt.py   TinyGPT.forward (:297-352) and everything it calls
 *   src/codonlm/training/objectives.py  multi-offset / termination losses (:6-105)
 * The reference has no FFI of its own (it is pure PyTorch); each entry point below
 * names the reference lines whose arithmetic it replaces.  The Python side
 * (genomics-lm_b200/codonlm_b200) binds these with ctypes and mirrors the reference's module
 * contract; INTEGRATION.md shows the stub a maintainer adds to the reference tree.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (sm_100a, compute capability 10.0) unless it says "host";
 *   - tensors are dense row-major; `ld*` are row pitches in ELEMENTS;
 *   - bf16 tensors consumed through TMA need 16-byte aligned bases and pitches (pitch % 8 == 0);
 *   - `stream` is a cudaStream_t; no entry point synchronises, allocates or frees;
 *   - return value 0 = ok, <0 = error (cgpt_last_error() has the text, thread-local);
 *   - "accumulate" outputs are added to (+=), everything else is overwritten.
 */
#ifndef CGPT_B200_H
#define CGPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGPT_VERSION 100 /* 0.1.0 */

#define CGPT_OK 0
#define CGPT_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define CGPT_ERR_CUDA (-2)    /* CUDA runtime or driver error */
#define CGPT_ERR_DEVICE (-3)  /* not a compute-capability 10.0 device */

typedef void* cgpt_stream_t;

int cgpt_version(void);
/* 0 iff `device` is a cc 10.0 GPU and the TMA driver entry point resolves. */
int cgpt_device_ok(int device);
/* bind the calling thread to `device` inside this library's CUDA runtime (call once per rank). */
int cgpt_set_device(int device);
const char* cgpt_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t cgpt_launch_count(void);

/* ---------------------------------------------------------------- integer scans ---------- */
/* seg[b,t] = #{t' <= t : idx[b,t'] == sep_id}      model_tiny_gpt.py:290 (cumsum(idx==sep)). */
int cgpt_segment_ids(const int64_t* idx, int32_t* seg, int B, int T, int sep_id, cgpt_stream_t stream);
/* start[b,t] = max{t' <= t : idx[b,t'] == sep_id} or 0: first position of t's segment.  Segment ids are
 * non-decreasing, so  j<=i && seg[i]==seg[j]  <=>  start[i] <= j <= i  (what the attention kernels use). */
int cgpt_segment_starts(const int64_t* idx, int32_t* start, int B, int T, int sep_id, cgpt_stream_t stream);
/* next[b,t] = min{t' >= t : yb[b,t'] in ids} or T   objectives.py:78-86 (flip/cummin/flip). */
int cgpt_next_in_set(const int64_t* yb, int32_t* next, int B, int T, const int64_t* ids_host, int n_ids,
                     cgpt_stream_t stream);
/* labels = bucket(next_stop - t; edges), T-sentinel -> n_edges, PAD -> ignore_index
 *                                                   objectives.py:87-91. */
int cgpt_termination_labels(const int64_t* yb, const int32_t* next_stop, int64_t* labels, int B, int T,
                            const int64_t* edges_host, int n_edges, int64_t ignore_index,
                            cgpt_stream_t stream);

/* ---------------------------------------------------------------- embedding -------------- */
/* x[b,t,:] = tok_w[idx[b,t],:] (+ pos_w[t,:])       model_tiny_gpt.py:306-309 (fp32). */
int cgpt_embed_fwd(const int64_t* idx, const float* tok_w, const float* pos_w /*nullable*/, float* x, int B,
                   int T, int d, int vocab, cgpt_stream_t stream);
/* dtok_w[v,:] += sum_{idx==v} dx ; dpos_w[t,:] += sum_b dx   (autograd of the above). */
int cgpt_embed_bwd(const int64_t* idx, const float* dx, float* dtok_w, float* dpos_w /*nullable*/, int B,
                   int T, int d, int vocab, cgpt_stream_t stream);

/* ---------------------------------------------------------------- LayerNorm -------------- */
/* nn.LayerNorm(d), eps, affine                      model_tiny_gpt.py:137,139,216.
 * fp32 in; writes bf16 and/or fp32 normalised output; saves mean/rstd (fp32 [M]). */
int cgpt_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16 /*nullable*/,
                       float* y_f32 /*nullable*/, float* mean, float* rstd, int M, int d, float eps,
                       cgpt_stream_t stream);
/* dx = (dres) + LN'(dy); dgamma,dbeta accumulate.  dy is bf16 (dy_is_f32=0) or fp32.
 * `dx_bf16` (nullable) receives a bf16 copy of dx for the GEMMs upstream; `dx_colsum` (nullable, [d], accumulate)
 * receives sum_rows dx = the bias gradient of the residual linear that produced x. */
int cgpt_layernorm_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, const float* mean,
                       const float* rstd, const float* dres /*nullable*/, float* dx, void* dx_bf16 /*nullable*/,
                       float* dgamma, float* dbeta, float* dx_colsum /*nullable*/, int M, int d, cgpt_stream_t stream);

/* ---------------------------------------------------------------- dense GEMM (tcgen05) --- */
/* D[M,N] = A·Bᵀ with bf16 operands, fp32 accumulation in TMEM, fused epilogue.
 * Replaces every nn.Linear on the path (model_tiny_gpt.py:70-73,85-93,132,143-147,51-57,235-239)
 * and their autograd (dgrad / wgrad).
 *   a_mn_major=0: A is [M,K] (pitch lda);  =1: A is stored [K,M] (pitch lda)   ("Aᵀ given")
 *   b_mn_major=0: B is [N,K] (pitch ldb);  =1: B is stored [K,N] (pitch ldb)
 * forward  y = x·Wᵀ      : A=x[M,K]      B=W[N,K]            (0,0)
 * dgrad    dx = dy·W     : A=dy[M,N']    B=W[N',K'] as [K,N]  (0,1)
 * wgrad    dW = dyᵀ·x    : A=dy[Mtok,N'] as [K,M]  B=x[Mtok,K'] as [K,N]  (1,1), split_k>1
 */
#define CGPT_EPI_NONE 0
#define CGPT_EPI_GELU 1    /* out = gelu_erf(v); aux_out (nullable) = gelu_erf'(v)   (nn.GELU(), :145)   */
#define CGPT_EPI_MUL_AUX 2 /* out = v * aux          (backward through GELU: aux = gelu_erf'(pre))      */
typedef struct {
  const void* a;
  const void* b;
  int a_mn_major, b_mn_major;
  int64_t lda, ldb;
  int M, N, K;
  int split_k;           /* >= 1; > 1 requires out_f32 && accumulate */
  const float* bias;     /* [N] fp32, nullable */
  int epilogue;          /* CGPT_EPI_* */
  const void* aux;       /* bf16 [M,N] pitch ldaux (MUL_AUX) */
  void* aux_out;         /* bf16 [M,N] pitch ldaux (GELU, nullable) */
  int64_t ldaux;
  const float* residual; /* fp32 [M,N] pitch ldc, nullable: v += residual */
  void* out;             /* bf16 or fp32 [M,N] pitch ldc */
  int out_f32;
  int accumulate;        /* out_f32 only: atomically out += v */
  int64_t ldc;
  float* colsum;         /* nullable, fp32 [N], accumulate: column sums of the (bf16-rounded) output; only with the
                            MUL_AUX epilogue on TMA-aligned operands (bias gradient of the layer in front of a GELU) */
} cgpt_gemm_args;
int cgpt_gemm_bf16(const cgpt_gemm_args* args_host, cgpt_stream_t stream);

/* ---------------------------------------------------------------- small elementwise ------ */
int cgpt_cast_f32_bf16(const float* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int64_t cols,
                       cgpt_stream_t stream); /* pad columns [cols, ld_out) are zeroed */
/* fp32 [rows, cols] -> bf16 [rows, 3*cols_pad] = [hi | lo | hi] with hi = bf16(x), lo = bf16(x - hi)
 * (zero padding up to cols_pad).  Against a partner laid out [hi | hi | lo] one bf16 GEMM over the tripled
 * reduction dimension yields hi*hi + lo*hi + hi*lo, i.e. ~16 mantissa bits: the fp32-accurate LM head on
 * tensor cores (model_tiny_gpt.py:327,336).  partner=1 writes the [hi | hi | lo] form. */
int cgpt_split3_f32_bf16(const float* in, int64_t ld_in, void* out, int64_t rows, int64_t cols, int64_t cols_pad,
                         int partner, cgpt_stream_t stream);
/* dst[r,c] += S[r,c] + S[r,c+col_off] + S[r+row_off,c] + S[r+row_off,c+col_off]  (r < rows, c < cols).
 * Folds the four hi/lo cross products of the stacked split-head weight-gradient GEMM
 * S = [g_hi | g_lo]^T [x_hi | x_lo] into the fp32 gradient of the LM head (backward of model_tiny_gpt.py:327,336). */
int cgpt_fold_quadrants_add(const float* s, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int row_off,
                            int col_off, cgpt_stream_t stream);
/* out[n] += sum_m x[m,n]   (bias gradients) */
int cgpt_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, cgpt_stream_t stream);
/* RoPE, half-split pairing i <-> i+hd/2            model_tiny_gpt.py:35-45, applied in place to the
 * q and k column blocks of packed qkv [B*T, (H+2Hk)*hd]; cos/sin fp32 [T, hd/2]; inverse=1 is the
 * transpose (backward). */
int cgpt_rope_qk(void* qkv, const float* cos_t, const float* sin_t, int B, int T, int H, int Hk, int hd,
                 int inverse, cgpt_stream_t stream);
/* SwiGLU gate: act = silu(g) * u with gu = [g | u] (hidden h each, pitch ldgu); model_tiny_gpt.py:57. */
int cgpt_swiglu_fwd(const void* gu, int64_t ldgu, void* act, int64_t ldact, int M, int h, cgpt_stream_t stream);
int cgpt_swiglu_bwd(const void* gu, int64_t ldgu, const void* dact, int64_t ldact, void* dgu, int M, int h,
                    cgpt_stream_t stream);

/* ---------------------------------------------------------------- attention (tcgen05) ---- */
/* softmax(QKᵀ·scale + mask)·V over packed qkv [B*T, (H+2Hk)*hd] (q | k | v column blocks);
 * mask[i,j] = j<=i && (window<=0 || i-j<window) && (seg_start==NULL || j>=seg_start[b,i])
 *                                                   model_tiny_gpt.py:103-131 + :273-295;
 * seg_start int32 [B,T] from cgpt_segment_starts();
 * GQA: query head h reads kv head h / (H/Hk)       (:94-96, without the repeat_interleave copy).
 * out bf16 [B*T, H*hd]; lse fp32 [B,H,T] (natural-log logsumexp of the scaled scores). */
/* dropout_p > 0 (training, :104,129): probabilities are dropped (scaled by 1/(1-p)) before P·V with a
 * Philox4x32-10 mask keyed by (seed, offset, b, h, i, j); backward regenerates the same mask. */
int cgpt_attn_fwd(const void* qkv, const int32_t* seg_start /*nullable*/, void* out, float* lse, int B, int T, int H,
                  int Hk, int hd, int window, float scale, float dropout_p, uint64_t seed, uint64_t offset,
                  cgpt_stream_t stream);
/* dqkv bf16 [B*T,(H+2Hk)*hd]; `ws` fp32 workspace of cgpt_attn_bwd_workspace() bytes. */
int64_t cgpt_attn_bwd_workspace(int B, int T, int H, int Hk, int hd);
int cgpt_attn_bwd(const void* qkv, const int32_t* seg_start, const void* out, const void* dout, const float* lse,
                  void* dqkv, void* ws, int B, int T, int H, int Hk, int hd, int window, float scale,
                  float dropout_p, uint64_t seed, uint64_t offset, cgpt_stream_t stream);
/* Same, and dqkv_colsum[(H+2Hk)*hd] (fp32, nullable) += column sums of dqkv: the gradients of the query | key | value
 * biases (nn.Linear biases at :85-93), taken inside the kernels instead of by another pass over dqkv. */
int cgpt_attn_bwd_colsum(const void* qkv, const int32_t* seg_start, const void* out, const void* dout,
                         const float* lse, void* dqkv, void* ws, float* dqkv_colsum, int B, int T, int H, int Hk,
                         int hd, int window, float scale, float dropout_p, uint64_t seed, uint64_t offset,
                         cgpt_stream_t stream);
/* Introspection path (use_sdpa=False, :116-131): dense probabilities att fp32 [B,H,T,T]. */
int cgpt_attn_probs(const void* qkv, const int32_t* seg_start, float* att, int B, int T, int H, int Hk, int hd,
                    int window, float scale, float dropout_p, uint64_t seed, uint64_t offset, cgpt_stream_t stream);

/* Incremental decode (one new token per sequence against a K/V cache) — the cached form of the reference's sampling
 * loops, which re-run the whole context per generated token (generate.py:14-27, query_model.py:186-213).
 * qkv_new bf16 [B, (H+2Hk)*hd] (RoPE already applied); k_cache / v_cache bf16 [B, Tmax, Hk*hd] hold positions 0..t-1
 * and receive position t; lo[b] (nullable) = first visible position (segment start, model_tiny_gpt.py:283-295);
 * out bf16 [B, H*hd] = softmax(q·K[lo..t]^T·scale)·V[lo..t].  With t_dev (nullable, device int32) the position is read
 * on the device instead of from `t`, so that a captured CUDA graph of the decode step can be replayed for every t. */
int cgpt_attn_decode(const void* qkv_new, void* k_cache, void* v_cache, const int32_t* lo, void* out, int B, int t,
                     const int32_t* t_dev, int Tmax, int H, int Hk, int hd, int window, float scale,
                     cgpt_stream_t stream);

/* Shape guidance: out = x + nn.Linear(3, d)(shape_embeddings)   model_tiny_gpt.py:226-229, 310-311 (also :377-378).
 * x, out fp32 [M, d] (d % 4 == 0); s fp32 [M, 3]; w fp32 [d, 3]; b fp32 [d].
 * Backward: dw [d, 3] += dx^T s, db [d] += column sums of dx, ds [M, 3] = dx w (nullable: only when the shape encoder is
 * trained, loop.py:695); the gradient of x is dx itself. */
int cgpt_shape_proj_fwd(const float* x, const float* s, const float* w, const float* b, float* out, int64_t M, int d,
                        cgpt_stream_t stream);
int cgpt_shape_proj_bwd(const float* dx, const float* s, const float* w, float* dw, float* db, float* ds, int64_t M,
                        int d, cgpt_stream_t stream);

/* ---------------------------------------------------------------- token feed -------------- */
/* One micro-batch (xb, yb) int64 [B, T_out] gathered on the device from a RESIDENT packed dataset — the dynamic
 * format of the reference (flat token array + per-sequence lengths, data_loading.py:212-225) — for the sequence
 * indices of a batch: xb[r, c] = seq[c], yb[r, c] = seq[c+1] for c < len-1, PAD = 0 elsewhere; with
 * T_out = max(len) - 1 this is MmapPackedDataset.fetch_batch (data_loading.py:297-315) without the host gather and
 * without the H2D copy of the batch (only the B indices travel). */
int cgpt_pack_lm_batch(const int32_t* tokens, const int64_t* offsets, const int64_t* lengths, const int64_t* indices,
                       int B, int T_out, int64_t* xb, int64_t* yb, cgpt_stream_t stream);

/* ---------------------------------------------------------------- LM / aux heads --------- */
/* out[M,N] = x[M,d]·w[N,d]ᵀ (+bias), fp32 FMA, N <= 128
 *                                                   head :217,327; termination_head :220-224,330. */
int cgpt_skinny_linear_fwd(const float* x, const float* w, const float* bias /*nullable*/, float* out, int M,
                           int N, int d, cgpt_stream_t stream);
/* dx (+)= dout·w ; dw += doutᵀ·x ; dbias += colsum(dout). */
int cgpt_skinny_linear_bwd(const float* dout, const float* x, const float* w, float* dx, int dx_accumulate,
                           float* dw, float* dbias /*nullable*/, int M, int N, int d, cgpt_stream_t stream);
/* Cross entropy, mean over kept rows               model_tiny_gpt.py:343-349; objectives.py:39-57,100-105.
 * Row r=(b,t) uses target tgt[b, t+shift]; it is kept iff t+shift < T, target != ignore_index and
 * (next_boundary==NULL || next_boundary[b,t] >= t+shift)   (offset_target_mask, objectives.py:13-23).
 * sums[0] = sum of weighted losses, sums[1] = sum of w[target] (fixed-order, deterministic: per-CTA partials, added up
 * in index order by the last CTA to finish); *mean_out (nullable) = sums[0] / sums[1] — the mean loss — or 0 when no
 * row is kept and zero_if_empty != 0 (objectives.py:100-105).  All three are WRITTEN, not accumulated.
 * row_lse[M] is saved for backward; row_ws is a [2*M] fp32 scratch (the per-CTA partial loss | weight sums). */
int cgpt_ce_fwd(const float* logits, const int64_t* targets, const int32_t* next_boundary /*nullable*/,
                const float* class_w /*nullable*/, float* sums, float* mean_out, float* row_lse, float* row_ws,
                int B, int T, int V, int shift, float smoothing, int64_t ignore_index, int zero_if_empty,
                cgpt_stream_t stream);
/* dlogits = coef * (*gscale) / sums[1] * dL_row/dlogits   (zero rows for dropped targets).
 * Optional bf16 by-product for the GEMMs that consume the gradient (nullable / bf16_mode 0 = none):
 * bf16_mode 1: dl_bf16[M, ld_bf16] = bf16(dlogits), pad columns zero; bf16_mode 2: dl_bf16[M, 3*ld_bf16] =
 * hi | lo | hi with hi = bf16(g), lo = bf16(g - hi) (the split operand of the fp32-accurate head). */
int cgpt_ce_bwd(const float* logits, const float* row_lse, const int64_t* targets,
                const int32_t* next_boundary, const float* class_w, const float* sums,
                const float* gscale /*device scalar, nullable = 1*/, float coef, float* dlogits,
                void* dl_bf16 /*nullable*/, int bf16_mode, int64_t ld_bf16, int B, int T, int V, int shift,
                float smoothing, int64_t ignore_index, cgpt_stream_t stream);

/* ---------------------------------------------------------------- dropout ---------------- */
/* out = (residual) + x * mask / (1-p)   nn.Dropout on the embedding (:312) and the MLP output (:57,147).
 * x, residual fp32; out fp32 or bf16; n % 4 == 0.  The mask is a pure function of (seed, offset, index):
 * backward calls this again on the gradient with the same (seed, offset). */
int cgpt_dropout(const float* x, const float* residual /*nullable*/, void* out, int out_bf16, int64_t n, float p,
                 uint64_t seed, uint64_t offset, cgpt_stream_t stream);

/* Device-resident dropout generator state (reference: torch's CUDA generator drives nn.Dropout / SDPA dropout_p,
 * model_tiny_gpt.py:104,147,312; tests/test_attention_dropout.py:62-78).  dev_state = {seed, base offset} (2 x u64 in
 * device memory) or NULL.  While set, every kernel that takes (seed, offset) reads the seed from dev_state[0] and adds
 * dev_state[1] to its by-value offset ON THE DEVICE, so a captured CUDA graph of a training step draws fresh masks on
 * every replay; cgpt_philox_advance adds `increment` to the base offset (one launch, part of the graph).
 * Process-wide (backward kernels are launched from autograd's thread); NULL restores by-value seeds. */
int cgpt_set_philox_state(const uint64_t* dev_state /*nullable*/);
int cgpt_philox_advance(uint64_t* dev_state, uint64_t increment, cgpt_stream_t stream);

/* ---------------------------------------------------------------- optimiser -------------- */
/* torch.optim.AdamW step on a flat fp32 buffer (loop.py:681-731 param groups are separate calls);
 * g is multiplied by grad_scale first (grad-accum / DDP mean, loop.py:145-150); optional bf16 shadow. */
/* dev_hyper (nullable): device array [lr, 1-beta1^step, sqrt(1-beta2^step)] that overrides lr/step, so a
 * captured CUDA graph of the whole step can be replayed while the host advances the schedule. */
int cgpt_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16 /*nullable*/, int64_t n,
               float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
               const float* dev_hyper, cgpt_stream_t stream);
/* The same step with the gradient read as bf16: the data-parallel trainer all-reduces bf16 gradient buckets (SURVEY
 * §8e) and the optimiser consumes the reduced buckets in place — no fp32 copy-back pass over the gradients. */
int cgpt_adamw_bf16grad(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16 /*nullable*/, int64_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        const float* dev_hyper, cgpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CGPT_B200_H */
