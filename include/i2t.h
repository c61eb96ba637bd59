/*
 * libi2t -- C ABI of the B200 (sm_100a) encoder-decoder hot path of image2text.
 *
 * The reference (iitmdinesh/image2text) is pure Python: it has no FFI / plugin interface, every
 * hot operation is a PyTorch library call.  Each entry point below therefore names the reference
 * CALL SITE (file:line, relative to the reference root) whose arithmetic it replaces; the Python
 * host in image2text_b200/ binds them with ctypes (see INTEGRATION.md for the stub a maintainer of
 * the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer (outputs, workspaces) is allocated by the caller,
 *     the library never frees or retains a pointer past the call;
 *   - all pointers are DEVICE pointers unless a parameter says "host";
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *     never synchronises;
 *   - return value: 0 on success, negative I2T_ERR_* otherwise; i2t_last_error() returns a
 *     thread-local message.  Nothing throws across the ABI;
 *   - dtype codes: I2T_F32 = 0, I2T_BF16 = 1.  Accumulation is always fp32;
 *   - row-major tensors, innermost dimension contiguous unless a stride argument says otherwise.
 */
#ifndef I2T_H_
#define I2T_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2T_F32 0
#define I2T_BF16 1

#define I2T_OK 0
#define I2T_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define I2T_ERR_CUDA (-2)      /* a CUDA runtime or driver call failed */

/* activation codes for the GEMM epilogues */
#define I2T_ACT_NONE 0
#define I2T_ACT_GELU_TANH 1    /* models/layers.py:477 nn.GELU('tanh'); HF gelu_new */
#define I2T_ACT_GELU_ERF 2     /* torchvision MLPBlock nn.GELU() */

/* attention mask modes (closed forms of SURVEY.md Appendix C after discrepancy D9) */
#define I2T_MASK_NONE 0        /* ViT self attention, cross attention */
#define I2T_MASK_CAUSAL 1      /* key j visible to query i  <=>  j <= i */
#define I2T_MASK_PROMPT 2      /* i < n_prompt: j <= i ;  i >= n_prompt: n_prompt <= j <= i */

int i2t_version(void);
const char* i2t_last_error(void);
/* number of kernel launches issued by this library in the calling process since load */
int64_t i2t_launch_count(void);

/* ---- LayerNorm: models/layers.py:357-358 (eps 1e-5), torchvision vision_transformer.py:96 (1e-6) --- */
/* y[r,:] = (x[r,:] - mean) * rstd * gamma + beta.  x rows are x_row_stride elements apart (lets the ViT
 * final norm read only token 0 of every image).  mean / rstd (fp32, rows) are optional (backward). */
int i2t_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int64_t rows, int64_t cols, int64_t x_row_stride, float eps, int x_dtype, int y_dtype,
                      void* stream);
/* dx, and dgamma/dbeta ACCUMULATED (+=) into fp32 buffers (may be NULL).  models/layers.py:357-358 backward. */
int i2t_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, int dy_dtype, int x_dtype,
                      int dx_dtype, void* stream);
/* The same with dx = (LayerNorm backward) + dx_add: the pre-LN block's residual join (x feeds both the LayerNorm branch and the
 * skip connection, models/layers.py:597-606), one pass instead of a separate add.  dx_add has dx's dtype and shape. */
int i2t_layernorm_bwd_add(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          const void* dx_add, void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, int dy_dtype,
                          int x_dtype, int dx_dtype, void* stream);

/* ---- dense contraction: every nn.Linear / Conv1D / conv_proj on the path
 *      (models/layers.py:452,469,482,484; models/decoder.py:256; nn.MultiheadAttention in/out proj
 *      models/layers.py:600-605; torchvision vision_transformer.py:277) -------------------------------- */
/* C[M,N] = act(op(A) * op(B) + bias[N]) + residual[M,N]
 *   a_kmajor=1: A is [M][K] (K contiguous, leading dim lda);  0: A is [K][M] (M contiguous)
 *   b_kmajor=1: B is [N][K] (K contiguous, leading dim ldb) -- nn.Linear weight;  0: B is [K][N] -- HF Conv1D
 *   accumulate=1: C += result (fp32 C only; used by weight gradients)
 *   colsum (fp32, N, optional): colsum[n] += sum_m C_new[m,n] is NOT provided here; see i2t_colsum.
 * fp32 operands run exact fp32 FMA tiles (the parity anchor); bf16 operands run tcgen05 tiles. */
int i2t_gemm(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
             int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
             int ab_dtype, int res_dtype, int c_dtype, void* stream);
/* Same with flags.  I2T_GEMM_B_STABLE: no preceding work in the stream writes B (decode-time weights): the kernel, launched
 * with programmatic stream serialization, requests its first B tiles before waiting for the previous kernel to finish. */
#define I2T_GEMM_B_STABLE 1
int i2t_gemm_ex(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
                int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
                int ab_dtype, int res_dtype, int c_dtype, int flags, void* stream);
/* out[n] += sum_m X[m,n]   (bias gradients) */
int i2t_colsum(const void* X, float* out, int64_t M, int64_t N, int64_t ldx, int x_dtype, void* stream);

/* ---- fused attention: models/layers.py:465 F.scaled_dot_product_attention, torchvision :113 ---------- */
/* q,k,v element (b,h,t,e) lives at ptr + b*batch_stride + t*row_stride + h*head_dim + e (so the packed
 * (B,T,3C) output of c_attn / in_proj is read in place); out is (B,Tq,H*head_dim).
 * lse (fp32, B*H*Tq, optional) receives the row log-sum-exp for the backward pass. */
int i2t_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                 int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                 int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int in_dtype,
                 int out_dtype, void* stream);
/* Backward of the above (autograd of reference models/layers.py:465).  dq,dk,dv are written with the same strides
 * as q,k,v; out/dout are (B,Tq,H*head_dim).  workspace: i2t_attn_bwd_workspace_bytes(...) bytes of device memory. */
int64_t i2t_attn_bwd_workspace_bytes(int64_t B, int64_t H, int64_t Tq, int64_t head_dim);
/* Debugging aid for the tcgen05 backward (sequences <= 256 rows): with I2T_ATTN_BWD_DEBUG=-1 in the environment CTA (0,0) of every
 * launch stamps its timeline in SM clock ticks; copies the 32 stamps of the latest launch to host_out (synchronises the device). */
int i2t_attn_bwd_trace(long long* host_out);
int i2t_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                 void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                 int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride, int64_t kv_batch_stride,
                 int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int dtype, void* stream);

/* ---- ViT patch embedding: torchvision vision_transformer.py:268-296 ---------------------------------- */
/* images (B,3,H,W) fp32 NCHW -> patches (B*nh*nw, 3*p*p), column = c*p*p + ky*p + kx (conv weight order) */
int i2t_patch_im2col(const float* images, void* patches, int64_t B, int64_t H, int64_t W, int64_t p, int out_dtype,
                     void* stream);
/* x[b,0,:] = cls + pos[0];  x[b,1+i,:] = patch_out[b*np+i,:] + pos[1+i]   (fp32 residual stream) */
int i2t_vit_assemble(const void* patch_out, const float* cls, const float* pos, float* x, int64_t B, int64_t np,
                     int64_t C, int patch_dtype, void* stream);

/* ---- LSH tail: models/layers.py:139-144,211-219, models/encoder.py:116-117 ---------------------------- */
/* feat (B,D) fp32 -> out (B,n_cls,E).  tables: device arrays of n_cls*n_res pointers (slot-major):
 * proj[i] -> (D,n_proj) fp32, grid[i] -> (num_bins[r]) fp32, emb[i] -> ((num_bins[r]+1)*n_proj, E) fp32.
 * bucket_out (int32, B*n_cls*n_res*n_proj, optional) receives the EmbeddingBag row indices. */
int i2t_lsh_tail(const float* feat, const void* const* proj, const void* const* grid, const void* const* emb,
                 const int32_t* num_bins, float* out, int32_t* bucket_out, int64_t B, int64_t D, int64_t n_cls,
                 int64_t n_res, int64_t n_proj, int64_t E, void* stream);

/* Backward of the tail's EmbeddingBag(mode="mean") lookups (models/layers.py:139-144): demb_host is a HOST array of
 * n_cls*n_res device pointers (slot-major) to fp32 gradient tables of the emb[i] shapes; the kernel ADDS
 * dout[b,s,:] / n_proj into the rows `bucket` (as written by i2t_lsh_tail) names.  The hashing has no gradient. */
int i2t_lsh_tail_bwd(const float* dout, const int32_t* bucket, void* const* demb_host, int64_t B, int64_t n_cls,
                     int64_t n_res, int64_t n_proj, int64_t E, void* stream);

/* ---- PEER tail core: models/layers.py:73-109 minus its dense projections (i2t_gemm) ---------------------------- */
/* ql / qr (M,H,U) fp32: scores of the left / right query units (query . W_left/right); key (M,H,D): key projection;
 * emb_in (E,D), emb_out (E,O).  Per (row, head): top-K of ql and qr, top-K of their K x K sums, softmax, expert id =
 * left * K + right (:93-96, literally), w_k = softmax_k * gelu_tanh(emb_in[id_k] . key); out (M,O) = sum_{h,k} w_k emb_out[id_k]
 * (the residual projection is added by the caller).  s_* (M,H,K) are saved for the backward pass.  M = batch * n_cls. */
int i2t_peer_lookup_fwd(const float* ql, const float* qr, const float* key, const float* emb_in, const float* emb_out, float* out,
                        int32_t* s_idx, float* s_score, float* s_dot, int32_t* s_lpos, int32_t* s_rpos, int64_t M, int64_t H,
                        int64_t U, int64_t K, int64_t D, int64_t O, void* stream);
/* Backward: dql, dqr (M,H,U) and dkey (M,H,D) are written; demb_in (E,D) / demb_out (E,O) are ACCUMULATED (zero them first;
 * dense like the reference's nn.Embedding gradients). */
int i2t_peer_lookup_bwd(const float* dout, const float* key, const float* emb_in, const float* emb_out, const int32_t* s_idx,
                        const float* s_score, const float* s_dot, const int32_t* s_lpos, const int32_t* s_rpos, float* dql,
                        float* dqr, float* dkey, float* demb_in, float* demb_out, int64_t M, int64_t H, int64_t U, int64_t K,
                        int64_t D, int64_t O, void* stream);

/* ---- decoder input: models/vision_encoder_decoder.py:84-88 + models/decoder.py:234-243 ----------------- */
/* x[b,t,:] = (t < n_prompt ? prompt[b,t,:] : wte[ids[b,t-n_prompt],:]) + wpe[t,:],  t < T */
int i2t_embed_fwd(const int64_t* ids, const float* prompt, const float* wte, const float* wpe, float* x, int64_t B,
                  int64_t T, int64_t n_prompt, int64_t S, int64_t C, void* stream);

/* ---- cross attention with a short key/value set: models/layers.py:537-542,600-605 --------------------- */
/* q (B,T,C) [row stride q_row_stride]; k,v element (b,h,s,e) at ptr + b*kv_batch_stride + s*kv_row_stride
 * + h*head_dim + e; S <= 64; no mask; out (B,T,C). */
int i2t_xattn_fwd(const void* q, const void* k, const void* v, void* out, int64_t B, int64_t H, int64_t T, int64_t S,
                  int64_t head_dim, int64_t q_row_stride, int64_t kv_batch_stride, int64_t kv_row_stride,
                  int in_dtype, int out_dtype, void* stream);

/* Backward of i2t_xattn_fwd.  dq has q's layout; dk/dv are fp32 buffers the caller zeroed, element (b,h,s,e) at
 * ptr + b*dkv_batch_stride + s*dkv_row_stride + h*head_dim + e, accumulated with atomicAdd. */
int i2t_xattn_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, float* dk, float* dv,
                  int64_t B, int64_t H, int64_t T, int64_t S, int64_t head_dim, int64_t q_row_stride,
                  int64_t kv_batch_stride, int64_t kv_row_stride, int64_t dkv_batch_stride, int64_t dkv_row_stride,
                  int dtype, void* stream);

/* 1 (default): bf16 contractions run on tcgen05 when the shape fits; 0: always the FMA kernel (A/B testing). */
void i2t_set_tensor_core_gemm(int enabled);
/* 1 (default): large bf16 GEMMs use the CTA-pair kernel (tcgen05 cta_group::2, 256-row tiles); 0: 128x128 tiles only. */
void i2t_set_gemm_cta_pair(int enabled);
/* 1 (default): the CTA-pair kernel's output tile leaves through shared memory + cp.async.bulk.tensor stores; 0: row stores. */
void i2t_set_gemm_tma_store(int enabled);
/* 1 (default): bf16 GEMMs with few output tiles and fp32 output (decode projections over a batch, weight gradients) split K
 * over the idle SMs and add the partial tiles with fp32 atomics (summation order not fixed); 0: never. */
void i2t_set_gemm_split_k(int enabled);
/* bf16 attention: 1 (default) tensor cores -- tcgen05/TMEM forward when head_dim == 64 and <= 384 keys, mma.sync otherwise
 * and for the backward; 2: mma.sync kernels only; 0: the fp32-math kernels (A/B testing). */
void i2t_set_tensor_core_attention(int mode);

/* ---- KV-cached decode step (no counterpart in the reference, which recomputes the prefix every token:
 *      models/vision_encoder_decoder.py:144-150).  pos_ptr is a DEVICE int32: the index of the token being processed
 *      (also its KV-cache slot); every kernel reads it on the device so one CUDA graph replays all steps. ---------- */
/* 1 (default): decode kernels are launched with programmatic dependent launch so that a kernel's weight copies overlap
 * the tail of its predecessor; 0: plain stream order. */
void i2t_set_pdl(int enabled);
/* x[b,:] = wte[ids[b*ids_ld + pos]] + wpe[n_prompt + pos]            (models/decoder.py:234-243) */
int i2t_dec_embed(const int64_t* ids, const float* wte, const float* wpe, float* x, const int32_t* pos_ptr, int64_t B,
                  int64_t C, int64_t ids_ld, int64_t n_prompt, void* stream);
int i2t_dec_advance(int32_t* pos_ptr, void* stream);
/* x[b,:] = rows[b, *pos_ptr, :] + wpe[*pos_ptr, :]: a soft-prompt row as the input of decode step *pos_ptr (HF decoders
 * treat the prompt rows as ordinary causal positions, models/decoder.py:343-360) */
int i2t_dec_embed_rows(const float* rows, int64_t batch_stride, const float* wpe, float* x, const int32_t* pos_ptr, int64_t B,
                       int64_t C, void* stream);
/* large-batch decode (projections as GEMMs): append columns [C,2C) / [2C,3C) of the packed fp32 (B,ld) qkv rows to row
 * *pos_ptr of the (B,Tmax,C) K / V caches (what i2t_dec_linear's qkv_split epilogue does for B <= 16) */
int i2t_dec_kv_append(const float* qkv, int64_t ld, void* kcache, void* vcache, int64_t cache_batch_stride, int64_t C,
                      int64_t B, int cache_dtype, const int32_t* pos_ptr, void* stream);
/* out[b,n] = act(LN?(x)[b,:] . W[n,:] + bias[n]) (+ residual[b,n]);  B <= 16, W is (N,K) fp32 or bf16.
 * qkv_split=1 (N == 3C): columns [0,C) -> out (B,ldo), [C,2C) -> kcache, [2C,3C) -> vcache at row *pos_ptr of the
 * (B,Tmax,C) caches (fused KV-cache append).  Replaces c_attn/c_proj/c_fc/lm_head at models/layers.py:452,469,482,484
 * and models/decoder.py:256 for one new token per sequence. */
int i2t_dec_linear(const float* x, const float* ln_gamma, const float* ln_beta, float ln_eps, const void* W,
                   const float* bias, const float* residual, float* out, int64_t ldo, int64_t B, int64_t N, int64_t K,
                   int act, int w_dtype, int qkv_split, void* kcache, void* vcache, int64_t cache_batch_stride, int64_t C,
                   int cache_dtype, const int32_t* pos_ptr, void* stream);
/* one query per (b,h) against a (B,Tmax,C)-layout cache; visible keys = [0, *len_ptr + len_add) (len_ptr may be NULL) */
int i2t_dec_attn(const float* q, int64_t q_ld, const void* kcache, const void* vcache, int64_t cache_batch_stride,
                 int64_t cache_row_stride, float* out, int64_t out_ld, const int32_t* len_ptr, int64_t len_add, int64_t B,
                 int64_t H, int64_t head_dim, int cache_dtype, void* stream);
/* Same, with two fusions for the batched (GEMM-mode) decode step: (1) knew / vnew (optional, fp32, element (b,h,e) at
 * ptr + b*new_ld + h*head_dim + e): the K / V row of the token being decoded, i.e. key len-1 -- rounded to the cache dtype, used
 * from registers AND stored into the cache by this kernel (replaces i2t_dec_kv_append); (2) out_dtype = I2T_BF16 writes the
 * result as bf16, ready to be the next GEMM's A operand (replaces a cast kernel). */
int i2t_dec_attn_append(const float* q, int64_t q_ld, void* kcache, void* vcache, int64_t cache_batch_stride,
                        int64_t cache_row_stride, void* out, int64_t out_ld, const int32_t* len_ptr, int64_t len_add, int64_t B,
                        int64_t H, int64_t head_dim, int cache_dtype, const float* knew, const float* vnew, int64_t new_ld,
                        int out_dtype, void* stream);
/* LayerNorm of the (rows, cols) fp32 residual stream for the batched decode step (reference models/layers.py:357-358 on one
 * token per sequence), launched with programmatic stream serialization: gamma / beta are fetched before the dependency wait.
 * zero_ptr (optional, fp32, zero_count elements, multiple of 4): zero-filled after the wait -- the output buffer of the split-K
 * projection that consumes this LayerNorm (call i2t_gemm on it with accumulate = 1). */
int i2t_dec_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int64_t cols, float eps,
                      int y_dtype, float* zero_ptr, int64_t zero_count, void* stream);
/* h = act(z), fp32 in, h_dtype out (nn.GELU(approximate='tanh'), reference models/layers.py:477,483, applied to the one-token
 * MLP pre-activation); PDL-aware variant of i2t_act_fwd for the batched decode step. */
int i2t_dec_act(const float* z, void* h, int64_t n, int act, int h_dtype, void* stream);

/* One whole decode step (every layer, LM head, sampler) as ONE cooperative launch: the same arithmetic as the
 * i2t_dec_* sequence, with the weights streamed through a shared-memory ring by a producer warp per CTA and ~1 us grid
 * barriers between the ~80 dependent stages (decode_mega.cu).  lin / att / sched are device tables of plain numbers:
 *   lin[op][20]  = {W, bias, ln_gamma, ln_beta, in, out, residual, N, K, act, mode(0 plain | 1 qkv split + KV append),
 *                   kcache, vcache, in_mode(0 buffer | 1 token embedding, in = wte), wpe, ldo, cache_batch_stride, 0,0,0}
 *   att[a][8]    = {k, v, batch_stride, row_stride, len_mode(0: *pos+1 | 1: const), len_const, 0, 0}
 *   sched[s][4]  = {kind(0 linear | 1 attention | 2 sample | 3 advance position), index, 0, 0}
 * bar: device uint32 scratch; error_flag: device int32, non-zero if an internal wait timed out (results invalid).
 * max_k: largest K of any linear (<= 3072).  B <= 8.  trace (optional, device int64[n_sched*4]): clock64 stamps of
 * CTA 0 per stage (begin, activations staged, computed, barrier passed) for profiling. */
int i2t_decode_mega(const int64_t* lin, const int64_t* att, const int32_t* sched, int64_t n_sched, int64_t n_ops, int64_t B,
                    int64_t C, int64_t H, int64_t V, int64_t n_prompt, int w_dtype, int64_t* ids, int64_t ids_ld,
                    int32_t* pos, float* q, float* y, float* logits, uint32_t* bar, int32_t* error_flag, float temperature,
                    int64_t top_k, const int32_t* ngrams, int64_t n_ngrams, const uint64_t* seed_ptr, int32_t* ticket,
                    int64_t max_k, int64_t* trace, void* stream);

/* Whole generate() loop as ONE cooperative launch, bf16 weights (decode_mega2.cu): n_prefill steps of sched_prefill (no LM
 * head; the prompt tokens beyond the first) then n_sample steps of sched_sample, starting at the device position *pos.
 * Same tables as i2t_decode_mega; lin[op][17] bit 0 marks the LM head: with top_k == 1 its epilogue applies the
 * no-repeat-n-gram ban and reduces the arg-max into `keys` (device uint64[24], scratch) instead of writing logits, and
 * the next step's embedding reads the pick from there.  Weights travel HBM -> registers as mma.sync A fragments,
 * prefetched before each grid barrier.  max_len: largest number of cached positions any step can see
 * (<= i2t_decode_mega2_max_keys()); max_k: widest linear input (<= 3072); n_embd <= 768.
 * Replaces models/vision_encoder_decoder.py:144-180 (the per-token loop, which re-runs the whole prefix). */
int i2t_decode_mega2_max_keys(void);
int i2t_decode_mega2(const int64_t* lin, const int64_t* att, const int32_t* sched_sample, int64_t n_sched_sample,
                     const int32_t* sched_prefill, int64_t n_sched_prefill, int64_t n_ops, int64_t n_att, int64_t n_prefill,
                     int64_t n_sample, int64_t B, int64_t C, int64_t H, int64_t V, int64_t n_prompt, int64_t* ids,
                     int64_t ids_ld, int32_t* pos, float* q, float* y, float* logits, uint32_t* bar, int32_t* error_flag,
                     uint64_t* keys, float temperature, int64_t top_k, const int32_t* ngrams, int64_t n_ngrams,
                     const uint64_t* seed_ptr, int64_t max_k, int64_t max_len, int64_t* trace, void* stream);

/* Whole generate() loop as ONE cooperative launch WITHOUT grid barriers, bf16 weights (decode_mega3.cu; the default for
 * up to 8 sequences).  Activations cross CTAs through exchange buffers the caller pre-fills with 0xFF bytes (NaN poison):
 * a producer stores its result, a consumer polls the words it needs until none is poison (3 generations per buffer,
 * `gen_stride` bytes apart; the writer of step t re-poisons generation t + 1).  Weights are read from `wpack`: one
 * contiguous stream per CTA (start offsets cta_base[grid]) in consumption order and mma fragment layout, written by
 * i2t_decode_mega3_pack (one call per linear op: W = bf16 [N][K], rows ldw elements apart, tile_off[tile] = byte offset of the tile's
 * chunks; a tile = 16 rows, i2t_decode_mega3_tile_bytes(K) bytes; tile u of an op belongs to CTA (u + rot) % grid with
 * grid = i2t_decode_mega3_grid(), a CTA's tiles in ascending order, ops in schedule order).  A producer warp streams them
 * through a shared-memory ring with cp.async.bulk, several stages ahead of the arithmetic.  tc (and tc_layout of the pack
 * call) = 1: the linear stages run on tcgen05 -- a tile's weights are packed as K-major 128B-swizzled UMMA atoms, one thread
 * issues tcgen05.mma M128 N16 K16 over the tile's whole K, accumulators live in TMEM, tcgen05.commit frees the ring slot;
 * 0: mma.sync tiles with K split over 8 warps.
 * Tables: lin[op][24], att[a][12], cmb[c][8] (combine stages of K-split projections), sched[s][4] as documented at the top
 * of decode_mega3.cu.  The caller also fills the
 * cache rows [*pos, *pos + steps) of every layer with 0xFF bytes, ids beyond the prompt with -1 and ctakeys
 * (uint64[3 * grid * 8]) with 0.  error_flag: 2 = a wait timed out, 3 = ring wait timed out, 5 = too many banned tokens.
 * Replaces models/vision_encoder_decoder.py:144-180 (the per-token loop, which re-runs the whole prefix). */
int i2t_decode_mega3_max_keys(void);
/* back-off (ns) between unsuccessful polls of an exchange buffer; 0 (default) = spin */
void i2t_set_decode_poll_sleep(int ns);
int i2t_decode_mega3_grid(void);
int64_t i2t_decode_mega3_tile_bytes(int64_t K);
int i2t_decode_mega3_pack(const void* W, int64_t N, int64_t K, int64_t ldw, void* dst, const int64_t* tile_off, int64_t tc_layout,
                          void* stream);
/* one launch that poisons / zeroes everything i2t_decode_mega3 expects: exch (0xFF), cache rows [pos0, pos0 + steps) of
 * n_rows = layers x sequences rows (row_pitch bytes each, pos_bytes per position), ids[:, P:ids_cols] = -1, ctakeys, *err */
int i2t_decode_mega3_prepare(void* exch, int64_t exch_bytes, void* kcache, void* vcache, int64_t n_rows, int64_t row_pitch,
                             int64_t pos_bytes, int64_t pos0, int64_t steps, int64_t* ids, int64_t ids_ld, int64_t B, int64_t P,
                             int64_t ids_cols, uint64_t* ctakeys, int64_t n_keys, int32_t* err, void* stream);
int i2t_decode_mega3(const int64_t* lin, const int64_t* att, const int64_t* cmb, const int32_t* sched, int64_t n_sched,
                     int64_t n_ops, int64_t n_att, int64_t n_cmb, int64_t n_prefill, int64_t n_sample, int64_t B, int64_t C, int64_t H, int64_t V,
                     int64_t n_prompt, int64_t* ids, int64_t ids_ld, int32_t* pos, float* logits, int64_t ldl,
                     uint32_t* bar, int32_t* error_flag, uint64_t* ctakeys, const void* wpack, const int64_t* cta_base,
                     int64_t gen_stride, float temperature, int64_t top_k, const int32_t* ngrams, int64_t n_ngrams,
                     const uint64_t* seed_ptr, int64_t max_k, int64_t max_len, int64_t* trace, int64_t trace_cta,
                     int64_t tc, void* stream);

/* ---- sampler: models/vision_encoder_decoder.py:152-180 + transformers NoRepeatNGramLogitsProcessor ----------------
 * logits (B,ldl) fp32 are modified in place (/temperature, banned -> -inf).  Tokens ids[b, 0..cur_len) are the history
 * (cur_len = *pos_ptr + 1 when pos_ptr != NULL, else the cur_len argument); the draw is written to ids[b, cur_len] when
 * write_token != 0.  top_k <= 0 means no top-k filter.  probs_out (B,V fp32, optional) receives the sampling
 * distribution.  nucleus_p in (0,1): top-p filter of vision_encoder_decoder.py:160-172 after the top-k filter (sorted
 * descending, keep while the cumulative probability <= max(p, p_max)); 0 or 1 = off.  seed_ptr (device uint64, optional) overrides seed so a captured graph can be re-seeded.
 * advance_pos != 0: the last CTA increments *pos_ptr (ticket = zeroed device int32). */
int i2t_sample(float* logits, int64_t ldl, int64_t B, int64_t V, int64_t* ids, int64_t ids_ld, int32_t* pos_ptr,
               int advance_pos, int64_t cur_len, float temperature, int64_t top_k, float nucleus_p, const int32_t* ngrams,
               int64_t n_ngrams, uint64_t seed, const uint64_t* seed_ptr, float* probs_out, int32_t* ticket, int write_token,
               void* stream);
/* 1 (default): top_k = 1 without nucleus filter / probs_out takes the one-pass arg-max kernel (ties -> lowest id; the stored
 * logits are banned in place but not divided by the temperature); 0: always the general sampler (A/B testing). */
void i2t_set_sampler_greedy_fast_path(int enabled);

/* ---- training-side memory-bound kernels ------------------------------------------------------------------------- */
/* h = act(z) / dz = dh * act'(z): nn.GELU('tanh') models/layers.py:477,483; torchvision nn.GELU(); HF gelu_new.
 * n (elements) must be a multiple of 4. */
int i2t_act_fwd(const void* z, void* h, int64_t n, int act, int z_dtype, int h_dtype, void* stream);
int i2t_act_bwd(const void* z, const void* dh, void* dz, int64_t n, int act, int z_dtype, int g_dtype, void* stream);
/* dwte[ids[b,s],:] += dx[b, n_prompt+s, :]  (autograd of models/decoder.py:234 / vision_encoder_decoder.py:85-88) */
int i2t_embed_bwd(const int64_t* ids, const float* dx, float* dwte, int64_t B, int64_t T, int64_t n_prompt, int64_t S,
                  int64_t C, void* stream);
/* out = g / (||g||_2 + 1e-6), norm over all n elements (models/functions.py:19-24).  acc: device double scratch. */
int i2t_gradnorm_scale(const void* g, void* out, double* acc, int64_t n, int dtype, void* stream);
/* Weighted (optionally distilled) LM loss, training/wrapper.py:80-96,120-151.  logits (B,T_logits,V); only the first
 * Tl positions of every sequence are used (labels (B,ld_labels) int64).  weights (B*Tl) and loss_rows (B*Tl) are fp32
 * scratch outputs; loss_out is a device fp32 scalar; dlogits (same shape/dtype as logits, optional) receives
 * dLoss/dlogits for the first Tl positions (other positions must be zeroed by the caller).  ld_logits / ld_teacher: row
 * pitch in elements (>= V) of logits+dlogits / teacher_logits -- the LM head writes rows padded to a multiple of 8 so that
 * its backward GEMMs meet TMA's 16-byte pitch rule.  grad_scale multiplies dlogits only (the loss value is unscaled): the
 * caller that will back-propagate `loss * s` passes s here and skips the rescaling pass over the (B,T,V) gradient. */
int i2t_lm_loss(const void* logits, const void* teacher_logits, const int64_t* labels, float* weights, float* loss_rows,
                float* loss_out, void* dlogits, int64_t B, int64_t T_logits, int64_t Tl, int64_t V, int64_t ld_labels,
                float temperature, float alpha, int inv_sqrt_position, int use_eos_weight, float eos_weight,
                int64_t eos_id, int64_t ignore_index, int64_t ld_logits, int64_t ld_teacher, float grad_scale, int dtype,
                void* stream);
/* Contrastive auxiliary loss, training/wrapper.py:98-118, after the similarity GEMM pred = H Ht^T (R x R fp32, R = B*L, H = the
 * decoder's hidden rows, Ht = wte[labels]): sum_r w_r * CE(pred[r, valid columns] / temperature, target r), w = the LM-loss
 * weights of get_weights (:80-96), columns whose label is ignore_index masked, infinite row losses dropped.  weights / loss_rows:
 * fp32 scratch (R); loss_out: device scalar; dpred (R x R, optional) = grad_scale * d loss / d pred. */
int i2t_contrastive_loss(const float* pred, const int64_t* labels, float* weights, float* loss_rows, float* loss_out, float* dpred,
                         int64_t B, int64_t L, int64_t ld_labels, float temperature, int inv_sqrt_position, int use_eos_weight,
                         float eos_weight, int64_t eos_id, int64_t ignore_index, float grad_scale, void* stream);
/* y = x / max(||x||_2, eps) per row and its backward: F.normalize(p=2, dim=-1) at models/encoder.py:118-119 */
int i2t_l2norm_fwd(const float* x, float* y, int64_t rows, int64_t cols, float eps, void* stream);
int i2t_l2norm_bwd(const float* x, const float* dy, float* dx, int64_t rows, int64_t cols, float eps, void* stream);
/* x *= *scale_ptr (device scalar) */
int i2t_scale_inplace(void* x, const float* scale_ptr, int64_t n, int dtype, void* stream);

/* ---- training-mode dropout ------------------------------------------------------------------------------
 * Masks are a pure function of (rng_state, site, element coordinates) -- Philox4x32-10, csrc/rng.cuh -- so nothing is
 * stored for the backward.  rng_state: device uint64[2] = {seed, step offset}; site: index of the dropout call inside one
 * forward pass.  The reference draws from torch's generator instead: same Bernoulli(1-p) / (1-p) scaling, another stream. */
/* Same contract as i2t_attn_fwd / i2t_attn_bwd with dropout on the attention probabilities (dropout_p of
 * F.scaled_dot_product_attention at reference models/layers.py:465; nn.MultiheadAttention(dropout=) at :537-542; HF
 * attn_pdrop).  Row index (b*H + h)*Tq + i and key j select the mask word (see csrc/rng.cuh drop_attn4). */
int i2t_attn_fwd_dropout(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                         int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                         int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int in_dtype,
                         int out_dtype, float p_drop, const void* rng_state, int64_t site, void* stream);
int i2t_attn_bwd_dropout(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                         void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                         int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride, int64_t kv_batch_stride,
                         int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int dtype, float p_drop,
                         const void* rng_state, int64_t site, void* stream);
/* The same for the packed self-attention buffer (q, k, v = the three C-wide segments of one (B*T, 3C) buffer, likewise dq / dk / dv),
 * plus the backward of the token-level q / k / v dropout of models/layers.py:454-461 that the forward applied to that buffer
 * (i2t_token_dropout with probability p_tok at site tok_site): the gradients come out multiplied by the (row, segment) masks. */
int i2t_attn_bwd_dropout_tok(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                             void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H, int64_t T, int64_t head_dim,
                             int64_t batch_stride, int64_t row_stride, int mask_mode, int64_t n_prompt, int dtype, float p_drop,
                             const void* rng_state, int64_t site, float p_tok, int64_t tok_site, void* stream);
/* out[i] = residual[i] + keep(i) * y[i] / (1-p)   (residual optional; out fp32; n % 4 == 0): resid_dropout
 * models/layers.py:469 + the residual add :596, _MLP.dropout :485 + :606, transformer.drop models/decoder.py:236-243 */
int i2t_dropout_add_fwd(const void* y, const float* residual, float* out, int64_t n, float p, const void* rng_state,
                        int64_t site, int y_dtype, void* stream);
/* g[i] = keep(i) * dy[i] / (1-p), cast to g_dtype: backward of the above w.r.t. y */
int i2t_dropout_bwd(const void* dy, void* g, int64_t n, float p, const void* rng_state, int64_t site, int dy_dtype,
                    int g_dtype, void* stream);
/* In place: x[row, s*seg + c] *= keep(row, s) / (1-p) for s < nseg: the (B,1,T,1) q/k/v masks of models/layers.py:454-461
 * on the packed (rows, 3C) buffer (forward), and on its gradient (backward). */
int i2t_token_dropout(void* x, int64_t rows, int64_t ld, int64_t seg, int64_t nseg, float p, const void* rng_state,
                      int64_t site, int dtype, void* stream);
/* rng_state[1] += 1 on the stream (fresh masks for the next step, also under CUDA-graph replay) */
int i2t_rng_advance(void* rng_state, void* stream);

/* ---- fused multi-tensor optimiser steps (fp32 state).  table: device int64[n_tensors*4] = {p, g, m, v} pointers per
 *      tensor; chunk c updates elements [chunk_off[c], chunk_off[c]+chunk_len[c]) of tensor chunk_tensor[c];
 *      step counts from 1; grad_scale multiplies the gradient on the fly (1.0 = reference semantics). ---------------- */
/* torch.optim.AdamW as constructed at trainer.py:169-172 (eps 1e-8, amsgrad off) */
int i2t_adamw_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                    int64_t n_chunks, double lr, double beta1, double beta2, double eps, double weight_decay,
                    int64_t step, double grad_scale, void* stream);
/* SNRAdam, models/optimizer.py:56-113 */
int i2t_snradam_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                      const int32_t* chunk_len, int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int64_t step, double grad_scale, void* stream);
/* momentum-distillation teacher: p_m = p_m*momentum + p*(1-momentum), training/wrapper.py:53-60; table = {p_m, p, s, 0}
 * with s = the bf16 copy of p_m the bf16 forward reads (refreshed in the same pass) or 0 */
int i2t_ema_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                  int64_t n_chunks, double momentum, void* stream);
/* dst (bf16) = src (fp32) over a tensor list, table = {src, dst, 0, 0}: refreshes the bf16 copies of the weights an
 * optimiser step wrote (the reference's autocast re-casts fp32 masters on every use, training/utils.py:96) */
int i2t_cast_bf16_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                        int64_t n_chunks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* I2T_H_ */
