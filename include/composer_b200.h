/*
 * composer_b200 — C ABI of the B200-native Transformer hot path.
 *
 * This is the drop-in boundary: the Python class `composer_b200.models.Transformer`
 * (which mirrors the reference's `composer.models.Transformer`,
 * composer/models/transformer.py:599-960) drives these entry points through
 * ctypes.  Everything here is plain C: ints, floats, raw device pointers and an
 * opaque engine handle.  No torch / CUDA C++ types cross the boundary; streams
 * are passed as `void*` (a `cudaStream_t`).
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure
 *     (-1 invalid argument / unsupported configuration, -2 CUDA error); the
 *     message is available from cb200_last_error() (thread-local).
 *   - nothing allocates device memory: the caller owns the parameter, gradient,
 *     optimizer, shadow and workspace arenas (sizes from the *_elems / *_bytes
 *     queries) and binds them once.
 *   - a handle is not thread-safe; use one handle per rank / host thread.
 *   - all work is enqueued on the given stream; no call synchronises the device
 *     except cb200_generate (see below).
 *
 * Each entry point cites the reference code it replaces.
 */
#ifndef COMPOSER_B200_H
#define COMPOSER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hyperparameters; names follow default_config.yml:32-45 and Transformer.__init__
 * (composer/models/transformer.py:610-614). */
typedef struct cb200_config {
    int32_t vocab_size;              /* cli._get_event_vocab_size, composer/cli.py:400-412 */
    int32_t embedding_size;
    int32_t window_size;
    int32_t decoder_layers_count;
    int32_t attention_head_count;
    float attention_dropout_rate;
    float residual_dropout_rate;
    float layer_normalization_epsilon;
    int32_t scale_attention;         /* transformer.py:345-348 */
    int32_t use_layer_normalization; /* transformer.py:583-591 */
} cb200_config;

const char* cb200_last_error(void);
int cb200_abi_version(void);
/* Number of kernels this library has launched in this process (graph replays included). */
long long cb200_launch_count(void);

/* ---- parameter arena ------------------------------------------------------
 * All trainable variables live in one flat fp32 arena in Keras creation order
 * with the reference's variable names (wte/weight, wpe/embeddings,
 * h_{i}/ln_1/gamma, ..., ln_f/beta; transformer.py:116, 257-269, 482-494, 551,
 * 563, 667-694).  Weights are stored [in, out] as in Conv1D (transformer.py:207). */
int64_t cb200_param_elems(const cb200_config* cfg);
int cb200_param_tensor_count(const cb200_config* cfg);
int cb200_param_tensor_info(const cb200_config* cfg, int index, char* name, int name_capacity, int64_t* offset,
                            int32_t* rows, int32_t* cols);

/* ---- engine ---------------------------------------------------------------- */
int cb200_engine_create(const cb200_config* cfg, void** engine);
int cb200_engine_destroy(void* engine);

/* bf16 shadow arena ([in,out] copies, [out,in] transposes and the padded wte^T). */
int64_t cb200_shadow_elems(void* engine);
/* Activation workspace for batches of up to B x T tokens (training keeps the
 * tensors backward needs; inference only the live ones). */
int64_t cb200_workspace_bytes(void* engine, int B, int T, int training);

/* Bind caller-owned device memory.  grads / adam_m / adam_v may be NULL for an
 * inference-only engine. */
int cb200_engine_bind(void* engine, float* params, float* grads, float* adam_m, float* adam_v, void* shadow,
                      void* workspace, int64_t workspace_bytes, int max_B, int max_T, int training);

/* Re-derive the bf16 shadows from the fp32 parameters (after loading a
 * checkpoint; cb200_adam_step does it itself). */
int cb200_refresh_shadows(void* engine, void* stream);

/* Transformer.call + loss + accuracy (transformer.py:696-833, 888, 916-918,
 * 924-926).  ids / labels: device int32 [B, T].  labels may be NULL (no loss).
 * training != 0 enables dropout (Philox keyed by seed, step) and keeps
 * activations for cb200_backward.  *loss_sum (device float) += sum over
 * positions of (logsumexp(z) - z[y]); *correct (device int) += #(argmax == y);
 * either may be NULL.  logits (device fp32 [B*T, vocab]) may be NULL.
 * grad_scale multiplies d(loss_sum)/d(logits) (pass 1 / (global token count)
 * for the reference's mean reduction). */
int cb200_forward(void* engine, const int32_t* ids, const int32_t* labels, int B, int T, int training, uint64_t seed,
                  uint32_t step, float grad_scale, float* loss_sum, int32_t* correct, float* logits, void* stream);

/* tape.gradient (transformer.py:920): accumulates into the gradient arena
 * (zero it with cb200_zero_grads first).  Stages, for overlapping the
 * data-parallel all-reduce with the rest of backward:
 *   stage -1            everything
 *   stage 0             head: tied logits matmul, ln_f
 *   stage 1..L          decoder block (L - stage + 1), i.e. blocks in reverse order
 *   stage L+1           embeddings */
int cb200_backward(void* engine, int stage, void* stream);
int cb200_zero_grads(void* engine, void* stream);

/* optimizers.Adam as TF-2 Keras applies it (transformer.py:887, 921): t is the
 * 1-based step; epsilon is added to sqrt(v).  grad_scale is applied to the
 * gradients first (e.g. 1/world_size after a sum all-reduce). */
int cb200_adam_step(void* engine, float learning_rate, float beta_1, float beta_2, float epsilon, int64_t t,
                    float grad_scale, void* stream);

/* ---- generation -------------------------------------------------------------
 * KV-cache decoding through the model's `past=` semantics (transformer.py:423-437,
 * 735-770) with the CLI's sampling rule (cli.py:663-676).
 * cache: device bf16 scratch of cb200_kv_cache_elems(engine, B, t_max) elements.  Its layout is private to the call
 * (a generation always starts from an empty cache): [L, B, H, t_max, 2, d_h] with swizzled 16-byte pieces for the
 * persistent cluster kernel, which is used when t_max is a multiple of 64 and the shape allows it (see
 * cb200_set_decode_impl), [L, 2, B, H, t_max, d_h] for the per-step kernels.
 * prompt: device int32 [B, prompt_len] (every sequence has the same prompt length).
 * out_ids: device int32 [B, n_new].  temperature <= 0 selects argmax.
 * seq_index_base: global index of sequence 0 (keys the Philox stream so that
 * results do not depend on the sharding).  uniforms_out (device fp32 [B, prompt_len-1+n_new],
 * may be NULL) receives the uniform draw used at every step (for parity tests).
 * last_logits (device fp32 [B, vocab], may be NULL) receives the logits of the
 * final step.  The call runs the whole generation (one persistent kernel, or a CUDA graph of one step replayed)
 * and synchronises the stream before returning. */
int64_t cb200_kv_cache_elems(void* engine, int B, int t_max);
int64_t cb200_decode_workspace_bytes(void* engine, int B);
int cb200_generate(void* engine, void* cache, int t_max, void* workspace, int64_t workspace_bytes,
                   const int32_t* prompt, int B, int prompt_len, int n_new, float temperature, uint64_t seed,
                   int64_t seq_index_base, int32_t* out_ids, float* uniforms_out, float* step_logits, void* stream);

/* ---- the model call with a KV cache the CALLER keeps (composer/models/transformer.py:696-833: `past` / `presents`) ----
 * cache: device bf16 [L, 2, B, H, t_max, d_h]; layer l's slice [2, B, H, :t, d_h] is exactly the reference's
 * `present` of that layer for a context of t tokens (transformer.py:419-432: split_heads(key), split_heads(value),
 * stacked), with room for t_max positions so that a step appends in place instead of re-concatenating (:423-426).
 *
 * cb200_prefill      = Transformer.call(inputs [B, T], past=None): one batched pass over the whole prompt (the training
 *                      forward kernels in inference mode); writes the presents of positions 0 .. T-1 and, when `logits`
 *                      is not NULL, the fp32 logits [B, T, vocab].  Needs cb200_engine_bind for >= B x T tokens.
 * cb200_decode_step  = Transformer.call(inputs[:, -1:], past=presents) (:735-737): ids [B] (device int32) are the
 *                      tokens at position `pos` = the length of the past; appends their k, v rows at `pos` and writes
 *                      the fp32 logits [B, vocab] of that position.  workspace: cb200_decode_workspace_bytes(engine, B).
 * cb200_generate (above) uses the same prefill for prompts longer than one token (its cache layout stays private);
 * cb200_set_decode_prefill(0) restores the token-by-token teacher forcing of round 1 (A/B and tests). */
int cb200_prefill(void* engine, const int32_t* ids, int B, int T, void* cache, int t_max, float* logits, void* stream);
int cb200_decode_step(void* engine, void* cache, int t_max, void* workspace, int64_t workspace_bytes, const int32_t* ids,
                      int B, int pos, float* logits, void* stream);
int cb200_set_decode_prefill(int enabled);

/* ---- single kernels (unit parity tests, ncu isolation) ---------------------- */
/* D[M,N] = A[M,K] B[N,K]^T (+ bias).  kind: 0 bias, 1 bias+gelu (out1 = gelu), 2 bias+dropout+residual(aux),
 * 3 acc * gelu'(aux), 4 weight gradient outf += A^T B with A [K,M], B [K,N], 6 bias with B stored [K,N]. */
int cb200_gemm(int kind, int M, int N, int K, const void* A, int lda, const void* B, int ldb, const float* bias,
               void* out0, int ld_out0, void* out1, int ld_out1, const void* aux, int ld_aux, float* outf, int ld_outf,
               float dropout_rate, uint64_t seed, uint32_t step, uint32_t site, uint32_t layer, void* stream);
/* Fused tied-logits matmul + softmax cross-entropy (+ gradient). */
int cb200_logits_ce(int M, int V, int E, const void* h, const void* wte, const int32_t* labels, void* dlogits,
                    int ld_dlogits, float grad_scale, float* loss_sum, int32_t* correct, float* logits, void* stream);
int cb200_embed_fwd(const int32_t* ids, const float* wte, const float* wpe, void* out, int B, int T, int E, int pos0,
                    int vocab, float dropout_rate, uint64_t seed, uint32_t step, void* stream);
int cb200_embed_bwd(const int32_t* ids, const void* dh, float* dwte, float* dwpe, int B, int T, int E, int pos0,
                    int vocab, float dropout_rate, uint64_t seed, uint32_t step, void* stream);
int cb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int rows, int E,
                        float eps, void* stream);
int cb200_layernorm_bwd(const void* dy_a, const void* dy_b, const void* x, const float* stats, const float* gamma,
                        const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int E, void* stream);
int cb200_bias_grad(const void* dy, void* g_out, float* dbias, int rows, int N, float dropout_rate, uint64_t seed,
                    uint32_t step, uint32_t site, uint32_t layer, void* stream);
int cb200_attention_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int D, float scale,
                        float dropout_rate, uint64_t seed, uint32_t step, uint32_t layer, void* stream);
int cb200_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                        float* dq_acc, void* dqkv, int B, int T, int H, int D, float scale, float dropout_rate,
                        uint64_t seed, uint32_t step, uint32_t layer, void* stream);
/* Attention forward has three implementations of the same arithmetic: 0 = tcgen05 / TMEM, probabilities kept in
 * TMEM and consumed by a TS-form MMA (default); 1 = warp-level mma.sync (round 1; kept for A/B measurements and as
 * a cross-check in the tests); 2 = tcgen05 / TMEM with the probabilities staged through shared memory. */
int cb200_set_attention_fwd_impl(int impl);
/* Diagnostic: device buffer of 12 x 512 int64 that receives the event timeline ((event << 40) | SM clock; word 0 of
 * each 512-word region = number of events) of every warp of one CTA of the attention backward kernel.  NULL disables. */
int cb200_set_attention_trace(void* buffer);
/* Selects how cb200_generate runs: 0 (default) = one persistent thread-block-cluster kernel for the whole
 * generation when the shape allows it (embedding_size 256 or 512, heads a multiple of 4), 1 = one CUDA graph
 * of per-layer kernels per step.  max_clusters > 0 caps the clusters of the persistent kernel (tests);
 * cluster_size 0 = automatic, 4 or 8 = CTAs per cluster. */
int cb200_set_decode_impl(int impl, int max_clusters, int cluster_size);
/* Diagnostic: device buffer of 64 int64 (15 phases; ring waits from 16; marks inside the linear phases from 24) that receives the cycles the first CTA of the persistent decode kernel
 * spent in each phase (embedding, ln_1, c_attn, attention, barrier, c_proj, barrier, ln_2 + c_fc, barrier,
 * mlp c_proj, barrier, ln_f + logits, barrier, sampling, barrier), summed over the generation.  NULL disables. */
int cb200_set_decode_profile(void* counters);
/* Clusters of `cluster_size` (4 or 8) CTAs, up to 16 sequences each, of the persistent decode kernel that are
 * co-resident on the current device; 0 when the engine's shape is not served at that cluster size. */
int cb200_decode_cluster_capacity(void* engine, int cluster_size);
/* Keep masks (1 = kept) exactly as the kernels draw them; for parity tests with dropout on. */
int cb200_attention_dropout_mask(uint8_t* mask, int B, int T, int H, float dropout_rate, uint64_t seed, uint32_t step,
                                 uint32_t layer, void* stream);
int cb200_rowmajor_dropout_mask(uint8_t* mask, int rows, int cols, float dropout_rate, uint64_t seed, uint32_t step,
                                uint32_t site, uint32_t layer, void* stream);
int cb200_adam(float* p, const float* g, float* m, float* v, void* shadow, int64_t n, float lr_t, float beta_1,
               float beta_2, float epsilon, float grad_scale, void* stream);
int cb200_decode_attention(const void* qkv, void* kcache, void* vcache, void* out, const int32_t* pos, int B, int H,
                           int D, int t_max, float scale, void* stream);

/* Skinny linear layer of the decode step: Y[B, N] = X[B, K] Wt^T + bias, Wt stored [N, K] (bf16).
 * epilogue: 0 bias, 1 bias + gelu, 2 bias + residual (res [B, ldres]). */
int cb200_decode_linear(int epilogue, const void* X, int ldx, const void* Wt, const float* bias, const void* res,
                        int ldres, void* Y, int ldy, int B, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMPOSER_B200_H */
