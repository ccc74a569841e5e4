/*
 * ggml_b200.h -- C ABI of libggml_b200.so, the B200 (sm_100a) quantized-inference backend for ggllm.cpp.
 *
 * Plain C: pointers, sizes and ggml type ids (enum ggml_type, ggml.h:241-262) only.  No torch / C++ types.
 * Pointers named *_dev are CUDA device pointers on the current device; everything else is host memory.
 * Errors follow the reference backend's convention (ggml-cuda.cu:22-51, ggml.h:204-210): CUDA failures print
 * to stderr and exit(1), contract violations abort(); there are no error codes to check.
 * There is NO CPU fallback: every entry point needs a CUDA device.
 *
 * Three layers are exported:
 *   1. b200_*            kernel-level operators (this file, part A): what ggml_cuda_op_* do for one graph node
 *                        (ggml-cuda.cu:2153-2518) but device-resident and with the CPU oracle's numerics.
 *   2. b200_falcon_*     the Falcon eval path (part B): loader upload + layer-range partition + falcon_eval,
 *                        i.e. the "rewired" libfalcon offload (libfalcon.cpp:1552-1959, 2011-2588, 4566).
 *   3. ggml_cuda_*       the reference's own operator surface (ggml-cuda.h:31-60), declared in
 *                        ggml_b200_cuda_surface.h, so that ggml.c / libfalcon.cpp built with -DGGML_USE_CUBLAS
 *                        link against this library unchanged.
 */
#ifndef GGML_B200_H
#define GGML_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ========================================= part A: kernel-level operators ================================= */

/* Replaces ggml_init_cublas (ggml-cuda.cu:1982-2041).  Selects `device`, creates the backend stream.
 * Returns the number of SMs.  Idempotent per device. */
int    b200_init(int device);
int    b200_device_count(void);
/* Run all subsequent b200_* work on this cudaStream_t (NULL = the backend's own stream). */
void   b200_set_stream(void * cuda_stream);
void   b200_synchronize(void);

/* CUDA-event timing on a stream (NULL = the backend stream): what bench.py times kernels with */
void * b200_event_create(void);
void   b200_event_destroy(void * ev);
void   b200_event_record(void * ev, void * cuda_stream);
void   b200_event_synchronize(void * ev);
float  b200_event_elapsed_ms(void * ev_start, void * ev_stop);
void   b200_stream_synchronize(void * cuda_stream);

/* Raw device memory helpers so that C / ctypes callers can stage buffers without another CUDA binding. */
void * b200_malloc(size_t bytes);
void   b200_free(void * p_dev);
void   b200_memcpy_h2d(void * dst_dev, const void * src, size_t bytes);
void   b200_memcpy_d2h(void * dst, const void * src_dev, size_t bytes);
void   b200_memset(void * p_dev, int value, size_t bytes);
/* pinned host memory: ggml_cuda_host_malloc / ggml_cuda_host_free (ggml-cuda.cu:2079-2103); may return NULL */
void * b200_host_malloc(size_t bytes);
void   b200_host_free(void * p);

/* ---- weights.  Replaces ggml_cuda_transform_tensor (ggml-cuda.cu:3030-3073): `blocks` is the tensor's raw
 * data exactly as stored in a GGML/GGCC file: M rows of K/blk blocks (block_q4_0 .. block_q6_K, or f16/f32 rows).
 * The device copy is re-laid out in planes (DESIGN.md "HBM layout"); the caller may free `blocks` on return. */
typedef struct b200_weight b200_weight;
b200_weight * b200_weight_upload(int ggml_type, int64_t K, int64_t M, const void * blocks);
/* well-formed pseudo-random blocks generated on the device (throughput runs on 40B/180B-sized synthetic models) */
b200_weight * b200_weight_random(int ggml_type, int64_t K, int64_t M, uint64_t seed);
void          b200_weight_free(b200_weight * w);      /* ggml_cuda_free_data, ggml-cuda.cu:3075-3092 */
size_t        b200_weight_device_bytes(const b200_weight * w);
/* dst[r][0..K) = dequantised row rows[r] (rows_dev == NULL: rows 0..nrows-1).  Bit-exact with
 * dequantize_row_q* (ggml.c:1509-1619, k_quants.c:344-877); this is ggml_get_rows (ggml.c:11975-12002). */
void          b200_dequantize_rows(const b200_weight * w, const int32_t * rows_dev, int nrows, float * dst_dev, int64_t dst_stride);

/* ---- quantised activations: the CPU mat-mul's INIT pass (ggml.c:11462-11476): rows quantised to the weight
 * type's vec_dot_type (Q8_0 / Q8_1 / Q8_K).  Codes and scales are bit-exact with the CPU on an x86 host. */
typedef struct b200_actq b200_actq;
b200_actq * b200_actq_alloc(int weight_ggml_type, int64_t K, int N);
void        b200_actq_free(b200_actq * a);
void        b200_quantize_act(const float * x_dev, int64_t x_stride, b200_actq * a);
/* test hook: copy out codes q[N*K], scales d[N*K/blk], Q8_1 sums s[N*K/32] (or NULL), block sums bs (or NULL) */
void        b200_actq_download(const b200_actq * a, int8_t * q, float * d, float * s, int16_t * bs);

/* ---- y[n][m] = sum_k W[m][k] x[n][k].  Replaces ggml_cuda_mul_mat (ggml-cuda.cu:2931-2951):
 * N == 1..b200_mmv_max_n(): fused dequantise + integer-dot mat-vec (replaces dequantize_mul_mat_vec*,
 *                            ggml-cuda.cu:475-845, 1121-1171)
 * larger N               : tcgen05 tensor-core GEMM with fused dequantisation (replaces to_fp16_cuda +
 *                            float_to_half + cublasGemmEx, ggml-cuda.cu:2353-2403)
 * x/y are fp32 device buffers with row strides in floats. */
void   b200_mul_mat(const b200_weight * w, const float * x_dev, int64_t x_stride, int N, float * y_dev, int64_t y_stride);
/* the two halves separately; epilogue: 0 none, 1 GELU (fp16-LUT semantics), 2 y = (dot + r1) + r2 */
void   b200_mul_mat_vec_q(const b200_weight * w, const b200_actq * a, float * y_dev, int64_t y_stride,
                          int epilogue, const float * r1_dev, const float * r2_dev);
/* the decode mat-vec with its producers folded into the prologue (N = 1, Q4_K / Q4_0 weights):
 *   gamma != NULL : y = W * Q( LayerNorm((ra + rb) + x) * gamma + beta ), ra/rb optional, x_out (optional) receives
 *                   the updated row (ra + rb) + x   -- libfalcon.cpp:2399-2400, 2166-2185 + ggml.c:10568-10595, 11462-11476
 *   gamma == NULL : y = W * Q(x)
 * returns 0 if the type / shape is not covered by the fused kernel (callers then use b200_layernorm + b200_mul_mat) */
int    b200_mul_mat_vec_fused(const b200_weight * w, const float * x_dev, const float * ra_dev, const float * rb_dev,
                              const float * gamma_dev, const float * beta_dev, float * x_out_dev, float * y_dev, int epilogue);
/* y = epilogue(W * a_in) for N = 1 and, from the SAME kernel, a_out = Q(y): the next mat-mul's INIT-pass quantisation
 * (ggml.c:11462-11476) done chunk by chunk as the CTAs that own a 256-value chunk finish (decode: ffn_up -> ffn_down,
 * libfalcon.cpp:2389-2394).  a_out must have K == rows of w and the activation type of the next weight.
 * returns 0 if the type / shape is not covered (rows % 256 != 0, or not Q4_K / Q4_0). */
int    b200_mul_mat_vec_q_chain(const b200_weight * w, const b200_actq * a_in, float * y_dev, int epilogue, b200_actq * a_out);
/* fp32 -> weight blocks in the FILE layout (18 B per 32 for Q4_0, 144 B per 256 for Q4_K) on the device, bit for bit what
 * quantize_row_q4_0_reference (ggml.c:927-962) / quantize_row_q4_K_reference (k_quants.c:542-605) write.
 * Returns 0 for a type without a device quantiser (quantise on the host then), 1 on success. */
int    b200_quantize_weights(int ggml_type, const float * x_dev, void * blocks_dev, int64_t n_elems);
int    b200_mmv_max_n(void);
/* the GEMM half alone, on fp16 activations x[n][k] already on the device (what b200_mul_mat does after quantising):
 * impl 1 = tcgen05 tensor-core kernel (returns 0 if the shape is not covered: N > 512 or K % 64 != 0),
 * impl 0 = CUDA-core kernel with identical operand rounding (the test reference for impl 1). */
int    b200_mul_mat_f16(const b200_weight * w, const void * x_f16_dev, int64_t x_stride, int N, float * y_dev, int64_t y_stride,
                        int epilogue_gelu, int impl);

/* ---- the other operators of the Falcon graph, CPU-oracle numerics */
/* y = norm(x) * g + b per row of n values (g, b may be NULL = plain ggml_norm, ggml.c:10540-10599) */
void   b200_layernorm(const float * x_dev, int64_t x_stride, const float * g_dev, const float * b_dev,
                      float * y_dev, int64_t y_stride, int n, int rows);
void   b200_gelu(const float * x_dev, float * y_dev, int64_t n);                        /* ggml.c:3461-3484 */
void   b200_add(const float * a_dev, const float * b_dev, float * y_dev, int64_t n);   /* ggml.c:8312- */
/* NeoX RoPE (mode 2) with dynamic NTK (ggml.c:12875-12898, 12957-12979) on x[n_tok][n_head][head_dim], in place */
void   b200_rope_neox(float * x_dev, int n_tok, int n_head, int head_dim, int64_t tok_stride, int n_past,
                      int n_ctx_rope, int dynamic_mode, float ntk_alpha, int freq_base);
/* RoPE(Q,K) + KV append + causal multi-query attention for one layer (libfalcon.cpp:2229-2366).
 * qkv_dev: [n_tok][(n_head + 2 n_head_kv) * head_dim] (rotated in place like ggml_rope_inplace);
 * k/v cache: [n_ctx][n_head_kv][head_dim] f32; out: [n_tok][n_head * head_dim]. */
void   b200_attention(float * qkv_dev, float * k_cache_dev, float * v_cache_dev, float * out_dev,
                      int n_head, int n_head_kv, int head_dim, int n_tok, int n_past, int n_ctx, int n_ctx_rope);

/* ---- the two fused decode-step nodes exactly as the eval path launches them (per-operator parity tests at the real model widths)
 * [x = (ra + rb) + x, written back when ra != NULL] ; a1 = Q(norm(x) * g1 + b1) ; a2 = Q(norm(x) * g2 + b2) (a2 / g2 / b2 optional):
 * residual adds (libfalcon.cpp:2399-2400) + LayerNorm(s) (ggml.c:10540-10599, libfalcon.cpp:2166-2185) + the next mat-mul's INIT
 * pass (ggml.c:11462-11476).  rows == 1 and n <= 8192: thread-block-cluster kernel; otherwise one CTA per row. */
void   b200_layernorm_q(float * x_dev, int64_t x_stride, const float * ra_dev, const float * rb_dev,
                        const float * g1_dev, const float * b1_dev, b200_actq * a1,
                        const float * g2_dev, const float * b2_dev, b200_actq * a2, int n, int rows);
/* b200_attention for ONE new token with the split-KV decode kernels; qout (optional) also receives the output row quantised for
 * the wo mat-mul.  Returns 1 when that quantisation ran inside the attention combine step, 0 when it needed its own kernel. */
int    b200_attention_decode(float * qkv_dev, float * k_cache_dev, float * v_cache_dev, float * out_dev,
                             int n_head, int n_head_kv, int head_dim, int n_past, int n_ctx, int n_ctx_rope, b200_actq * qout);

/* ---- sampling on the device: falcon_main's default chain (examples/falcon/falcon_main.cpp:945-975) over a logits row in HBM:
 * llama_sample_repetition_penalty over the last repeat_last_n ids, then temp <= 0 ? llama_sample_token_greedy :
 * llama_sample_top_k -> llama_sample_top_p -> llama_sample_temperature -> llama_sample_token (libfalcon.cpp:3281-3307, 3433-3447,
 * 3094-3150, 3269-3279, 3449-3468).  The draw reproduces std::discrete_distribution over std::mt19937(seed), so the same seed samples
 * the same ids as the reference.  top_k must be 1..1024 (the reference's "<= 0 means the whole vocabulary" is not supported),
 * repeat_last_n <= 256; tail-free / typical / mirostat / frequency and presence penalties (all off by default) are not implemented. */
typedef struct { int32_t top_k; float top_p; float temp; float repeat_penalty; int32_t repeat_last_n; uint32_t seed; } b200_sampling_params;
typedef struct b200_sampler b200_sampler;
/* last_tokens[0..n_last): the ids already generated / in the prompt, oldest first; the last repeat_last_n of them seed the window.  NULL on bad parameters. */
b200_sampler * b200_sampler_create(const b200_sampling_params * p, const int32_t * last_tokens, int n_last);
int32_t        b200_sampler_sample(b200_sampler * s, const float * logits_dev, int n_vocab);   /* samples, appends the id to the window, returns it */
void           b200_sampler_free(b200_sampler * s);

/* ========================================= part B: Falcon eval path ====================================== */

typedef struct b200_falcon b200_falcon;

typedef struct {
    int32_t n_vocab, n_embd, n_head, n_head_kv, n_layer;
    int32_t falcon_type;      /* 7 | 40: selects the single- or dual-LayerNorm layer (libfalcon.cpp:1578-1593, 2177) */
    int32_t n_ctx;            /* KV capacity (falcon_context_params.n_ctx, libfalcon.h:91) */
    int32_t n_batch;          /* largest n_tokens of one eval (libfalcon.h:92) */
    /* layer-range pipeline (replaces tensor_split / n_gpu_layers, libfalcon.h:93-97): this process owns layers
     * [layer_first, layer_last) of the model; rank/world describe its place in the NCCL pipeline. */
    int32_t layer_first, layer_last;
    int32_t rank, world;
} b200_falcon_params;

/* Create the device-resident model shell (KV cache, activation arena, CUDA graphs).  Weights are attached
 * afterwards, tensor by tensor, under the reference's GGCC tensor names (libfalcon.cpp:1764-1861), e.g.
 * "transformer.h.3.mlp.dense_h_to_4h.weight".  Tensors of layers outside [layer_first, layer_last) are ignored. */
b200_falcon * b200_falcon_create(const b200_falcon_params * params);
void          b200_falcon_set_tensor(b200_falcon * f, const char * name, int ggml_type, int n_dims,
                                     const int64_t * ne, const void * data);
/* random-init tensor of the named shape generated on the device (synthetic throughput models) */
void          b200_falcon_set_tensor_random(b200_falcon * f, const char * name, int ggml_type, uint64_t seed);
/* load every tensor of a GGCC v10 file (format: libfalcon.cpp:770-973) through the GPU-direct path: mmap -> pinned ring buffers ->
 * cudaMemcpyAsync -> on-the-fly planar repack, several host threads, two buffers in flight each, one synchronize at the end
 * (replaces libfalcon.cpp:1196-1270 + the blocking per-tensor copy of ggml-cuda.cu:3030-3073).  The file is validated while it is
 * read (bounds, types, shapes, names); returns 0 on success, -1 on an unreadable / malformed / mismatching file. */
int           b200_falcon_load_ggcc(b200_falcon * f, const char * path);
/* seconds the last b200_falcon_load_ggcc took (header parse to the final synchronize) and the quantised-matrix bytes it streamed */
double        b200_falcon_load_seconds(const b200_falcon * f, size_t * bytes);
/* hparams of a GGCC file without loading it (fills n_vocab..falcon_type); returns 0 on success */
int           b200_ggcc_read_hparams(const char * path, b200_falcon_params * out);
void          b200_falcon_free(b200_falcon * f);
size_t        b200_falcon_weight_bytes(const b200_falcon * f);   /* algorithmic bytes of the resident quantised matrices */

/* NCCL pipeline plumbing (world > 1): rank 0 creates the id, every rank passes the same 128 bytes. */
void          b200_nccl_unique_id(void * id128);
void          b200_falcon_init_pipeline(b200_falcon * f, const void * id128);

/* falcon_eval (libfalcon.cpp:4566): n_tokens token ids at position n_past.  Host buffers in and out:
 * token ids H2D and logits D2H are part of the call.  logits: n_vocab floats of the last token, or
 * n_tokens * n_vocab if all_logits (falcon_context_params.logits_all).  n_ctx_rope = the rope's 4th parameter
 * (n_max_real_ctx ? that : n_ctx, libfalcon.cpp:2229-2230); 0 = use n_ctx.
 * In a pipeline every rank calls it; only the last rank's `logits` are written.  Returns 0 on success. */
int           b200_falcon_eval(b200_falcon * f, const int32_t * tokens, int n_tokens, int n_past, int n_ctx_rope,
                               float * logits, int all_logits);
/* device-resident decode step for throughput measurement: token id already on the device, logits stay on the
 * device (b200_falcon_logits_dev).  Same kernels/graph as b200_falcon_eval minus the two PCIe copies. */
int           b200_falcon_decode_dev(b200_falcon * f, const int32_t * token_dev, int n_past, int n_ctx_rope);   /* 0 = ok, 1 = n_past outside [0, n_ctx) */
const float * b200_falcon_logits_dev(const b200_falcon * f);
/* greedy generation without leaving the device (single GPU): feeds `first_token` at position n_past, then n_steps times
 * "decode, arg-max of the logits (lowest index on ties), use it as the next token".  tokens_out[i] = token sampled after
 * step i.  What falcon_main does with top_k = 1 (llama_sample_token_greedy, libfalcon.cpp:3464-3473), minus the 260 KB
 * logits D2H and the host scan per token.  Returns 0 on success. */
int           b200_falcon_generate_greedy(b200_falcon * f, int32_t first_token, int n_past, int n_steps, int n_ctx_rope, int32_t * tokens_out);
/* KV cache rows [pos, pos + n) of one (global) layer index, host buffers of n * n_head_kv * head_dim floats each (either may be NULL).
 * The building block of session save / restore (falcon_copy_state_data / falcon_set_state_data, libfalcon.cpp:4313-4490) over the
 * device-resident cache.  Return 0 on success, 1 if the layer is not on this rank or the range leaves [0, n_ctx). */
int           b200_falcon_kv_read(b200_falcon * f, int layer, int pos, int n, float * k_out, float * v_out);
int           b200_falcon_kv_write(b200_falcon * f, int layer, int pos, int n, const float * k_in, const float * v_in);
/* session file over the device KV cache (what falcon_save_session_file / falcon_load_session_file keep of the KV state,
 * libfalcon.cpp:4490-4563), in this library's own container: positions [0, n_tokens) of every local layer.  save: 0 / -1.
 * load: the number of positions restored (continue evaluating at that n_past), -1 on a missing / truncated / mismatching file. */
int           b200_falcon_save_kv(b200_falcon * f, const char * path, int n_tokens);
int           b200_falcon_load_kv(b200_falcon * f, const char * path);
/* pseudo-random K / V rows for positions [pos, pos + n) of every local layer, generated on the device (long-context throughput runs) */
int           b200_falcon_kv_fill_random(b200_falcon * f, int pos, int n, uint64_t seed);
/* b200_falcon_generate_greedy with the sampling chain above run on the device after every step (every rank of a pipeline calls it with
 * the same arguments; the last rank samples).  Returns 0 on success, 1 on bad arguments. */
int           b200_falcon_generate(b200_falcon * f, const b200_sampling_params * p, const int32_t * last_tokens, int n_last,
                                   int32_t first_token, int n_past, int n_steps, int n_ctx_rope, int32_t * tokens_out);
/* the cudaStream_t the eval path runs on (for event timing) */
void *        b200_falcon_stream(b200_falcon * f);
/* roofline probe: every resident quantised mat-vec of this rank (4 per layer + lm_head) launched back to back,
 * `reps` passes, timed with CUDA events on the eval stream.  Returns total ms; fills the launch count and the
 * algorithmic weight bytes streamed in that region. */
float         b200_falcon_profile_matvec(b200_falcon * f, int reps, int * n_launches, size_t * bytes);
/* number of kernel launches (graph nodes) issued by the most recent eval on this rank */
int           b200_falcon_last_launches(const b200_falcon * f);
int           b200_attention_long_launches(void);     /* diagnostics: launches (eager or captured) of the long-context decode attention kernels so far */
/* CUDA-event time (ms) of the most recent eval's device work on this rank */
float         b200_falcon_last_ms(const b200_falcon * f);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
