// ref_harness.cpp -- thin C driver over the UNMODIFIED reference's public C API (libfalcon.h:149-335).
//
// TEST INFRASTRUCTURE ONLY.  Compiled (oracle/Makefile, target `ref`) together with the reference objects
// into oracle/_ref/libfalcon_ref.so (CPU build) and oracle/_ref/libfalcon_hook.so (-DGGML_USE_CUBLAS build whose
// ggml_cuda_* symbols are resolved by our libggml_b200.so).  It contains no arithmetic of its own: it
// only calls falcon_init_backend / falcon_init_from_file / falcon_eval / falcon_get_logits, the same
// sequence examples/falcon/falcon_main.cpp:146-159, 822-843 performs, so that Python (ctypes) does not have
// to pass C++ structs by value.
#include "libfalcon.h"
#include "ggml.h"
#include <cstdio>
#include <cstring>
#include <vector>

extern "C" {

// returns an opaque handle or NULL.  n_gpu_layers is ignored by the CPU build.
void * refh_load(const char * path, int n_ctx, int n_batch, int n_gpu_layers, int logits_all) {
    static bool backend_ready = false;
    if (!backend_ready) { falcon_init_backend(); backend_ready = true; }
    falcon_context_params p = falcon_context_default_params();
    p.n_ctx = n_ctx;
    p.n_batch = n_batch;
    p.n_gpu_layers = n_gpu_layers;
    p.seed = 1;
    p.f16_kv = false;              // Falcon always runs an f32 KV cache (examples/falcon_common.cpp:786)
    p.logits_all = logits_all != 0;
    p.use_mmap = true;
    return falcon_init_from_file(path, p);
}

// evaluates `n` tokens at position n_past; copies n_vocab (or n*n_vocab if the context was loaded with
// logits_all) floats to `logits`.  Returns 0 on success (falcon_eval's convention, libfalcon.cpp:4566).
int refh_eval(void * h, const int * tokens, int n, int n_past, int n_threads, int n_max_real_ctx, float * logits, int logits_all) {
    falcon_context * ctx = (falcon_context *) h;
    falcon_evaluation_config cfg;
    cfg.n_tokens = n;
    cfg.n_past = n_past;
    cfg.n_threads = n_threads;
    cfg.n_max_real_ctx = n_max_real_ctx;
    int rc = falcon_eval(ctx, tokens, cfg);
    if (rc != 0) return rc;
    const int nv = falcon_n_vocab(ctx);
    memcpy(logits, falcon_get_logits(ctx), sizeof(float) * (size_t) nv * (logits_all ? n : 1));
    return 0;
}

int refh_n_vocab(void * h) { return falcon_n_vocab((falcon_context *) h); }
void refh_free(void * h) { llama_free((falcon_context *) h); }

// quantise a GGCC file with the reference's own quantiser driver (libfalcon.cpp:3533-3743); ftype = enum llama_ftype
int refh_quantize(const char * in, const char * out, int ftype, int nthread) {
    llama_model_quantize_params qp = llama_model_quantize_default_params();
    qp.ftype = (enum llama_ftype) ftype;
    qp.nthread = nthread;
    return falcon_model_quantize(in, out, &qp);
}

// BASELINE config 1: Y = W x for n_mats rotating Q4_0 weight matrices [M rows][K] and ONE f32 activation column, through the
// reference's own ggml_mul_mat + ggml_graph_compute -- the graph examples/benchmark/benchmark-matmult.cpp:181-202 builds, with the
// shape as a parameter and the caller's blocks instead of ggml_quantize_q4_0(1.0f) (its sum check is numerically broken).
// `blocks` holds n_mats * M * K/32 block_q4_0 (18 B each).  Every iteration computes all n_mats graphs once, so consecutive calls
// never reuse a matrix out of cache when n_mats * M * K * 0.5625 B exceeds the LLC.  Returns the mean microseconds per mat-vec;
// *best_us = the fastest single call; y_out (M floats) = the product with matrix 0.
double refh_matvec_bench(int K, int M, int n_mats, int iters, int n_threads, const void * blocks, const float * x, float * y_out, double * best_us) {
    ggml_time_init();
    const size_t wbytes = (size_t) M * (K / 32) * 18;
    struct ggml_init_params ip = { (size_t) n_mats * (wbytes + (size_t) M * 4 + 4096) + (size_t) K * 4 + (64u << 20), NULL, false };
    struct ggml_context * ctx = ggml_init(ip);
    if (!ctx) return -1.0;
    struct ggml_tensor * xt = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, 1);
    memcpy(xt->data, x, (size_t) K * 4);
    std::vector<struct ggml_cgraph> graphs(n_mats);
    for (int i = 0; i < n_mats; i++) {
        struct ggml_tensor * w = ggml_new_tensor_2d(ctx, GGML_TYPE_Q4_0, K, M);
        memcpy(w->data, (const char *) blocks + (size_t) i * wbytes, wbytes);
        graphs[i] = ggml_build_forward(ggml_mul_mat(ctx, w, xt));
        graphs[i].n_threads = n_threads;
    }
    for (int i = 0; i < n_mats; i++) ggml_graph_compute(ctx, &graphs[i]);      // warm-up: thread pool, work buffer
    memcpy(y_out, graphs[0].nodes[graphs[0].n_nodes - 1]->data, (size_t) M * 4);
    double total = 0.0, best = 1e30;
    for (int it = 0; it < iters; it++)
        for (int i = 0; i < n_mats; i++) {
            const int64_t t0 = ggml_time_us();
            ggml_graph_compute(ctx, &graphs[i]);
            const double us = (double) (ggml_time_us() - t0);
            total += us; if (us < best) best = us;
        }
    if (best_us) *best_us = best;
    ggml_free(ctx);
    return total / ((double) iters * n_mats);
}

// falcon_main's default sampling chain over a caller-provided logits row (examples/falcon/falcon_main.cpp:945-975): repetition penalty
// over last_tokens, then greedy (temp <= 0) or top-k -> top-p -> temperature -> llama_sample_token, which draws from the context's
// std::mt19937 (reseed with refh_set_seed).  Only calls the reference's public llama_sample_* functions, in falcon_main's order.
void refh_set_seed(void * h, int seed) { llama_set_rng_seed((falcon_context *) h, seed); }
int refh_sample(void * h, const float * logits, int n_vocab, const int * last_tokens, int n_last, int top_k, float top_p, float temp, float repeat_penalty) {
    falcon_context * ctx = (falcon_context *) h;
    std::vector<falcon_token_data> candidates;
    candidates.reserve(n_vocab);
    for (falcon_token id = 0; id < n_vocab; id++) candidates.emplace_back(falcon_token_data{ id, logits[id], 0.0f });
    falcon_token_data_array cp = { candidates.data(), candidates.size(), false };
    llama_sample_repetition_penalty(ctx, &cp, last_tokens, (size_t) n_last, repeat_penalty);
    if (temp <= 0) return llama_sample_token_greedy(ctx, &cp);
    llama_sample_top_k(ctx, &cp, top_k <= 0 ? n_vocab : top_k, 1);
    llama_sample_top_p(ctx, &cp, top_p, 1);
    llama_sample_temperature(ctx, &cp, temp);
    return llama_sample_token(ctx, &cp);
}

// the reference's own timing table (libfalcon.cpp:4700-4714)
void refh_print_timings(void * h) { falcon_print_timings((falcon_context *) h); }

} // extern "C"
