// attention.cu -- multi-query / grouped-query attention over the f32 KV cache.
//
// Restates, as ONE kernel per layer, the seven graph nodes the reference pins to the CPU
// (libfalcon.cpp:2285-2366: K view/permute, mul_mat(K,Q) with cuda_op_directive=0, scale, diag_mask_inf, soft_max,
// mul_mat(V,P), permute+cpy) with the CPU numerics:
//   s[p]  = (q . k_p) * (1/sqrt(head_dim))                           ggml.c:10911-11102, libfalcon.cpp:2313-2317
//   mask  : p > n_past + t  ->  -inf                                 ggml.c:12342-12348
//   e[p]  = f16->f32(LUT_exp[f32->f16(s[p] - max)]), sum in double, y = e * (float)(1/sum)     ggml.c:12427-12449
//   out   = sum_p V[p] * y[p]
// GQA head -> kv head map is the f32 mat-mul's  h / (n_head / n_head_kv)  (ggml.c:11074).
//
// KV cache layout on the device: K and V both [n_ctx][n_head_kv][head_dim] f32 per layer (the reference's K layout,
// libfalcon.cpp:2238-2242; its transposed, ping-ponged V copy with the O(n_past) re-copy per token,
// libfalcon.cpp:2256-2281, is replaced by a plain append -- values are identical).
//
// v1 kernel: one CTA per (query head, query token); scores live in shared memory (T floats).  Exact oracle
// semantics (global max before exp).  Fine for decode; the prompt path gets a tiled tensor-core kernel later.
#include "kernels.h"

#define ATT_THREADS 128

__device__ __forceinline__ float exp_f16lut(float v) {      // table_exp_f16[f16(v)], ggml.c:4281-4290
    return __half2float(__float2half_rn(expf(__half2float(__float2half_rn(v)))));
}

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const float * __restrict__ qkv, const float * __restrict__ kc, const float * __restrict__ vc,
                                                               float * __restrict__ out, int64_t out_stride, AttnParams p) {
    extern __shared__ __align__(16) float sm[];
    trace_begin(p.trace);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // wo's mat-vec may start prefetching its weights
    const int h = blockIdx.x, t = blockIdx.y, D = p.head_dim;
    const int n_past = p.n_past_dev ? *p.n_past_dev : p.n_past;
    const int T = n_past + t + 1;                            // keys visible to this query (causal)
    const int kvh = h / (p.n_head / p.n_head_kv);
    float * q = sm;                                          // [D]
    float * s = sm + D;                                      // [T]
    __shared__ float red_f[ATT_THREADS / 32];
    __shared__ double red_d[ATT_THREADS / 32];
    __shared__ float part[ATT_THREADS];

    for (int i = threadIdx.x; i < D; i += ATT_THREADS) q[i] = qkv[(size_t) t * p.qkv_stride + (size_t) h * D + i];
    __syncthreads();
    const float scale = 1.0f / sqrtf((float) D);
    const size_t kv_row = (size_t) p.n_head_kv * D;

    // scores: one thread per key, 16-byte loads along the head dimension
    float lmax = -INFINITY;
    for (int k = threadIdx.x; k < T; k += ATT_THREADS) {
        const float4 * kr = reinterpret_cast<const float4 *>(kc + (size_t) k * kv_row + (size_t) kvh * D);
        float acc = 0.f;
        for (int i = 0; i < D / 4; i++) {
            const float4 kv = kr[i];
            acc += kv.x * q[4 * i] + kv.y * q[4 * i + 1] + kv.z * q[4 * i + 2] + kv.w * q[4 * i + 3];
        }
        acc = __fmul_rn(acc, scale);
        s[k] = acc;
        lmax = fmaxf(lmax, acc);
    }
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) red_f[threadIdx.x >> 5] = lmax;
    __syncthreads();
    float gmax = red_f[0];
    for (int w = 1; w < ATT_THREADS / 32; w++) gmax = fmaxf(gmax, red_f[w]);

    double lsum = 0.0;
    for (int k = threadIdx.x; k < T; k += ATT_THREADS) { const float e = exp_f16lut(__fsub_rn(s[k], gmax)); s[k] = e; lsum += (double) e; }
    lsum = warp_sum_d(lsum);
    if ((threadIdx.x & 31) == 0) red_d[threadIdx.x >> 5] = lsum;
    __syncthreads();
    double gsum = 0.0;
    for (int w = 0; w < ATT_THREADS / 32; w++) gsum += red_d[w];
    const float inv = (float) (1.0 / gsum);

    // out[i] = sum_k V[k][i] * (e[k] * inv): thread = (key parity group, dim) so that a warp reads 128 contiguous bytes of V
    const int i = threadIdx.x % D, grp = threadIdx.x / D, ngrp = ATT_THREADS / D;     // D = 64 -> 2 groups
    float acc = 0.f;
    for (int k = grp; k < T; k += ngrp) acc += vc[(size_t) k * kv_row + (size_t) kvh * D + i] * __fmul_rn(s[k], inv);
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < D) {
        float r = part[threadIdx.x];
        for (int g = 1; g < ngrp; g++) r += part[g * D + threadIdx.x];
        out[(size_t) t * out_stride + (size_t) h * D + threadIdx.x] = r;
    }
    trace_end(p.trace);
}

size_t attention_scratch_bytes(const AttnParams &) { return 0; }

void launch_attention(const float * qkv, const float * k_cache, const float * v_cache, float * out, int64_t out_stride,
                      const AttnParams & p, float *, cudaStream_t stream) {
    if (p.n_tok <= 0) return;
    B200_ASSERT(p.head_dim % 4 == 0 && ATT_THREADS % p.head_dim == 0);
    // shared memory is sized for the worst case so that a captured graph stays valid while n_past grows
    const int t_max = p.n_past_dev ? p.n_ctx : p.n_past + p.n_tok;
    const size_t smem = (size_t) (p.head_dim + t_max) * sizeof(float);
    static bool set = false;
    if (!set) { B200_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        B200_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, B200_CARVEOUT)); set = true; }
    B200_ASSERT(smem <= 200 * 1024);
    dim3 grid((unsigned) p.n_head, (unsigned) p.n_tok);
    AttnParams pt = p; pt.trace = b200_trace_slot("attention");
    attention_kernel<<<grid, ATT_THREADS, smem, stream>>>(qkv, k_cache, v_cache, out, out_stride, pt);
    B200_CUDA_CHECK(cudaGetLastError());
}
