// ln_tail.cuh -- the residual adds that close a layer + the next LayerNorm(s) + activation quantisation, run by the LAST
// CTA of the mat-vec that produces the final addend (wo), instead of by a kernel of its own: the row is 8192 values, so
// the stand-alone kernel (ops.cu layernorm_q_reg_kernel) is pure launch + dependent-load latency (7 us between the end of
// wo and the first qkv row, profiles/r1_decode_timeline.md).  Same arithmetic as that kernel, value for value:
//   x = (ra + rb) + x (libfalcon.cpp:2399-2400);  mean, variance in double (ggml.c:10568-10595);
//   y = ((x - mean) * scale) * gamma + beta (libfalcon.cpp:2166-2185);  Q8 quantisation (ggml.c:11462-11476)
#pragma once
#include "actquant.cuh"

struct LnTail {
    float * x; const float * ra;                 // x updated in place; rb is the calling kernel's own output row
    const float * g1, * b1, * g2, * b2; ActQ A1, A2; int has2;
    int n; unsigned * ctr;                       // ctr: zero-initialised arrival counter (re-armed by the last CTA); nullptr = no tail
};

// sv: shared memory for n floats, red: NT / 32 doubles.  n % 256 == 0.
template <int NT>
__device__ __forceinline__ void ln_tail_run(const LnTail & t, const float * rb, float * sv, double * red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto total = [&](double v) -> double {
        v = warp_sum_d(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) r += red[w];
        return r;
    };
    double s = 0.0;
    for (int e = tid * 4; e < t.n; e += NT * 4) {
        const float4 a = __ldcg(reinterpret_cast<const float4 *>(t.ra + e)), c = __ldcg(reinterpret_cast<const float4 *>(rb + e));
        float4 p = __ldcg(reinterpret_cast<const float4 *>(t.x + e));
        p.x = __fadd_rn(__fadd_rn(a.x, c.x), p.x); p.y = __fadd_rn(__fadd_rn(a.y, c.y), p.y); p.z = __fadd_rn(__fadd_rn(a.z, c.z), p.z); p.w = __fadd_rn(__fadd_rn(a.w, c.w), p.w);
        *reinterpret_cast<float4 *>(t.x + e) = p;
        *reinterpret_cast<float4 *>(sv + e) = p;
        s += (double) p.x; s += (double) p.y; s += (double) p.z; s += (double) p.w;
    }
    const float mean = (float) (total(s) / t.n);
    double s2 = 0.0;
    for (int e = tid * 4; e < t.n; e += NT * 4) {
        float4 p = *reinterpret_cast<float4 *>(sv + e);
        p.x = __fsub_rn(p.x, mean); p.y = __fsub_rn(p.y, mean); p.z = __fsub_rn(p.z, mean); p.w = __fsub_rn(p.w, mean);
        *reinterpret_cast<float4 *>(sv + e) = p;
        s2 += (double) __fmul_rn(p.x, p.x); s2 += (double) __fmul_rn(p.y, p.y); s2 += (double) __fmul_rn(p.z, p.z); s2 += (double) __fmul_rn(p.w, p.w);
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float) (total(s2) / t.n), 1e-5f)));
    __syncthreads();
    for (int e = tid * 8; e < t.n; e += NT * 8) {                 // a warp covers 256 consecutive values: whole quantisation blocks
        float v[8], y[8];
        const float4 p = *reinterpret_cast<float4 *>(sv + e), q = *reinterpret_cast<float4 *>(sv + e + 4);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && !t.has2) break;
            const float * g = pass ? t.g2 : t.g1, * b = pass ? t.b2 : t.b1;
            const float4 ga = __ldg(reinterpret_cast<const float4 *>(g + e)), gb = __ldg(reinterpret_cast<const float4 *>(g + e + 4));
            const float4 ba = __ldg(reinterpret_cast<const float4 *>(b + e)), bb = __ldg(reinterpret_cast<const float4 *>(b + e + 4));
            const float gg[8] = { ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w }, bv[8] = { ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w };
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = __fadd_rn(__fmul_rn(__fmul_rn(v[i], scale), gg[i]), bv[i]);
            const ActQ & A = pass ? t.A2 : t.A1;
            if (A.type == T_Q8_K) quantize_chunk8<T_Q8_K>(y, lane, A, 0, e);
            else if (A.type == T_Q8_1) quantize_chunk8<T_Q8_1>(y, lane, A, 0, e);
            else quantize_chunk8<T_Q8_0>(y, lane, A, 0, e);
        }
    }
}
