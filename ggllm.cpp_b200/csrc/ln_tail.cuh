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

// Every load of a phase is issued before the first use (fully unrolled, CH chunks of 8 values per thread), so the tail
// costs three dependent L2 round trips, not three per loop iteration.  red: NT / 32 doubles of shared memory.
// n == CH * NT * 8 or less (whole warps drop out together), n % 256 == 0.
template <int NT, int CH>
__device__ __forceinline__ void ln_tail_body(const LnTail & t, const float * rb, double * red) {
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
    float v[CH][8];
    bool ok[CH];
    float4 la[CH][2], lc[CH][2], lp[CH][2];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int e = (c * NT + tid) * 8;
        ok[c] = e < t.n;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (ok[c]) { la[c][h] = __ldcg(reinterpret_cast<const float4 *>(t.ra + e) + h); lc[c][h] = __ldcg(reinterpret_cast<const float4 *>(rb + e) + h);
                         lp[c][h] = __ldcg(reinterpret_cast<const float4 *>(t.x + e) + h); }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int e = (c * NT + tid) * 8;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[c]) {
                const float4 a = la[c][h], cc = lc[c][h]; p = lp[c][h];
                p.x = __fadd_rn(__fadd_rn(a.x, cc.x), p.x); p.y = __fadd_rn(__fadd_rn(a.y, cc.y), p.y); p.z = __fadd_rn(__fadd_rn(a.z, cc.z), p.z); p.w = __fadd_rn(__fadd_rn(a.w, cc.w), p.w);
                *(reinterpret_cast<float4 *>(t.x + e) + h) = p;
                s += (double) p.x; s += (double) p.y; s += (double) p.z; s += (double) p.w;
            }
            v[c][4 * h] = p.x; v[c][4 * h + 1] = p.y; v[c][4 * h + 2] = p.z; v[c][4 * h + 3] = p.w;
        }
    }
    const float mean = (float) (total(s) / t.n);
    double s2 = 0.0;
#pragma unroll
    for (int c = 0; c < CH; c++) if (ok[c]) {
#pragma unroll
        for (int i = 0; i < 8; i++) { v[c][i] = __fsub_rn(v[c][i], mean); s2 += (double) __fmul_rn(v[c][i], v[c][i]); }
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float) (total(s2) / t.n), 1e-5f)));
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !t.has2) break;
        const float * g = pass ? t.g2 : t.g1, * b = pass ? t.b2 : t.b1;
        const ActQ & A = pass ? t.A2 : t.A1;
        float4 lg[CH][2], lb[CH][2];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int e = (c * NT + tid) * 8;
#pragma unroll
            for (int h = 0; h < 2; h++) if (ok[c]) { lg[c][h] = __ldg(reinterpret_cast<const float4 *>(g + e) + h); lb[c][h] = __ldg(reinterpret_cast<const float4 *>(b + e) + h); }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int e = (c * NT + tid) * 8;
            if (!ok[c]) continue;                          // whole warps drop out together (n % 256 == 0)
            const float gg[8] = { lg[c][0].x, lg[c][0].y, lg[c][0].z, lg[c][0].w, lg[c][1].x, lg[c][1].y, lg[c][1].z, lg[c][1].w };
            const float bv[8] = { lb[c][0].x, lb[c][0].y, lb[c][0].z, lb[c][0].w, lb[c][1].x, lb[c][1].y, lb[c][1].z, lb[c][1].w };
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = __fadd_rn(__fmul_rn(__fmul_rn(v[c][i], scale), gg[i]), bv[i]);
            if (A.type == T_Q8_K) quantize_chunk8<T_Q8_K>(y, lane, A, 0, e);
            else if (A.type == T_Q8_1) quantize_chunk8<T_Q8_1>(y, lane, A, 0, e);
            else quantize_chunk8<T_Q8_0>(y, lane, A, 0, e);
        }
    }
}
template <int NT>
__device__ __forceinline__ bool ln_tail_run(const LnTail & t, const float * rb, double * red) {
    const int ch = (t.n + NT * 8 - 1) / (NT * 8);
    if (ch <= 1) ln_tail_body<NT, 1>(t, rb, red);
    else if (ch <= 2) ln_tail_body<NT, 2>(t, rb, red);
    else if (ch <= 4) ln_tail_body<NT, 4>(t, rb, red);
    else if (ch <= 8) ln_tail_body<NT, 8>(t, rb, red);
    else return false;
    return true;
}
// gamma / beta of the coming LayerNorm do not depend on anything: one CTA pulls them into L2 when the kernel starts
__device__ __forceinline__ void ln_tail_prefetch(const LnTail & t) {
    for (int e = threadIdx.x * 32; e < t.n; e += blockDim.x * 32) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(t.g1 + e)); asm volatile("prefetch.global.L2 [%0];" :: "l"(t.b1 + e));
        if (t.has2) { asm volatile("prefetch.global.L2 [%0];" :: "l"(t.g2 + e)); asm volatile("prefetch.global.L2 [%0];" :: "l"(t.b2 + e)); }
    }
}
