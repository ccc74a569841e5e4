// decode_mega.cu -- all transformer layers of ONE decode step (n_tokens == 1) as a single persistent kernel.
//
// Why: with one kernel per graph node the HBM stream stops at every kernel boundary (tail of the old grid, launch
// and pipeline ramp of the new one, then the activation prologue): ~4 us x 5 boundaries per layer on a layer that
// needs 59 us of pure weight streaming (profiles/r1_decode_timeline.md).  Here one CTA per SM stays resident for the
// whole step, every CTA owns a fixed slice of the rows of every matrix, and the only things that order CTAs are four
// counters per layer in global memory.  The register ring of the NEXT matrix is always filled before a CTA waits on a
// counter, so the weights keep flowing while activations are exchanged through L2.
//
// Per layer (libfalcon.cpp:2143-2400; numerics identical to the stand-alone kernels, whose device code is shared):
//   P1  [two 256-thread halves, rows interleaved]   wait D(l-1); LayerNorm(s) of the residual row + Q8 quantisation in
//       registers (ops.cu layernorm_q semantics); qkv rows, arrive A; ffn_up rows (+GELU LUT), arrive B.  The ring
//       runs through the qkv -> ffn_up switch without draining (same K).
//   P2  [whole CTA = one query head]                wait A; RoPE(q), RoPE(k_new), K/V append; exact-max LUT softmax
//       attention over the f32 cache (attention.cu semantics); arrive C
//   P3  [whole CTA]                                 ffn_down ring was filled before P2; wait B; quantise gelu(up);
//       ffn_down rows -> kept in shared memory
//   P4  [two halves]                                wait C (long satisfied); quantise the attention row; wo rows;
//       epilogue x[row] = (down[row] + wo[row]) + x[row]  (libfalcon.cpp:2399-2400) by the CTA that owns the row in
//       both matrices; arrive D
// Only D is a barrier every CTA really waits on; A, B and C are normally satisfied by the time they are read.
//
// Activations that other CTAs produced inside this kernel are read with ld.global.cg (L2; L1 is not coherent).
// The launch is cooperative (co-residency of all CTAs is required by the counters).
#include "kernels.h"
#include "mmv_fast.cuh"

#define MG_NT 512
#define MG_HALF 256

struct MegaLayer { WPlanes qkv, up, down, wo; const float * ga, * ba, * gm, * bm; float * kc, * vc; };
struct MegaArgs {
    const MegaLayer * layers; int n_layer;
    float * x, * qkv, * up, * att;             // residual row [E] (updated in place), fp32 scratch rows
    unsigned * flags;                          // [n_layer][4] = A, B, C, D arrival counters, zeroed before the launch
    const int * n_past_dev; int n_past, n_ctx;
    int E, FF, H, HKV, D, dual;
    float theta_scale;
    unsigned long long * trace;                // 5 slots per layer (ln, qkv_up, attn, down, wo) or nullptr
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned * p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_inc(unsigned * p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(p) : "memory"); }
__device__ __forceinline__ float4 ld_cg4(const float * p) { float4 r; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ float ld_cg(const float * p) { float r; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory"); return r; }

// the group (whole CTA: id 0 / 512 threads, a half: id 1 + half / 256 threads) has finished its stores -> one arrival
__device__ __forceinline__ void group_arrive(unsigned * flag, int bar_id, int nthr, int tg) {
    named_sync(bar_id, nthr);
    if (tg == 0) { __threadfence(); red_release_inc(flag); }
}
__device__ __forceinline__ void group_wait(const unsigned * flag, unsigned target, int bar_id, int nthr, int tg) {
    if (tg == 0) { while (ld_acquire_u32(flag) < target) { } __threadfence(); }
    named_sync(bar_id, nthr);
}
__device__ __forceinline__ void mg_trace(unsigned long long * base, int idx, bool end) {
    if (base && threadIdx.x == 0) { if (end) atomicMax(base + 2 * idx + 1, gtimer()); else atomicMin(base + 2 * idx, gtimer()); }
}

__device__ __forceinline__ float mg_exp_lut(float v) { return __half2float(__float2half_rn(expf(__half2float(__float2half_rn(v))))); }   // ggml.c:4281-4290
// pair i of a head at position p (ops.cu rope_pair): theta = p * theta_scale^i by repeated fp32 multiplication
__device__ __forceinline__ void mg_rope(const float x0, const float x1, int i, int p, float theta_scale, float & y0, float & y1) {
    float theta = (float) p;
    for (int k = 0; k < i; k++) theta = __fmul_rn(theta, theta_scale);
    const float c = cosf(theta), s = sinf(theta);
    y0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
    y1 = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
}

// 32 fp32 values of piece g from a row other CTAs wrote (L2 loads)
template <int TYPE>
__device__ __forceinline__ void load_piece_cg(const float * row, int g, bool valid, float (&v)[32]) {
    int ea, eb; FX<TYPE>::seg(g, ea, eb);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (valid) { a = ld_cg4(row + ea + 4 * i); b = ld_cg4(row + eb + 4 * i); }
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
        v[16 + 4 * i] = b.x; v[17 + 4 * i] = b.y; v[18 + 4 * i] = b.z; v[19 + 4 * i] = b.w;
    }
}
// y = v * gamma + beta on the piece's two segments, then the mat-mul's activation quantisation
template <int TYPE>
__device__ __forceinline__ typename FX<TYPE>::XR affine_quant(const float (&v)[32], const float * gamma, const float * beta, int g, bool valid, int lane) {
    int ea, eb; FX<TYPE>::seg(g, ea, eb);
    float y[32];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, ba = ga, bb = ga;
        if (valid) {
            ga = __ldg(reinterpret_cast<const float4 *>(gamma + ea) + i); gb = __ldg(reinterpret_cast<const float4 *>(gamma + eb) + i);
            ba = __ldg(reinterpret_cast<const float4 *>(beta + ea) + i);  bb = __ldg(reinterpret_cast<const float4 *>(beta + eb) + i);
        }
        y[4 * i]      = __fadd_rn(__fmul_rn(v[4 * i], ga.x), ba.x);      y[4 * i + 1]  = __fadd_rn(__fmul_rn(v[4 * i + 1], ga.y), ba.y);
        y[4 * i + 2]  = __fadd_rn(__fmul_rn(v[4 * i + 2], ga.z), ba.z);  y[4 * i + 3]  = __fadd_rn(__fmul_rn(v[4 * i + 3], ga.w), ba.w);
        y[16 + 4 * i] = __fadd_rn(__fmul_rn(v[16 + 4 * i], gb.x), bb.x); y[17 + 4 * i] = __fadd_rn(__fmul_rn(v[17 + 4 * i], gb.y), bb.y);
        y[18 + 4 * i] = __fadd_rn(__fmul_rn(v[18 + 4 * i], gb.z), bb.z); y[19 + 4 * i] = __fadd_rn(__fmul_rn(v[19 + 4 * i], gb.w), bb.w);
    }
    if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; i++) y[i] = 0.f;
    }
    return FX<TYPE>::quant_x(y, g, lane);
}

__device__ __forceinline__ void split_rows(int M, int part, int nparts, int & r0, int & r1) {       // balanced to +-1
    const int per = M / nparts, rem = M % nparts;
    r0 = part * per + min(part, rem); r1 = r0 + per + (part < rem ? 1 : 0);
}

// shared memory map (bytes): see mega_smem_bytes()
struct MegaSmem {
    MegaLayer layer[2];
    float part_half[2][2 * 8 * 4];      // per half: [2 buffers][8 warps][G]
    float part_full[2 * 16 * 4];
    double red1[2][8], red2[2][8];      // LayerNorm reductions per half
    float red_f[16]; double red_d[16];  // attention reductions
    float pv[MG_NT];
    float q[128], k[128], v[128];       // roped q, roped new k, new v of this CTA's head
    float dn[128];                      // this CTA's ffn_down rows (E / gridDim <= 128)
};

template <int TYPE, int JD>
__global__ void __launch_bounds__(MG_NT, 1) falcon_decode_mega_kernel(const MegaArgs a) {
    using T = FX<TYPE>;
    constexpr int DE = 8, DD = 8 / JD;                  // ring depth: D * J = 8 pieces in flight per thread
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MegaSmem & S = *reinterpret_cast<MegaSmem *>(smem_raw);
    float * scores = reinterpret_cast<float *>(smem_raw + sizeof(MegaSmem));       // [n_ctx]

    const int tid = threadIdx.x, lane = tid & 31, half = tid >> 8, th = tid & (MG_HALF - 1);
    const int cta = blockIdx.x, ncta = gridDim.x, vcta = 2 * cta + half, nv = 2 * ncta;
    const int hbar = 1 + half;
    const int n_past = a.n_past_dev ? *a.n_past_dev : a.n_past;

    // layer descriptors travel through shared memory one layer ahead, so no phase starts with a dependent global load
    auto fetch_layer = [&](int l) {
        if (l < a.n_layer) {
            const uint32_t * src = reinterpret_cast<const uint32_t *>(a.layers + l);
            uint32_t * dst = reinterpret_cast<uint32_t *>(&S.layer[l & 1]);
            for (int i = tid; i < (int) (sizeof(MegaLayer) / 4); i += MG_NT) dst[i] = __ldg(src + i);
        }
    };
    fetch_layer(0);
    __syncthreads();

    const int PE = S.layer[0].qkv.nb * T::PPB, PF = S.layer[0].down.nb * T::PPB;
    int ge[1]; bool ve[1];                               // piece of the K = E matrices (one per thread of a half)
    ve[0] = th < PE; ge[0] = ve[0] ? th : PE - 1;
    int gd[JD]; bool vd[JD];                             // pieces of ffn_down (whole CTA)
#pragma unroll
    for (int j = 0; j < JD; j++) { const int g = j * MG_NT + tid; vd[j] = g < PF; gd[j] = vd[j] ? g : PF - 1; }

    const int aux0[1] = { 0 }; int auxd[JD];
#pragma unroll
    for (int j = 0; j < JD; j++) auxd[j] = 0;
    int gc_half = 0, gc_full = 0;
    int d0, d1; split_rows(a.E, cta, ncta, d0, d1);      // rows of ffn_down AND wo this CTA owns
    const int dh = d0 + (d1 - d0 + 1) / 2;
    const int o0 = half ? dh : d0, o1 = half ? d1 : dh;  // the half's wo rows

    typename T::WR we[DE][1];
    typename T::WR wd[DD][JD];

    for (int l = 0; l < a.n_layer; l++) {
        const MegaLayer & L = S.layer[l & 1];
        unsigned * flag = a.flags + 4 * l;
        unsigned long long * tr = a.trace ? a.trace + 2 * 5 * l : nullptr;
        int q0, q1, u0, u1;
        split_rows(L.qkv.M, vcta, nv, q0, q1);
        split_rows(L.up.M, vcta, nv, u0, u1);

        // ------------------------------------------------------------------ P1: LayerNorm -> qkv -> ffn_up
        mg_trace(tr, 0, false);
        WP pq[1], pu[1];
        pq[0] = T::wp(L.qkv, ge[0]); pu[0] = T::wp(L.up, ge[0]);
        if (q1 > q0) ring_fill<TYPE, 1, DE>(we, pq, q0, q1); else ring_fill<TYPE, 1, DE>(we, pu, u0, u1);
        if (l > 0) group_wait(a.flags + 4 * (l - 1) + 3, (unsigned) nv, hbar, MG_HALF, th);
        typename T::XR xa[1], xm[1];
        {
            float v[32];
            load_piece_cg<TYPE>(a.x, ge[0], ve[0], v);
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i++) s += (double) v[i];                  // invalid pieces hold zeros
            s = warp_sum_d(s);
            if (lane == 0) S.red1[half][th >> 5] = s;
            named_sync(hbar, MG_HALF);
            double t = 0.0;
#pragma unroll
            for (int wi = 0; wi < 8; wi++) t += S.red1[half][wi];
            const float mean = (float) (t / a.E);
            double s2 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i++) { v[i] = __fsub_rn(v[i], mean); s2 += ve[0] ? (double) __fmul_rn(v[i], v[i]) : 0.0; }
            s2 = warp_sum_d(s2);
            if (lane == 0) S.red2[half][th >> 5] = s2;
            named_sync(hbar, MG_HALF);
            t = 0.0;
#pragma unroll
            for (int wi = 0; wi < 8; wi++) t += S.red2[half][wi];
            const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float) (t / a.E), 1e-5f)));     // ggml.c:10568-10595
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __fmul_rn(v[i], scale);
            xm[0] = affine_quant<TYPE>(v, L.gm, L.bm, ge[0], ve[0], lane);
            if (a.dual) xa[0] = affine_quant<TYPE>(v, L.ga, L.ba, ge[0], ve[0], lane); else xa[0] = xm[0];
        }
        mg_trace(tr, 0, true); mg_trace(tr, 1, false);
        {
            float * yq = a.qkv;
            ring_run<TYPE, MG_HALF, 1, DE, true>(we, pq, q0, q1, pu, u0, u1, xa, aux0, S.part_half[half], gc_half, hbar, th,
                                                 [&](int row, float v) { yq[row] = v; }, [] {});
            group_arrive(flag + 0, hbar, MG_HALF, th);
            float * yu = a.up;
            ring_run<TYPE, MG_HALF, 1, DE, false>(we, pu, u0, u1, pu, 0, 0, xm, aux0, S.part_half[half], gc_half, hbar, th,
                                                  [&](int row, float v) { yu[row] = gelu_lut(v); }, [] {});
            group_arrive(flag + 1, hbar, MG_HALF, th);
        }
        mg_trace(tr, 1, true);

        // ------------------------------------------------------------------ P2: attention (ffn_down's ring is filled first)
        WP pd[JD];
#pragma unroll
        for (int j = 0; j < JD; j++) pd[j] = T::wp(L.down, gd[j]);
        ring_fill<TYPE, JD, DD>(wd, pd, d0, d1);
        mg_trace(tr, 2, false);
        group_wait(flag + 0, (unsigned) nv, 0, MG_NT, tid);
        fetch_layer(l + 1);                          // every thread is past layer l-1 (whose slot this overwrites); visible long before it is used
        {
            const int Dh = a.D, hd2 = Dh / 2, T_ = n_past + 1;
            const int grp_heads = a.H / a.HKV;
            for (int h = cta; h < a.H; h += ncta) {
                const int kvh = h / grp_heads;                                 // ggml.c:11074
                const float * qsrc = a.qkv + (size_t) h * Dh, * ksrc = a.qkv + (size_t) (a.H + kvh) * Dh, * vsrc = a.qkv + (size_t) (a.H + a.HKV + kvh) * Dh;
                if (tid < hd2) { mg_rope(ld_cg(qsrc + tid), ld_cg(qsrc + tid + hd2), tid, n_past, a.theta_scale, S.q[tid], S.q[tid + hd2]); }
                else if (tid >= 64 && tid < 64 + hd2) { const int i = tid - 64; mg_rope(ld_cg(ksrc + i), ld_cg(ksrc + i + hd2), i, n_past, a.theta_scale, S.k[i], S.k[i + hd2]); }
                else if (tid >= 128 && tid < 128 + Dh) { S.v[tid - 128] = ld_cg(vsrc + tid - 128); }
                __syncthreads();
                const size_t kv_row = (size_t) a.HKV * Dh;
                if (h % grp_heads == 0 && tid < Dh) {                          // K / V append, libfalcon.cpp:2238-2281
                    L.kc[(size_t) n_past * kv_row + (size_t) kvh * Dh + tid] = S.k[tid];
                    L.vc[(size_t) n_past * kv_row + (size_t) kvh * Dh + tid] = S.v[tid];
                }
                const float scale = 1.0f / sqrtf((float) Dh);
                float lmax = -INFINITY;
                for (int k = tid; k < T_; k += MG_NT) {
                    const float4 * kr = k == n_past ? reinterpret_cast<const float4 *>(S.k) : reinterpret_cast<const float4 *>(L.kc + (size_t) k * kv_row + (size_t) kvh * Dh);
                    float acc = 0.f;
                    for (int i = 0; i < Dh / 4; i++) {
                        const float4 kv = kr[i];
                        acc += kv.x * S.q[4 * i] + kv.y * S.q[4 * i + 1] + kv.z * S.q[4 * i + 2] + kv.w * S.q[4 * i + 3];
                    }
                    acc = __fmul_rn(acc, scale);
                    scores[k] = acc;
                    lmax = fmaxf(lmax, acc);
                }
                lmax = warp_max(lmax);
                if (lane == 0) S.red_f[tid >> 5] = lmax;
                __syncthreads();
                float gmax = S.red_f[0];
                for (int w = 1; w < MG_NT / 32; w++) gmax = fmaxf(gmax, S.red_f[w]);
                double lsum = 0.0;
                for (int k = tid; k < T_; k += MG_NT) { const float e = mg_exp_lut(__fsub_rn(scores[k], gmax)); scores[k] = e; lsum += (double) e; }
                lsum = warp_sum_d(lsum);
                if (lane == 0) S.red_d[tid >> 5] = lsum;
                __syncthreads();
                double gsum = 0.0;
                for (int w = 0; w < MG_NT / 32; w++) gsum += S.red_d[w];
                const float inv = (float) (1.0 / gsum);                        // ggml.c:12427-12449
                const int i = tid % Dh, grp = tid / Dh, ngrp = MG_NT / Dh;
                float acc = 0.f;
                for (int k = grp; k < T_; k += ngrp) {
                    const float vv = k == n_past ? S.v[i] : L.vc[(size_t) k * kv_row + (size_t) kvh * Dh + i];
                    acc += vv * __fmul_rn(scores[k], inv);
                }
                S.pv[tid] = acc;
                __syncthreads();
                if (tid < Dh) {
                    float r = S.pv[tid];
                    for (int g = 1; g < ngrp; g++) r += S.pv[g * Dh + tid];
                    a.att[(size_t) h * Dh + tid] = r;
                }
                __syncthreads();
            }
        }
        group_arrive(flag + 2, 0, MG_NT, tid);
        mg_trace(tr, 2, true);

        // ------------------------------------------------------------------ P3: ffn_down
        mg_trace(tr, 3, false);
        group_wait(flag + 1, (unsigned) nv, 0, MG_NT, tid);
        {
            typename T::XR xd[JD];
#pragma unroll
            for (int j = 0; j < JD; j++) {
                float v[32];
                load_piece_cg<TYPE>(a.up, gd[j], vd[j], v);
                xd[j] = T::quant_x(v, gd[j], lane);
            }
            float * dn = S.dn;
            ring_run<TYPE, MG_NT, JD, DD, false>(wd, pd, d0, d1, pd, 0, 0, xd, auxd, S.part_full, gc_full, 0, tid,
                                                 [&](int row, float v) { dn[row - d0] = v; }, [] {});
        }
        mg_trace(tr, 3, true);

        // ------------------------------------------------------------------ P4: wo + residual
        mg_trace(tr, 4, false);
        WP po[1]; po[0] = T::wp(L.wo, ge[0]);
        ring_fill<TYPE, 1, DE>(we, po, o0, o1);
        __syncthreads();                                                       // S.dn complete for both halves
        group_wait(flag + 2, (unsigned) ncta, hbar, MG_HALF, th);
        {
            typename T::XR xo[1];
            float v[32];
            load_piece_cg<TYPE>(a.att, ge[0], ve[0], v);
            xo[0] = T::quant_x(v, ge[0], lane);
            float * x = a.x; const float * dn = S.dn;
            ring_run<TYPE, MG_HALF, 1, DE, false>(we, po, o0, o1, po, 0, 0, xo, aux0, S.part_half[half], gc_half, hbar, th,
                                                  [&](int row, float v) { x[row] = __fadd_rn(__fadd_rn(dn[row - d0], v), ld_cg(x + row)); }, [] {});
        }
        group_arrive(flag + 3, hbar, MG_HALF, th);
        mg_trace(tr, 4, true);
    }
}

static size_t mega_smem_bytes(int n_ctx) { return sizeof(MegaSmem) + (size_t) n_ctx * 4; }

template <int TYPE, int JD>
static void launch_mega_t(const MegaArgs & a, cudaStream_t stream) {
    auto kern = falcon_decode_mega_kernel<TYPE, JD>;
    const size_t smem = mega_smem_bytes(a.n_ctx);
    static size_t set_for = 0;
    if (set_for < smem) { B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set_for = smem; }
    int dev, nsm; B200_CUDA_CHECK(cudaGetDevice(&dev)); B200_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) nsm); cfg.blockDim = dim3(MG_NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;   // all CTAs co-resident: the counters need it
    cfg.attrs = attr; cfg.numAttrs = getenv("B200_MEGA_NOCOOP") ? 0 : 1;
    B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
}

// Which models the persistent decode kernel covers (everything else takes the per-node path of engine.cu)
bool decode_mega_supports(int wtype, int E, int FF, int H, int HKV, int D, int n_ctx, int n_sm) {
    if (wtype != T_Q4_K && wtype != T_Q4_0) return false;
    const int blk = wtype == T_Q4_K ? 256 : 32, ppb = wtype == T_Q4_K ? 8 : 1;
    if (E % blk || FF % blk) return false;
    const int PE = E / blk * ppb, PF = FF / blk * ppb;
    if (PE > MG_HALF || PF > 2 * MG_NT) return false;
    if (D > 128 || D % 4 || MG_NT % D || D / 2 > 64 || H % HKV) return false;
    if ((E + n_sm - 1) / n_sm > 128) return false;
    if (mega_smem_bytes(n_ctx) > 200 * 1024) return false;
    return true;
}

void launch_decode_mega(int wtype, const void * layers_dev, int n_layer, float * x, float * qkv, float * up, float * att, unsigned * flags,
                        const int * n_past_dev, int n_past, int n_ctx, int E, int FF, int H, int HKV, int D, int dual, float theta_scale, cudaStream_t stream) {
    MegaArgs a;
    a.layers = reinterpret_cast<const MegaLayer *>(layers_dev); a.n_layer = n_layer;
    a.x = x; a.qkv = qkv; a.up = up; a.att = att; a.flags = flags;
    a.n_past_dev = n_past_dev; a.n_past = n_past; a.n_ctx = n_ctx;
    a.E = E; a.FF = FF; a.H = H; a.HKV = HKV; a.D = D; a.dual = dual; a.theta_scale = theta_scale;
    a.trace = nullptr;
    static const char * nm[5] = { "mg_ln", "mg_qkv_up", "mg_attn", "mg_down", "mg_wo" };
    for (int l = 0; l < n_layer; l++) for (int i = 0; i < 5; i++) { unsigned long long * s = b200_trace_slot(nm[i]); if (l == 0 && i == 0) a.trace = s; }
    B200_CUDA_CHECK(cudaMemsetAsync(flags, 0, (size_t) n_layer * 4 * sizeof(unsigned), stream));
    const int blk = wtype == T_Q4_K ? 256 : 32, ppb = wtype == T_Q4_K ? 8 : 1;
    const bool j2 = FF / blk * ppb > MG_NT;
    if (wtype == T_Q4_K) { if (j2) launch_mega_t<T_Q4_K, 2>(a, stream); else launch_mega_t<T_Q4_K, 1>(a, stream); }
    else if (wtype == T_Q4_0) { if (j2) launch_mega_t<T_Q4_0, 2>(a, stream); else launch_mega_t<T_Q4_0, 1>(a, stream); }
    else B200_ASSERT(!"decode_mega: unsupported weight type");
}

size_t decode_mega_layer_bytes() { return sizeof(MegaLayer); }
void decode_mega_fill_layer(void * dst_host, const WPlanes & qkv, const WPlanes & up, const WPlanes & down, const WPlanes & wo,
                            const float * ga, const float * ba, const float * gm, const float * bm, float * kc, float * vc) {
    MegaLayer L; L.qkv = qkv; L.up = up; L.down = down; L.wo = wo; L.ga = ga; L.ba = ba; L.gm = gm; L.bm = bm; L.kc = kc; L.vc = vc;
    memcpy(dst_host, &L, sizeof(L));
}
