// actquant.cuh -- the per-lane body of activation quantisation, shared by the standalone kernel (actquant.cu)
// and the producers that fuse it into their epilogue (LayerNorm, GELU).  See actquant.cu for the contracts.
#pragma once
#include "formats.cuh"

// Each lane owns 8 consecutive values v[0..8) starting at element k0 of activation row n; a 256-block is one
// full warp, a 32-block is 4 consecutive lanes.  All lanes of a WARP must call this together (the shuffles name every lane);
// blocks whose lanes pass store == false (whole blocks past the end of a row) compute on their zeros and write nothing.
template <int TYPE>
__device__ __forceinline__ void quantize_chunk8(const float (&v)[8], int lane, const ActQ & A, int n, int k0, bool store = true) {
    constexpr int LANES = TYPE == T_Q8_K ? 32 : 4;                  // lanes per block
    float amax = 0.f, vmax = 0.f; int imax = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { const float ax = fabsf(v[j]); if (ax > amax) { amax = ax; vmax = v[j]; imax = j; } }
    imax += (lane % LANES) * 8;
    // block reduce; ties keep the FIRST element, like the sequential `if (ax > amax)` scan of the reference
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, amax, o), ov = __shfl_xor_sync(0xffffffffu, vmax, o);
        const int oi = __shfl_xor_sync(0xffffffffu, imax, o);
        if (oa > amax || (oa == amax && oi < imax)) { amax = oa; vmax = ov; imax = oi; }
    }

    int q[8]; int sum = 0; float d;
    if (TYPE == T_Q8_K) {
        if (amax == 0.f) { d = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) q[j] = 0;
        } else {
            const float iscale = __fdiv_rn(-128.f, vmax);
#pragma unroll
            for (int j = 0; j < 8; j++) { q[j] = min(127, __float2int_rn(__fmul_rn(iscale, v[j]))); sum += q[j]; }
            d = __fdiv_rn(1.f, iscale);
        }
    } else {
        const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) { q[j] = __float2int_rn(__fmul_rn(v[j], id)); sum += q[j]; }
        d = __fdiv_rn(amax, 127.f);
        if (TYPE == T_Q8_0) d = __half2float(__float2half_rn(d));
    }
    const uint32_t lo = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
    const uint32_t hi = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((uint32_t) (q[7] & 0xff) << 24);
    if (store) *reinterpret_cast<uint2 *>(A.q + (size_t) n * A.K + k0) = make_uint2(lo, hi);
    if (A.h && store) {                                           // GEMM operand: d * q rounded once to fp16 (what actq_to_f16_kernel computes)
        __half2 hh[4];
#pragma unroll
        for (int j = 0; j < 4; j++) hh[j] = __floats2half2_rn(__fmul_rn(d, (float) q[2 * j]), __fmul_rn(d, (float) q[2 * j + 1]));
        *reinterpret_cast<uint4 *>(A.h + (size_t) n * A.K + k0) = *reinterpret_cast<const uint4 *>(hh);
    }

    if (TYPE == T_Q8_K) {
        const int s16 = sum + __shfl_xor_sync(0xffffffffu, sum, 1);              // 16 codes = 2 lanes
        if (store && (lane & 1) == 0) A.bs[(size_t) n * (A.K / 16) + k0 / 16] = (int16_t) s16;
        if (store && lane == 0) A.d[(size_t) n * (A.K / 256) + k0 / 256] = d;
    } else {
        int s32 = sum + __shfl_xor_sync(0xffffffffu, sum, 1);
        s32 += __shfl_xor_sync(0xffffffffu, s32, 2);
        if (store && (lane & 3) == 0) {
            A.d[(size_t) n * (A.K / 32) + k0 / 32] = d;
            A.bs[(size_t) n * (A.K / 32) + k0 / 32] = (int16_t) s32;     // not part of block_q8_0/1: lets the mat-vec fold the -8 / -16 code offsets
            if (TYPE == T_Q8_1) A.s[(size_t) n * (A.K / 32) + k0 / 32] = __fmul_rn(d, (float) s32);
        }
    }
}
