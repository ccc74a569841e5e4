// mmv.cu -- decode mat-vec: y[m] = sum_k W[m][k] * x[k] for block-quantised W and a Q8-quantised x.
//
// Replaces dequantize_mul_mat_vec<> / dequantize_mul_mat_vec_q{2..6}_k (ggml-cuda.cu:475-845, 1121-1171, one warp
// per row, 2-4 byte scalar loads, fp32 activations) and follows the arithmetic of the CPU twin instead
// (ggml_vec_dot_q*_q8_*, ggml.c:2342-3340 and k_quants.c:1005-2790): int8 x int4..6 block dots in int32 (dp4a),
// one fp32 multiply-accumulate per (sub-)block.  The integer part is exact; only the fp32 summation order
// differs from the CPU (which itself differs between its scalar and AVX2 bodies).
//
// Shape of the kernel (HBM-bound; see DESIGN.md "mmv"):
//   * persistent grid of min(SMs*k, rows) CTAs; warp w of the grid owns rows w, w+W, ...
//   * the activation codes (K bytes) are staged once per CTA into shared memory by a 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier); scales/block sums (tiny) by ordinary loads
//   * every lane streams 16-byte pieces of the quant plane with ld.global.nc.L1::no_allocate, U pieces in
//     flight before the first use; consecutive lanes read consecutive 16 B => 512 B per warp-instruction
//   * lane-local dp4a accumulation, one warp-shuffle reduction per row, lane 0 stores (+ fused epilogue)
#include "kernels.h"

#define MMV_THREADS 256

struct XS {                 // activation row in shared memory
    const int8_t * q; const float * d; const float * s; const int16_t * bs;
};

__device__ __forceinline__ int dot16_u(const uint32_t w0, const uint32_t w1, const uint32_t w2, const uint32_t w3, const uint4 x) {
    int s = dp4a_us(w0, (int) x.x, 0); s = dp4a_us(w1, (int) x.y, s); s = dp4a_us(w2, (int) x.z, s); return dp4a_us(w3, (int) x.w, s);
}
__device__ __forceinline__ uint4 lds16(const int8_t * p) { return *reinterpret_cast<const uint4 *>(p); }

template <int TYPE> struct MV;

// ---------------------------------------------------------------- Q4_K  (k_quants.c:1999-2055)
template <> struct MV<T_Q4_K> {
    static constexpr int PPB = 8;                // 16-byte pieces per 256-weight block
    struct Regs { uint4 q, h; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 128 + pc * 16);
        r.h = ldg_v4(W.p[1] + row * W.stride[1] + (size_t) b * 16);
        return r;
    }
    __device__ static void scales(const uint4 h, int p, int & sc0, int & m0, int & sc1, int & m1) {
        // 6-bit (scale,min) pairs 2p and 2p+1 of the 12-byte field held in h.y h.z h.w (get_scale_min_k4)
        if (p < 2) {
            const int sh = 16 * p;
            sc0 = (h.y >> sh) & 63; sc1 = (h.y >> (sh + 8)) & 63;
            m0 = (h.z >> sh) & 63;  m1 = (h.z >> (sh + 8)) & 63;
        } else {
            const int sh = 16 * (p - 2);
            const uint32_t a = h.w >> sh, lo = h.y >> sh, hi = h.z >> sh;
            sc0 = (a & 0xF) | (((lo >> 6) & 3) << 4);        sc1 = ((a >> 8) & 0xF) | (((lo >> 14) & 3) << 4);
            m0 = ((a >> 4) & 0xF) | (((hi >> 6) & 3) << 4);  m1 = ((a >> 12) & 0xF) | (((hi >> 14) & 3) << 4);
        }
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const int p = pc >> 1, half = pc & 1;
        const int e0 = b * 256 + 64 * p + 16 * half;                     // low nibbles: e0.., high nibbles: e0+32..
        const uint4 xl = lds16(x.q + e0), xh = lds16(x.q + e0 + 32);
        const int il = dot16_u(r.q.x & 0x0F0F0F0F, r.q.y & 0x0F0F0F0F, r.q.z & 0x0F0F0F0F, r.q.w & 0x0F0F0F0F, xl);
        const int ih = dot16_u((r.q.x >> 4) & 0x0F0F0F0F, (r.q.y >> 4) & 0x0F0F0F0F, (r.q.z >> 4) & 0x0F0F0F0F, (r.q.w >> 4) & 0x0F0F0F0F, xh);
        int sc0, m0, sc1, m1; scales(r.h, p, sc0, m0, sc1, m1);
        const int isum = sc0 * il + sc1 * ih;
        const int msum = m0 * x.bs[b * 16 + 4 * p + half] + m1 * x.bs[b * 16 + 4 * p + 2 + half];
        const float xd = x.d[b];
        const float d = f16_bits_to_f32((uint16_t) (r.h.x & 0xffff)), dmin = f16_bits_to_f32((uint16_t) (r.h.x >> 16));
        return (d * xd) * (float) isum - (dmin * xd) * (float) msum;
    }
};

// ---------------------------------------------------------------- Q5_K  (k_quants.c:2340-2400)
template <> struct MV<T_Q5_K> {
    static constexpr int PPB = 8;
    struct Regs { uint4 q, qh, h; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 128 + pc * 16);
        r.qh = ldg_v4(W.p[1] + row * W.stride[1] + (size_t) b * 32 + (pc & 1) * 16);
        r.h = ldg_v4(W.p[2] + row * W.stride[2] + (size_t) b * 16);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const int p = pc >> 1, half = pc & 1;
        const int e0 = b * 256 + 64 * p + 16 * half;
        const uint4 xl = lds16(x.q + e0), xh = lds16(x.q + e0 + 32);
        const int s0 = 2 * p, s1 = 2 * p + 1;
#define LO5(w, hw) (((w) & 0x0F0F0F0F) | ((((hw) >> s0) & 0x01010101) << 4))
#define HI5(w, hw) ((((w) >> 4) & 0x0F0F0F0F) | ((((hw) >> s1) & 0x01010101) << 4))
        const int il = dot16_u(LO5(r.q.x, r.qh.x), LO5(r.q.y, r.qh.y), LO5(r.q.z, r.qh.z), LO5(r.q.w, r.qh.w), xl);
        const int ih = dot16_u(HI5(r.q.x, r.qh.x), HI5(r.q.y, r.qh.y), HI5(r.q.z, r.qh.z), HI5(r.q.w, r.qh.w), xh);
#undef LO5
#undef HI5
        int sc0, m0, sc1, m1; MV<T_Q4_K>::scales(r.h, p, sc0, m0, sc1, m1);
        const int isum = sc0 * il + sc1 * ih;
        const int msum = m0 * x.bs[b * 16 + 4 * p + half] + m1 * x.bs[b * 16 + 4 * p + 2 + half];
        const float xd = x.d[b];
        const float d = f16_bits_to_f32((uint16_t) (r.h.x & 0xffff)), dmin = f16_bits_to_f32((uint16_t) (r.h.x >> 16));
        return (d * xd) * (float) isum - (dmin * xd) * (float) msum;
    }
};

// ---------------------------------------------------------------- Q6_K  (k_quants.c:2748-2789)
template <> struct MV<T_Q6_K> {
    static constexpr int PPB = 8;
    struct Regs { uint4 ql, qh, sc; uint32_t d; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.ql = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 128 + pc * 16);            // = n*64 + c*16
        r.qh = ldg_stream_v4(W.p[1] + row * W.stride[1] + (size_t) b * 64 + (pc >> 2) * 32 + (pc & 1) * 16);
        r.sc = ldg_v4(W.p[2] + row * W.stride[2] + (size_t) b * 16);
        r.d = ldg_u16(W.p[3] + row * W.stride[3] + (size_t) b * 2);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const int n = pc >> 2, c = pc & 3;
        const int el = 128 * n + ((c & 2) ? 32 : 0) + 16 * (c & 1);      // element offset (in block) of the low-nibble group; high-nibble group = el + 64
        const int ls = (c & 2) ? 2 : 0, hs = ls + 4;
        const uint4 xl = lds16(x.q + b * 256 + el), xh = lds16(x.q + b * 256 + el + 64);
#define LO6(w, hw) (((w) & 0x0F0F0F0F) | ((((hw) >> ls) & 0x03030303) << 4))
#define HI6(w, hw) ((((w) >> 4) & 0x0F0F0F0F) | ((((hw) >> hs) & 0x03030303) << 4))
        int il = dot16_u(LO6(r.ql.x, r.qh.x), LO6(r.ql.y, r.qh.y), LO6(r.ql.z, r.qh.z), LO6(r.ql.w, r.qh.w), xl);
        int ih = dot16_u(HI6(r.ql.x, r.qh.x), HI6(r.ql.y, r.qh.y), HI6(r.ql.z, r.qh.z), HI6(r.ql.w, r.qh.w), xh);
#undef LO6
#undef HI6
        il -= 32 * x.bs[b * 16 + (el >> 4)];                             // codes are stored +32
        ih -= 32 * x.bs[b * 16 + ((el + 64) >> 4)];
        const int8_t * sc = reinterpret_cast<const int8_t *>(&r.sc);
        const int isum = (int) sc[el >> 4] * il + (int) sc[(el + 64) >> 4] * ih;
        return (f16_bits_to_f32((uint16_t) r.d) * x.d[b]) * (float) isum;
    }
};

// ---------------------------------------------------------------- Q3_K  (k_quants.c:1684-1745)
template <> struct MV<T_Q3_K> {
    static constexpr int PPB = 4;                // 16 bytes of qs = 64 weights
    struct Regs { uint4 q, hm; uint32_t s0, s1, s2, d; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 64 + pc * 16);              // = n*32 + c*16
        r.hm = ldg_stream_v4(W.p[1] + row * W.stride[1] + (size_t) b * 32 + (pc & 1) * 16);
        const uint8_t * s = W.p[2] + row * W.stride[2] + (size_t) b * 12;
        r.s0 = ldg_u32(s); r.s1 = ldg_u32(s + 4); r.s2 = ldg_u32(s + 8);
        r.d = ldg_u16(W.p[3] + row * W.stride[3] + (size_t) b * 2);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const int n = pc >> 1, c = pc & 1;
        const uint8_t sb[12] = { (uint8_t) r.s0, (uint8_t) (r.s0 >> 8), (uint8_t) (r.s0 >> 16), (uint8_t) (r.s0 >> 24),
                                 (uint8_t) r.s1, (uint8_t) (r.s1 >> 8), (uint8_t) (r.s1 >> 16), (uint8_t) (r.s1 >> 24),
                                 (uint8_t) r.s2, (uint8_t) (r.s2 >> 8), (uint8_t) (r.s2 >> 16), (uint8_t) (r.s2 >> 24) };
        int isum = 0;
#pragma unroll
        for (int quad = 0; quad < 4; quad++) {
            const int el = 128 * n + 32 * quad + 16 * c;
            const uint4 xv = lds16(x.q + b * 256 + el);
            const int hb = 4 * n + quad;
#define C3(w, hw) ((((w) >> (2 * quad)) & 0x03030303) | ((((hw) >> hb) & 0x01010101) << 2))
            int i = dot16_u(C3(r.q.x, r.hm.x), C3(r.q.y, r.hm.y), C3(r.q.z, r.hm.z), C3(r.q.w, r.hm.w), xv);
#undef C3
            i -= 4 * x.bs[b * 16 + (el >> 4)];                          // code = (q2 | hbit<<2) - 4
            isum += (q3_scale(sb, el >> 4) - 32) * i;
        }
        return (f16_bits_to_f32((uint16_t) r.d) * x.d[b]) * (float) isum;
    }
};

// ---------------------------------------------------------------- Q2_K  (k_quants.c:1267-1305)
template <> struct MV<T_Q2_K> {
    static constexpr int PPB = 4;
    struct Regs { uint4 q, sc; uint32_t dm; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 64 + pc * 16);
        r.sc = ldg_v4(W.p[1] + row * W.stride[1] + (size_t) b * 16);
        r.dm = ldg_u32(W.p[2] + row * W.stride[2] + (size_t) b * 4);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const int n = pc >> 1, c = pc & 1;
        const uint8_t * sc = reinterpret_cast<const uint8_t *>(&r.sc);
        int isum = 0, msum = 0;
#pragma unroll
        for (int quad = 0; quad < 4; quad++) {
            const int el = 128 * n + 32 * quad + 16 * c;
            const uint4 xv = lds16(x.q + b * 256 + el);
#define C2(w) (((w) >> (2 * quad)) & 0x03030303)
            const int i = dot16_u(C2(r.q.x), C2(r.q.y), C2(r.q.z), C2(r.q.w), xv);
#undef C2
            const int s = sc[el >> 4];
            isum += (s & 0xF) * i;
            msum += (s >> 4) * x.bs[b * 16 + (el >> 4)];
        }
        const float xd = x.d[b];
        return (xd * f16_bits_to_f32((uint16_t) (r.dm & 0xffff))) * (float) isum - (xd * f16_bits_to_f32((uint16_t) (r.dm >> 16))) * (float) msum;
    }
};

// ---------------------------------------------------------------- legacy 32-weight blocks (ggml.c:2591-2609, 2716-2733, 2952-2974, 3208-3230, 3321-3333)
__device__ __forceinline__ uint32_t spread4(uint32_t bits4) { return ((bits4 & 0xF) * 0x00204081u) & 0x01010101u; }   // bit i -> byte i

template <> struct MV<T_Q4_0> {
    static constexpr int PPB = 1;
    struct Regs { uint4 q; uint32_t d; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 16);
        r.d = ldg_u16(W.p[1] + row * W.stride[1] + (size_t) b * 2);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int, const XS & x) {
        const uint4 xl = lds16(x.q + b * 32), xh = lds16(x.q + b * 32 + 16);
        int s = dot16_u(r.q.x & 0x0F0F0F0F, r.q.y & 0x0F0F0F0F, r.q.z & 0x0F0F0F0F, r.q.w & 0x0F0F0F0F, xl);
        s += dot16_u((r.q.x >> 4) & 0x0F0F0F0F, (r.q.y >> 4) & 0x0F0F0F0F, (r.q.z >> 4) & 0x0F0F0F0F, (r.q.w >> 4) & 0x0F0F0F0F, xh);
        s -= 8 * x.bs[b];
        return ((float) s * f16_bits_to_f32((uint16_t) r.d)) * x.d[b];
    }
};
template <> struct MV<T_Q4_1> {
    static constexpr int PPB = 1;
    struct Regs { uint4 q; uint32_t dm; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 16);
        r.dm = ldg_u32(W.p[1] + row * W.stride[1] + (size_t) b * 4);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int, const XS & x) {
        const uint4 xl = lds16(x.q + b * 32), xh = lds16(x.q + b * 32 + 16);
        int s = dot16_u(r.q.x & 0x0F0F0F0F, r.q.y & 0x0F0F0F0F, r.q.z & 0x0F0F0F0F, r.q.w & 0x0F0F0F0F, xl);
        s += dot16_u((r.q.x >> 4) & 0x0F0F0F0F, (r.q.y >> 4) & 0x0F0F0F0F, (r.q.z >> 4) & 0x0F0F0F0F, (r.q.w >> 4) & 0x0F0F0F0F, xh);
        return (f16_bits_to_f32((uint16_t) (r.dm & 0xffff)) * x.d[b]) * (float) s + f16_bits_to_f32((uint16_t) (r.dm >> 16)) * x.s[b];
    }
};
template <> struct MV<T_Q5_0> {
    static constexpr int PPB = 1;
    struct Regs { uint4 q; uint32_t qh, d; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 16);
        r.qh = ldg_u32(W.p[1] + row * W.stride[1] + (size_t) b * 4);
        r.d = ldg_u16(W.p[2] + row * W.stride[2] + (size_t) b * 2);
        return r;
    }
    __device__ static int idot(const uint4 q, uint32_t qh, int b, const XS & x) {
        const uint4 xl = lds16(x.q + b * 32), xh = lds16(x.q + b * 32 + 16);
        int s = dot16_u((q.x & 0x0F0F0F0F) | (spread4(qh) << 4), (q.y & 0x0F0F0F0F) | (spread4(qh >> 4) << 4),
                        (q.z & 0x0F0F0F0F) | (spread4(qh >> 8) << 4), (q.w & 0x0F0F0F0F) | (spread4(qh >> 12) << 4), xl);
        s += dot16_u(((q.x >> 4) & 0x0F0F0F0F) | (spread4(qh >> 16) << 4), ((q.y >> 4) & 0x0F0F0F0F) | (spread4(qh >> 20) << 4),
                     ((q.z >> 4) & 0x0F0F0F0F) | (spread4(qh >> 24) << 4), ((q.w >> 4) & 0x0F0F0F0F) | (spread4(qh >> 28) << 4), xh);
        return s;
    }
    __device__ static float dot(const Regs & r, int b, int, const XS & x) {
        const int s = idot(r.q, r.qh, b, x) - 16 * x.bs[b];
        return (f16_bits_to_f32((uint16_t) r.d) * x.d[b]) * (float) s;
    }
};
template <> struct MV<T_Q5_1> {
    static constexpr int PPB = 1;
    struct Regs { uint4 q; uint32_t qh, dm; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 16);
        r.qh = ldg_u32(W.p[1] + row * W.stride[1] + (size_t) b * 4);
        r.dm = ldg_u32(W.p[2] + row * W.stride[2] + (size_t) b * 4);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int, const XS & x) {
        const int s = MV<T_Q5_0>::idot(r.q, r.qh, b, x);
        return (f16_bits_to_f32((uint16_t) (r.dm & 0xffff)) * x.d[b]) * (float) s + f16_bits_to_f32((uint16_t) (r.dm >> 16)) * x.s[b];
    }
};
template <> struct MV<T_Q8_0> {
    static constexpr int PPB = 2;                // 16 int8 weights per piece
    struct Regs { uint4 q; uint32_t d; };
    __device__ static Regs load(const WPlanes & W, size_t row, int b, int pc) {
        Regs r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) b * 32 + pc * 16);
        r.d = ldg_u16(W.p[1] + row * W.stride[1] + (size_t) b * 2);
        return r;
    }
    __device__ static float dot(const Regs & r, int b, int pc, const XS & x) {
        const uint4 xv = lds16(x.q + b * 32 + pc * 16);
        int s = dp4a_ss((int) r.q.x, (int) xv.x, 0); s = dp4a_ss((int) r.q.y, (int) xv.y, s);
        s = dp4a_ss((int) r.q.z, (int) xv.z, s); s = dp4a_ss((int) r.q.w, (int) xv.w, s);
        return (float) s * (f16_bits_to_f32((uint16_t) r.d) * x.d[b]);
    }
};

// fp16-LUT-equivalent GELU (ggml.c:3461-3484): f16 in, fp32 formula, f16 out
__device__ __forceinline__ float gelu_f16lut(float v) {
    const float f = __half2float(__float2half_rn(v));
    const float g = 0.5f * f * (1.0f + tanhf(0.79788456080286535587989211986876f * f * (1.0f + 0.044715f * f * f)));
    return __half2float(__float2half_rn(g));
}

template <int TYPE, int U>
__global__ void __launch_bounds__(MMV_THREADS) mmv_kernel(const WPlanes W, const ActQ A, float * __restrict__ y, int64_t y_stride, const MmvEpilogue epi) {
    using T = MV<TYPE>;
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.y;                                   // activation row (column of Y)
    const int K = W.K, ablk = TYPE >= T_Q2_K ? 256 : 32;
    const int nd = K / ablk, nbs = TYPE >= T_Q2_K ? K / 16 : K / 32;
    uint64_t * bar = reinterpret_cast<uint64_t *>(smem);
    int8_t * xq = reinterpret_cast<int8_t *>(smem + 16);
    float * xd = reinterpret_cast<float *>(smem + 16 + K);
    float * xs = xd + nd;                                       // Q8_1 only (nd entries), otherwise unused
    int16_t * xbs = reinterpret_cast<int16_t *>((TYPE == T_Q4_1 || TYPE == T_Q5_1) ? (xs + nd) : xs);

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t) K);
        tma_load_1d(xq, A.q + (size_t) n * K, (uint32_t) K, bar);      // activation tile: global -> shared via TMA
    }
    for (int i = threadIdx.x; i < nd; i += MMV_THREADS) {
        xd[i] = A.d[(size_t) n * nd + i];
        if (TYPE == T_Q4_1 || TYPE == T_Q5_1) xs[i] = A.s[(size_t) n * nd + i];
    }
    for (int i = threadIdx.x; i < nbs; i += MMV_THREADS) xbs[i] = A.bs[(size_t) n * nbs + i];
    __syncthreads();                                            // barrier init + scale arrays visible
    const XS x = { xq, xd, xs, xbs };

    const int lane = threadIdx.x & 31;
    const int warps_per_cta = MMV_THREADS / 32;
    const int gw = blockIdx.x * warps_per_cta + (threadIdx.x >> 5), nw = gridDim.x * warps_per_cta;
    const int P = W.nb * T::PPB;                                // 16-byte pieces per row
    bool staged = false;

    for (int row = gw; row < W.M; row += nw) {
        float acc = 0.f;
        for (int g0 = 0; g0 < P; g0 += 32 * U) {
            typename T::Regs regs[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int g = g0 + u * 32 + lane;
                if (g < P) regs[u] = T::load(W, (size_t) row, g / T::PPB, g % T::PPB);
            }
            if (!staged) { mbar_wait(bar, 0); staged = true; }  // first weight loads are already in flight
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int g = g0 + u * 32 + lane;
                if (g < P) acc += T::dot(regs[u], g / T::PPB, g % T::PPB, x);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc;
            if (epi.kind == EPI_GELU) v = gelu_f16lut(v);
            else if (epi.kind == EPI_ADD2) v = (v + epi.r1[(size_t) n * y_stride + row]) + epi.r2[(size_t) n * y_stride + row];
            y[(size_t) n * y_stride + row] = v;
        }
    }
    if (!staged) mbar_wait(bar, 0);                             // never exit with a bulk copy still landing in our smem
}

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) { int dev; B200_CUDA_CHECK(cudaGetDevice(&dev)); B200_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev)); }
    return g_num_sms;
}

template <int TYPE, int U>
static void launch_typed(const WPlanes & W, const ActQ & A, float * y, int64_t y_stride, MmvEpilogue epi, cudaStream_t stream) {
    const int K = W.K, ablk = TYPE >= T_Q2_K ? 256 : 32;
    const size_t smem = 16 + (size_t) K + (size_t) (K / ablk) * 4 * ((TYPE == T_Q4_1 || TYPE == T_Q5_1) ? 2 : 1) + (size_t) (TYPE >= T_Q2_K ? K / 16 : K / 32) * 2 + 16;
    static bool attr_set = false;
    if (!attr_set) { B200_CUDA_CHECK(cudaFuncSetAttribute(mmv_kernel<TYPE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_set = true; }
    const int warps = MMV_THREADS / 32;
    int ctas = num_sms() * 4;                                   // persistent: 4 CTAs x 256 threads per SM when shared memory allows
    const int need = (W.M + warps - 1) / warps;
    if (ctas > need) ctas = need;
    dim3 grid((unsigned) ctas, (unsigned) A.N);
    mmv_kernel<TYPE, U><<<grid, MMV_THREADS, smem, stream>>>(W, A, y, y_stride, epi);
    B200_CUDA_CHECK(cudaGetLastError());
}

void launch_mmv(const WPlanes & W, const ActQ & A, float * y, int64_t y_stride, MmvEpilogue epi, cudaStream_t stream) {
    B200_ASSERT(A.K == W.K && A.type == act_type_for(W.type));
    B200_ASSERT(W.K % 32 == 0 && W.K <= 190 * 1024);
    switch (W.type) {
        case T_Q4_K: launch_typed<T_Q4_K, 4>(W, A, y, y_stride, epi, stream); break;
        case T_Q5_K: launch_typed<T_Q5_K, 2>(W, A, y, y_stride, epi, stream); break;
        case T_Q6_K: launch_typed<T_Q6_K, 2>(W, A, y, y_stride, epi, stream); break;
        case T_Q3_K: launch_typed<T_Q3_K, 2>(W, A, y, y_stride, epi, stream); break;
        case T_Q2_K: launch_typed<T_Q2_K, 2>(W, A, y, y_stride, epi, stream); break;
        case T_Q4_0: launch_typed<T_Q4_0, 4>(W, A, y, y_stride, epi, stream); break;
        case T_Q4_1: launch_typed<T_Q4_1, 4>(W, A, y, y_stride, epi, stream); break;
        case T_Q5_0: launch_typed<T_Q5_0, 4>(W, A, y, y_stride, epi, stream); break;
        case T_Q5_1: launch_typed<T_Q5_1, 4>(W, A, y, y_stride, epi, stream); break;
        case T_Q8_0: launch_typed<T_Q8_0, 4>(W, A, y, y_stride, epi, stream); break;
        default: B200_ASSERT(!"launch_mmv: unsupported weight type");
    }
}

// ---- f16 / f32 weights with fp32 activations (ggml.c:10911-11102, 11104-11316): one warp per row, 16-byte loads
template <bool F16>
__global__ void __launch_bounds__(256) mmv_f_kernel(const WPlanes W, const float * __restrict__ x, int64_t x_stride, float * __restrict__ y, int64_t y_stride) {
    const int n = blockIdx.y, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    const float * xr = x + (size_t) n * x_stride;
    for (int row = gw; row < W.M; row += nw) {
        float acc = 0.f;
        const uint8_t * wr = W.p[0] + (size_t) row * W.stride[0];
        if (F16) {
            for (int k = lane * 8; k < W.K; k += 256) {
                if (k + 8 <= W.K) {
                    const uint4 w = ldg_stream_v4(wr + (size_t) k * 2);
                    const __half2 * h = reinterpret_cast<const __half2 *>(&w);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 f = __half22float2(h[j]);
                        // the CPU path rounds the activations to fp16 first (ggml.c:11232-11251)
                        acc += f.x * __half2float(__float2half_rn(xr[k + 2 * j])) + f.y * __half2float(__float2half_rn(xr[k + 2 * j + 1]));
                    }
                } else for (int j = k; j < W.K; j++) acc += f16_bits_to_f32(reinterpret_cast<const uint16_t *>(wr)[j]) * __half2float(__float2half_rn(xr[j]));
            }
        } else {
            for (int k = lane * 4; k < W.K; k += 128) {
                if (k + 4 <= W.K) {
                    const uint4 w = ldg_stream_v4(wr + (size_t) k * 4);
                    acc += __uint_as_float(w.x) * xr[k] + __uint_as_float(w.y) * xr[k + 1] + __uint_as_float(w.z) * xr[k + 2] + __uint_as_float(w.w) * xr[k + 3];
                } else for (int j = k; j < W.K; j++) acc += reinterpret_cast<const float *>(wr)[j] * xr[j];
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) y[(size_t) n * y_stride + row] = acc;
    }
}
void launch_mmv_f(const WPlanes & W, const float * x, int64_t x_stride, int N, float * y, int64_t y_stride, cudaStream_t stream) {
    int ctas = num_sms() * 4; const int need = (W.M + 7) / 8; if (ctas > need) ctas = need;
    dim3 grid((unsigned) ctas, (unsigned) N);
    if (W.type == T_F16) mmv_f_kernel<true><<<grid, 256, 0, stream>>>(W, x, x_stride, y, y_stride);
    else if (W.type == T_F32) mmv_f_kernel<false><<<grid, 256, 0, stream>>>(W, x, x_stride, y, y_stride);
    else B200_ASSERT(!"launch_mmv_f: f16/f32 only");
    B200_CUDA_CHECK(cudaGetLastError());
}
