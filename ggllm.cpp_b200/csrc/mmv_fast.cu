// mmv_fast.cu -- the decode mat-vec for the types the BASELINE configs use (Q4_K, Q4_0), tuned for issue slots.
//
// ncu on the generic ring kernel (profiles/r1_mmv_q4k_ring.md) showed the mat-vec is NOT limited by HBM latency but
// by instruction issue and the L1/shared-memory data path: 101 warp-instructions per 512 B of weights, most of them
// address arithmetic, nibble shifts, 6-bit scale unpacking and shared-memory reads of the activation.  This kernel
// removes them instead of hiding them:
//   * the activation lives in REGISTERS.  Thread t of a CTA always handles the same 16-byte "piece" position of a
//     row (piece g = j*NT + t), so the 32 activation codes, their block sums and scale that piece needs are loaded
//     once (TMA bulk copy global->shared, then one read into registers) and reused for every row the CTA owns
//   * weights go HBM -> registers with one coalesced 16-byte ld.global.nc per piece (a warp reads 512 contiguous
//     bytes of the row), double-buffered: the next group of rows is in flight while the current one is computed
//   * high nibbles are dotted in place (x16, exact) instead of shifted; the four 6-bit scale/min values of a
//     sub-block pair come pre-expanded in one 32-bit word and are applied with two dp2a
//   * the K dimension is split across the warps of the CTA; G rows are reduced together with a transposing
//     butterfly (6 shuffles per 4 rows) and a fixed-order sum over warps through shared memory -> deterministic
// Arithmetic is identical to mmv.cu / the CPU's integer block dots (ggml.c:2591-2609, k_quants.c:1999-2055).
#include "kernels.h"

struct Epi { int kind; const float * r1; const float * r2; unsigned long long * trace; };

// Where the activation row comes from (FastX, kernels.h):
//   mode 0: already quantised (ActQ, written by quantize_act / layernorm_q)
//   mode 1: fp32 row x[K]; every CTA quantises it itself while its first weight rows are in flight
//   mode 2: fp32 row -> [x = (ra + rb) + x] -> LayerNorm(gamma, beta) -> quantise, all in the prologue (J == 1 only):
//           the residual adds that close the previous layer (libfalcon.cpp:2399-2400), the LayerNorm
//           (ggml.c:10568-10595 + libfalcon.cpp:2166-2185) and the mat-mul's INIT pass (ggml.c:11462-11476) without a
//           kernel of their own.  CTA 0 writes the updated residual row to x_out.
// In modes 1/2 the 8 threads that share a Q8_K block hold exactly its 256 values (32 each), so the block maximum is
// three shuffles away and the int8 codes are produced directly in the registers the dot products read.

__device__ __forceinline__ int dot16(const uint32_t w0, const uint32_t w1, const uint32_t w2, const uint32_t w3, const uint4 x) {
    int s = dp4a_us(w0, (int) x.x, 0); s = dp4a_us(w1, (int) x.y, s); s = dp4a_us(w2, (int) x.z, s); return dp4a_us(w3, (int) x.w, s);
}
__device__ __forceinline__ int dp2a_lo_su(int pair16, uint32_t bytes) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair16), "r"(bytes), "r"(0)); return d; }
__device__ __forceinline__ int dp2a_hi_su(int pair16, uint32_t bytes) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair16), "r"(bytes), "r"(0)); return d; }
__device__ __forceinline__ float gelu_lut(float v) {      // fp16-LUT semantics, ggml.c:3461-3484
    const float f = __half2float(__float2half_rn(v));
    const float g = 0.5f * f * (1.0f + tanhf(0.79788456080286535587989211986876f * f * (1.0f + 0.044715f * f * f)));
    return __half2float(__float2half_rn(g));
}

template <int TYPE> struct FX;

template <> struct FX<T_Q4_K> {
    static constexpr int PPB = 8;                                   // pieces per block
    struct XR { uint4 xl, xh; int bs; float xd; };                  // activation state of one piece position
    struct WR { uint4 q; uint32_t sm, dd; };
    __device__ static XR load_x(const int8_t * xq, const ActQ & A, int n, int g) {
        const int b = g >> 3, pc = g & 7, p = pc >> 1, half = pc & 1;
        XR r;
        const int e0 = b * 256 + 64 * p + 16 * half;
        r.xl = *reinterpret_cast<const uint4 *>(xq + e0);
        r.xh = *reinterpret_cast<const uint4 *>(xq + e0 + 32);
        const int16_t * bs = A.bs + (size_t) n * (A.K / 16) + b * 16 + 4 * p + half;
        r.bs = ((int) bs[0] & 0xffff) | ((int) bs[2] << 16);
        r.xd = A.d[(size_t) n * (A.K / 256) + b];
        return r;
    }
    // element offsets (in the row) of the two 16-value segments piece g multiplies
    __device__ static void seg(int g, int & ea, int & eb) { const int b = g >> 3, pc = g & 7; ea = b * 256 + 64 * (pc >> 1) + 16 * (pc & 1); eb = ea + 32; }
    // v[0..16) = segment a, v[16..32) = segment b of this thread's piece; the 8 lanes of a block quantise it together
    // (quantize_row_q8_K_reference, k_quants.c:899-934: signed value of largest magnitude, first one on ties)
    __device__ static XR quant_x(const float (&v)[32], int g, int lane) {
        int ea, eb; seg(g, ea, eb);
        float amax = 0.f, vmax = 0.f; int imax = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) { const float ax = fabsf(v[i]); const int idx = (i < 16 ? ea : eb - 16) + i; if (ax > amax || (ax == amax && ax > 0.f && idx < imax)) { amax = ax; vmax = v[i]; imax = idx; } }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, amax, o), ov = __shfl_xor_sync(0xffffffffu, vmax, o);
            const int oi = __shfl_xor_sync(0xffffffffu, imax, o);
            if (oa > amax || (oa == amax && oi < imax)) { amax = oa; vmax = ov; imax = oi; }
        }
        XR r;
        const bool zero = amax == 0.f;
        const float iscale = zero ? 0.f : __fdiv_rn(-128.f, vmax);
        r.xd = zero ? 0.f : __fdiv_rn(1.f, iscale);
        int s0 = 0, s1 = 0;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {                                         // codes are packed as they are produced: nothing but v[] stays live
            uint32_t pk = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int q = zero ? 0 : min(127, __float2int_rn(__fmul_rn(iscale, v[4 * i + k])));
                if (i < 4) s0 += q; else s1 += q;
                pk |= (uint32_t) (q & 0xff) << (8 * k);
            }
            w[i] = pk;
        }
        r.xl = make_uint4(w[0], w[1], w[2], w[3]); r.xh = make_uint4(w[4], w[5], w[6], w[7]);
        r.bs = (s0 & 0xffff) | (s1 << 16);
        (void) lane;
        return r;
    }
    __device__ static WR load_w(const WPlanes & W, size_t row, int g) {
        WR r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) g * 16);
        r.sm = ldg_u32(W.p[1] + row * W.stride[1] + (size_t) (g >> 1) * 4);
        r.dd = ldg_u32(W.p[2] + row * W.stride[2] + (size_t) (g >> 3) * 4);
        return r;
    }
    __device__ static float dot(const WR & w, const XR & x) {
        const int il = dot16(w.q.x & 0x0F0F0F0F, w.q.y & 0x0F0F0F0F, w.q.z & 0x0F0F0F0F, w.q.w & 0x0F0F0F0F, x.xl);
        const int ih = dot16(w.q.x & 0xF0F0F0F0, w.q.y & 0xF0F0F0F0, w.q.z & 0xF0F0F0F0, w.q.w & 0xF0F0F0F0, x.xh) >> 4;
        const int isum = dp2a_lo_su((il & 0xffff) | (ih << 16), w.sm);       // sc0*il + sc1*ih   (|il|,|ih| <= 16*15*127 < 2^15)
        const int msum = dp2a_hi_su(x.bs, w.sm);                             // m0*bs_lo + m1*bs_hi
        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&w.dd));
        return (dm.x * x.xd) * (float) isum - (dm.y * x.xd) * (float) msum;
    }
};

template <> struct FX<T_Q4_0> {
    static constexpr int PPB = 1;
    struct XR { uint4 xl, xh; int bs; float xd; };
    struct WR { uint4 q; uint32_t d; };
    __device__ static XR load_x(const int8_t * xq, const ActQ & A, int n, int g) {
        XR r;
        r.xl = *reinterpret_cast<const uint4 *>(xq + g * 32);
        r.xh = *reinterpret_cast<const uint4 *>(xq + g * 32 + 16);
        r.bs = A.bs[(size_t) n * (A.K / 32) + g];
        r.xd = A.d[(size_t) n * (A.K / 32) + g];
        return r;
    }
    __device__ static void seg(int g, int & ea, int & eb) { ea = g * 32; eb = ea + 16; }
    // a piece is a whole 32-value block: the x86 body of quantize_row_q8_0 (ggml.c:1201-1237), thread-local
    __device__ static XR quant_x(const float (&v)[32], int, int) {
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(v[i]));
        const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
        XR r;
        r.xd = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
        int s = 0;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t pk = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) { const int q = __float2int_rn(__fmul_rn(v[4 * i + k], id)); s += q; pk |= (uint32_t) (q & 0xff) << (8 * k); }
            w[i] = pk;
        }
        r.xl = make_uint4(w[0], w[1], w[2], w[3]); r.xh = make_uint4(w[4], w[5], w[6], w[7]);
        r.bs = s;
        return r;
    }
    __device__ static WR load_w(const WPlanes & W, size_t row, int g) {
        WR r;
        r.q = ldg_stream_v4(W.p[0] + row * W.stride[0] + (size_t) g * 16);
        r.d = ldg_u16(W.p[1] + row * W.stride[1] + (size_t) g * 2);
        return r;
    }
    __device__ static float dot(const WR & w, const XR & x) {
        int s = dot16(w.q.x & 0x0F0F0F0F, w.q.y & 0x0F0F0F0F, w.q.z & 0x0F0F0F0F, w.q.w & 0x0F0F0F0F, x.xl);
        s += dot16(w.q.x & 0xF0F0F0F0, w.q.y & 0xF0F0F0F0, w.q.z & 0xF0F0F0F0, w.q.w & 0xF0F0F0F0, x.xh) >> 4;
        s -= 8 * x.bs;                                                       // codes are stored +8
        return ((float) s * f16_bits_to_f32((uint16_t) w.d)) * x.xd;
    }
};

// G row sums per lane -> the total of row r in every lane of the 8-lane group r (G == 4), 16-lane group (G == 2) or warp
template <int G> __device__ __forceinline__ float transpose_reduce(const float (&acc)[G], int lane, int & row_of_lane) {
    float w;
    if (G == 4) {
        const bool hi = lane & 16;
        float v0 = hi ? acc[2] : acc[0], v1 = hi ? acc[3] : acc[1];
        v0 += __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[2], 16);
        v1 += __shfl_xor_sync(0xffffffffu, hi ? acc[1] : acc[3], 16);
        const bool mid = lane & 8;
        w = mid ? v1 : v0;
        w += __shfl_xor_sync(0xffffffffu, mid ? v0 : v1, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        row_of_lane = (hi ? 2 : 0) + (mid ? 1 : 0);
    } else if (G == 2) {
        const bool hi = lane & 16;
        w = hi ? acc[G - 1] : acc[0];
        w += __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[G - 1], 16);
        w += __shfl_xor_sync(0xffffffffu, w, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        row_of_lane = hi ? 1 : 0;
    } else {
        w = acc[0];
        w += __shfl_xor_sync(0xffffffffu, w, 16);
        w += __shfl_xor_sync(0xffffffffu, w, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        row_of_lane = 0;
    }
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w;
}

// D = rows in flight per thread (register ring), reduced G = min(D, 4) rows at a time.  D * J = 8 pieces = 192 B in
// flight per thread at all times (>= 96 KB per SM): the ring is refilled one row at a time, right after that row's
// slot has been consumed, so the depth never drops while a group is being computed.
template <int TYPE, int NT, int J, int D, int MODE>
__global__ void __launch_bounds__(NT, NT == 256 ? (D <= 6 ? 3 : 2) : 1) mmv_fast_kernel(const WPlanes W, const FastX X, float * __restrict__ y, int64_t y_stride, const Epi epi) {
    using T = FX<TYPE>;
    constexpr int NW = NT / 32, G = (D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1;
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t * bar = reinterpret_cast<uint64_t *>(smem);
    float * partial = reinterpret_cast<float *>(smem + 16);           // [2][NW][G]
    double * red = reinterpret_cast<double *>(smem + 16 + 2 * NW * 4 * 4);      // [NW] block reduction scratch (mode 2)
    int8_t * xq = reinterpret_cast<int8_t *>(smem + 16 + 2 * NW * 4 * 4 + NW * 8);
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = W.nb * T::PPB;
    const ActQ & A = X.A;

    trace_begin(epi.trace);
    if (tid == 0 && MODE == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // rows [row0, row1) of this CTA, balanced to +-1
    const int per = W.M / gridDim.x, rem = W.M % gridDim.x;
    const int row0 = blockIdx.x * per + min((int) blockIdx.x, rem), row1 = row0 + per + ((int) blockIdx.x < rem ? 1 : 0);
    const int nrows = row1 - row0;

    int gidx[J]; bool valid[J];
#pragma unroll
    for (int j = 0; j < J; j++) { const int g = j * NT + tid; valid[j] = g < P; gidx[j] = valid[j] ? g : P - 1; }

    typename T::WR w[D][J];
#pragma unroll
    for (int s = 0; s < D; s++) {                                            // weights are in flight before the activation arrives
        const int row = min(row0 + s, row1 - 1);
#pragma unroll
        for (int j = 0; j < J; j++) w[s][j] = T::load_w(W, (size_t) row, gidx[j]);
    }

    // everything above touched only weights; the activation row is produced by the previous kernel(s) of the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
    typename T::XR xr[J];
    if (MODE == 0) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t) W.K);
            tma_load_1d(xq, A.q + (size_t) n * W.K, (uint32_t) W.K, bar);   // activation codes: global -> shared by TMA
        }
        __syncthreads();                                                      // barrier initialised (thread 0 did it before issuing)
        mbar_wait(bar, 0);
#pragma unroll
        for (int j = 0; j < J; j++) xr[j] = T::load_x(xq, A, n, gidx[j]);
    } else {
        const float * xrow = X.x + (size_t) n * X.x_stride;
        float mean = 0.f, scale = 1.f;
#pragma unroll
        for (int j = 0; j < J; j++) {
            int ea, eb; T::seg(gidx[j], ea, eb);
            float v[32];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 a = *reinterpret_cast<const float4 *>(xrow + ea + 4 * i), b = *reinterpret_cast<const float4 *>(xrow + eb + 4 * i);
                v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
                v[16 + 4 * i] = b.x; v[17 + 4 * i] = b.y; v[18 + 4 * i] = b.z; v[19 + 4 * i] = b.w;
            }
            if (MODE == 2 && J == 1) {
                if (X.ra) {                                                   // x = (ra + rb) + x, libfalcon.cpp:2399-2400
                    const float * ra = X.ra + (size_t) n * X.x_stride, * rb = X.rb + (size_t) n * X.x_stride;
#pragma unroll
                    for (int i = 0; i < 16; i++) { v[i] = __fadd_rn(__fadd_rn(ra[ea + i], rb[ea + i]), v[i]); v[16 + i] = __fadd_rn(__fadd_rn(ra[eb + i], rb[eb + i]), v[16 + i]); }
                }
                if (X.x_out && blockIdx.x == 0 && valid[j]) {
                    float * xo = X.x_out + (size_t) n * X.x_stride;
#pragma unroll
                    for (int i = 0; i < 16; i++) { xo[ea + i] = v[i]; xo[eb + i] = v[16 + i]; }
                }
                // LayerNorm over the whole row: this CTA's NT threads hold all K values (32 each)
                double s = 0.0;
                if (valid[j]) {
#pragma unroll
                    for (int i = 0; i < 32; i++) s += (double) v[i];
                }
                s = warp_sum_d(s);
                if (lane == 0) red[warp] = s;
                __syncthreads();
                double t = 0.0;
#pragma unroll
                for (int wi = 0; wi < NW; wi++) t += red[wi];
                mean = (float) (t / W.K);
                __syncthreads();
                double s2 = 0.0;
                if (valid[j]) {
#pragma unroll
                    for (int i = 0; i < 32; i++) { const float c = __fsub_rn(v[i], mean); s2 += (double) __fmul_rn(c, c); }
                }
                s2 = warp_sum_d(s2);
                if (lane == 0) red[warp] = s2;
                __syncthreads();
                t = 0.0;
#pragma unroll
                for (int wi = 0; wi < NW; wi++) t += red[wi];
                scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float) (t / W.K), 1e-5f)));
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    v[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[i], mean), scale), X.gamma[ea + i]), X.beta[ea + i]);
                    v[16 + i] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[16 + i], mean), scale), X.gamma[eb + i]), X.beta[eb + i]);
                }
            }
            if (!valid[j]) {
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = 0.f;
            }
            xr[j] = T::quant_x(v, gidx[j], lane);
        }
        __syncthreads();                                                      // `partial` / `red` are reused below
    }

    float acc[G];
    for (int base = 0; base < nrows; base += D) {
        // Programmatic dependent launch: release the next kernel of the stream once this CTA has issued its last weight
        // loads.  Triggering earlier would park the dependent grid's CTAs at the head of the hardware queue for the whole
        // duration of this kernel and keep the small attention kernels of the other stream from being scheduled
        // (measured: profiles/r1_decode_timeline.md).
        if (base + 2 * D >= nrows) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#pragma unroll
        for (int s = 0; s < D; s++) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < J; j++) { const float d = T::dot(w[s][j], xr[j]); a += valid[j] ? d : 0.f; }
            acc[s % G] = a;
            const int nxt = row0 + base + D + s;                             // refill this slot D rows ahead
            if (nxt < row1) {
#pragma unroll
                for (int j = 0; j < J; j++) w[s][j] = T::load_w(W, (size_t) nxt, gidx[j]);
            }
            if ((s % G) == G - 1) {
                const int gi = (base + s) / G;                               // group index; rows gi*G .. gi*G+G-1 (relative)
                int rl;
                const float v0 = transpose_reduce<G>(acc, lane, rl);
                float * part = partial + (gi & 1) * NW * G;
                if ((lane & (G == 4 ? 7 : G == 2 ? 15 : 31)) == 0) part[warp * G + rl] = v0;
                __syncthreads();
                if (tid < G) {
                    const int row = row0 + gi * G + tid;
                    if (row < row1) {
                        float v = 0.f;
#pragma unroll
                        for (int wi = 0; wi < NW; wi++) v += part[wi * G + tid];      // fixed order: deterministic
                        if (epi.kind == EPI_GELU) v = gelu_lut(v);
                        else if (epi.kind == EPI_ADD2) v = (v + epi.r1[(size_t) n * y_stride + row]) + epi.r2[(size_t) n * y_stride + row];
                        y[(size_t) n * y_stride + row] = v;
                    }
                }
            }
        }
    }
    trace_end(epi.trace);
}

static int fast_num_sms() {
    static int n = 0;
    if (!n) { int dev; B200_CUDA_CHECK(cudaGetDevice(&dev)); B200_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); }
    return n;
}

template <int TYPE, int NT, int J, int D, int MODE>
static void launch_mode(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, Epi epi, cudaStream_t stream) {
    const size_t smem = 16 + 2 * (NT / 32) * 4 * 4 + (NT / 32) * 8 + (size_t) W.K;
    static bool set = false;
    if (!set) { B200_CUDA_CHECK(cudaFuncSetAttribute(mmv_fast_kernel<TYPE, NT, J, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        // same L1/shared split as the small kernels of the other stream: an SM cannot host kernels with different carve-outs at once
        B200_CUDA_CHECK(cudaFuncSetAttribute(mmv_fast_kernel<TYPE, NT, J, D, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, B200_CARVEOUT)); set = true; }
    int ctas = fast_num_sms() * (NT == 256 ? (D <= 6 ? 3 : 2) : 1);
    if (ctas > (W.M + 3) / 4) ctas = (W.M + 3) / 4;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) ctas, (unsigned) X.N); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;          // PDL: may start while the previous kernel of the stream drains
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = getenv("B200_NO_PDL") ? 0 : 1;
    B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mmv_fast_kernel<TYPE, NT, J, D, MODE>, W, X, y, y_stride, epi));
}
template <int TYPE, int NT, int J, int D>
static void launch_cfg(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, Epi epi, cudaStream_t stream) {
    if (X.mode == 0) launch_mode<TYPE, NT, J, D, 0>(W, X, y, y_stride, epi, stream);
    else if (X.mode == 1) launch_mode<TYPE, NT, J, D, 1>(W, X, y, y_stride, epi, stream);
    else if (J == 1) launch_mode<TYPE, NT, J, D, J == 1 ? 2 : 1>(W, X, y, y_stride, epi, stream);
    else B200_ASSERT(!"mmv_fast: LayerNorm prologue needs J == 1");
}

template <int TYPE>
static bool launch_type(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, Epi epi, cudaStream_t stream) {
    const int P = W.nb * FX<TYPE>::PPB;
    if (W.K > 64 * 1024) return false;
    if (X.mode == 2 && P > 512) return false;                 // the fused LayerNorm needs the whole row inside one CTA pass (J == 1)
    if (P <= 256) { if (getenv("B200_MMV_D6")) launch_cfg<TYPE, 256, 1, 6>(W, X, y, y_stride, epi, stream); else launch_cfg<TYPE, 256, 1, 8>(W, X, y, y_stride, epi, stream); }
    else if (P <= 512) launch_cfg<TYPE, 512, 1, 8>(W, X, y, y_stride, epi, stream);
    else if (P <= 1024) launch_cfg<TYPE, 512, 2, 4>(W, X, y, y_stride, epi, stream);
    else if (P <= 2048) launch_cfg<TYPE, 512, 4, 2>(W, X, y, y_stride, epi, stream);
    else return false;
    return true;
}

// returns false if the shape / type is not covered (the caller then uses the generic ring kernel of mmv.cu)
bool launch_mmv_fast_x(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, MmvEpilogue e, cudaStream_t stream) {
    const char * nm = W.M > 40000 ? "mmv_lmhead" : W.K > 16384 ? "mmv_down" : W.M > 16384 ? "mmv_up" : W.M > 8192 ? "mmv_qkv" : "mmv_wo";
    const Epi epi = { e.kind, e.r1, e.r2, b200_trace_slot(nm) };
    switch (W.type) {
        case T_Q4_K: return launch_type<T_Q4_K>(W, X, y, y_stride, epi, stream);
        case T_Q4_0: return launch_type<T_Q4_0>(W, X, y, y_stride, epi, stream);
    }
    return false;
}
bool mmv_fast_supports(int wtype, int K, int mode) {
    if (wtype != T_Q4_K && wtype != T_Q4_0) return false;
    const int P = wtype == T_Q4_K ? K / 32 : K / 32;
    return K <= 64 * 1024 && P <= (mode == 2 ? 512 : 2048);
}
bool launch_mmv_fast(const WPlanes & W, const ActQ & A, float * y, int64_t y_stride, MmvEpilogue e, cudaStream_t stream) {
    FastX X{}; X.mode = 0; X.A = A; X.N = A.N;
    return launch_mmv_fast_x(W, X, y, y_stride, e, stream);
}
