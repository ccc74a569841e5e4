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

#include "mmv_fast.cuh"

// D = rows in flight per thread (register ring), reduced G = min(D, 4) rows at a time.  D * J = 8 pieces = 192 B in
// flight per thread at all times (>= 96 KB per SM): the ring is refilled one row at a time, right after that row's
// slot has been consumed, so the depth never drops while a group is being computed.
// `cta` of `nctas` CTAs work on this matrix (a launch may carry two matrices, see mmv_fast2_kernel)
template <int TYPE, int NT, int J, int D, int MODE>
__device__ __forceinline__ void mmv_body(const WPlanes & W, const FastX & X, float * __restrict__ y, int64_t y_stride, const Epi & epi,
                                         const int cta, const int nctas, const int n, uint8_t * smem) {
    using T = FX<TYPE>;
    constexpr int NW = NT / 32;
    uint64_t * bar = reinterpret_cast<uint64_t *>(smem);
    float * partial = reinterpret_cast<float *>(smem + 16);           // [2][NW][G]
    double * red = reinterpret_cast<double *>(smem + 16 + 2 * NW * 4 * 4);      // [NW] block reduction scratch (mode 2)
    int8_t * xq = reinterpret_cast<int8_t *>(smem + ((16 + 2 * NW * 4 * 4 + NW * 8 + 15) & ~15));      // 16-byte aligned (TMA destination)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = W.nb * T::PPB;
    const ActQ & A = X.A;

    trace_begin(epi.trace);
    if (tid == 0 && MODE == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // rows [row0, row1) of this CTA, balanced to +-1
    const int per = W.M / nctas, rem = W.M % nctas;
    const int row0 = cta * per + min(cta, rem), row1 = row0 + per + (cta < rem ? 1 : 0);
    const int nrows = row1 - row0;

    int gidx[J]; bool valid[J];
    WP wp[J];
    int aux[J];
#pragma unroll
    for (int j = 0; j < J; j++) { const int g = j * NT + tid; valid[j] = g < P; gidx[j] = valid[j] ? g : P - 1; wp[j] = T::wp(W, gidx[j]); aux[j] = piece_aux<TYPE>(gidx[j]); }

    typename T::WR w[D][J];
    ring_fill<TYPE, J, D>(w, wp, row0, row1);                                // weights are in flight before the activation arrives
    const L2PF pf = l2pf_of(W, X.l2_dist);
    if (pf.dist > 0 && tid == 0) l2_prefetch_rows(pf, min(row0 + D, row1), min(row0 + D + pf.dist, row1));

    // everything above touched only weights; the activation row is produced by the previous kernel(s) of the stream
    if (!epi.late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    typename T::XR xr[J];
    if (MODE == 0) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t) W.K);
            tma_load_1d(xq, A.q + (size_t) n * W.K, (uint32_t) W.K, bar);   // activation codes: global -> shared by TMA
        }
        __syncthreads();                                                      // barrier initialised (thread 0 did it before issuing)
        mbar_wait(bar, 0);
#pragma unroll
        for (int j = 0; j < J; j++) xr[j] = valid[j] ? T::load_x(xq, A, n, gidx[j]) : zero_xr<TYPE>();
    } else if constexpr (T::HAS_PROLOGUE_QUANT) {
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
                if (X.x_out && cta == 0 && valid[j]) {
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
            xr[j] = T::quant_x(v, gidx[j], lane);       // an all-zero piece quantises to an all-zero XR: it adds exactly 0 to every row
        }
        __syncthreads();                                                      // `partial` / `red` are reused below
    }

    // Programmatic dependent launch: the next kernel of the stream is released once this CTA is about to issue its last
    // weight loads.  Triggering earlier would park the dependent grid's CTAs at the head of the hardware queue for the
    // whole duration of this kernel and keep the small attention kernels of the other stream from being scheduled
    // (measured: profiles/r1_decode_timeline.md).
    int gcount = 0;
    const int ekind = epi.kind; const float * r1p = epi.r1, * r2p = epi.r2;
    ring_run<TYPE, NT, J, D, false>(w, wp, row0, row1, wp, 0, 0, xr, aux, partial, gcount, 0, tid,
        [&](int row, float v) {
            if (ekind == EPI_GELU) v = gelu_lut(v);
            else if (ekind == EPI_ADD2) v = (v + r1p[(size_t) n * y_stride + row]) + r2p[(size_t) n * y_stride + row];
            y[(size_t) n * y_stride + row] = v;
        },
        [&]() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }, pf);
    if (epi.late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");      // see MmvEpilogue::late_wait
    if (epi.qctr) {
        // Quantise the finished output row for the next mat-mul, 256 values at a time.  A chunk's rows belong to up to
        // three CTAs; each adds its row count to the chunk's counter once its own rows are stored and fenced, and the CTA
        // that completes the count quantises the chunk (one warp) and re-arms the counter for the next launch.
        __shared__ int s_do;
        __threadfence();
        __syncthreads();
        if (nrows > 0) {
            for (int b = row0 >> 8; b <= (row1 - 1) >> 8; b++) {
                if (tid == 0) {
                    const unsigned cnt = (unsigned) (min(row1, (b + 1) << 8) - max(row0, b << 8));
                    const unsigned old = atomicAdd(epi.qctr + b, cnt);
                    s_do = old + cnt == 256u;
                    if (s_do) epi.qctr[b] = 0;
                }
                __syncthreads();
                if (s_do && warp == 0) {
                    __threadfence();
                    const float * src = y + (size_t) n * y_stride + (b << 8) + lane * 8;
                    const float4 p = __ldcg(reinterpret_cast<const float4 *>(src)), q = __ldcg(reinterpret_cast<const float4 *>(src) + 1);
                    const float v[8] = { p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w };
                    if (epi.qA.type == T_Q8_K) quantize_chunk8<T_Q8_K>(v, lane, epi.qA, n, (b << 8) + lane * 8);
                    else quantize_chunk8<T_Q8_0>(v, lane, epi.qA, n, (b << 8) + lane * 8);
                }
                __syncthreads();
            }
        }
    }
    trace_end(epi.trace);
}

template <int TYPE, int NT, int J, int D, int MODE>
// 96 registers for the 256-thread shapes: two CTAs per SM leave a quarter of the register file to the small attention
// kernels of the other stream, which then run beside ffn_up instead of after it
__global__ void __launch_bounds__(NT) __maxnreg__(NT <= 256 ? 96 : 128) mmv_fast_kernel(const WPlanes W, const FastX X, float * __restrict__ y, int64_t y_stride, const Epi epi) {
    extern __shared__ __align__(16) uint8_t smem[];
    mmv_body<TYPE, NT, J, D, MODE>(W, X, y, y_stride, epi, (int) blockIdx.x, (int) gridDim.x, (int) blockIdx.y, smem);
}

static int fast_num_sms() {
    static int n = 0;
    if (!n) { int dev; B200_CUDA_CHECK(cudaGetDevice(&dev)); B200_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); }
    return n;
}

template <int TYPE, int NT, int J, int D, int MODE>
static void launch_mode(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, Epi epi, cudaStream_t stream) {
    const size_t smem = ((16 + 2 * (NT / 32) * 4 * 4 + (NT / 32) * 8 + 15) & ~15) + (size_t) ((W.K + 15) & ~15);
    static bool set = false;
    if (!set) { B200_CUDA_CHECK(cudaFuncSetAttribute(mmv_fast_kernel<TYPE, NT, J, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        // same L1/shared split as the small kernels of the other stream: an SM cannot host kernels with different carve-outs at once
        B200_CUDA_CHECK(cudaFuncSetAttribute(mmv_fast_kernel<TYPE, NT, J, D, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, B200_CARVEOUT)); set = true; }
    int ctas = fast_num_sms() * (NT == 128 ? 4 : NT == 160 || NT == 192 ? 3 : NT == 256 ? 2 : 1);
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
    else if constexpr (FX<TYPE>::HAS_PROLOGUE_QUANT) {
        if (X.mode == 1) launch_mode<TYPE, NT, J, D, 1>(W, X, y, y_stride, epi, stream);
        else if (J == 1) launch_mode<TYPE, NT, J, D, J == 1 ? 2 : 1>(W, X, y, y_stride, epi, stream);
        else B200_ASSERT(!"mmv_fast: LayerNorm prologue needs J == 1");
    } else B200_ASSERT(!"mmv_fast: this weight type takes quantised activations only (mode 0)");
}

template <int TYPE>
static bool launch_type(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, Epi epi, cudaStream_t stream) {
    const int P = W.nb * FX<TYPE>::PPB;
    if (W.K > 64 * 1024) return false;
    if (X.mode == 2 && P > 512) return false;                 // the fused LayerNorm needs the whole row inside one CTA pass (J == 1)
    constexpr int D1 = FX<TYPE>::D256;                        // ring depth for one piece per thread; D * J stays constant
    if (P <= 128 && D1 == 4) launch_cfg<TYPE, 128, 1, D1>(W, X, y, y_stride, epi, stream);       // 64-weight pieces (Q3_K): K = 8192 is 128 pieces
    else if (P > 128 && P <= 160 && D1 == 8 && !getenv("B200_NO_NT160")) launch_cfg<TYPE, 160, 1, D1>(W, X, y, y_stride, epi, stream);   // Falcon-7B: K = 4544 is 142 pieces
    else if (P <= 256) launch_cfg<TYPE, 256, 1, D1>(W, X, y, y_stride, epi, stream);
    // Falcon-180B (K = 14848: 464 pieces): two 256-thread CTAs at 96 registers instead of one 512-thread CTA at 128 leave a quarter of the
    // register file to the attention kernels of the other stream, as the K = 8192 shape does
    else if (P > 256 && P <= 512 && D1 == 8 && X.mode == 0 && !getenv("B200_NO_NT256J2")) launch_cfg<TYPE, 256, 2, D1 / 2>(W, X, y, y_stride, epi, stream);
    else if (P <= 512) launch_cfg<TYPE, 512, 1, D1>(W, X, y, y_stride, epi, stream);
    // Falcon-7B's ffn_down (K = 18176: 568 pieces): 3 pieces per thread of a 192-thread CTA use 568 of 576 slots; the 512 x 2 shape
    // below would leave 45 % of its lanes without a piece (and its 128-register CTAs own the whole register file)
    else if (P > 512 && P <= 576 && D1 == 8 && X.mode == 0 && !getenv("B200_NO_NT192")) launch_cfg<TYPE, 192, 3, 2>(W, X, y, y_stride, epi, stream);
    else if (P <= 1024) launch_cfg<TYPE, 512, 2, D1 / 2>(W, X, y, y_stride, epi, stream);
    else if (P <= 2048 && D1 == 8) launch_cfg<TYPE, 512, 4, 2>(W, X, y, y_stride, epi, stream);
    else return false;
    return true;
}

// returns false if the shape / type is not covered (the caller then uses the generic ring kernel of mmv.cu)
bool launch_mmv_fast_x(const WPlanes & W, const FastX & X, float * y, int64_t y_stride, MmvEpilogue e, cudaStream_t stream) {
    const char * nm = W.M > 40000 ? "mmv_lmhead" : W.K > 16384 ? "mmv_down" : W.M > 16384 ? "mmv_up" : W.M > 8192 ? "mmv_qkv" : "mmv_wo";
    Epi epi = { e.kind, e.r1, e.r2, b200_trace_slot(nm), ActQ{}, nullptr, getenv("B200_NO_LATE_WAIT") ? 0 : e.late_wait };
    if (e.qout && e.qctr) {
        B200_ASSERT(X.N == 1 && W.M % 256 == 0 && e.qout->K == W.M && (e.qout->type == T_Q8_K || e.qout->type == T_Q8_0));
        epi.qA = *e.qout; epi.qctr = e.qctr;
    }
    static int dist_bytes = -1;
    // HBM -> L2 prefetch ahead of the ring: off by default, measured slower at every distance (profiles/r1_decode_timeline.md)
    if (dist_bytes < 0) { const char * s = getenv("B200_L2PF_KB"); dist_bytes = (s ? atoi(s) : 0) * 1024; }
    FastX Xp = X;
    Xp.l2_dist = W.stride[0] ? (int) (dist_bytes / W.stride[0]) : 0;
    switch (W.type) {
        case T_Q4_K: return launch_type<T_Q4_K>(W, Xp, y, y_stride, epi, stream);
        case T_Q4_0: return launch_type<T_Q4_0>(W, Xp, y, y_stride, epi, stream);
        case T_Q3_K: return X.mode == 0 && !getenv("B200_Q3K_GENERIC") && launch_type<T_Q3_K>(W, Xp, y, y_stride, epi, stream);
    }
    return false;
}
// Which (type, K, activation mode) the fused single-stream decode path of engine.cu may rely on.  Q3_K has a fast mat-vec
// (used through launch_mmv) but is NOT listed: its kernel is issue-bound, and the two-stream per-node path, which runs the
// attention and MLP branches' mat-vecs concurrently, is faster for it (189 vs 180 tok/s, Falcon-40B).
bool mmv_fast_supports(int wtype, int K, int mode) {
    if (wtype != T_Q4_K && wtype != T_Q4_0) return false;
    const int P = wtype == T_Q4_K ? K / 32 : K / 32;
    return K <= 64 * 1024 && P <= (mode == 2 ? 512 : 2048);
}
// true when the shape chosen for W (3 CTAs of 192 threads at 96 registers) leaves no room on an SM for a side-stream kernel:
// the decode step then schedules wo BEFORE this mat-vec (engine.cu).  The 512-thread shapes leave ~14k registers: one attention CTA fits.
bool mmv_fast_fills_sm(const WPlanes & W) {
    if ((W.type != T_Q4_K && W.type != T_Q4_0) || getenv("B200_NO_NT192")) return false;
    const int P = W.K / 32;
    return P > 512 && P <= 576;
}
bool launch_mmv_fast(const WPlanes & W, const ActQ & A, float * y, int64_t y_stride, MmvEpilogue e, cudaStream_t stream) {
    FastX X{}; X.mode = 0; X.A = A; X.N = A.N;
    return launch_mmv_fast_x(W, X, y, y_stride, e, stream);
}
