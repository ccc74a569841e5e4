// attention_tc.cu -- causal grouped-query attention for a batch of N > 1 new tokens on the 5th-generation tensor cores.
//
// Same contract as attention.cu / attention_prefill.cu (libfalcon.cpp:2285-2366, ggml.c:12389-12458): scores scaled by
// 1/sqrt(64), the row's GLOBAL maximum subtracted before an fp16-LUT exp, probabilities normalised by (float)(1/sum).
// The global maximum is why this is a TWO-PASS kernel and not an online softmax: pass 1 only finds the row maxima
// (Q K^T on the tensor cores, nothing stored), pass 2 recomputes each score tile, turns it into e = LUT(s - max) --
// which is an fp16 value by construction, so the fp16 operand of the second product is EXACT -- and accumulates
// O += e V in TMEM; O is scaled by 1/sum at the end.  Nothing but Q, K, V is read and only O is written: the
// [rows x T] score matrix of attention_prefill.cu (0.5 GB per layer at 512 x 2048) never exists.
//
// One CTA = 128 query rows of one KV head (row = token * G + head_in_group: the G query heads that share the KV head
// are stacked, so a K / V tile serves all of them) x all visible keys, 128 keys per tile:
//   S[128 x 128] = Q[128 x 64] K^T      tcgen05.mma.kind::f16  M 128, N 128, 4 x K 16     -> TMEM columns [0, 128)
//   O[128 x 64] += P[128 x 128] V       tcgen05.mma.kind::f16  M 128, N  64, 8 x K 16     -> TMEM columns [128, 192)
// Operands are converted fp32 -> fp16 by the CTA's 128 threads straight into the K-major SWIZZLE_128B layout
// (V transposed on the way: the cache is [key][d], the B operand wants [d][key]); threads t and t + 128 own TMEM lane t =
// row t and split its columns (conversion work, softmax and the exp LUT are what bounds this kernel, not the MMAs).
// 82 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM overlap each other's load / MMA / softmax phases.
// Precision: Q, K, V rounded to fp16 (11 bits), fp32 accumulation; P exact.  Tolerance: tests/test_falcon_gpu.py (GEMM path).
#include "kernels.h"

namespace {

constexpr int AT_M = 128, AT_N = 128, AT_D = 64, AT_THREADS = 256;      // two threads per query row: thread t and t + 128 split the columns
constexpr int SQ = 0, SK = 16384, SV = 32768, SP = 49152, SBAR = 81920, SX = 81984;      // byte offsets in the (1024-aligned) shared memory
constexpr size_t AT_SMEM = 1024 + 81984 + 2 * 128 * 4;                                   // SX: per-row exchange between the two column halves

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart (as gemm_tc.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t) 1 << 16;
    d |= (uint64_t) (1024 >> 4) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) 2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t instr_desc_f16(int n) { return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (AT_M >> 4) << 24); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
// table_exp_f16[f16(v)] (ggml.c:4281-4290) with the fast exponential: ex2.approx is within 2 ulp (fp32) of expf, so the fp16-rounded
// result differs from the table only when the exact value sits within ~1e-6 (relative) of an fp16 rounding boundary (about one
// entry in a thousand, by one fp16 ulp) -- inside this path's fp16-operand tolerance; the decode kernels keep expf
__device__ __forceinline__ float exp_lut(float v) { return __half2float(__float2half_rn(__expf(__half2float(__float2half_rn(v))))); }

struct AttnTcArgs {
    const float * qkv; const float * kc; const float * vc; float * out;
    int n_head_kv, G, n_tok, n_past, T, rows;            // T = n_past + n_tok; rows = G * n_tok per KV head
    int64_t qkv_stride, out_stride;
};

// 64 fp32 values (or zeros) -> one 128-byte row of a K-major SWIZZLE_128B tile
// (half h of the row: chunks 4h .. 4h+3, i.e. 32 of the 64 values)
__device__ __forceinline__ void store_row_f16(uint8_t * tile, int r, const float * src, bool valid, int h) {
    uint8_t * row = tile + r * 128;
    const int sw = r & 7;
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
        const int c = 4 * h + cc;
        uint4 h = make_uint4(0, 0, 0, 0);
        if (valid) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(src) + 2 * c), b = __ldg(reinterpret_cast<const float4 *>(src) + 2 * c + 1);
            h = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
        }
        *reinterpret_cast<uint4 *>(row + ((c ^ sw) << 4)) = h;
    }
}

__global__ void __launch_bounds__(AT_THREADS, 2) attention_tc_kernel(const AttnTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint64_t * bar_s = reinterpret_cast<uint64_t *>(smem + SBAR), * bar_pv = bar_s + 1;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(bar_s + 2);
    float * xch = reinterpret_cast<float *>(smem + SX);                   // [2][128]
    const int tt = threadIdx.x, t = tt & 127, hf = tt >> 7, warp = (tt >> 5) & 3;     // row t (= TMEM lane), column half hf, TMEM lane quarter
    const int g = blockIdx.y, r0 = blockIdx.x * AT_M;
    const int row = r0 + t;
    const bool row_ok = row < a.rows;
    const int tok = row / a.G, head = g * a.G + row % a.G;
    const int vis = row_ok ? a.n_past + tok + 1 : 0;                      // causal: keys < vis (ggml.c:12342-12348)
    const int t_last = (min(r0 + AT_M, a.rows) - 1) / a.G;
    const int kmax = a.n_past + t_last + 1, ntiles = (kmax + AT_N - 1) / AT_N;
    const int vis_min = r0 + AT_M <= a.rows ? a.n_past + r0 / a.G + 1 : 0;   // keys visible to EVERY row of the tile (0 if it has padding rows)
    const float scale = 1.0f / sqrtf((float) AT_D);

    if (tt == 0) { mbar_init(bar_s, 1); mbar_init(bar_pv, 1); mbar_fence_init(); }
    if (tt < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    store_row_f16(smem + SQ, t, a.qkv + (size_t) tok * a.qkv_stride + (size_t) head * AT_D, row_ok, hf);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t) (warp * 32) << 16);  // this warp's 32 lanes
    const uint32_t q_addr = smem_u32(smem + SQ), k_addr = smem_u32(smem + SK), v_addr = smem_u32(smem + SV), p_addr = smem_u32(smem + SP);
    const uint32_t id_s = instr_desc_f16(AT_N), id_o = instr_desc_f16(AT_D);
    uint32_t ns = 0, npv = 0;                                              // completed uses of the two barriers (phase = count & 1)

    auto load_k = [&](int k0) {
        const int key = k0 + t;
        store_row_f16(smem + SK, t, a.kc + ((size_t) key * a.n_head_kv + g) * AT_D, key < a.T, hf);
    };
    auto s_mma = [&]() {                                                   // S = Q K^T, thread 0 only
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_D / 16; k++) tc_mma_f16(tmem_base, umma_desc(q_addr + k * 32), umma_desc(k_addr + k * 32), id_s, k != 0);
        tc_commit(bar_s);
    };

    // ---------------------------------------------------------------- pass 1: row maxima
    float m = -INFINITY;
    for (int kt = 0; kt < ntiles; kt++) {
        load_k(kt * AT_N);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tt == 0) s_mma();
        mbar_wait(bar_s, ns & 1); ns++;
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int c = 2 * hf + cc;
            uint32_t v[32];
            tmem_ld32(tmem_row + (uint32_t) (c * 32), v);
#pragma unroll
            for (int j = 0; j < 32; j++) { const float s = __fmul_rn(__uint_as_float(v[j]), scale); if (kt * AT_N + c * 32 + j < vis) m = fmaxf(m, s); }
        }
    }
    xch[hf * 128 + t] = m;                                                 // the row maximum over both column halves
    __syncthreads();
    m = fmaxf(xch[t], xch[128 + t]);

    // ---------------------------------------------------------------- pass 2: e = LUT(s - max), O += e V
    float l = 0.f;
    for (int kt = 0; kt < ntiles; kt++) {
        if (kt > 0) { mbar_wait(bar_pv, npv & 1); npv++; }                 // the previous tile's P V product has read sV / sP
        const int k0 = kt * AT_N, key = k0 + t;
        load_k(k0);
        {   // V tile, transposed: element (d, key) of the [64 x 128] K-major operand = two [64 x 64] sub-tiles of 8 KB
            const float * src = a.vc + ((size_t) key * a.n_head_kv + g) * AT_D;
            const bool ok = key < a.T;
            uint8_t * base = smem + SV + (t >> 6) * 8192 + (t & 7) * 2;
            const int kc8 = (t & 63) >> 3;
#pragma unroll
            for (int cc = 0; cc < 8; cc++) {
                const int c = 8 * hf + cc;                                     // this thread's 32 of the key's 64 dims
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) f = __ldg(reinterpret_cast<const float4 *>(src) + c);
                const float fv[4] = { f.x, f.y, f.z, f.w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int d = 4 * c + i;
                    *reinterpret_cast<__half *>(base + d * 128 + ((kc8 ^ (d & 7)) << 4)) = __float2half_rn(fv[i]);
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tt == 0) s_mma();
        mbar_wait(bar_s, ns & 1); ns++;
        tc_fence_after();
        const bool full = k0 + AT_N <= vis_min;                           // every key of the tile is visible to every row: no mask tests
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int c = 2 * hf + cc;
            uint32_t v[32];
            tmem_ld32(tmem_row + (uint32_t) (c * 32), v);
            uint8_t * prow = smem + SP + (c >> 1) * 16384 + t * 128;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float s = __fmul_rn(__uint_as_float(v[8 * q + j]), scale);
                    e[j] = (full || k0 + c * 32 + 8 * q + j < vis) ? exp_lut(__fsub_rn(s, m)) : 0.f;
                    l += e[j];
                }
                const int ci = (c & 1) * 4 + q;
                *reinterpret_cast<uint4 *>(prow + ((ci ^ (t & 7)) << 4)) = make_uint4(pack_h2(e[0], e[1]), pack_h2(e[2], e[3]), pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tt == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < AT_N / 16; k++)
                tc_mma_f16(tmem_base + AT_N, umma_desc(p_addr + (k >> 2) * 16384 + (k & 3) * 32), umma_desc(v_addr + (k >> 2) * 8192 + (k & 3) * 32), id_o, (kt | k) != 0);
            tc_commit(bar_pv);
        }
    }
    mbar_wait(bar_pv, npv & 1);
    tc_fence_after();
    xch[hf * 128 + t] = l;                                                 // row sum over both column halves, fixed order
    __syncthreads();
    l = xch[t] + xch[128 + t];
    {
        const float inv = (float) (1.0 / (double) l);                      // ggml.c:12427-12449
        float * dst = a.out + (size_t) tok * a.out_stride + (size_t) head * AT_D + hf * 32;
        uint32_t v[32];
        tmem_ld32(tmem_row + (uint32_t) (AT_N + hf * 32), v);              // warp-collective: every lane loads, valid rows store
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(dst + j) = make_float4(__fmul_rn(__uint_as_float(v[j]), inv), __fmul_rn(__uint_as_float(v[j + 1]), inv),
                                                                   __fmul_rn(__uint_as_float(v[j + 2]), inv), __fmul_rn(__uint_as_float(v[j + 3]), inv));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tt < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256) : "memory");
}

} // namespace

extern "C" int b200_mmv_max_n(void);

// false: shape not covered (head_dim != 64) -> the caller uses the CUDA-core kernels of attention_prefill.cu
bool launch_attention_tc(const float * qkv, const float * k_cache, const float * v_cache, float * out, int64_t out_stride,
                         const AttnParams & p, cudaStream_t stream) {
    // Batches small enough for the integer mat-vec path (N <= 8) keep the fp32 CUDA-core attention, so that path stays at the
    // CPU's reassociation level end to end; the tensor-core batch path (fp16 GEMM operands) gets fp16 attention operands.
    if (p.head_dim != AT_D || p.n_past_dev != nullptr || (p.n_tok <= b200_mmv_max_n() && !getenv("B200_ATTN_TC")) || getenv("B200_ATTN_SIMT")) return false;
    if ((p.qkv_stride % 4) != 0 || (out_stride % 4) != 0) return false;
    AttnTcArgs a;
    a.qkv = qkv; a.kc = k_cache; a.vc = v_cache; a.out = out;
    a.n_head_kv = p.n_head_kv; a.G = p.n_head / p.n_head_kv; a.n_tok = p.n_tok; a.n_past = p.n_past; a.T = p.n_past + p.n_tok;
    a.rows = a.G * p.n_tok; a.qkv_stride = p.qkv_stride; a.out_stride = out_stride;
    static bool set = false;
    if (!set) { B200_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) AT_SMEM)); set = true; }
    dim3 grid((unsigned) ((a.rows + AT_M - 1) / AT_M), (unsigned) p.n_head_kv);
    attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(a);
    B200_CUDA_CHECK(cudaGetLastError());
    return true;
}
