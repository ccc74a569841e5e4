// sampling.cu -- the reference's default sampling chain on the device, so that only a token id leaves the GPU per step.
//
// What falcon_main does on the host with the 260 KB logits row of every token (examples/falcon/falcon_main.cpp:945-975):
//   llama_sample_repetition_penalty  (libfalcon.cpp:3281-3307)  logits of ids in the last-n window: l <= 0 ? l * penalty : l / penalty
//   temp <= 0 : llama_sample_token_greedy (:3433-3447)          first maximum
//   else        llama_sample_top_k (:3094-3118)  ->  llama_sample_top_p (:3121-3150, including its "last_idx = i" cut, which drops the
//               candidate that crosses p)  ->  llama_sample_temperature (:3269-3279)  ->  llama_sample_token (:3449-3468):
//               softmax in fp32, then std::discrete_distribution over the probabilities driven by the context's std::mt19937
// is restated here bit for bit, INCLUDING the random draw: the kernel carries the MT19937 state (seeded like std::mt19937(seed)),
// builds the double-precision cumulative table of libstdc++'s discrete_distribution (normalise by the double sum, partial sums, last
// entry forced to 1.0), draws generate_canonical<double, 53> from two 32-bit outputs and takes the lower bound -- so a run with the
// same seed samples the same ids as the reference, up to expf differing by an ulp between the device and glibc (a draw within
// ~1e-7 of a table boundary).  Not implemented (all off by default in the reference): tail-free, typical, mirostat, frequency /
// presence penalties, logit bias.
// One CTA of 1024 threads: penalty + k rounds of block arg-max (k <= 1024, default 40) + a sequential tail on thread 0.
#include "kernels.h"

struct SamplerState {                   // device resident
    uint32_t mt[624]; int mti;
    int32_t window[B200_SAMPLER_MAX_WINDOW]; int wlen, wcap;      // the last-n ids the repetition penalty looks at (ring, oldest first when read from wpos)
    int wpos;
};

namespace {

__device__ uint32_t mt_next(SamplerState * s) {                       // MT19937 (std::mt19937): regenerate every 624 outputs, then temper
    if (s->mti >= 624) {
        for (int i = 0; i < 624; i++) {
            const uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->mti = 0;
    }
    uint32_t y = s->mt[s->mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

__global__ void sampler_init_kernel(SamplerState * s, uint32_t seed, const int32_t * window, int n, int cap) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t) i;
    s->mti = 624;
    s->wcap = cap; s->wlen = n < cap ? n : cap; s->wpos = 0;
    for (int i = 0; i < s->wlen; i++) s->window[i] = window[n - s->wlen + i];
    if (s->wlen == cap) s->wpos = 0; else s->wpos = s->wlen;          // next slot to write
}

__global__ void __launch_bounds__(1024) sample_kernel(const float * __restrict__ logits, int n_vocab, SamplerParams p, SamplerState * st,
                                                      float * __restrict__ work, int32_t * out, int32_t * hist, int * step) {
    __shared__ float sv[32]; __shared__ int si[32];
    __shared__ float cand_l[1024]; __shared__ int cand_id[1024];
    __shared__ int32_t win[B200_SAMPLER_MAX_WINDOW]; __shared__ int s_wlen;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_wlen = (p.repeat_penalty != 1.0f) ? st->wlen : 0;
    __syncthreads();
    const int wlen = s_wlen;
    for (int i = tid; i < wlen; i += 1024) win[i] = st->window[i];
    __syncthreads();
    // ---- repetition penalty while copying the row
    for (int i = tid; i < n_vocab; i += 1024) {
        float l = logits[i];
        bool hit = false;
        for (int j = 0; j < wlen; j++) hit |= win[j] == i;
        if (hit) l = l <= 0.f ? __fmul_rn(l, p.repeat_penalty) : __fdiv_rn(l, p.repeat_penalty);
        work[i] = l;
    }
    __syncthreads();
    const int k = p.temp <= 0.f ? 1 : p.top_k;
    // ---- the k largest logits, descending (equal values: lowest id first), by k rounds of block arg-max
    for (int r = 0; r < k; r++) {
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < n_vocab; i += 1024) { const float v = work[i]; if (v > best || (v == best && i < bi)) { best = v; bi = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = best; si[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = sv[lane]; bi = si[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { cand_l[r] = best; cand_id[r] = bi; if (bi < n_vocab) work[bi] = -INFINITY; }
        }
        __syncthreads();
    }
    if (tid != 0) return;
    // ---- sequential tail (k values)
    int n = k, pick = 0;
    if (p.temp > 0.f) {
        if (p.top_p < 1.0f) {                                          // llama_sample_top_p: softmax, cumulative sum, cut
            const float max_l = cand_l[0];
            float cum = 0.f;
            for (int i = 0; i < n; i++) cum = __fadd_rn(cum, expf(__fsub_rn(cand_l[i], max_l)));
            float run = 0.f; int last = n;
            for (int i = 0; i < n; i++) {
                run = __fadd_rn(run, __fdiv_rn(expf(__fsub_rn(cand_l[i], max_l)), cum));
                if (run > p.top_p && i >= 1) { last = i; break; }
            }
            n = last;
        }
        for (int i = 0; i < n; i++) cand_l[i] = __fdiv_rn(cand_l[i], p.temp);       // llama_sample_temperature
        if (n >= 2) {                                                  // llama_sample_token: softmax, discrete_distribution(probs)(rng)
            const float max_l = cand_l[0];
            float cum = 0.f;
            for (int i = 0; i < n; i++) cum = __fadd_rn(cum, expf(__fsub_rn(cand_l[i], max_l)));
            double sum = 0.0;
            for (int i = 0; i < n; i++) { cand_l[i] = __fdiv_rn(expf(__fsub_rn(cand_l[i], max_l)), cum); sum += (double) cand_l[i]; }
            const uint32_t x0 = mt_next(st), x1 = mt_next(st);
            double u = ((double) x0 + (double) x1 * 4294967296.0) / 18446744073709551616.0;      // generate_canonical<double, 53>
            if (u >= 1.0) u = 0.99999999999999988898;                  // nextafter(1.0, 0.0)
            double acc = 0.0;
            pick = n - 1;
            for (int i = 0; i < n; i++) {
                acc += (double) cand_l[i] / sum;
                const double cp = i == n - 1 ? 1.0 : acc;              // the table's last entry is forced to 1.0
                if (!(cp < u)) { pick = i; break; }                    // std::lower_bound: first entry >= u
            }
        }
    }
    const int id = cand_id[pick];
    *out = id;
    if (hist) { const int s = *step; hist[s] = id; *step = s + 1; }
    if (st->wcap > 0) {                                                // the window slides: drop the oldest id, append the new one
        if (st->wlen < st->wcap) st->window[st->wlen++] = id;
        else { for (int i = 1; i < st->wcap; i++) st->window[i - 1] = st->window[i]; st->window[st->wcap - 1] = id; }
    }
}

} // namespace

SamplerState * sampler_state_alloc() { SamplerState * s = nullptr; B200_CUDA_CHECK(cudaMalloc(&s, sizeof(SamplerState))); return s; }
void sampler_state_free(SamplerState * s) { if (s) B200_CUDA_CHECK(cudaFree(s)); }
// window: the n most recent token ids (oldest first) the repetition penalty starts from; at most `cap` (= repeat_last_n) are kept
void launch_sampler_init(SamplerState * s, uint32_t seed, const int32_t * window_dev, int n, int cap, cudaStream_t stream) {
    B200_ASSERT(cap >= 0 && cap <= B200_SAMPLER_MAX_WINDOW);
    sampler_init_kernel<<<1, 1, 0, stream>>>(s, seed, window_dev, n, cap);
    B200_CUDA_CHECK(cudaGetLastError());
}
void launch_sample(const float * logits, int n_vocab, const SamplerParams & p, SamplerState * st, float * work, int32_t * out, int32_t * hist, int * step, cudaStream_t stream) {
    B200_ASSERT(p.top_k >= 1 && p.top_k <= 1024);
    sample_kernel<<<1, 1024, 0, stream>>>(logits, n_vocab, p, st, work, out, hist, step);
    B200_CUDA_CHECK(cudaGetLastError());
}
