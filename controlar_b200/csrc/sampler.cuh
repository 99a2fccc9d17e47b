// sampler.cuh — CFG combine + temperature + top-k/top-p + soft-max + multinomial/arg-max in ONE kernel,
// plus the embedding gather (+ layer-0 control add) of the *next* decode step.
// Replaces (reference file:line):
//   CFG combine                 autoregressive/models/generate.py:89-90,103-107
//   sample()                    generate.py:59-74
//   top_k_top_p_filtering()     generate.py:17-56
//   tok_embeddings(idx)         gpt_t2i.py:445     } fused tail: h for the next position
//   h += cs*ctrl[0][:, p+1]     gpt_t2i.py:466     }
// torch.multinomial(p, 1) == argmax(p / q), q ~ Exp(1) (SURVEY.md §7 hard-part 4).  q comes either from a
// caller-provided buffer (parity tests) or from Philox4x32-10 keyed by (seed; step, row, index).
#pragma once
#include "common.cuh"

constexpr int SMP_THREADS = 1024;

__device__ __forceinline__ uint32_t float_order_key(float f) {   // larger float -> larger key
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                              uint32_t c3, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float exp1_from_bits(uint32_t x) {
    // u in (0,1]: (x + 1) * 2^-32 ;  q = -log(u) ~ Exp(1)
    const float u = ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    return -logf(u);
}

struct SampleArgs {
    const float* logits;    // [b_eff, V]
    int V; int B;           // B images; b_eff = 2B when cfg
    int use_cfg; int cfg_on; float cfg_scale;
    float inv_temp; int top_k; float top_p; int sample_logits;
    const float* noise;     // [B, V] (or [steps, B, V] when noise_per_step) or null
    int noise_per_step;
    uint32_t seed_lo, seed_hi; int step;        // Philox sub-stream = index of the token being produced
    int cfg_interval;       // device-side cfg_flag: off when step-1 > cfg_interval >= 0  (generate.py:121-122)
    int* idx_out;           // [B] (or tokens_out + step when tokens_ld > 0)
    int tokens_ld;
    float* probs_out;       // [B, V] or null
    // fused next-step embedding (decode loop only; null => skip)
    void* h_out; const void* tok_emb; const void* ctrl0; int d; int n_img; int T; float cs; int dtype;
    int* tok_buf;           // [b_eff] int32 tokens consumed by teacher-free decode
    int* pos_ptr;           // device scalar: position of the token being produced is *pos_ptr + 1
    int* done_ctr;          // ticket: the last block to finish advances *pos_ptr
};

template <typename T>
__device__ __forceinline__ void write_next_h(const SampleArgs& a, int b_row, int tok, int pos_next) {
    // h = tok_embeddings[tok] (+ cs * ctrl0[b][pos_next - T + 1])    gpt_t2i.py:445,466
    const T* e = (const T*)a.tok_emb + (size_t)tok * a.d;
    T* h = (T*)a.h_out + (size_t)b_row * a.d;
    const int p = pos_next - a.T + 1;
    const T* c = (a.ctrl0 && p >= 0 && p < a.n_img) ? (const T*)a.ctrl0 + ((size_t)b_row * a.n_img + p) * a.d : nullptr;
    for (int k = threadIdx.x; k < a.d; k += blockDim.x) {
        float v = tof(e[k]);
        if (c) v = rnd<T>(v + rnd<T>(a.cs * tof(c[k])));
        h[k] = fromf<T>(v);
    }
}

// one CTA per image
__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(SampleArgs a) {
    extern __shared__ float z[];                  // [V]
    __shared__ unsigned int hist[256];
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    __shared__ unsigned int sel_prefix, sel_k;
    __shared__ float s_thr;
    __shared__ int s_tok;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    const int pos = a.pos_ptr ? ld_cg(a.pos_ptr) : 0;
    const int step = a.pos_ptr ? (pos - a.T + 1) : a.step;   // index of the token being produced
    bool cfg_on = a.cfg_on != 0;
    if (a.cfg_interval > -1 && step - 1 > a.cfg_interval) cfg_on = false;

    // ---- CFG combine + temperature
    const float* lc = a.logits + (size_t)b * V;
    const float* lu = a.logits + (size_t)(b + a.B) * V;
    for (int i = tid; i < V; i += SMP_THREADS) {
        float v = __ldcg(lc + i);
        if (a.use_cfg && cfg_on) { const float u = __ldcg(lu + i); v = u + (v - u) * a.cfg_scale; }
        z[i] = v * a.inv_temp;
    }
    __syncthreads();

    // ---- top-k: k-th largest by 4x8-bit radix select; keep z >= thr (ties kept, generate.py:37)
    if (a.top_k > 0 && a.top_k < V) {
        if (tid == 0) { sel_prefix = 0; sel_k = (unsigned)a.top_k; }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < 256; i += SMP_THREADS) hist[i] = 0;
            __syncthreads();
            const unsigned prefix = sel_prefix;
            const unsigned pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (int i = tid; i < V; i += SMP_THREADS) {
                const unsigned key = float_order_key(z[i]);
                if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned k = sel_k, bin = 255;
                for (;; --bin) {
                    if (hist[bin] >= k) break;
                    k -= hist[bin];
                    if (bin == 0) break;
                }
                sel_k = k;
                sel_prefix = prefix | (bin << shift);
            }
            __syncthreads();
        }
        if (tid == 0) {
            const unsigned key = sel_prefix;
            const unsigned u = (key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key;
            s_thr = __uint_as_float(u);
        }
        __syncthreads();
        const float thr = s_thr;
        for (int i = tid; i < V; i += SMP_THREADS) if (z[i] < thr) z[i] = -INFINITY;
        __syncthreads();
    }

    // ---- soft-max (max, sum of exp)
    float mx = -INFINITY;
    for (int i = tid; i < V; i += SMP_THREADS) mx = fmaxf(mx, z[i]);
    mx = warp_max(mx);
    if (lane == 0) red_f[warp] = mx;
    __syncthreads();
    mx = red_f[0];
    for (int w = 1; w < SMP_THREADS / 32; ++w) mx = fmaxf(mx, red_f[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < V; i += SMP_THREADS) { const float e = expf(z[i] - mx); z[i] = e; sum += e; }
    sum = warp_sum(sum);
    if (lane == 0) red_f[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < SMP_THREADS / 32; ++w) sum += red_f[w];
    __syncthreads();

    // ---- nucleus (top-p), generate.py:40-55: in descending order a token is removed iff the cumulative
    // probability of the tokens strictly before it exceeds top_p (first always kept).  Equivalently token x is
    // kept iff f(p_x) <= top_p with f(v) = mass of tokens with probability > v; f is a non-increasing step
    // function, so the kept set is {p >= tau*}; tau* is bracketed by 30 bisection steps (tokens within 2^-30 of
    // the boundary count as ties and are kept; torch.sort's order among exact ties is unspecified anyway).
    float keep_above = -1.f;
    if (a.top_p < 1.0f) {
        float lo = 0.f, hi = 1.0f;                       // f(lo) > top_p >= f(hi)
        for (int it = 0; it < 30; ++it) {
            const float mid = 0.5f * (lo + hi);
            float ma = 0.f;
            for (int i = tid; i < V; i += SMP_THREADS) { const float p = z[i] / sum; if (p > mid) ma += p; }
            ma = warp_sum(ma);
            if (lane == 0) red_f[warp] = ma;
            __syncthreads();
            ma = 0.f;
            for (int w = 0; w < SMP_THREADS / 32; ++w) ma += red_f[w];
            __syncthreads();
            if (ma <= a.top_p) hi = mid; else lo = mid;
        }
        keep_above = lo;
        float s2 = 0.f;
        for (int i = tid; i < V; i += SMP_THREADS) {
            if (!(z[i] / sum > keep_above)) z[i] = 0.f;
            s2 += z[i];
        }
        s2 = warp_sum(s2);
        if (lane == 0) red_f[warp] = s2;
        __syncthreads();
        s2 = 0.f;
        for (int w = 0; w < SMP_THREADS / 32; ++w) s2 += red_f[w];
        __syncthreads();
        sum = s2;                                        // soft-max over the kept logits only
    }
    for (int i = tid; i < V; i += SMP_THREADS) z[i] = z[i] / sum;
    __syncthreads();
    if (a.probs_out) for (int i = tid; i < V; i += SMP_THREADS) a.probs_out[(size_t)b * V + i] = z[i];

    // ---- draw: arg-max of p (greedy) or of p / q (exponential race); lowest index wins ties
    float best = -1.f; int besti = 0x7fffffff;
    const float* nz = a.noise ? a.noise + ((size_t)(a.noise_per_step ? step : 0) * a.B + b) * V : nullptr;
    if (a.sample_logits && a.noise == nullptr) {
        for (int i4 = tid; i4 < (V + 3) / 4; i4 += SMP_THREADS) {
            uint32_t r[4];
            philox4x32_10(a.seed_lo, a.seed_hi, (uint32_t)i4, (uint32_t)b, (uint32_t)step, 0x43415231u, r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i4 * 4 + j;
                if (i < V) {
                    const float s = z[i] / exp1_from_bits(r[j]);
                    if (s > best) { best = s; besti = i; }
                }
            }
        }
    } else {
        for (int i = tid; i < V; i += SMP_THREADS) {
            float s = z[i];
            if (a.sample_logits) s = s / nz[i];
            if (s > best) { best = s; besti = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) { red_f[warp] = best; red_i[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
        float bb = red_f[0]; int bi = red_i[0];
        for (int w = 1; w < SMP_THREADS / 32; ++w)
            if (red_f[w] > bb || (red_f[w] == bb && red_i[w] < bi)) { bb = red_f[w]; bi = red_i[w]; }
        s_tok = bi;
        if (a.tokens_ld > 0) a.idx_out[(size_t)b * a.tokens_ld + step] = bi;
        else a.idx_out[b] = bi;
        if (a.tok_buf) { a.tok_buf[b] = bi; if (a.use_cfg) a.tok_buf[b + a.B] = bi; }
    }
    __syncthreads();

    // ---- fused tail: next step's input rows (cond half b, uncond half b+B)
    if (a.h_out) {
        const int tok = s_tok;
        if (a.dtype == CAR_BF16) {
            write_next_h<bf16>(a, b, tok, pos + 1);
            if (a.use_cfg) write_next_h<bf16>(a, b + a.B, tok, pos + 1);
        } else {
            write_next_h<float>(a, b, tok, pos + 1);
            if (a.use_cfg) write_next_h<float>(a, b + a.B, tok, pos + 1);
        }
    }
    // ---- the last block to finish advances the device-side position / step counters
    if (a.done_ctr) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const int old = atomicAdd(a.done_ctr, 1);
            if (old == (int)gridDim.x - 1) {
                *a.done_ctr = 0;
                if (a.pos_ptr) *a.pos_ptr = pos + 1;
            }
        }
    }
}
