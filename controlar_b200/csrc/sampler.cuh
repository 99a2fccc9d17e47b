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

__device__ __forceinline__ float exp1_from_bits(uint32_t x);
// Exp(1) draw of element i of row b at token `step` (kept out of line: it sits in a fully unrolled loop)
__device__ __noinline__ float exp1_noise(uint32_t seed_lo, uint32_t seed_hi, int i, int b, int step) {
    uint32_t r[4];
    philox4x32_10(seed_lo, seed_hi, (uint32_t)(i >> 2), (uint32_t)b, (uint32_t)step, 0x43415231u, r);
    return exp1_from_bits(r[i & 3]);
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
    int pos_val;            // position when pos_ptr is null (persistent decode kernel)
    float* ssq_rows;        // optional [16]: sum of squares of the written h rows (index = row), else null
    int h_reps; long long h_rep_stride;
    long long* dbg_ts;      // dev (PK_TRACE builds): globaltimer stamps of the sampler's stages, written by thread 0   // extra replicas of h_out (persistent kernel), 0/1 = none
};

template <typename T>
__device__ __forceinline__ void write_next_h(const SampleArgs& a, int b_row, int tok, int pos_next, float* red32) {
    // h = tok_embeddings[tok] (+ cs * ctrl0[b][pos_next - T + 1])    gpt_t2i.py:445,466
    const T* e = (const T*)a.tok_emb + (size_t)tok * a.d;
    T* h = (T*)a.h_out + (size_t)b_row * a.d;
    const int p = pos_next - a.T + 1;
    const T* c = (a.ctrl0 && p >= 0 && p < a.n_img) ? (const T*)a.ctrl0 + ((size_t)b_row * a.n_img + p) * a.d : nullptr;
    float ss = 0.f;
    for (int k = threadIdx.x; k < a.d; k += blockDim.x) {
        float v = tof(e[k]);
        if (c) v = rnd<T>(v + rnd<T>(a.cs * tof(c[k])));
        const T hv = fromf<T>(v);
        h[k] = hv;
        for (int rep = 1; rep < a.h_reps; ++rep) h[(size_t)rep * a.h_rep_stride + k] = hv;
        ss += v * v;
    }
    if (a.ssq_rows) {     // deterministic block reduction (fixed thread->element map, fixed tree)
        ss = warp_sum(ss);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red32[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red32[w];
            a.ssq_rows[b_row] = t;
        }
    }
}

// Block-wide deterministic reductions (fixed tree): `red` holds THREADS/32 slots.
template <int THREADS> __device__ __forceinline__ float smp_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += red[w];
    return s;
}
template <int THREADS> __device__ __forceinline__ float smp_block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = red[0];
#pragma unroll
    for (int w = 1; w < THREADS / 32; ++w) s = fmaxf(s, red[w]);
    return s;
}
template <int THREADS> __device__ __forceinline__ unsigned smp_block_count(unsigned v, unsigned* red) {
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += red[w];
    return s;
}

// One CTA per image (b).  The row lives in registers: thread t owns elements t, t+THREADS, ... (EPT of them).
//   top-k : exact k-th largest by a two-level selection (2048-bin histogram over a monotone linear map of the values, then exact
//           ranks of the few candidates of the boundary bin on the order-preserving integer keys); ties at the threshold are kept
//   soft-max / nucleus / race only touch the kept elements (k of V), Philox is evaluated per kept element.
#ifdef PK_TRACE
#define SMP_STAMP(k) do { if (a.dbg_ts != nullptr && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); a.dbg_ts[k] = t_; } } while (0)
#else
#define SMP_STAMP(k) do { } while (0)
#endif

constexpr int SMP_SCRATCH = 2304 * 8;     // bytes of shared scratch sample_body needs (selection histogram + candidates, later the compacted race list)

template <int THREADS, int EPT>
__device__ __forceinline__ void sample_body(const SampleArgs& a, const int b, unsigned char* scratch) {
    __shared__ float red_f[THREADS / 32];
    __shared__ unsigned red_u[THREADS / 32];
    __shared__ int red_i[THREADS / 32];
    __shared__ int s_tok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    const int pos = a.pos_ptr ? ld_cg(a.pos_ptr) : a.pos_val;
    const int step = a.pos_ptr ? (pos - a.T + 1) : a.step;   // index of the token being produced
    bool cfg_on = a.cfg_on != 0;
    if (a.cfg_interval > -1 && step - 1 > a.cfg_interval) cfg_on = false;

    SMP_STAMP(0);
    // ---- CFG combine + temperature (generate.py:103-107, :60)
    float zr[EPT];
    const float* lc = a.logits + (size_t)b * V;
    const float* lu = a.logits + (size_t)(b + a.B) * V;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        const int i = tid + j * THREADS;
        float v = -INFINITY;
        if (i < V) {
            v = __ldcg(lc + i);
            if (a.use_cfg && cfg_on) { const float u = __ldcg(lu + i); v = u + (v - u) * a.cfg_scale; }
            v *= a.inv_temp;
        }
        zr[j] = v;
    }
    SMP_STAMP(1);
    // ---- top-k threshold (ties at the threshold are kept, generate.py:37).  The row is turned into its order-preserving
    // integer keys in place (the map is a bijection), bisected, and turned back — one register array, not two.
    if (a.top_k > 0 && a.top_k < V) {
        // Exact radix-style selection in two levels.  Level 1: a 2048-bin histogram over a MONOTONE linear map of the value range
        // (bin = int((z - lo) * scale): rounding, scaling and truncation are all monotone, so every element of a higher bin is
        // strictly larger than every element of a lower one; linear bins spread a bell-shaped row over ~all bins, where bins
        // on the raw float bits would pile it into a handful and serialise the shared-memory atomics).  Level 2: the bin that
        // contains the k-th largest element holds a few dozen candidates; their exact rank is counted on the order-preserving
        // integer keys.  Falls back to the bit-wise bisection when that bin is crowded (degenerate rows).
        constexpr int NBIN = 2048, BPT = NBIN / THREADS, NCAND = 1024;
        static_assert(NBIN % THREADS == 0, "bins per thread");
        static_assert((NBIN + NCAND) * 4 <= SMP_SCRATCH, "selection scratch");
        unsigned* const s_hist = reinterpret_cast<unsigned*>(scratch);
        unsigned* const s_cand = s_hist + NBIN;
        __shared__ unsigned s_cnt, s_bin, s_krem, s_thr;
        float lo = INFINITY, hi = -INFINITY;
#pragma unroll
        for (int j = 0; j < EPT; ++j) if (tid + j * THREADS < V) { lo = fminf(lo, zr[j]); hi = fmaxf(hi, zr[j]); }
        hi = smp_block_max<THREADS>(hi, red_f);
        lo = -smp_block_max<THREADS>(-lo, red_f);
        const float scale = hi > lo ? (float)(NBIN - 1) / (hi - lo) : 0.f;
#pragma unroll
        for (int i = 0; i < BPT; ++i) s_hist[tid * BPT + i] = 0u;
        if (tid == 0) { s_cnt = 0u; s_thr = 0u; }
        __syncthreads();
        int mybin[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            mybin[j] = min(max((int)((zr[j] - lo) * scale), 0), NBIN - 1);
            if (tid + j * THREADS < V) atomicAdd(&s_hist[mybin[j]], 1u);
        }
        __syncthreads();
        SMP_STAMP(5);
        {   // bin (from the top) in which the cumulative count reaches k
            unsigned loc[BPT], mine = 0u;
#pragma unroll
            for (int i = 0; i < BPT; ++i) { loc[i] = s_hist[tid * BPT + i]; mine += loc[i]; }
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += v; }
            if (lane == 0) red_u[warp] = incl;
            __syncthreads();
            unsigned above = incl - mine;
            for (int w = warp + 1; w < THREADS / 32; ++w) above += red_u[w];
            const unsigned kk = (unsigned)a.top_k;
#pragma unroll
            for (int i = BPT - 1; i >= 0; --i) {
                if (above < kk && kk <= above + loc[i]) { s_bin = (unsigned)(tid * BPT + i); s_krem = kk - above; }
                above += loc[i];
            }
        }
        __syncthreads();
        SMP_STAMP(6);
        const int bsel = (int)s_bin;
        const unsigned krem = s_krem;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            if (tid + j * THREADS < V && mybin[j] == bsel) {
                const unsigned pc = atomicAdd(&s_cnt, 1u);
                if (pc < (unsigned)NCAND) s_cand[pc] = float_order_key(zr[j]);
            }
        }
        __syncthreads();
        SMP_STAMP(7);
        const unsigned ncand = s_cnt;
        unsigned cand = 0u;
        if (ncand <= (unsigned)NCAND) {
            // exact rank inside the bin: the k_rem-th largest candidate (duplicates counted) is the one with
            // #{greater} < k_rem <= #{greater or equal}
            for (unsigned c = tid; c < ncand; c += THREADS) {
                const unsigned key = s_cand[c];
                unsigned gt = 0u, ge = 0u;
                for (unsigned q = 0; q < ncand; ++q) { const unsigned kq = s_cand[q]; gt += kq > key ? 1u : 0u; ge += kq >= key ? 1u : 0u; }
                if (gt < krem && krem <= ge) s_thr = key;                  // (ties write the same value)
            }
            __syncthreads();
            cand = s_thr;
        } else {
            // crowded bin: bit-wise bisection on the integer keys of the whole row, two bits per step
            __shared__ unsigned red_c[2][THREADS / 32][3];
#pragma unroll 1
            for (int shift = 30, it = 0; shift >= 0; shift -= 2, ++it) {
                const unsigned t1 = cand | (1u << shift), t2 = cand | (2u << shift), t3 = cand | (3u << shift);
                unsigned c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    const unsigned k = (tid + j * THREADS) < V ? float_order_key(zr[j]) : 0u;
                    c1 += k >= t1 ? 1u : 0u; c2 += k >= t2 ? 1u : 0u; c3 += k >= t3 ? 1u : 0u;
                }
                c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
                if (lane == 0) { red_c[it & 1][warp][0] = c1; red_c[it & 1][warp][1] = c2; red_c[it & 1][warp][2] = c3; }
                __syncthreads();
                unsigned s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                for (int w = 0; w < THREADS / 32; ++w) { s1 += red_c[it & 1][w][0]; s2 += red_c[it & 1][w][1]; s3 += red_c[it & 1][w][2]; }
                const unsigned kk = (unsigned)a.top_k;                 // counts are non-increasing in the threshold
                cand = s3 >= kk ? t3 : (s2 >= kk ? t2 : (s1 >= kk ? t1 : cand));
            }
        }
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const unsigned key = (tid + j * THREADS) < V ? float_order_key(zr[j]) : 0u;
            if (key < cand) zr[j] = -INFINITY;
        }
    }
    SMP_STAMP(2);
    // ---- soft-max over the kept elements
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < EPT; ++j) mx = fmaxf(mx, zr[j]);
    mx = smp_block_max<THREADS>(mx, red_f);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        const float e = zr[j] == -INFINITY ? 0.f : expf(zr[j] - mx);
        zr[j] = e;
        sum += e;
    }
    sum = smp_block_sum<THREADS>(sum, red_f);
    // ---- nucleus (top-p), generate.py:40-55: in descending order a token is removed iff the cumulative probability
    // of the tokens strictly before it exceeds top_p (first always kept).  Equivalently token x is kept iff
    // f(p_x) <= top_p with f(v) = mass of tokens with probability > v; f is a non-increasing step function, so the
    // kept set is {p >= tau*}; tau* is bracketed by 30 bisection steps (tokens within 2^-30 of the boundary count as
    // ties and are kept; torch.sort's order among exact ties is unspecified anyway).
    if (a.top_p < 1.0f) {
        float lo = 0.f, hi = 1.0f;                       // f(lo) > top_p >= f(hi)
        for (int it = 0; it < 30; ++it) {
            const float mid = 0.5f * (lo + hi);
            float ma = 0.f;
#pragma unroll
            for (int j = 0; j < EPT; ++j) { const float p = zr[j] / sum; if (p > mid) ma += p; }
            ma = smp_block_sum<THREADS>(ma, red_f);
            if (ma <= a.top_p) hi = mid; else lo = mid;
        }
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < EPT; ++j) { if (!(zr[j] / sum > lo)) zr[j] = 0.f; s2 += zr[j]; }
        sum = smp_block_sum<THREADS>(s2, red_f);         // soft-max over the kept logits only
    }
    if (a.probs_out) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) { const int i = tid + j * THREADS; if (i < V) a.probs_out[(size_t)b * V + i] = zr[j] / sum; }
    }
    SMP_STAMP(3);
    // ---- draw: arg-max of p (greedy) or of p / q (exponential race); lowest index wins ties
    float best = -1.f; int besti = 0x7fffffff;
    const float* nz = a.noise ? a.noise + ((size_t)(a.noise_per_step ? step : 0) * a.B + b) * V : nullptr;
    bool raced = false;
    if (a.sample_logits && nz == nullptr) {
        // Philox costs ~100 instructions per element and the kept elements (top_k of V) are scattered over all lanes, so the
        // element-order loop below would run the generator in every (warp, j) iteration for a few lanes each.  Compact the
        // kept (index, probability) pairs into shared memory first and race over the dense list: the arg-max with the
        // lowest-index tie-break does not depend on the order, so the token is the same one.
        constexpr int CAP = 2304;                       // top_k <= 2000 plus ties; larger kept sets take the plain loop
        static_assert(CAP * 8 <= SMP_SCRATCH, "race list scratch");
        int* const s_ci = reinterpret_cast<int*>(scratch);               // (the selection scratch is dead: block reductions lie in between)
        float* const s_cp = reinterpret_cast<float*>(scratch + CAP * 4);
        __shared__ unsigned s_cn;
        if (tid == 0) s_cn = 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int i = tid + j * THREADS;
            const bool kept = i < V && zr[j] > 0.f;
            const unsigned bal = __ballot_sync(0xffffffffu, kept);
            if (bal != 0u) {                            // (warp-uniform)
                unsigned base = 0u;
                if (lane == 0) base = atomicAdd(&s_cn, (unsigned)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned at = base + (unsigned)__popc(bal & ((1u << lane) - 1u));
                if (kept && at < (unsigned)CAP) { s_ci[at] = i; s_cp[at] = zr[j] / sum; }
            }
        }
        __syncthreads();
        const unsigned nk = s_cn;
        if (nk <= (unsigned)CAP) {                      // (CTA-uniform)
            raced = true;
            for (unsigned c = tid; c < nk; c += THREADS) {
                const int i = s_ci[c];
                const float sv = s_cp[c] / exp1_noise(a.seed_lo, a.seed_hi, i, b, step);
                if (sv > best || (sv == best && i < besti)) { best = sv; besti = i; }
            }
        }
    }
    if (!raced) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int i = tid + j * THREADS;
            if (i < V && zr[j] > 0.f) {
                float sv = zr[j] / sum;
                if (a.sample_logits) {
                    float q;
                    if (nz) q = nz[i];
                    else q = exp1_noise(a.seed_lo, a.seed_hi, i, b, step);
                    sv = sv / q;
                }
                if (sv > best || (sv == best && i < besti)) { best = sv; besti = i; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __syncthreads();
    if (lane == 0) { red_f[warp] = best; red_i[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
        float bb = red_f[0]; int bi = red_i[0];
        for (int w = 1; w < THREADS / 32; ++w)
            if (red_f[w] > bb || (red_f[w] == bb && red_i[w] < bi)) { bb = red_f[w]; bi = red_i[w]; }
        s_tok = bi;
        if (a.tokens_ld > 0) a.idx_out[(size_t)b * a.tokens_ld + step] = bi;
        else a.idx_out[b] = bi;
        if (a.tok_buf) { a.tok_buf[b] = bi; if (a.use_cfg) a.tok_buf[b + a.B] = bi; }
    }
    __syncthreads();

    SMP_STAMP(4);
    // ---- fused tail: next step's input rows (cond half b, uncond half b+B)
    if (a.h_out) {
        const int tok = s_tok;
        if (a.dtype == CAR_BF16) {
            write_next_h<bf16>(a, b, tok, pos + 1, red_f);
            if (a.use_cfg) write_next_h<bf16>(a, b + a.B, tok, pos + 1, red_f);
        } else {
            write_next_h<float>(a, b, tok, pos + 1, red_f);
            if (a.use_cfg) write_next_h<float>(a, b + a.B, tok, pos + 1, red_f);
        }
    }
    // ---- the last block to finish advances the device-side position
    if (a.done_ctr) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const int old = atomicAdd(a.done_ctr, 1);
            if (old == a.B - 1) {
                *a.done_ctr = 0;
                if (a.pos_ptr) *a.pos_ptr = pos + 1;
            }
        }
    }
}

constexpr int SMP_EPT = 16;                       // V <= SMP_THREADS * SMP_EPT = 16384
__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(SampleArgs a) {
    __shared__ __align__(16) unsigned char smp_scratch[SMP_SCRATCH];
    sample_body<SMP_THREADS, SMP_EPT>(a, blockIdx.x, smp_scratch);
}
