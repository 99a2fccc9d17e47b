// attention.cuh — KV-cache attention of the LlamaGen block.
//   attn_decode : one query row per (sequence, head) against the cache [0, pos]  — the HBM-bound part of a
//                 decode step.  Replaces F.scaled_dot_product_attention with the math backend forced
//                 (autoregressive/models/generate.py:120, gpt_t2i.py:282-286) and the bool mask
//                 causal_mask[:, None, input_pos] edited by generate.py:184-193.
//   attn_prefill: the T prefix rows (causal + emb_mask columns + forced diagonal), gpt_t2i.py:448,282-286.
// Cache layout is the reference's KVCache: [B_eff, H, S, 64] (gpt_t2i.py:220-235).
// The reference soft-maxes over all S slots with -inf on masked ones; skipping masked slots is identical.
#pragma once
#include "common.cuh"

constexpr int AD_WARPS = 4;
constexpr int AD_THREADS = AD_WARPS * 32;
constexpr int AD_PART = 68;   // m, l, pad, pad, acc[64]

template <typename T> struct RowLanes;                 // lanes that share one 64-element row with 16-byte loads
template <> struct RowLanes<bf16> { static constexpr int LPR = 8, EPL = 8; };
template <> struct RowLanes<float> { static constexpr int LPR = 16, EPL = 4; };

template <typename T> __device__ __forceinline__ void load_row_frag(const T* p, float (&f)[RowLanes<T>::EPL]);
template <> __device__ __forceinline__ void load_row_frag<bf16>(const bf16* p, float (&f)[8]) {
    const uint4 v = ldg_cg128(p);
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
}
template <> __device__ __forceinline__ void load_row_frag<float>(const float* p, float (&f)[4]) {
    const uint4 v = ldg_cg128(p);
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}

// grid = (B_eff*H, nsplit), block = 128.  Split s handles keys [s*chunk, min(n,(s+1)*chunk)), n = pos+1.
// The last CTA to finish a (b,h) combines the nsplit partials in index order (deterministic) and writes
// out[b][h*64 + e] rounded to T (the SDPA output cast).
template <typename T>
__global__ void __launch_bounds__(AD_THREADS, 6)
attn_decode_kernel(const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                   const int* __restrict__ emb_mask, int mask_ld, const int* __restrict__ pos_ptr, int H, int S,
                   int Tpre, int nsplit, int flags, float* __restrict__ part, int* __restrict__ tickets, T* __restrict__ out) {
    constexpr int LPR = RowLanes<T>::LPR, EPL = RowLanes<T>::EPL, RPW = 32 / LPR;
    constexpr int UNR = 4;
    __shared__ float sm_m[AD_WARPS * RPW], sm_l[AD_WARPS * RPW];
    __shared__ float sm_acc[AD_WARPS * RPW][64];
    __shared__ int sm_last;

    const int bh = blockIdx.x, b = bh / H, hd = bh - b * H;
    const int split = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane / LPR, cl = lane % LPR;          // row slot within the warp, column chunk
    // `pos` is only ever advanced by the sampler, which never triggers its dependents early, so it is stable for
    // every kernel of the step — including before pdl_wait().
    pdl_launch_dependents();
    const int pos = ld_cg(pos_ptr), n = pos + 1;
    int chunk = (n + nsplit - 1) / nsplit;
    chunk = (chunk + 7) & ~7;
    const int k0 = split * chunk, k1 = min(n, k0 + chunk);
    // pull this CTA's slice of the cache towards L2 while the QKV GEMM is still running (rows < pos are final)
    if (flags & 1) {
        const char* kb = reinterpret_cast<const char*>(kc + ((size_t)bh * S + k0) * 64);
        const char* vb = reinterpret_cast<const char*>(vc + ((size_t)bh * S + k0) * 64);
        const int lines = max(0, min(k1, pos) - k0) * 64 * (int)sizeof(T) / 128;
        for (int i = tid; i < lines; i += AD_THREADS) { prefetch_l2(kb + (size_t)i * 128); prefetch_l2(vb + (size_t)i * 128); }
    }
    pdl_wait();

    float qf[EPL];
    {
        const T* qp = q + (size_t)b * H * 64 + hd * 64 + cl * EPL;
#pragma unroll
        for (int e = 0; e < EPL; ++e) qf[e] = ld_cg(qp + e);
    }
    const T* kbase = kc + ((size_t)bh * S) * 64 + cl * EPL;
    const T* vbase = vc + ((size_t)bh * S) * 64 + cl * EPL;
    const int* mrow = emb_mask ? emb_mask + (size_t)b * mask_ld : nullptr;

    float m_run = -INFINITY, l_run = 0.f, acc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = 0.f;

    const int stride = AD_WARPS * RPW;                    // rows per CTA sweep
    for (int rb = k0 + warp * RPW; rb < k1; rb += stride * UNR) {      // warp-uniform trip count (full-mask shuffles)
        const int r0 = rb + sub;
        float kf[UNR][EPL], vf[UNR][EPL];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int r = r0 + u * stride;
            ok[u] = r < k1 && (r >= Tpre || mrow == nullptr || mrow[r] != 0);
            if (ok[u]) { load_row_frag<T>(kbase + (size_t)r * 64, kf[u]); load_row_frag<T>(vbase + (size_t)r * 64, vf[u]); }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            float s = 0.f;
            if (ok[u]) {
#pragma unroll
                for (int e = 0; e < EPL; ++e) s = fmaf(qf[e], kf[u][e], s);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (ok[u]) {
                s *= 0.125f;                              // 1/sqrt(head_dim=64)
                const float m_new = fmaxf(m_run, s);
                const float corr = __expf(m_run - m_new);  // exp(-inf) = 0 on the first row
                const float p = __expf(s - m_new);
                l_run = l_run * corr + p;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, vf[u][e], acc[e] * corr);
                m_run = m_new;
            }
        }
    }
    // ---- merge the AD_WARPS*RPW row slots of this CTA
    const int slot = warp * RPW + sub;
    if (cl == 0) { sm_m[slot] = m_run; sm_l[slot] = l_run; }
#pragma unroll
    for (int e = 0; e < EPL; ++e) sm_acc[slot][cl * EPL + e] = acc[e];
    __syncthreads();
    float* my_part = part + ((size_t)bh * nsplit + split) * AD_PART;
    if (tid < 64) {
        float M = -INFINITY;
        for (int s = 0; s < AD_WARPS * RPW; ++s) M = fmaxf(M, sm_m[s]);
        float L = 0.f, a = 0.f;
        for (int s = 0; s < AD_WARPS * RPW; ++s) {
            const float w = (sm_m[s] == -INFINITY) ? 0.f : __expf(sm_m[s] - M);
            L += sm_l[s] * w;
            a += sm_acc[s][tid] * w;
        }
        if (tid == 0) { my_part[0] = M; my_part[1] = L; }
        my_part[4 + tid] = a;
    }
    // ---- last-arriving split combines
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&tickets[bh], 1);
        sm_last = (old == nsplit - 1);
    }
    __syncthreads();
    if (!sm_last) return;
    __threadfence();
    if (tid < 64) {
        const float* pp = part + (size_t)bh * nsplit * AD_PART;
        float M = -INFINITY;
        for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(pp + s * AD_PART));
        float L = 0.f, a = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float ms = __ldcg(pp + s * AD_PART);
            const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
            L += __ldcg(pp + s * AD_PART + 1) * w;
            a += __ldcg(pp + s * AD_PART + 4 + tid) * w;
        }
        out[(size_t)b * H * 64 + hd * 64 + tid] = fromf<T>(a / L);
    }
    if (tid == 0) tickets[bh] = 0;
}

// ---------------------------------------------------------------------------------------------------------
// prefill attention: one warp per (b, h, query row i); keys [0, i] from the cache; Tq <= 256.
// mask(i, s) = s <= i and (s >= Tpre or emb_mask[b][s]) or s == i   (generate.py:184-193)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
attn_prefill_kernel(const T* __restrict__ q /*[B*Tq][H*64]*/, const T* __restrict__ kc, const T* __restrict__ vc,
                    const int* __restrict__ emb_mask, int mask_ld, int B, int H, int S, int Tq, int Tpre,
                    T* __restrict__ out /*[B*Tq][H*64]*/) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * 4 + warp;
    if (item >= (long long)B * H * Tq) return;
    const int i = (int)(item % Tq);
    const int hd = (int)((item / Tq) % H);
    const int b = (int)(item / ((long long)Tq * H));
    const T* qp = q + ((size_t)b * Tq + i) * H * 64 + hd * 64;
    const T* kb = kc + (((size_t)b * H + hd) * S) * 64;
    const T* vb = vc + (((size_t)b * H + hd) * S) * 64;
    const int* mrow = emb_mask ? emb_mask + (size_t)b * mask_ld : nullptr;
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int s = lane + 32 * j;
        sc[j] = -INFINITY;
        if (s <= i && (s == i || s >= Tpre || mrow == nullptr || mrow[s] != 0)) {
            float d = 0.f;
            for (int e = 0; e < 64; ++e) d = fmaf(tof(qp[e]), tof(kb[(size_t)s * 64 + e]), d);
            sc[j] = d * 0.125f;
            mx = fmaxf(mx, sc[j]);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = (sc[j] == -INFINITY) ? 0.f : expf(sc[j] - mx);
        sum += sc[j];
    }
    sum = warp_sum(sum);
    float o0 = 0.f, o1 = 0.f;
    for (int s = 0; s <= i; ++s) {
        const float p = __shfl_sync(0xffffffffu, sc[s >> 5], s & 31);
        if (p != 0.f) {
            o0 = fmaf(p, tof(vb[(size_t)s * 64 + 2 * lane]), o0);
            o1 = fmaf(p, tof(vb[(size_t)s * 64 + 2 * lane + 1]), o1);
        }
    }
    T* op = out + ((size_t)b * Tq + i) * H * 64 + hd * 64 + 2 * lane;
    op[0] = fromf<T>(o0 / sum);
    op[1] = fromf<T>(o1 / sum);
}
