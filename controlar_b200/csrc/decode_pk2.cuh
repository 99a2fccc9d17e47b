// decode_pk2.cuh — the conditional-decoding loop (N-1 decode steps + CFG + sampling) as ONE persistent cooperative kernel,
// second generation: TWO independent chains per SM.
//
// Replaces, per generated token (reference file:line):
//   decode_one_token + decode_n_tokens      autoregressive/models/generate.py:95-131
//   Transformer.forward (decode branch)     autoregressive/models/gpt_t2i.py:444-470
//   TransformerBlock / Attention / FeedForward / RMSNorm / KVCache.update   gpt_t2i.py:187-306
//   sample / top_k_top_p_filtering / CFG    generate.py:17-74,103-107
//
// Why two chains (profiles/r1_decode_persistent.md): a decode step is a chain of 5 all-to-all dependent phases per layer; at
// one 512-thread CTA per SM the SM issues 1.1 instructions per clock and HBM is 28 % busy — the loop waits on L2 round trips,
// not on bandwidth.  The images of a batch are independent, so the batch is cut into micro-batches (<= 8 rows each: 4 images
// x {cond, uncond}) and every SM hosts one 256-thread CTA of EACH micro-batch.  The chains overlap each other's latency; the
// weights are streamed once from HBM and once more from L2 (the chains run within a layer of each other; one grid barrier per
// token re-aligns them).  Eight rows also fit the mma A tile with the upper half zero, so a packet carries one row pair.
//
// Per CTA: 8 warps; weights stream through a 4 x 20 KB shared-memory ring (cp.async.bulk + full/empty mbarriers, producer =
// lane 0 of the last warp, non-blocking); blocks are consumed one at a time (two accumulator chains per block), partial sums
// of the 8 warps meet in a [warp][block][64] buffer, one CTA barrier, then a fixed-order reduction + epilogue.
// Activations cross CTAs as 8-byte tagged packets {bf16 pair, tag} polled with strong loads (see decode_persistent.cuh).
#pragma once
#include "decode_persistent.cuh"
#include "pk_plan.h"

constexpr int P2_WARPS = 8, P2_THREADS = P2_WARPS * 32;
constexpr int P2_NSLOT = 4;                           // ring slots (power of two)
constexpr int P2_UNIT_KS = 40;                        // k32-steps per streamed unit: 5 per warp
constexpr int P2_KPW = P2_UNIT_KS / P2_WARPS;         // 5
constexpr int P2_SLOT_BYTES = P2_UNIT_KS * 512;       // 20 KB
constexpr int P2_MAXSUB = 3;                          // units per 8-column block: K <= 3 * 40 * 32 = 3840
constexpr int P2_MAXA = P2_MAXSUB * P2_KPW;           // 15 k-steps per warp
constexpr int P2_MAXBLK = 14;                         // 8-column blocks per CTA and phase (rows of the reduction buffer)
constexpr int P2_RED = 64;                            // floats per (warp, block): 8 rows x 8 columns
constexpr int P2_MAXSEG = PKP_MAXSEG;
constexpr int P2_NMB = 2;                             // micro-batches (CTAs per SM)
constexpr int P2_LIST = 3072;                         // compacted sampler candidates (top-k + ties); more -> un-compacted race
constexpr int P2_SMEM_RING = P2_NSLOT * P2_SLOT_BYTES;                  // 81920
constexpr int P2_SMEM_RED = P2_WARPS * P2_MAXBLK * P2_RED * 4;          // 28672 (attention scratch, sampler histogram/list alias it)
constexpr int P2_SMEM_MISC = 4096;
constexpr int P2_SMEM_TOTAL = P2_SMEM_RING + P2_SMEM_RED + P2_SMEM_MISC;
static_assert(P2_LIST * 8 <= P2_SMEM_RED, "sampler list must fit the reduction buffer");
static_assert(P2_WARPS * 2 * 68 * 4 <= P2_SMEM_RED, "attention scratch must fit the reduction buffer");

struct P2Params {
    int dim, F, V, L, H, T, S, n_img, b_eff, B;       // B = images, b_eff = 2B with CFG
    int nmb; int img_lo[P2_NMB], img_cnt[P2_NMB];     // micro-batch mb handles images [img_lo, img_lo + img_cnt)
    float eps, cs;
    const bf16* tok_emb; const bf16* norm_w; const uint4* w_out;
    const uint4* const* wqkv; const uint4* const* wo; const uint4* const* w13; const uint4* const* w2;
    const bf16* const* attn_norm; const bf16* const* ffn_norm;
    bf16* const* kc; bf16* const* vc;
    const bf16* ctrl[3]; int has_ctrl;
    const float* rope; const int* emb_mask;
    float* logits;                                    // [b_eff][V] (global rows)
    const int* part;                                  // [4][Gc + 1] block offsets per CTA rank: qkv blocks, d-column blocks, w1/w3 pairs, head blocks
    uint2* h2[P2_NMB][2]; uint2* h1[P2_NMB][2]; uint2* att[P2_NMB][2]; uint2* act[P2_NMB][2]; uint2* qkv[P2_NMB][2]; uint2* partial[P2_NMB][2];
    int part_slots;
    unsigned int tag_base;
    unsigned int* bar; unsigned int bar_base;
    SampleArgs smp;
    int n_steps;
    const int* forced; int forced_ld;                 // teacher forcing (parity tests)
    float* trace;                                     // optional [n_steps][b_eff][V]
    int exp_flags;
    long long* dbg; int dbg_step;
    int* nanflag;                                     // dev (PK_TRACE): [nmb][8 steps][L + 1][16] first non-finite value seen per (step, layer, site)
};

// identity of this CTA's micro-batch: local row r (< M) <-> global row of the batch
struct P2Mb {
    int mb, c, Gc;            // micro-batch, rank inside it, CTAs per micro-batch
    int img_lo, cnt, M;       // images, rows (= cnt or 2 cnt)
    int B, use_cfg;
    __device__ __forceinline__ int grow(int r) const { return (use_cfg && r >= cnt) ? B + img_lo + (r - cnt) : img_lo + r; }
};

// A-fragment packet layout of an [8][K] activation tile: for k32-step s, lane (g, t) and half q one 16-byte packet
// {pair(row g, k = 32s + 8t + 4q), tag, pair(row g, k + 2), tag} at 16-byte index (2s + q) * 32 + lane.  uint2 index of (row, k):
__device__ __forceinline__ size_t p2_a_index(int r, int k) {
    const int s = k >> 5, t = (k >> 3) & 3, p = (k >> 1) & 3;
    return ((size_t)((s * 2 + (p >> 1)) * 32 + r * 4 + t)) * 2 + (p & 1);
}

#ifdef PK_TRACE
#define P2_NANCHK(site, cond) do { if (P.nanflag != nullptr && nf_step < 8 && (cond)) atomicOr(P.nanflag + ((mb.mb * 8 + nf_step) * (P.L + 1) + l) * 16 + (site), 1); } while (0)
#else
#define P2_NANCHK(site, cond) do { } while (0)
#endif

struct P2Stream {
    PkCursor c;
    unsigned int issued;
    int lo[5], hi[5];
};

struct P2Smem {
    unsigned char* ring;
    float* red;               // [P2_WARPS][P2_MAXBLK][64]
    float* ssq;               // [P2_WARPS][8]
    uint64_t* full;           // [P2_NSLOT]
    uint64_t* empty;          // [P2_NSLOT]
    P2Stream* st;
    uint32_t* own;            // [2][32] residual-stream pairs of the d-column blocks this CTA owns
    uint32_t* qrow;           // [3][P2_MAXSEG][32]
    PkAttnPlan* plan;
};

__device__ __forceinline__ bool p2_mbar_test(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(pk_smem(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool p2_mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(pk_smem(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void p2_mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pk_smem(b)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// weight stream
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int p2_phase_ks(const P2Params& P, int phase) { return (phase == 3 ? P.F : P.dim) >> 5; }
__device__ __forceinline__ void p2_cursor_next_phase(const P2Params& P, PkCursor& c) {
    if (c.phase == 4) { c.phase = 0; c.l = 0; ++c.step; if (c.step >= P.n_steps - 1) c.done = true; }
    else if (c.phase == 3) { c.phase = 0; ++c.l; }
    else ++c.phase;
}
__device__ __forceinline__ void p2_cursor_skip_empty(const P2Params& P, const P2Stream& st, PkCursor& c) {
    while (!c.done) {
        if (c.phase < 4 && c.l >= P.L) c.phase = 4;
        if (st.lo[c.phase] < st.hi[c.phase]) { c.blk = st.lo[c.phase]; c.sub = 0; return; }
        p2_cursor_next_phase(P, c);
    }
}
__device__ __forceinline__ const uint4* p2_cursor_take(const P2Params& P, const P2Stream& st, PkCursor& c, uint32_t& bytes) {
    const int KS = p2_phase_ks(P, c.phase);
    const uint4* W = c.phase == 0 ? P.wqkv[c.l] : c.phase == 1 ? P.wo[c.l] : c.phase == 2 ? P.w13[c.l] : c.phase == 3 ? P.w2[c.l] : P.w_out;
    const int ks0 = c.sub * P2_UNIT_KS, nks = min(P2_UNIT_KS, KS - ks0);
    const uint4* src = W + ((size_t)c.blk * KS + ks0) * 32;
    bytes = (uint32_t)nks * 512u;
    if ((c.sub + 1) * P2_UNIT_KS < KS) { ++c.sub; return src; }
    c.sub = 0;
    if (++c.blk < st.hi[c.phase]) return src;
    p2_cursor_next_phase(P, c);
    p2_cursor_skip_empty(P, st, c);
    return src;
}
// producer (one thread): issue every unit whose ring slot has been released by all 8 warps; never blocks
__device__ __noinline__ void p2_producer_advance(const P2Params& P, unsigned char* ring, uint64_t* full, uint64_t* empty, P2Stream* stp) {
    P2Stream& st = *stp;
    while (!st.c.done) {
        const unsigned int use = st.issued / P2_NSLOT, slot = st.issued % P2_NSLOT;
        if (use > 0 && !p2_mbar_test(&empty[slot], (use - 1) & 1)) break;
        uint32_t bytes;
        const uint4* src = p2_cursor_take(P, st, st.c, bytes);
        pk_mbar_expect(&full[slot], bytes);
        pk_bulk_g2s(ring + (size_t)slot * P2_SLOT_BYTES, src, bytes, &full[slot]);
        ++st.issued;
    }
}

// ---------------------------------------------------------------------------------------------------------
// A operand: poll the tagged packets of this warp's k-steps into mma fragments.  k-step i of warp w is
// s = (i / 5) * 40 + (i % 5) * 8 + w  (unit i / 5 of a block holds k-steps [40 u, 40 u + 40))
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ constexpr int p2_kstep(int i, int warp) { return (i / P2_KPW) * P2_UNIT_KS + (i % P2_KPW) * P2_WARPS + warp; }

template <int I0, int CNT>
__device__ __forceinline__ void p2_poll_round(const unsigned char* __restrict__ base, int nst, int warp, int lane, unsigned int tag,
                                              bool need, uint32_t (&a)[P2_MAXA][4]) {
    uint4 v[CNT][2];
    const unsigned char* rec[CNT];
#pragma unroll
    for (int u = 0; u < CNT; ++u) rec[u] = base + (size_t)((I0 + u < nst) ? p2_kstep(I0 + u, warp) : warp) * 1024 + lane * 16;
#pragma unroll
    for (int u = 0; u < CNT; ++u) { v[u][0] = pk_ld128(rec[u]); v[u][1] = pk_ld128(rec[u] + 512); }
    unsigned int spins = 0;
    while (true) {
        bool any_bad = false;
#pragma unroll
        for (int u = 0; u < CNT; ++u) {
            const unsigned int b = need ? ((v[u][0].y ^ tag) | (v[u][0].w ^ tag) | (v[u][1].y ^ tag) | (v[u][1].w ^ tag)) : 0u;
            if (I0 + u < nst && b != 0u) {
                any_bad = true;
                v[u][0] = pk_ld128(rec[u]); v[u][1] = pk_ld128(rec[u] + 512);
            }
        }
        if (!any_bad) break;
        __nanosleep(40);
        pk_spin_check(spins);
    }
#pragma unroll
    for (int u = 0; u < CNT; ++u) {
        const bool in = (I0 + u < nst) && need;
        a[I0 + u][0] = in ? v[u][0].x : 0u; a[I0 + u][1] = in ? v[u][0].z : 0u;
        a[I0 + u][2] = in ? v[u][1].x : 0u; a[I0 + u][3] = in ? v[u][1].z : 0u;
    }
}

// arrival hint before the full poll: warp 0 watches the first packet of 32 of the K/8 producer blocks with back-off
__device__ __forceinline__ void p2_prepoll(const uint2* buf, int K, unsigned int tag, int rank) {
    const int nblk = K >> 3;
    if (threadIdx.x < 32) {
        const uint2* pkt = buf + p2_a_index(0, (int)((rank * 7u + threadIdx.x * (unsigned)max(1, nblk >> 5)) % (unsigned)nblk) * 8);
        unsigned int spins = 0;
        while (!__all_sync(0xffffffffu, pk_ld64(pkt).y == tag)) { __nanosleep(100); pk_spin_check(spins); }
    }
    __syncthreads();
}

__device__ __forceinline__ void mma_bf16_16816_lo(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    const uint32_t z = 0u;     // rows 8..15 of the A tile are zero (a micro-batch has at most 8 rows)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(z), "r"(a2), "r"(z), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------------
// GEMM phase.  kind: 0 qkv (+RoPE, KV append) | 1 wo (+residual) | 2 w1/w3 (+SwiGLU) | 3 w2 (+residual, control add) | 4 head
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint2* p2_a_buf(const P2Params& P, int mb, int kind, int par) {
    return kind == 1 ? P.att[mb][par] : kind == 2 ? P.h1[mb][par] : kind == 3 ? P.act[mb][par] : P.h2[mb][par];
}
__device__ __forceinline__ const bf16* p2_ctrl_next(const P2Params& P, int l) {
    const int step3 = P.L / 3;
    return (P.has_ctrl && (l + 1) < P.L && (l + 1) % step3 == 0) ? P.ctrl[(l + 1) / step3] : nullptr;
}

__device__ __forceinline__ void p2_gemm_phase(const P2Params& P, const P2Smem& sm, const P2Mb& mb, const int kind, const int l, const int pos,
                                              const unsigned int tag, const int blk_lo, const int blk_hi, unsigned int& cons, float* trace_rows,
                                              long long* dbg, const int nf_step = 0) {
    const int par = l & 1;
    const bool NORM = (kind == 0 || kind == 2 || kind == 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int K = kind == 3 ? P.F : P.dim;
    const int KS = K >> 5;
    const int nsub = (KS + P2_UNIT_KS - 1) / P2_UNIT_KS;
    const int nst = KS / P2_WARPS;                     // host guarantees KS % 8 == 0
    const int M = mb.M;
    const int nblk = blk_hi - blk_lo;
    if (nblk <= 0) return;                             // this CTA owns no columns of this phase
#ifdef PK_TRACE
    const bool stamp = dbg != nullptr && tid == 0;
#else
    constexpr bool stamp = false;
#endif
    if (stamp) dbg[0] = pk_now();
    const bool is_prod = tid == P2_THREADS - 32;

    // epilogue identity: block ej of a pass of 4, element eq = row eg x column 2 ecp + er
    const int ej = tid >> 6, eq = tid & 63, ei = eq >> 1, er = eq & 1, eg = ei >> 2, ecp = ei & 3;
    uint32_t ctl = 0;
    if (kind == 1 && l == 0 && er == 0 && ej < nblk && ej < 2) {
        // layer 0: the residual stream of this CTA's columns is the embedding row the sampler wrote (H2[0])
        const int n = (blk_lo + ej) * 8 + 2 * ecp;
        const uint2* pp = P.h2[mb.mb][0] + p2_a_index(eg, n);
        uint2 v;
        unsigned int spins = 0;
        do { v = pk_ld64(pp); if (!(eg < M) || v.y == tag) break; pk_spin_check(spins); } while (true);
        sm.own[ej * 32 + ei] = v.x;
    }
    if (kind == 3 && er == 0 && ej < nblk) {
        const bf16* ctrl = p2_ctrl_next(P, l);
        const int n = (blk_lo + ej) * 8 + 2 * ecp;
        const int p = pos - P.T + 1;
        if (ctrl != nullptr && p >= 0 && p < P.n_img && eg < M)
            ctl = __ldg(reinterpret_cast<const unsigned int*>(ctrl + ((size_t)mb.grow(eg) * P.n_img + p) * P.dim + n));
    }

    // ---- A fragments (+ RMSNorm)
    uint32_t a[P2_MAXA][4];
    uint4 nwv[P2_KPW];
    {
        const uint2* abuf = p2_a_buf(P, mb.mb, kind, par);
        const unsigned char* base = reinterpret_cast<const unsigned char*>(abuf);
        const bool need = g < M;
        p2_prepoll(abuf, K, tag, mb.c);
        p2_poll_round<0, P2_KPW>(base, nst, warp, lane, tag, need, a);
        if (kind == 3) {
            p2_poll_round<P2_KPW, P2_KPW>(base, nst, warp, lane, tag, need, a);
            p2_poll_round<2 * P2_KPW, P2_KPW>(base, nst, warp, lane, tag, need, a);
        } else {
#pragma unroll
            for (int i = P2_KPW; i < P2_MAXA; ++i) { a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0u; }
        }
    }
    {
        bool badA = false;
#pragma unroll
        for (int i = 0; i < P2_MAXA; ++i)
#pragma unroll
            for (int p = 0; p < 4; ++p) { float x, y; unpack_bf16x2(a[i][p], x, y); badA = badA || !isfinite(x) || !isfinite(y); }
        P2_NANCHK(kind, badA);       // sites 0..4: the A operand of GEMM kind was already non-finite
    }
    if (stamp) dbg[1] = pk_now();
    {
        const bf16* nw = kind == 0 ? P.attn_norm[l] : kind == 2 ? P.ffn_norm[l] : P.norm_w;
#pragma unroll
        for (int i = 0; i < P2_KPW; ++i)
            nwv[i] = (NORM && i < nst) ? __ldg(reinterpret_cast<const uint4*>(nw + (warp + i * P2_WARPS) * 32 + t * 8)) : make_uint4(0, 0, 0, 0);
    }
    if (NORM) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < P2_KPW; ++i)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float x, y;
                unpack_bf16x2(a[i][p], x, y); s = fmaf(x, x, s); s = fmaf(y, y, s);
            }
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (t == 0) sm.ssq[warp * 8 + g] = s;
    }
    __syncthreads();                                   // ssq partials visible; the previous phase's readers of `red` are done
    if (NORM) {
        float qs = 0.f;
#pragma unroll
        for (int w = 0; w < P2_WARPS; ++w) qs += sm.ssq[w * 8 + g];
        const float rs = rsqrtf(qs / (float)K + P.eps);
#pragma unroll
        for (int i = 0; i < P2_KPW; ++i) {
            if (i < nst) {
                const uint32_t wi[4] = {nwv[i].x, nwv[i].y, nwv[i].z, nwv[i].w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    // RMSNorm.forward (gpt_t2i.py:193-198): (x.float() * rstd).type_as(x) * weight
                    const __nv_bfloat162 w2 = *reinterpret_cast<const __nv_bfloat162*>(&wi[p]);
                    float x, y;
                    unpack_bf16x2(a[i][p], x, y);
                    __nv_bfloat162 v = __floats2bfloat162_rn(x * rs, y * rs);
                    v = __hmul2(v, w2);
                    a[i][p] = *reinterpret_cast<uint32_t*>(&v);
                }
            }
        }
    }
    if (stamp) dbg[2] = pk_now();
#ifdef PK_TRACE
    if ((P.exp_flags & 1) && kind == 0) __nanosleep(20000);                  // dev: slow consumers
    if ((P.exp_flags & 2) && warp == P2_WARPS - 1) __nanosleep(20000);        // dev: slow producer warp
    if ((P.exp_flags & 4) && warp == 0) __nanosleep(20000);                   // dev: one slow consumer warp
#endif

    // ---- MMA: one block at a time, two accumulator chains, weights from the ring
    {
        const uint32_t ring_s = pk_smem(sm.ring) + lane * 16 + warp * 512;
        for (int j = 0; j < nblk; ++j) {
            float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int sub = 0; sub < P2_MAXSUB; ++sub) {
                if (sub < nsub) {
                    const unsigned int u = cons + (unsigned int)(j * nsub + sub);
                    const unsigned int slot = u & (P2_NSLOT - 1);
                    {
                        unsigned int spins = 0;
                        while (!p2_mbar_try(&sm.full[slot], (u / P2_NSLOT) & 1)) {
                            if (is_prod) p2_producer_advance(P, sm.ring, sm.full, sm.empty, sm.st);
                            pk_spin_check(spins);
                        }
                    }
                    const uint32_t sbase = ring_s + slot * P2_SLOT_BYTES;
#pragma unroll
                    for (int ii = 0; ii < P2_KPW; ++ii) {
                        const int i = sub * P2_KPW + ii;
                        if (i < nst) {
                            uint4 wf;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wf.x), "=r"(wf.y), "=r"(wf.z), "=r"(wf.w) : "r"(sbase + ii * (P2_WARPS * 512)));
                            { float x0, x1; unpack_bf16x2(wf.x, x0, x1); P2_NANCHK(13, !isfinite(x0) || !isfinite(x1) || fabsf(x0) > 4.f); }
                            mma_bf16_16816_lo(accA, a[i][0], a[i][1], wf.x, wf.y);
                            mma_bf16_16816_lo(accB, a[i][2], a[i][3], wf.z, wf.w);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) p2_mbar_arrive(&sm.empty[slot]);
                    if (is_prod) p2_producer_advance(P, sm.ring, sm.full, sm.empty, sm.st);
                }
            }
            *reinterpret_cast<float2*>(sm.red + (size_t)(warp * P2_MAXBLK + j) * P2_RED + g * 8 + 2 * t) = make_float2(accA[0] + accB[0], accA[1] + accB[1]);
        }
        cons += (unsigned int)(nblk * nsub);
    }
    __syncthreads();
    if (is_prod) p2_producer_advance(P, sm.ring, sm.full, sm.empty, sm.st);
    if (stamp) dbg[3] = pk_now();

    // ---- fixed-order cross-warp reduction + epilogue, 4 blocks (kind 2: 4 w1/w3 pairs) per pass
    const int per_pass = kind == 2 ? 8 : 4;
    for (int jb = 0; jb < nblk; jb += per_pass) {
        const int jj = kind == 2 ? jb + 2 * ej : jb + ej;
        const bool active = jj < nblk;
        float val = 0.f;
        if (active) {
#pragma unroll
            for (int w = 0; w < P2_WARPS; ++w) val += sm.red[(size_t)(w * P2_MAXBLK + jj) * P2_RED + eq];
            if (kind == 2) {
                float val3 = 0.f;
#pragma unroll
                for (int w = 0; w < P2_WARPS; ++w) val3 += sm.red[(size_t)(w * P2_MAXBLK + jj + 1) * P2_RED + eq];
                // FeedForward.forward gpt_t2i.py:217: w2(silu(w1 x) * w3 x), every intermediate in bf16
                val = rnd<bf16>(silu_f(rnd<bf16>(val))) * rnd<bf16>(val3);
            }
        }
        const float v0 = val;
        const float v1 = __shfl_xor_sync(0xffffffffu, val, 1);
        P2_NANCHK(8 + kind, active && eg < M && !isfinite(val));        // sites 8..12: the output of GEMM kind (before the epilogue) is non-finite
        if (active && er == 0 && eg < M) {
            const int r = eg;
            if (kind == 0) {
                const int n = (blk_lo + jj) * 8 + 2 * ecp;
                const int sec = n / P.dim, w = n - sec * P.dim, head = w >> 6, el = w & 63;
                float a0 = rnd<bf16>(v0), a1 = rnd<bf16>(v1);
                if (sec < 2) {   // apply_rotary_emb gpt_t2i.py:522-532 (interleaved pairs, fp32, then cast)
                    const float2 cs2 = __ldg(reinterpret_cast<const float2*>(P.rope + ((size_t)pos * 32 + (el >> 1)) * 2));
                    const float x0 = a0 * cs2.x - a1 * cs2.y, x1 = a1 * cs2.x + a0 * cs2.y;
                    a0 = x0; a1 = x1;
                }
                const uint32_t pk = pk_pack(a0, a1);
                // packets for the attention phase: [sec][row][head][el/8][(el%8)/2]
                pk_st64(P.qkv[mb.mb][par] + ((((size_t)sec * 8 + r) * P.H + head) * 8 + (el >> 3)) * 4 + ((el & 7) >> 1), pk, tag);
                if (sec > 0) {   // KVCache.update gpt_t2i.py:227-235 (read by later tokens; ordered by the per-token barrier)
                    bf16* cache = sec == 1 ? P.kc[l] : P.vc[l];
                    *reinterpret_cast<uint32_t*>(cache + (((size_t)mb.grow(r) * P.H + head) * P.S + pos) * 64 + el) = pk;
                }
            } else if (kind == 1 || kind == 3) {
                const int n = (blk_lo + jj) * 8 + 2 * ecp;
                float h0, h1;
                unpack_bf16x2(sm.own[(jj & 1) * 32 + ei], h0, h1);
                float o0 = rnd<bf16>(h0 + rnd<bf16>(v0)), o1 = rnd<bf16>(h1 + rnd<bf16>(v1));   // h + drop_path(...) gpt_t2i.py:305-306
                if (kind == 3) {   // gpt_t2i.py:466 — h += cs * ctrl[:, pos - T + 1] ahead of the next layer group
                    const int p = pos - P.T + 1;
                    if (p2_ctrl_next(P, l) != nullptr && p >= 0 && p < P.n_img) {
                        float c0, c1;
                        unpack_bf16x2(ctl, c0, c1);
                        o0 = rnd<bf16>(o0 + rnd<bf16>(P.cs * c0)); o1 = rnd<bf16>(o1 + rnd<bf16>(P.cs * c1));
                    }
                }
                const uint32_t pk = pk_pack(o0, o1);
                sm.own[(jj & 1) * 32 + ei] = pk;
                pk_st64((kind == 3 ? P.h2[mb.mb][par ^ 1] : P.h1[mb.mb][par]) + p2_a_index(r, n), pk, kind == 3 ? tag + 1u : tag);
            } else if (kind == 2) {
                const int n = ((blk_lo >> 1) + (jj >> 1)) * 8 + 2 * ecp;   // activation column
                pk_st64(P.act[mb.mb][par] + p2_a_index(r, n), pk_pack(v0, v1), tag);
            } else {
                const int n = (blk_lo + jj) * 8 + 2 * ecp;
                // logits = output(norm(h)).float()  gpt_t2i.py:469-470 (bf16 head output, then fp32)
                const float2 o = make_float2(rnd<bf16>(v0), rnd<bf16>(v1));
                *reinterpret_cast<float2*>(P.logits + (size_t)mb.grow(r) * P.V + n) = o;
                if (trace_rows != nullptr) *reinterpret_cast<float2*>(trace_rows + (size_t)mb.grow(r) * P.V + n) = o;
            }
        }
    }
    if (stamp) dbg[4] = pk_now();
}

// ---------------------------------------------------------------------------------------------------------
// attention phase: the flattened (row, head, key) space of the micro-batch is cut into Gc equal ranges (pk_plan.h)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void p2_attn_finalize(const P2Params& P, const P2Mb& mb, float Mx, float Ls, float a, const PkSegPlan& sgm, int e,
                                                 unsigned int tag, int par) {
    uint2* pb = P.partial[mb.mb][par] + ((size_t)sgm.bh * P.part_slots) * 66;
    if (!sgm.owner) {
        uint2* mine = pb + (size_t)(mb.c - sgm.first_cta) * 66;
        pk_st64(mine + 2 + e, __float_as_uint(a), tag);
        if (e == 0) pk_st64(mine, __float_as_uint(Mx), tag);
        if (e == 1) pk_st64(mine + 1, __float_as_uint(Ls), tag);
        return;
    }
    if (sgm.ks > 0) {   // combine the helpers' partials (CTAs first_cta .. c - 1) in index order, then ours
        const int nh = mb.c - sgm.first_cta;
        float Mc = -INFINITY, Lc = 0.f, ac = 0.f;
        for (int hI = 0; hI <= nh; ++hI) {
            float mh, lh, ah;
            if (hI < nh) {
                const uint2* src = pb + (size_t)hI * 66;
                uint2 pm, pl, pa;
                unsigned int spins = 0;
                do {
                    pm = pk_ld64(src); pl = pk_ld64(src + 1); pa = pk_ld64(src + 2 + e);
                    if (pm.y == tag && pl.y == tag && pa.y == tag) break;
                    __nanosleep(32);
                    pk_spin_check(spins);
                } while (true);
                mh = __uint_as_float(pm.x); lh = __uint_as_float(pl.x); ah = __uint_as_float(pa.x);
            } else { mh = Mx; lh = Ls; ah = a; }
            const float m_new = fmaxf(Mc, mh);
            const float wa = Mc == -INFINITY ? 0.f : __expf(Mc - m_new);
            const float wb = mh == -INFINITY ? 0.f : __expf(mh - m_new);
            Lc = Lc * wa + lh * wb; ac = ac * wa + ah * wb; Mc = m_new;
        }
        Ls = Lc; a = ac;
    }
    const float o = rnd<bf16>(a / Ls);           // SDPA output in the model dtype
    const float o1 = __shfl_down_sync(0xffffffffu, o, 1);
    if ((e & 1) == 0) pk_st64(P.att[mb.mb][par] + p2_a_index(sgm.b, sgm.hd * 64 + e), pk_pack(o, o1), tag);
}

__device__ __forceinline__ void p2_attn_phase(const P2Params& P, const P2Smem& sm, const P2Mb& mb, int layer, int pos, unsigned int tag, int par,
                                              long long* dbg, const int nf_step = 0) {
    const int l = layer;
    constexpr int EPL = 8, UNR = 8, ENT = 68;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane >> 3, cl = lane & 7;
    const PkAttnPlan& pl = *sm.plan;
    if (!pl.active) return;
    const int n = pos + 1;                                 // keys 0 .. pos; key `pos` is the token being decoded
    const int nseg = pl.nseg;
#ifdef PK_TRACE
    const bool stamp = dbg != nullptr && tid == 0;
#else
    constexpr bool stamp = false;
#endif
    if (stamp) dbg[0] = pk_now();
    const uint2* qkvb = P.qkv[mb.mb][par];
    const bf16* kc = P.kc[layer];
    const bf16* vc = P.vc[layer];
    float* sc = sm.red;                                    // [P2_WARPS][2][ENT]

    // q (and, for owner segments, this token's k and v) of every segment -> shared memory, polled in parallel:
    // warp sg, lanes 0-7 q, 8-15 k, 16-23 v (lane & 7 = 16-byte chunk = 4 packets)
    if (warp < nseg && lane < 24) {
        const PkSegPlan& q = pl.seg[warp];
        const int sec = lane >> 3;
        if (sec == 0 || q.owner) {
            const uint2* qp = qkvb + (((size_t)(sec * 8 + q.b) * P.H + q.hd) * 8 + (lane & 7)) * 4;
            uint4 v0, v1;
            unsigned int spins = 0;
            do {
                v0 = pk_ld128(qp); v1 = pk_ld128(qp + 2);
                if (v0.y == tag && v0.w == tag && v1.y == tag && v1.w == tag) break;
                __nanosleep(32);
                pk_spin_check(spins);
            } while (true);
            *reinterpret_cast<uint4*>(sm.qrow + (sec * P2_MAXSEG + warp) * 32 + (lane & 7) * 4) = make_uint4(v0.x, v0.z, v1.x, v1.z);
        }
    }
    if (stamp) dbg[1] = pk_now();
    __syncthreads();                                       // q/k/v rows visible; the scratch aliases the previous phase's reduction buffer

#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        float* ent = sc + (size_t)(warp * 2 + part) * ENT;
        const PkPart pt = pl.part[warp][part];
        if (pt.k0 >= pt.k1) { if (lane == 0) ent[1] = -INFINITY; continue; }     // (warp-uniform)
        const int bh = pt.bh, k0 = pt.k0, k1 = pt.k1;
        const int sg = bh - pl.pair_lo;
        const int grow = mb.grow(pt.b);
        const int hd = bh - pt.b * P.H;
        float qf[EPL];
        {
            const uint4 qq = *reinterpret_cast<const uint4*>(sm.qrow + sg * 32 + cl * 4);
            unpack_bf16x2(qq.x, qf[0], qf[1]); unpack_bf16x2(qq.y, qf[2], qf[3]);
            unpack_bf16x2(qq.z, qf[4], qf[5]); unpack_bf16x2(qq.w, qf[6], qf[7]);
        }
        const size_t rowbase = ((size_t)grow * P.H + hd) * P.S;
        const bf16* kbase = kc + rowbase * 64 + cl * EPL;
        const bf16* vbase = vc + rowbase * 64 + cl * EPL;
        const int* mrow = P.emb_mask ? P.emb_mask + (size_t)grow * P.T : nullptr;
        float m_run = -INFINITY, l_run = 0.f, acc[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
        for (int rb = k0; rb < k1; rb += 4 * UNR) {        // warp-uniform trip count
            uint4 kraw[UNR], vraw[UNR];
            int msk[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int r = rb + sub + 4 * u;
                const int rr = min(r, k1 - 1);              // slots past the end re-read the last row (unconditional loads)
                kraw[u] = ldg_cg128(kbase + (size_t)rr * 64); vraw[u] = ldg_cg128(vbase + (size_t)rr * 64);
                msk[u] = (mrow != nullptr && rr < P.T) ? __ldg(mrow + rr) : 1;
            }
            float sc8[UNR];
            float mbk = -INFINITY;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int r = rb + sub + 4 * u;
                if (r == n - 1) {   // the newest key / value is this token's: from the QKV packets, not from the cache
                    kraw[u] = *reinterpret_cast<const uint4*>(sm.qrow + (1 * P2_MAXSEG + sg) * 32 + cl * 4);
                    vraw[u] = *reinterpret_cast<const uint4*>(sm.qrow + (2 * P2_MAXSEG + sg) * 32 + cl * 4);
                }
                float kf[EPL];
                unpack_bf16x2(kraw[u].x, kf[0], kf[1]); unpack_bf16x2(kraw[u].y, kf[2], kf[3]);
                unpack_bf16x2(kraw[u].z, kf[4], kf[5]); unpack_bf16x2(kraw[u].w, kf[6], kf[7]);
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < EPL; ++e) s = fmaf(qf[e], kf[e], s);
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
                sc8[u] = (r < k1 && msk[u] != 0) ? s * 0.125f : -INFINITY;      // 1/sqrt(head_dim = 64); masked / past-the-end -> weight 0
                mbk = fmaxf(mbk, sc8[u]);
            }
            const float m_new = fmaxf(m_run, mbk);
            if (m_new != -INFINITY) {                     // (uniform over the 8 lanes of a row slot)
                const float corr = __expf(m_run - m_new);   // exp(-inf) = 0 on the first block
                l_run *= corr;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[e] *= corr;
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    float vf[EPL];
                    unpack_bf16x2(vraw[u].x, vf[0], vf[1]); unpack_bf16x2(vraw[u].y, vf[2], vf[3]);
                    unpack_bf16x2(vraw[u].z, vf[4], vf[5]); unpack_bf16x2(vraw[u].w, vf[6], vf[7]);
                    const float p = __expf(sc8[u] - m_new);
                    l_run += p;
#pragma unroll
                    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, vf[e], acc[e]);
                }
                m_run = m_new;
            }
        }
        // merge the warp's four row slots (lanes 8 apart), fixed order
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const float m_o = __shfl_xor_sync(0xffffffffu, m_run, o);
            const float l_o = __shfl_xor_sync(0xffffffffu, l_run, o);
            const float m_new = fmaxf(m_run, m_o);
            const float wA = m_run == -INFINITY ? 0.f : __expf(m_run - m_new);
            const float wB = m_o == -INFINITY ? 0.f : __expf(m_o - m_new);
            l_run = l_run * wA + l_o * wB;
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const float a_o = __shfl_xor_sync(0xffffffffu, acc[e], o);
                acc[e] = acc[e] * wA + a_o * wB;
            }
            m_run = m_new;
        }
        if (sub == 0) {
            if (cl == 0) { ent[0] = __int_as_float(bh); ent[1] = m_run; ent[2] = l_run; }
#pragma unroll
            for (int e = 0; e < EPL; ++e) ent[4 + cl * EPL + e] = acc[e];
        }
    }
    __syncthreads();
    if (stamp) dbg[2] = pk_now();
    // ---- finalise: two warps per segment (thread e = dimension) merge the entries of their pair in warp order
    for (int sg = warp >> 1; sg < nseg; sg += P2_WARPS / 2) {
        const int e = tid & 63;
        const PkSegPlan& q = pl.seg[sg];
        float Mx = -INFINITY;
        for (int w = q.w0; w <= q.w1; ++w) Mx = fmaxf(Mx, sc[(w * 2 + (int)((q.part_mask >> w) & 1u)) * ENT + 1]);
        float Ls = 0.f, a = 0.f;
        for (int w = q.w0; w <= q.w1; ++w) {
            const float* en = sc + (w * 2 + (int)((q.part_mask >> w) & 1u)) * ENT;
            const float mi = en[1];
            const float wt = (mi == -INFINITY) ? 0.f : __expf(mi - Mx);
            Ls += en[2] * wt;
            a += en[4 + e] * wt;
        }
        P2_NANCHK(6, !isfinite(a) || !isfinite(Ls) || !(Ls > 0.f && q.owner || !q.owner));   // site 6: attention partial of this CTA
        p2_attn_finalize(P, mb, Mx, Ls, a, q, e, tag, par);
    }
    if (stamp) dbg[3] = pk_now();
}

// ---------------------------------------------------------------------------------------------------------
// sampler: CFG combine + temperature + exact top-k (3-pass radix select) + soft-max + top-p + exponential race.
// One 256-thread CTA per image; the logits rows stay in L2 and are re-read by every pass (64 KB per row).
// ---------------------------------------------------------------------------------------------------------
struct P2SampSh { unsigned int bin, krem, cnt; float fa, fb; int ia; };

template <typename F>
__device__ __forceinline__ void p2_row_pass(const SampleArgs& a, const float* lc, const float* lu, bool cfg, F&& f) {
    const int V = a.V;
    for (int i4 = threadIdx.x * 4; i4 < V; i4 += P2_THREADS * 4) {
        float4 c = __ldcg(reinterpret_cast<const float4*>(lc + i4));
        if (cfg) {
            const float4 u = __ldcg(reinterpret_cast<const float4*>(lu + i4));
            c.x = u.x + (c.x - u.x) * a.cfg_scale; c.y = u.y + (c.y - u.y) * a.cfg_scale;
            c.z = u.z + (c.z - u.z) * a.cfg_scale; c.w = u.w + (c.w - u.w) * a.cfg_scale;
        }
        f(i4 + 0, c.x * a.inv_temp); f(i4 + 1, c.y * a.inv_temp); f(i4 + 2, c.z * a.inv_temp); f(i4 + 3, c.w * a.inv_temp);
    }
}

// block-wide: find the bin (from the top) where the cumulative count reaches k; hist has `nbins` = 8 * P2_THREADS entries
__device__ __forceinline__ void p2_find_bin(const unsigned int* hist, unsigned int k, unsigned int* wsum, P2SampSh* sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int loc[8], mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = hist[tid * 8 + j]; mine += loc[j]; }
    // suffix sum over threads (counts of the bins above this thread's range)
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += v; }
    if (lane == 0) wsum[warp] = incl;
    __syncthreads();
    unsigned int above = incl - mine;
    for (int w = warp + 1; w < P2_WARPS; ++w) above += wsum[w];
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        if (above < k && k <= above + loc[j]) { sh->bin = (unsigned int)(tid * 8 + j); sh->krem = k - above; }
        above += loc[j];
    }
    __syncthreads();
}

__device__ __forceinline__ float p2_block_sum(float v, float* red8) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < P2_WARPS; ++w) s += red8[w];
    return s;
}
__device__ __forceinline__ float p2_block_max(float v, float* red8) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = red8[0];
#pragma unroll
    for (int w = 1; w < P2_WARPS; ++w) s = fmaxf(s, red8[w]);
    return s;
}

// returns the sampled token (valid in every thread); `scratch` = the reduction buffer (>= 24 KB), `sh`/`red8`/`wsum` small shared
__device__ __noinline__ int p2_sample(const SampleArgs& a, int img, int step, float* scratch, P2SampSh* sh, float* red8, unsigned int* wsum, int* dbgflag = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    bool cfg = a.use_cfg != 0 && a.cfg_on != 0;
    if (a.cfg_interval > -1 && step - 1 > a.cfg_interval) cfg = false;      // generate.py:121-122
    const float* lc = a.logits + (size_t)img * V;
    const float* lu = a.logits + (size_t)(img + a.B) * V;
    unsigned int* hist = reinterpret_cast<unsigned int*>(scratch);          // [2048]
    unsigned int thr_key = 0u;
    float mx = -INFINITY;
    const bool do_topk = a.top_k > 0 && a.top_k < V;
    // ---- pass 1: maximum (+ histogram of the top 11 key bits)
    if (do_topk) { for (int i = tid; i < 2048; i += P2_THREADS) hist[i] = 0u; }
    __syncthreads();
    p2_row_pass(a, lc, lu, cfg, [&](int, float z) {
        mx = fmaxf(mx, z);
        if (do_topk) atomicAdd(&hist[float_order_key(z) >> 21], 1u);
    });
    mx = p2_block_max(mx, red8);                                           // (its barriers also publish the histogram)
    if (do_topk) {
        // exact k-th largest (ties at the threshold are kept, generate.py:37): radix select 11 + 11 + 10 bits
        p2_find_bin(hist, (unsigned int)a.top_k, wsum, sh);
        const unsigned int b1 = sh->bin, k2 = sh->krem;
        __syncthreads();
        for (int i = tid; i < 2048; i += P2_THREADS) hist[i] = 0u;
        __syncthreads();
        p2_row_pass(a, lc, lu, cfg, [&](int, float z) {
            const unsigned int key = float_order_key(z);
            if ((key >> 21) == b1) atomicAdd(&hist[(key >> 10) & 2047u], 1u);
        });
        __syncthreads();
        p2_find_bin(hist, k2, wsum, sh);
        const unsigned int b2 = sh->bin, k3 = sh->krem;
        __syncthreads();
        for (int i = tid; i < 2048; i += P2_THREADS) hist[i] = 0u;
        __syncthreads();
        const unsigned int pre = (b1 << 11) | b2;
        p2_row_pass(a, lc, lu, cfg, [&](int, float z) {
            const unsigned int key = float_order_key(z);
            if ((key >> 10) == pre) atomicAdd(&hist[key & 1023u], 1u);
        });
        __syncthreads();
        p2_find_bin(hist, k3, wsum, sh);
        thr_key = (pre << 10) | sh->bin;
        __syncthreads();
    }
    // ---- pass 2: soft-max denominator over the kept elements + per-thread kept count (deterministic order)
    float sum = 0.f;
    unsigned int mycnt = 0;
    p2_row_pass(a, lc, lu, cfg, [&](int, float z) {
        if (float_order_key(z) >= thr_key) { sum += expf(z - mx); ++mycnt; }
    });
    sum = p2_block_sum(sum, red8);
    if (dbgflag != nullptr && tid == 0) { if (!isfinite(mx)) atomicOr(dbgflag, 1); if (!isfinite(sum)) atomicOr(dbgflag, 2); }
    // exclusive scan of the kept counts -> deterministic positions in the candidate list
    unsigned int incl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned int off = incl - mycnt, total = 0;
    for (int w = 0; w < P2_WARPS; ++w) { if (w < warp) off += wsum[w]; total += wsum[w]; }
    const bool listed = total <= (unsigned int)P2_LIST;
    uint2* list = reinterpret_cast<uint2*>(scratch);                        // {index, exp(z - mx)} (the histogram is dead)
    __syncthreads();
    if (listed) {
        unsigned int o = off;
        p2_row_pass(a, lc, lu, cfg, [&](int i, float z) {
            if (float_order_key(z) >= thr_key) { list[o] = make_uint2((unsigned int)i, __float_as_uint(expf(z - mx))); ++o; }
        });
        __syncthreads();
    }
    // ---- nucleus (top-p), generate.py:40-55: in descending order a token is removed iff the probability mass strictly before
    // it exceeds top_p (the first is always kept), i.e. the kept set is {p >= tau*}; tau* = the smallest probability v with
    // mass{p > v} <= top_p, found EXACTLY by a 32-step bisection on the order-preserving integer key of e = exp(z - mx)
    unsigned int p_key = 0u;                                                // keep e with key >= p_key
    if (a.top_p < 1.0f) {
        const float target = a.top_p * sum;
        unsigned int lo = 0u;                                               // largest key with mass{key' > key} > target ... bisect bitwise
        // find the smallest key K with mass{e : key(e) > K} <= target; kept = {key >= K'} where K' = smallest listed key >= ... = K itself if present
        unsigned int K = 0u;
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned int cand = K | (1u << bit);                      // test threshold cand - 1: mass{key > cand - 1} = mass{key >= cand}
            float ma = 0.f;
            if (listed) { for (unsigned int j = tid; j < total; j += P2_THREADS) { const float e = __uint_as_float(list[j].y); if (float_order_key(e) >= cand) ma += e; } }
            else p2_row_pass(a, lc, lu, cfg, [&](int, float z) { if (float_order_key(z) >= thr_key) { const float e = expf(z - mx); if (float_order_key(e) >= cand) ma += e; } });
            ma = p2_block_sum(ma, red8);
            if (ma > target) K = cand;                                      // mass at or above cand still exceeds top_p: the boundary key is >= cand
        }
        (void)lo;
        // K = the largest key with mass{key >= K} > target (0 if none): the token(s) with key K are the first whose preceding mass
        // (mass{key > K}) is <= target -> kept; everything below K is removed
        p_key = K;
        float s2 = 0.f;
        if (listed) { for (unsigned int j = tid; j < total; j += P2_THREADS) { const float e = __uint_as_float(list[j].y); if (float_order_key(e) >= p_key) s2 += e; } }
        else p2_row_pass(a, lc, lu, cfg, [&](int, float z) { if (float_order_key(z) >= thr_key) { const float e = expf(z - mx); if (float_order_key(e) >= p_key) s2 += e; } });
        sum = p2_block_sum(s2, red8);                                       // soft-max over the kept logits only
    }
    // ---- draw: arg-max of p (greedy) or of p / q (exponential race, == torch.multinomial); lowest index wins ties
    float best = -1.f; int besti = 0x7fffffff;
    const float* nz = a.noise ? a.noise + ((size_t)(a.noise_per_step ? step : 0) * a.B + img) * V : nullptr;
    auto consider = [&](int i, float e) {
        if (float_order_key(e) < p_key || !(e > 0.f)) return;
        float s = e / sum;
        if (a.sample_logits) { const float q = nz ? nz[i] : exp1_noise(a.seed_lo, a.seed_hi, i, img, step); s = s / q; }
        if (s > best || (s == best && i < besti)) { best = s; besti = i; }
    };
    if (listed) { for (unsigned int j = tid; j < total; j += P2_THREADS) consider((int)list[j].x, __uint_as_float(list[j].y)); }
    else p2_row_pass(a, lc, lu, cfg, [&](int i, float z) { if (float_order_key(z) >= thr_key) consider(i, expf(z - mx)); });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __syncthreads();
    if (lane == 0) { red8[warp] = best; reinterpret_cast<int*>(wsum)[warp] = besti; }
    __syncthreads();
    float bb = red8[0]; int bi = reinterpret_cast<int*>(wsum)[0];
#pragma unroll
    for (int w = 1; w < P2_WARPS; ++w) {
        const float fb = red8[w]; const int ib = reinterpret_cast<int*>(wsum)[w];
        if (fb > bb || (fb == bb && ib < bi)) { bb = fb; bi = ib; }
    }
    __syncthreads();
    return bi;
}

// next-token input rows as H2 packets: h = tok_embeddings[tok] (+ cs * ctrl0[row][pos_next - T + 1])  gpt_t2i.py:445,466
__device__ __forceinline__ void p2_write_embedding(const P2Params& P, uint2* h2, unsigned int tag, int lrow, int grow, int tok, int pos_next) {
    const bf16* e = P.tok_emb + (size_t)tok * P.dim;
    const int p = pos_next - P.T + 1;
    const bf16* c = (P.has_ctrl && p >= 0 && p < P.n_img) ? P.ctrl[0] + ((size_t)grow * P.n_img + p) * P.dim : nullptr;
    for (int k2 = threadIdx.x; k2 < (P.dim >> 1); k2 += P2_THREADS) {
        float v0, v1;
        unpack_bf16x2(*reinterpret_cast<const uint32_t*>(e + 2 * k2), v0, v1);
        if (c) {
            float c0, c1;
            unpack_bf16x2(*reinterpret_cast<const uint32_t*>(c + 2 * k2), c0, c1);
            v0 = rnd<bf16>(v0 + rnd<bf16>(P.cs * c0)); v1 = rnd<bf16>(v1 + rnd<bf16>(P.cs * c1));
        }
        pk_st64(h2 + p2_a_index(lrow, 2 * k2), pk_pack(v0, v1), tag);
    }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P2_THREADS, 2) pk2_decode_kernel(const __grid_constant__ P2Params P) {
    extern __shared__ __align__(128) unsigned char p2_smem_raw[];
    __shared__ int s_tok;
    __shared__ int s_lo[5], s_hi[5];
    __shared__ P2SampSh s_samp;
    __shared__ float s_red8[P2_WARPS];
    __shared__ unsigned int s_wsum[P2_WARPS];
    P2Smem sm;
    {
        unsigned char* q = p2_smem_raw;
        sm.ring = q; q += P2_SMEM_RING;
        sm.red = reinterpret_cast<float*>(q); q += P2_SMEM_RED;
        sm.ssq = reinterpret_cast<float*>(q); q += P2_WARPS * 8 * 4;                 // 256
        sm.full = reinterpret_cast<uint64_t*>(q); q += 64;
        sm.empty = reinterpret_cast<uint64_t*>(q); q += 64;
        sm.st = reinterpret_cast<P2Stream*>(q); q += 128;
        sm.own = reinterpret_cast<uint32_t*>(q); q += 2 * 32 * 4;                    // 256
        sm.qrow = reinterpret_cast<uint32_t*>(q); q += 3 * P2_MAXSEG * 32 * 4;       // 2304
        sm.plan = reinterpret_cast<PkAttnPlan*>(q);                                  // <= 1024
    }
    static_assert(sizeof(PkAttnPlan) <= P2_SMEM_MISC - (256 + 64 + 64 + 128 + 256 + 2304), "plan does not fit");
    static_assert(sizeof(P2Stream) <= 128, "stream cursor does not fit");
    const int tid = threadIdx.x;
    P2Mb mb;
    mb.mb = (int)blockIdx.x % P.nmb; mb.c = (int)blockIdx.x / P.nmb; mb.Gc = (int)gridDim.x / P.nmb;
    mb.img_lo = P.img_lo[mb.mb]; mb.cnt = P.img_cnt[mb.mb]; mb.use_cfg = P.smp.use_cfg; mb.B = P.B;
    mb.M = mb.use_cfg ? 2 * mb.cnt : mb.cnt;
    const int Gc = mb.Gc;
    if (tid == 0) {
        for (int s = 0; s < P2_NSLOT; ++s) { pk_mbar_init(&sm.full[s], 1); pk_mbar_init(&sm.empty[s], P2_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == P2_THREADS - 32) {
        const int* pt = P.part;
        P2Stream& st = *sm.st;
        st.lo[0] = pt[0 * (Gc + 1) + mb.c]; st.hi[0] = pt[0 * (Gc + 1) + mb.c + 1];
        st.lo[1] = pt[1 * (Gc + 1) + mb.c]; st.hi[1] = pt[1 * (Gc + 1) + mb.c + 1];
        st.lo[2] = 2 * pt[2 * (Gc + 1) + mb.c]; st.hi[2] = 2 * pt[2 * (Gc + 1) + mb.c + 1];
        st.lo[3] = st.lo[1]; st.hi[3] = st.hi[1];
        st.lo[4] = pt[3 * (Gc + 1) + mb.c]; st.hi[4] = pt[3 * (Gc + 1) + mb.c + 1];
        for (int i = 0; i < 5; ++i) { s_lo[i] = st.lo[i]; s_hi[i] = st.hi[i]; }
        st.c.step = 0; st.c.l = 0; st.c.phase = 0; st.c.blk = 0; st.c.sub = 0; st.c.done = (P.n_steps <= 1) || mb.M == 0;
        st.issued = 0;
        p2_cursor_skip_empty(P, st, st.c);
        p2_producer_advance(P, sm.ring, sm.full, sm.empty, sm.st);          // primes the ring (the first P2_NSLOT units need no release)
    }
    __syncthreads();
    unsigned int cons = 0;
    unsigned int gen = P.bar_base;
    const unsigned int tstride = (unsigned int)P.L + 1u;

    for (int step = 0; step < P.n_steps; ++step) {
        const int pos = P.T - 1 + step;                    // logits of this position are sampled now
        const unsigned int tag0 = P.tag_base + (unsigned int)step * tstride + 1u;   // tag(step, 0)
#ifdef PK_TRACE
        const bool dbg_step = P.dbg != nullptr && step == P.dbg_step;
#else
        constexpr bool dbg_step = false;
#endif
        long long* const dbg_cta = P.dbg + (size_t)blockIdx.x * 64;
        if (dbg_step && tid == 0) dbg_cta[0] = pk_now();
        // ---------------- sampler (+ embedding of the chosen token for position pos + 1) ----------------
        if (mb.c < mb.cnt) {
            const int img = mb.img_lo + mb.c;
#ifdef PK_TRACE
            int* sflag = (P.nanflag != nullptr && step < 8) ? P.nanflag + ((mb.mb * 8 + step) * (P.L + 1) + 0) * 16 + 14 : nullptr;
#else
            int* sflag = nullptr;
#endif
            const int tok_s = p2_sample(P.smp, img, step, sm.red, &s_samp, s_red8, s_wsum, sflag);
            if (sflag != nullptr && tid == 0 && (tok_s < 0 || tok_s >= P.V)) atomicOr(sflag, 4);
            if (tid == 0) {
                P.smp.idx_out[(size_t)img * P.smp.tokens_ld + step] = tok_s;
                s_tok = P.forced != nullptr ? __ldg(P.forced + (size_t)img * P.forced_ld + step) : tok_s;
            }
            __syncthreads();
            if (step + 1 < P.n_steps) {
                const int tok = s_tok;
                p2_write_embedding(P, P.h2[mb.mb][0], tag0, mb.c, img, tok, pos + 1);
                if (mb.use_cfg) p2_write_embedding(P, P.h2[mb.mb][0], tag0, mb.c + mb.cnt, P.B + img, tok, pos + 1);
            }
        }
        if (dbg_step && tid == 0) dbg_cta[1] = pk_now();
        if (step + 1 == P.n_steps) break;
        const int p = pos + 1;                             // position being decoded
        if (mb.M > 0) {
            // per-token attention work split (depends on the context length only)
            if (tid < 2 * PKP_WARPS + PKP_MAXSEG) pkp_fill(*sm.plan, tid, mb.c, Gc, mb.M * P.H, P.H, p + 1);
            __syncthreads();
            for (int l = 0; l <= P.L; ++l) {
                const int par = l & 1;
                const unsigned int tag = tag0 + (unsigned int)l;
                const int nph = l < P.L ? 5 : 1;
                for (int ph = 0; ph < nph; ++ph) {
                    long long* dbg = (dbg_step && l == 3) ? dbg_cta + 8 + 8 * ph : nullptr;
                    if (l < P.L && ph == 1) { p2_attn_phase(P, sm, mb, l, p, tag, par, dbg, step); continue; }
                    const int kind = l == P.L ? 4 : (ph == 0 ? 0 : ph - 1);
                    p2_gemm_phase(P, sm, mb, kind, l, p, tag, s_lo[kind], s_hi[kind], cons,
                                  (kind == 4 && P.trace != nullptr) ? P.trace + (size_t)(step + 1) * P.b_eff * P.V : nullptr, dbg, step);
                }
            }
        }
        if (dbg_step && tid == 0) dbg_cta[3] = pk_now();
        pk_grid_sync(P.bar, gen);                          // logits complete; KV rows of this token ordered; the two chains re-aligned
        if (dbg_step && tid == 0) dbg_cta[4] = pk_now();
    }
}
