// car_api.cu — C-ABI entry points (include/controlar_b200.h) and the host-side chaining of the AR kernels.
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "common.cuh"
#include "gemm_skinny.cuh"
#include "attention.cuh"
#include "sampler.cuh"
#include "misc.cuh"
#include "decode_persistent.cuh"
#include "gemm_dense.cuh"
#include "gemm_tc5.cuh"
#include "train.cuh"
#include "train_bwd.cuh"
#include "t5.cuh"
#include <algorithm>

thread_local std::string g_car_err;
std::atomic<long long> g_car_launches{0};

// development knobs (environment, read once): CAR_MEGA (0 = per-kernel graph chain instead of the persistent decode kernel),
// CAR_DBG (trace step; needs a -DPK_TRACE build), CAR_EXP (experiment bits of decode_persistent.cuh), CAR_TC5 (0 = mma.sync dense
// GEMM in the prefill), CAR_PDL, CAR_L2PF, CAR_NSPLIT, CAR_NB_QKV/WO/W13/W2/HEAD (per-kernel chain)
struct Tune { int pdl, l2pf, nsplit, nb[5], skip, empty, dbg, mega, mega_pf; };
static long long* g_dbg = nullptr;   // [5 kernels][8] clock stamps (CAR_DBG=1)
__global__ void empty_kernel(int) {}
static const Tune& tune() {
    static Tune t = [] {
        Tune x;
        auto gi = [](const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; };
        x.mega = gi("CAR_MEGA", 1); x.mega_pf = gi("CAR_MEGA_PF", 1); x.dbg = gi("CAR_DBG", 0); x.skip = gi("CAR_SKIP", 0); x.empty = gi("CAR_EMPTY", 0);
        x.pdl = gi("CAR_PDL", 1); x.l2pf = gi("CAR_L2PF", 1); x.nsplit = gi("CAR_NSPLIT", 0);
        x.nb[EPI_STORE] = 0; x.nb[EPI_QKV] = gi("CAR_NB_QKV", 0); x.nb[EPI_RESID] = gi("CAR_NB_RESID", 0);
        x.nb[EPI_SWIGLU] = gi("CAR_NB_W13", 0); x.nb[EPI_LOGITS] = gi("CAR_NB_HEAD", 0);
        return x;
    }();
    return t;
}

// ---------------------------------------------------------------------------------------------------------
// structures
// ---------------------------------------------------------------------------------------------------------
struct CarModel {
    CarModelDesc d;
    // borrowed originals
    const void *tok_emb, *norm, *output, *cap_fc1, *cap_fc2, *label_table, *cond_fc1, *cond_fc2, *ctl_fc1[3], *ctl_fc2[3];
    std::vector<const void*> attention_norm, wqkv, wo, ffn_norm, w1, w3, w2;
    // owned GEMM-ready copies: bf16 -> fragment-packed; fp32 -> plain (only w13 is an owned interleaved copy)
    std::vector<void*> g_wqkv, g_wo, g_w13, g_w2;
    void *g_output, *g_cap_fc1, *g_cap_fc2, *g_cond_fc1, *g_cond_fc2, *g_ctl_fc1[3], *g_ctl_fc2[3];
    unsigned int pack_gen = 0;       // bumped by every (re)pack: states refresh their device pointer tables when it moves
    std::vector<void*> owned;
    size_t esize() const { return d.dtype == CAR_BF16 ? 2 : 4; }
};

struct CarState {
    CarModel* m;
    int b_eff, S, N, T;
    std::vector<void*> kc, vc;
    const float* rope;
    int* emb_mask;       // [b_eff][T] or null (= all ones)
    int* emb_mask_store;
    // decode scratch (b_eff rows)
    void *h, *q, *attn, *act;
    float* logits;       // [b_eff][V]
    int *tok, *pos, *done_ctr, *tickets, *tokens;
    float* attn_part;
    int nsplit;
    // control tokens [3][b_eff][N][d]
    void* ctrl[3];
    bool has_ctrl;
    float cs;
    // prefill scratch
    void *hP, *qP, *attnP, *actP, *t1, *t2;
    void *qkvP, *gP, *uP;     // dense prefill path (bf16): qkv [rows][3d], w1 / w3 outputs [rows][F]
    bool prefilled;
    // decode graph
    cudaGraphExec_t gexec;
    cudaStream_t cap_stream;   // capture happens on a private stream (the legacy default stream cannot capture)
    bool graph_ok; unsigned int graph_pack_gen;
    CarSampling gsp;
    const float* gnoise;
    // persistent decode kernel (decode_persistent.cuh)
    void** pk_ptrs;          // device arrays of per-layer pointers [8][L]
    int* pk_part;            // [4][grid + 1] block offsets per CTA
    uint2 *pk_h2[2], *pk_h1[2], *pk_att[2], *pk_act[2], *pk_qkv[2], *pk_partial[2];
    int pk_part_slots, pk_grid; bool pk_ok; unsigned int pk_ptrs_gen;
    unsigned int* pk_bar; unsigned int pk_bar_count, pk_tag_gen;
    size_t pk_pkt_bytes; void* pk_pkt_base;
    long long* pk_step_ts;   // caller-provided device buffer [N] for per-step timestamps (car_state_set_step_timer) or null
    std::vector<void*> owned;
};

static int alloc_dev(std::vector<void*>& owned, void** p, size_t bytes) {
    CAR_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    owned.push_back(*p);
    return CAR_OK;
}

// Device buffers that other SMs POLL (packet tags, barrier counters) are initialised with SM stores, not cudaMemset: on B200 a
// recycled allocation that was zeroed by cudaMemset has been observed to still return the previous owner's packets to strong
// polling loads (tests/test_ar_gpu.py run as a whole failed deterministically until tags were made unique per state;
// profiles/r2_stale_tags.md).  Two defences: this fill kernel, and tags that are unique process-wide (pk_alloc_tags).
__global__ void fill_u32_kernel(unsigned int* __restrict__ p, unsigned int v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
static int fill_u32(void* p, unsigned int v, size_t bytes, cudaStream_t st) {
    const size_t n = bytes / 4;
    CAR_LAUNCH(fill_u32_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, st, (unsigned int*)p, v, n);
    return CAR_OK;
}
// Process-wide tag allocator of the persistent decode kernel: every launch gets a fresh, never-reused range of packet tags, so
// a packet left in memory by ANY earlier launch or state can never be mistaken for a current one.  (2^32 tags last for ~10^5
// full-size generate() calls; on wrap-around every state re-zeroes its packet buffers before its next launch.)
static std::atomic<unsigned int> g_pk_tag_next{1u};
static std::atomic<unsigned int> g_pk_tag_gen{0u};
static unsigned int pk_alloc_tags(unsigned int span, unsigned int* gen_out) {
    for (;;) {
        unsigned int base = g_pk_tag_next.load();
        if (base > 0xF0000000u || base + span < base) {            // wrap: new generation, tags restart at 1
            unsigned int expected = base;
            if (g_pk_tag_next.compare_exchange_strong(expected, 1u)) g_pk_tag_gen.fetch_add(1u);
            continue;
        }
        if (g_pk_tag_next.compare_exchange_weak(base, base + span)) { *gen_out = g_pk_tag_gen.load(); return base; }
    }
}

extern "C" const char* car_last_error(void) { return g_car_err.c_str(); }
extern "C" int car_version(void) { return 100; }
extern "C" int64_t car_launch_count(int32_t reset) {
    long long v = g_car_launches.load();
    if (reset) g_car_launches.store(0);
    return v;
}

// SM count of the CURRENT device (the Python handles make the tensors' device current around every call)
static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); cached[dev] = n > 0 ? n : 148; }
    return cached[dev];
}

// ---------------------------------------------------------------------------------------------------------
// skinny GEMM dispatch
// ---------------------------------------------------------------------------------------------------------
template <int NB, int U, bool NORM>
static int launch_skinny_bf16_inst(cudaStream_t st, const bf16* A, int lda, const void* Wp, const bf16* nw, float eps,
                                   int M, int nblk, int K, const EpiParams& ep) {
    const size_t smem = skinny_smem_bytes(K, NB);
    static DevOnce once;
    if (once.first()) {
        CAR_CUDA(cudaFuncSetAttribute(skinny_gemm_bf16<NB, U, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CAR_CUDA(cudaFuncSetAttribute(skinny_gemm_bf16<NB, U, NORM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    if (smem > 200 * 1024) CAR_FAIL(CAR_ERR_UNSUPPORTED, "K too large for the shared-memory activation tile");
    dim3 grid((nblk + NB - 1) / NB, (M + 15) / 16);
    if (grid.y == 1) {   // decode shape: never more than one wave of CTAs; a CTA strides over its column groups
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, skinny_gemm_bf16<NB, U, NORM>, SK_THREADS, smem);
        const unsigned cap = (unsigned)sm_count() * (unsigned)std::max(1, std::min(occ, 2));
        if (grid.x > cap) grid.x = cap;
    }
    const int flags = tune().l2pf ? 1 : 0;
    if (tune().empty) { CAR_LAUNCH(empty_kernel, grid, SK_THREADS, smem, st, 0); return CAR_OK; }
    if (tune().pdl) CAR_LAUNCH_PDL((skinny_gemm_bf16<NB, U, NORM>), grid, dim3(SK_THREADS), smem, st, A, lda, (const uint4*)Wp, nw, eps, K, nblk, flags, ep);
    else CAR_LAUNCH((skinny_gemm_bf16<NB, U, NORM>), grid, SK_THREADS, smem, st, A, lda, (const uint4*)Wp, nw, eps, K, nblk, flags, ep);
    return CAR_OK;
}

template <int NB, bool NORM>
static int launch_skinny_bf16_u(cudaStream_t st, int U, const bf16* A, int lda, const void* Wp, const bf16* nw, float eps,
                                int M, int nblk, int K, const EpiParams& ep) {
    switch (U) {
        case 5: if (NB <= 2) return launch_skinny_bf16_inst<NB, (NB <= 2 ? 5 : 4), NORM>(st, A, lda, Wp, nw, eps, M, nblk, K, ep);
        case 7: if (NB <= 2) return launch_skinny_bf16_inst<NB, (NB <= 2 ? 7 : 4), NORM>(st, A, lda, Wp, nw, eps, M, nblk, K, ep);
        case 2: if (NB == 8) return launch_skinny_bf16_inst<NB, (NB == 8 ? 2 : 4), NORM>(st, A, lda, Wp, nw, eps, M, nblk, K, ep);
        default: return launch_skinny_bf16_inst<NB, (NB == 8 ? 2 : 4), NORM>(st, A, lda, Wp, nw, eps, M, nblk, K, ep);
    }
}

// W: bf16 -> packed (pack_weight_bf16_kernel) ; fp32 -> plain [N][K].  N = 8*nblk rows of W.
static int launch_skinny(cudaStream_t st, int dtype, const void* A, int lda, const void* W, const void* nw, float eps,
                         int M, int N, int K, EpiParams ep, bool norm) {
    if (K % 64 != 0 || N % 8 != 0) CAR_FAIL(CAR_ERR_UNSUPPORTED, "skinny GEMM needs K % 64 == 0 and N % 8 == 0");
    if (M <= 0) return CAR_OK;
    ep.M = M;
    const int nblk = N / 8;
    if (dtype == CAR_F32) {
        dim3 grid((nblk + 1) / 2, (M + 15) / 16);
        if (norm) CAR_LAUNCH((skinny_gemm_f32<true>), grid, SK_THREADS, 0, st, (const float*)A, lda, (const float*)W, (const float*)nw, eps, K, nblk, ep);
        else CAR_LAUNCH((skinny_gemm_f32<false>), grid, SK_THREADS, 0, st, (const float*)A, lda, (const float*)W, (const float*)nw, eps, K, nblk, ep);
        return CAR_OK;
    }
    const int mtiles = (M + 15) / 16;
    const int nsteps = ((K >> 5) + SK_WARPS - 1) / SK_WARPS;
    int U = 4;
    if (nsteps % 7 == 0) U = 7; else if (nsteps % 5 == 0) U = 5;
    int NB;
    if (mtiles > 1) NB = 4;                       // M-tiled (prefill): maximise reuse of the activation tile
    else {
        // one wave: ceil(nblk / NB) <= 148 where possible (register use makes these 1 CTA / SM kernels)
        const int want = (nblk + sm_count() - 1) / sm_count();
        NB = want > 4 ? 8 : (want > 2 ? 4 : (want > 1 ? 2 : 1));
    }
    if (mtiles == 1 && tune().nb[ep.kind] > 0) NB = tune().nb[ep.kind];
    if (ep.kind == EPI_SWIGLU && NB == 1) NB = 2;
    if (NB == 8) U = 2; else if (NB * U > 16) U = 4;   // register budget
    const bf16* Ab = (const bf16*)A; const bf16* nwb = (const bf16*)nw;
#define CAR_SK(NB_) (norm ? launch_skinny_bf16_u<NB_, true>(st, U, Ab, lda, W, nwb, eps, M, nblk, K, ep) \
                          : launch_skinny_bf16_u<NB_, false>(st, U, Ab, lda, W, nwb, eps, M, nblk, K, ep))
    switch (NB) {
        case 8: return CAR_SK(8);
        case 4: return CAR_SK(4);
        case 2: return CAR_SK(2);
        default: return CAR_SK(1);
    }
#undef CAR_SK
}

// ---------------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------------
static int pack_one(CarModel* m, cudaStream_t st, const void* w, const void* w3, int N, int K, bool interleave, void** dst,
                    bool allocate) {
    // N = rows of the logical (possibly interleaved) matrix
    if (m->d.dtype == CAR_BF16) {
        if (allocate) CAR_TRY(alloc_dev(m->owned, dst, (size_t)N * K * 2));
        const int nblk = N / 8;
        const long long total = (long long)nblk * (K / 32) * 32;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 4096);
        CAR_LAUNCH(pack_weight_bf16_kernel, blocks, 256, 0, st, (const bf16*)w, (const bf16*)w3, (uint4*)*dst, nblk, K, interleave ? 1 : 0);
    } else {
        if (!interleave) { *dst = const_cast<void*>(w); return CAR_OK; }
        if (allocate) CAR_TRY(alloc_dev(m->owned, dst, (size_t)N * K * 4));
        CAR_LAUNCH(interleave_rows_f32_kernel, 2048, 256, 0, st, (const float*)w, (const float*)w3, (float*)*dst, N / 2, K);
    }
    return CAR_OK;
}

static int model_pack_all(CarModel* m, const CarWeights* w, cudaStream_t st, bool allocate) {
    const CarModelDesc& d = m->d;
    const int L = d.n_layer;
    m->tok_emb = w->tok_embeddings; m->norm = w->norm; m->output = w->output;
    m->cap_fc1 = w->cap_fc1; m->cap_fc2 = w->cap_fc2; m->label_table = w->label_table;
    m->cond_fc1 = w->cond_fc1; m->cond_fc2 = w->cond_fc2;
    for (int j = 0; j < 3; ++j) { m->ctl_fc1[j] = w->ctl_fc1[j]; m->ctl_fc2[j] = w->ctl_fc2[j]; }
    m->attention_norm.assign(w->attention_norm, w->attention_norm + L);
    m->wqkv.assign(w->wqkv, w->wqkv + L); m->wo.assign(w->wo, w->wo + L);
    m->ffn_norm.assign(w->ffn_norm, w->ffn_norm + L);
    m->w1.assign(w->w1, w->w1 + L); m->w3.assign(w->w3, w->w3 + L); m->w2.assign(w->w2, w->w2 + L);
    if (allocate) { m->g_wqkv.assign(L, nullptr); m->g_wo.assign(L, nullptr); m->g_w13.assign(L, nullptr); m->g_w2.assign(L, nullptr); }
    for (int l = 0; l < L; ++l) {
        CAR_TRY(pack_one(m, st, m->wqkv[l], nullptr, 3 * d.dim, d.dim, false, &m->g_wqkv[l], allocate));
        CAR_TRY(pack_one(m, st, m->wo[l], nullptr, d.dim, d.dim, false, &m->g_wo[l], allocate));
        CAR_TRY(pack_one(m, st, m->w1[l], m->w3[l], 2 * d.ffn_dim, d.dim, true, &m->g_w13[l], allocate));
        CAR_TRY(pack_one(m, st, m->w2[l], nullptr, d.dim, d.ffn_dim, false, &m->g_w2[l], allocate));
    }
    ++m->pack_gen;
    CAR_TRY(pack_one(m, st, m->output, nullptr, d.vocab_size, d.dim, false, &m->g_output, allocate));
    if (d.model_type == 1) {
        if (!m->cap_fc1 || !m->cap_fc2) CAR_FAIL(CAR_ERR_ARG, "t2i model needs cap_fc1/cap_fc2");
        CAR_TRY(pack_one(m, st, m->cap_fc1, nullptr, d.dim, d.caption_dim, false, &m->g_cap_fc1, allocate));
        CAR_TRY(pack_one(m, st, m->cap_fc2, nullptr, d.dim, d.dim, false, &m->g_cap_fc2, allocate));
    } else if (!m->label_table) CAR_FAIL(CAR_ERR_ARG, "c2i model needs label_table");
    CAR_TRY(pack_one(m, st, m->cond_fc1, nullptr, d.dim, d.dim, false, &m->g_cond_fc1, allocate));
    CAR_TRY(pack_one(m, st, m->cond_fc2, nullptr, d.dim, d.dim, false, &m->g_cond_fc2, allocate));
    for (int j = 0; j < 3; ++j) {
        CAR_TRY(pack_one(m, st, m->ctl_fc1[j], nullptr, d.dim, d.dim, false, &m->g_ctl_fc1[j], allocate));
        CAR_TRY(pack_one(m, st, m->ctl_fc2[j], nullptr, d.dim, d.dim, false, &m->g_ctl_fc2[j], allocate));
    }
    return CAR_OK;
}

extern "C" int car_model_create(const CarModelDesc* desc, const CarWeights* w, void* stream, CarModel** out) {
    if (!desc || !w || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    const CarModelDesc& d = *desc;
    if (d.dtype != CAR_BF16 && d.dtype != CAR_F32) CAR_FAIL(CAR_ERR_ARG, "dtype must be CAR_BF16 or CAR_F32");
    if (d.n_head <= 0 || d.dim != d.n_head * 64) CAR_FAIL(CAR_ERR_UNSUPPORTED, "head_dim must be 64");
    if (d.n_layer <= 0 || d.n_layer % 3 != 0) CAR_FAIL(CAR_ERR_UNSUPPORTED, "n_layer must be a multiple of 3 (gpt_t2i.py:320,457)");
    if (d.dim % 64 || d.ffn_dim % 64 || d.vocab_size % 8) CAR_FAIL(CAR_ERR_UNSUPPORTED, "dim, ffn_dim must be multiples of 64; vocab of 8");
    if (d.model_type == 1 && d.caption_dim % 64) CAR_FAIL(CAR_ERR_UNSUPPORTED, "caption_dim must be a multiple of 64");
    if (d.cls_token_num < 1 || d.cls_token_num > 256) CAR_FAIL(CAR_ERR_UNSUPPORTED, "cls_token_num must be in [1,256]");
    CarModel* m = new CarModel();
    m->d = d;
    int r = model_pack_all(m, w, (cudaStream_t)stream, true);
    if (r != CAR_OK) { for (void* p : m->owned) cudaFree(p); delete m; return r; }
    *out = m;
    return CAR_OK;
}

extern "C" int car_model_repack(CarModel* m, const CarWeights* w, void* stream) {
    if (!m || !w) CAR_FAIL(CAR_ERR_ARG, "null argument");
    return model_pack_all(m, w, (cudaStream_t)stream, false);
}

extern "C" int car_model_destroy(CarModel* m) {
    if (!m) return CAR_OK;
    for (void* p : m->owned) cudaFree(p);
    delete m;
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------------------
// block ownership of the persistent decode kernel: counts per CTA balanced on streamed bytes per token
static void pk_partition(const CarModelDesc& d, int G, std::vector<int>& table) {
    const int nQ = 3 * d.dim / 8, nD = d.dim / 8, nP = d.ffn_dim / 8, nH = d.vocab_size / 8;
    const double L = d.n_layer;
    const double wQ = L * 16.0 * d.dim, wD = L * 16.0 * (d.dim + d.ffn_dim), wP = L * 32.0 * d.dim, wH = 16.0 * d.dim;
    std::vector<double> load(G, 0.0);
    std::vector<int> cQ(G, 0), cD(G, 0), cP(G, 0), cH(G, 0);
    for (int i = 0; i < nD; ++i) { cD[i % G] += 1; load[i % G] += wD; }
    auto spread = [&](int n, double w, std::vector<int>& cnt) {
        for (int i = 0; i < n; ++i) {
            int best = 0;
            for (int c = 1; c < G; ++c) if (load[c] < load[best] - 1e-9) best = c;
            cnt[best] += 1; load[best] += w;
        }
    };
    spread(nP, wP, cP); spread(nQ, wQ, cQ);
    (void)wH;
    for (int i = 0; i < nH; ++i) cH[i % G] += 1;      // the head ends in a grid barrier: its latency, not its bytes, is what counts
    table.assign(4 * (G + 1), 0);
    const std::vector<int>* cs[4] = {&cQ, &cD, &cP, &cH};
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < G; ++c) table[k * (G + 1) + c + 1] = table[k * (G + 1) + c] + (*cs[k])[c];
}

// per-layer device pointers of the persistent kernel: packed weights (library-owned), norm weights (borrowed from the module: they
// move when a parameter tensor is replaced, hence the refresh in launch_pk after car_model_repack), KV caches
static std::vector<const void*> pk_pointer_table(const CarState* s) {
    const int L = s->m->d.n_layer;
    std::vector<const void*> hp(8 * L);
    for (int l = 0; l < L; ++l) {
        hp[0 * L + l] = s->m->g_wqkv[l]; hp[1 * L + l] = s->m->g_wo[l]; hp[2 * L + l] = s->m->g_w13[l]; hp[3 * L + l] = s->m->g_w2[l];
        hp[4 * L + l] = s->m->attention_norm[l]; hp[5 * L + l] = s->m->ffn_norm[l]; hp[6 * L + l] = s->kc[l]; hp[7 * L + l] = s->vc[l];
    }
    return hp;
}

static int pk_state_setup(CarState* s) {
    const CarModelDesc& d = s->m->d;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int G = sms, L = d.n_layer;
    s->pk_grid = G;
    const int nbh = s->b_eff * d.n_head;
    // shapes the kernel is instantiated for (else car_generate falls back to the per-kernel graph chain)
    s->pk_ok = s->b_eff <= 16 && d.dim % 32 == 0 && d.dim <= 16 * 3 * 32 && d.dim <= PK_UNIT_KS * 32 /* one ring unit per block of the K = dim GEMMs */ && d.ffn_dim % 32 == 0 && d.ffn_dim <= 16 * 7 * 32 &&
               d.dim / 8 <= 2 * G && d.vocab_size <= 16384 && d.ffn_dim % 8 == 0 && nbh <= 5 * G &&
               2 * ((d.ffn_dim / 32 + PK_UNIT_KS - 1) / PK_UNIT_KS) <= PK_NSLOT && L <= PK_MAXL;
    if (!s->pk_ok) return CAR_OK;
    std::vector<const void*> hp = pk_pointer_table(s);
    std::vector<int> table;
    pk_partition(d, G, table);
    s->pk_part_slots = G / std::max(1, nbh) + 3;
    // packet buffers: two parities of H2, H1, ATT (K = dim), ACT (K = ffn), QKV, PARTIAL; one allocation, zeroed (tag 0 = never)
    const size_t a_d = (size_t)(d.dim / 32) * 2048, a_f = (size_t)(d.ffn_dim / 32) * 2048;
    const size_t qkv_b = (size_t)3 * 16 * d.n_head * 8 * 4 * 8, part_b = (size_t)nbh * s->pk_part_slots * 66 * 8;
    const size_t total = 2 * (3 * a_d + a_f + qkv_b + part_b);
    CAR_TRY(alloc_dev(s->owned, (void**)&s->pk_ptrs, hp.size() * sizeof(void*)));
    CAR_TRY(alloc_dev(s->owned, (void**)&s->pk_part, table.size() * sizeof(int)));
    CAR_TRY(alloc_dev(s->owned, (void**)&s->pk_bar, 64));
    CAR_TRY(alloc_dev(s->owned, &s->pk_pkt_base, total));
    s->pk_pkt_bytes = total;
    unsigned char* q = (unsigned char*)s->pk_pkt_base;
    for (int par = 0; par < 2; ++par) {
        s->pk_h2[par] = (uint2*)q; q += a_d; s->pk_h1[par] = (uint2*)q; q += a_d; s->pk_att[par] = (uint2*)q; q += a_d;
        s->pk_act[par] = (uint2*)q; q += a_f; s->pk_qkv[par] = (uint2*)q; q += qkv_b; s->pk_partial[par] = (uint2*)q; q += part_b;
    }
    CAR_CUDA(cudaMemcpy(s->pk_ptrs, hp.data(), hp.size() * sizeof(void*), cudaMemcpyHostToDevice));
    s->pk_ptrs_gen = s->m->pack_gen;
    CAR_CUDA(cudaMemcpy(s->pk_part, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
    CAR_TRY(fill_u32(s->pk_bar, 0u, 64, nullptr));
    CAR_TRY(fill_u32(s->pk_pkt_base, 0u, total, nullptr));
    CAR_CUDA(cudaStreamSynchronize(nullptr));          // state creation is rare; the caller's stream may not be ordered after the NULL stream
    return CAR_OK;
}

extern "C" int car_state_create(CarModel* m, int32_t b_eff, int32_t S, int32_t N, void* const* k_cache, void* const* v_cache,
                                const float* rope_table, CarState** out) {
    if (!m || !k_cache || !v_cache || !rope_table || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    const CarModelDesc& d = m->d;
    const int T = d.cls_token_num;
    if (b_eff <= 0 || N <= 0 || S < T + N) CAR_FAIL(CAR_ERR_ARG, "need b_eff > 0, N > 0, S >= T + N");
    if (N > d.block_size) CAR_FAIL(CAR_ERR_ARG, "N exceeds block_size (RoPE table rows)");
    CarState* s = new CarState();
    s->m = m; s->b_eff = b_eff; s->S = S; s->N = N; s->T = T;
    s->kc.assign(k_cache, k_cache + d.n_layer); s->vc.assign(v_cache, v_cache + d.n_layer);
    s->rope = rope_table; s->emb_mask = nullptr; s->has_ctrl = false; s->cs = 1.f; s->prefilled = false;
    s->gexec = nullptr; s->graph_ok = false; s->gnoise = nullptr; s->cap_stream = nullptr;
    if (cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; CAR_FAIL(CAR_ERR_CUDA, "cudaStreamCreateWithFlags failed"); }
    const size_t es = m->esize();
    const size_t dd = d.dim, F = d.ffn_dim, V = d.vocab_size;
    const size_t MP = (size_t)b_eff * T, MC = (size_t)b_eff * N;
    s->nsplit = std::max(1, std::min(16, (4 * sm_count() + b_eff * d.n_head - 1) / (b_eff * d.n_head)));
    if (tune().nsplit > 0) s->nsplit = std::min(32, tune().nsplit);
    int r = CAR_OK;
    auto A = [&](void** p, size_t bytes) { if (r == CAR_OK) r = alloc_dev(s->owned, p, bytes); };
    A(&s->h, b_eff * dd * es); A(&s->q, b_eff * dd * es); A(&s->attn, b_eff * dd * es); A(&s->act, b_eff * F * es);
    A((void**)&s->logits, b_eff * V * 4); A((void**)&s->tok, b_eff * 4); A((void**)&s->pos, 4 * 4);
    A((void**)&s->tickets, (size_t)b_eff * d.n_head * 4);
    A((void**)&s->tokens, (size_t)b_eff * N * 4);
    A((void**)&s->attn_part, (size_t)b_eff * d.n_head * s->nsplit * AD_PART * 4);
    for (int j = 0; j < 3; ++j) A(&s->ctrl[j], MC * dd * es);
    A(&s->hP, MP * dd * es); A(&s->qP, MP * dd * es); A(&s->attnP, MP * dd * es); A(&s->actP, MP * F * es);
    A(&s->t1, std::max(MC, MP) * dd * es); A(&s->t2, std::max(MC, MP) * dd * es);
    s->qkvP = s->gP = s->uP = nullptr;
    if (d.dtype == CAR_BF16) { A(&s->qkvP, MP * 3 * dd * es); A(&s->gP, MP * F * es); A(&s->uP, MP * F * es); }
    A((void**)&s->emb_mask, MP * 4);
    if (r == CAR_OK && cudaMemset(s->tickets, 0, (size_t)b_eff * d.n_head * 4) != cudaSuccess) r = CAR_ERR_CUDA;
    if (r == CAR_OK && cudaMemset(s->pos, 0, 16) != cudaSuccess) r = CAR_ERR_CUDA;
    if (r != CAR_OK) { for (void* p : s->owned) cudaFree(p); delete s; return r; }
    s->done_ctr = s->pos + 1;
    // persistent decode kernel resources (bf16 only)
    s->pk_ptrs = nullptr; s->pk_part = nullptr; s->pk_bar = nullptr; s->pk_grid = 0; s->pk_ok = false;
    s->pk_step_ts = nullptr;
    s->pk_bar_count = 0; s->pk_tag_gen = g_pk_tag_gen.load(); s->pk_pkt_base = nullptr; s->pk_pkt_bytes = 0;
    if (d.dtype == CAR_BF16) {
        int r2 = pk_state_setup(s);
        if (r2 != CAR_OK) { for (void* p : s->owned) cudaFree(p); delete s; return r2; }
    }
    s->emb_mask_store = s->emb_mask;
    s->emb_mask = nullptr;                           // all-ones until car_state_set_emb_mask
    s->gsp = CarSampling{};
    *out = s;
    return CAR_OK;
}

extern "C" int car_state_set_emb_mask(CarState* s, const int32_t* emb_mask_dev, void* stream) {
    if (!s) CAR_FAIL(CAR_ERR_ARG, "null state");
    if (!emb_mask_dev) { s->emb_mask = nullptr; return CAR_OK; }
    CAR_CUDA(cudaMemcpyAsync(s->emb_mask_store, emb_mask_dev, (size_t)s->b_eff * s->T * 4, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    s->emb_mask = s->emb_mask_store;
    return CAR_OK;
}

extern "C" int car_state_set_step_timer(CarState* s, int64_t* step_ns_dev) {
    if (!s) CAR_FAIL(CAR_ERR_ARG, "null state");
    s->pk_step_ts = (long long*)step_ns_dev;
    return CAR_OK;
}

extern "C" int car_state_destroy(CarState* s) {
    if (!s) return CAR_OK;
    if (s->gexec) cudaGraphExecDestroy(s->gexec);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    for (void* p : s->owned) if (p) cudaFree(p);
    delete s;
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// kernel chains
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int launch_attn_decode(CarState* s, int l, cudaStream_t st) {
    const CarModelDesc& d = s->m->d;
    dim3 grid(s->b_eff * d.n_head, s->nsplit);
    const int flags = tune().l2pf ? 1 : 0;
    if (tune().pdl)
        CAR_LAUNCH_PDL((attn_decode_kernel<T>), grid, dim3(AD_THREADS), 0, st, (const T*)s->q, (const T*)s->kc[l], (const T*)s->vc[l],
                       (const int*)s->emb_mask, s->T, (const int*)s->pos, d.n_head, s->S, s->T, s->nsplit, flags, s->attn_part,
                       s->tickets, (T*)s->attn);
    else
        CAR_LAUNCH((attn_decode_kernel<T>), grid, AD_THREADS, 0, st, (const T*)s->q, (const T*)s->kc[l], (const T*)s->vc[l],
                   (const int*)s->emb_mask, s->T, (const int*)s->pos, d.n_head, s->S, s->T, s->nsplit, flags, s->attn_part,
                   s->tickets, (T*)s->attn);
    return CAR_OK;
}

template <typename T>
static int launch_attn_prefill(CarState* s, int l, cudaStream_t st) {
    const CarModelDesc& d = s->m->d;
    const long long items = (long long)s->b_eff * d.n_head * s->T;
    CAR_LAUNCH((attn_prefill_kernel<T>), (unsigned)((items + 3) / 4), 128, 0, st, (const T*)s->qP, (const T*)s->kc[l],
               (const T*)s->vc[l], (const int*)s->emb_mask, s->T, s->b_eff, d.n_head, s->S, s->T, s->T, (T*)s->attnP);
    return CAR_OK;
}

static EpiParams epi_base(int kind) {
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = kind;
    ep.rpb = 1;
    return ep;
}

// ---------------------------------------------------------------------------------------------------------
// dense (M >= 128 rows) bf16 path of the prefill: tiled tensor-core GEMM on the ORIGINAL [N][K] weights
// ---------------------------------------------------------------------------------------------------------
static int dense_linear(cudaStream_t st, const void* A, int lda, const void* W, int M, int N, int K, int act, const void* resid, int ldr,
                        void* out, int ldo) {
    static DevOnce once;
    if (once.first()) {
        CAR_CUDA(cudaFuncSetAttribute(dense_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
    }
    if (M <= 0 || N <= 0) return CAR_OK;
    static const bool use_tc5 = [] { const char* e = getenv("CAR_TC5"); return e ? atoi(e) != 0 : true; }();
    if (use_tc5 && K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0 && (resid == nullptr || ldr % 8 == 0) &&
        ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && t5_encoder() != nullptr) {
        // tcgen05 path (gemm_tc5.cuh): TMA tensor-map loads, accumulator in TMEM, persistent warp-specialised CTAs
        static DevOnce once5;
        if (once5.first()) CAR_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
        alignas(64) CUtensorMap mapA, mapB;
        if (!t5_make_map(&mapA, A, M, K, lda) || !t5_make_map(&mapB, W, N, K, K)) CAR_FAIL(CAR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        Tc5P q;
        memset(&q, 0, sizeof(q));
        q.M = M; q.N = N; q.K = K;
        q.resid = (const bf16*)resid; q.ldr = ldr; q.C = (bf16*)out; q.ldc = ldo; q.act = act == ACT_GELU_TANH ? 1 : 0;
        static const bool use_x2 = [] { const char* e = getenv("CAR_TC5X2"); return e ? atoi(e) != 0 : true; }();
        // 2-CTA tiles (cta_group::2, 256 x 256 per CTA pair): twice the math per operand byte pulled from L2.  Measured (B200,
        // scripts/bench_gemm.py, TFLOP/s 1-CTA -> 2-CTA): 8192^3 894 -> 1345; 16384 x 1280 x 1280 700 -> 870; 1920 x 3840 x 1280
        // 503 -> 598; 1920 x 3584 x 1280 546 -> 554; but 1920 x 1280 x 3584 (40 pair tiles on 74 pairs) 375 -> 265: with fewer
        // pair tiles than ~1.3 waves the coarser tiling idles SMs, so small grids stay on the 128 x 128 kernel.
        const int ptiles_x2 = ((M + T2_BM - 1) / T2_BM) * ((N + T2_BN - 1) / T2_BN);
        if (use_x2 && M >= T2_BM && N >= T2_BN && ptiles_x2 >= 100) {
            static DevOnce once52;
            if (once52.first()) CAR_CUDA(cudaFuncSetAttribute(gemm_tc5x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
            const int pairs = std::max(1, std::min(ptiles_x2, sm_count() / 2));
            CAR_LAUNCH(gemm_tc5x2_kernel, 2 * pairs, T5_THREADS, T5_SMEM, st, mapA, mapB, q);
            return CAR_OK;
        }
        const int ntiles = ((M + T5_BM - 1) / T5_BM) * ((N + T5_BN - 1) / T5_BN);
        CAR_LAUNCH(gemm_tc5_kernel, std::min(ntiles, sm_count()), T5_THREADS, T5_SMEM, st, mapA, mapB, q);
        return CAR_OK;
    }
    DenseP p;
    memset(&p, 0, sizeof(p));
    p.A = (const bf16*)A; p.B = (const bf16*)W; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = K; p.C = out; p.ldc = ldo; p.alpha = 1.f;
    p.amode = A_PLAIN; p.act = act; p.resid = (const bf16*)resid; p.ldr = ldr; p.out_mode = 0;
    dim3 grid((N + DG_BN - 1) / DG_BN, (M + DG_BM - 1) / DG_BM, 1);
    CAR_LAUNCH(dense_gemm_kernel, grid, DG_THREADS, DG_SMEM, st, p);
    return CAR_OK;
}
static bool use_dense(const CarState* s, int rows) { return s->m->d.dtype == CAR_BF16 && rows >= 64 && s->qkvP != nullptr; }

// one prefill block on the dense path (same rounding points as the skinny chain)
static int enqueue_block_dense(CarState* s, int l, cudaStream_t st) {
    CarModel* m = s->m;
    const CarModelDesc& d = m->d;
    const int dim = d.dim, F = d.ffn_dim, rows = s->b_eff * s->T;
    CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), rows, 256, 0, st, (const bf16*)s->hP, (const bf16*)m->attention_norm[l], (bf16*)s->t1, dim, d.norm_eps);
    CAR_TRY(dense_linear(st, s->t1, dim, m->wqkv[l], rows, 3 * dim, dim, ACT_NONE, nullptr, 0, s->qkvP, 3 * dim));
    CAR_LAUNCH(rope_kv_write_kernel, sm_count() * 8, 256, 0, st, (const bf16*)s->qkvP, s->rope, (bf16*)s->qP, (bf16*)s->kc[l], (bf16*)s->vc[l], rows, s->T, dim,
               d.n_head, s->S);
    static const bool fa = [] { const char* e = getenv("CAR_PREFILL_FA"); return e ? atoi(e) != 0 : true; }();
    if (fa) {   // prefix attention on the tensor cores (attention.cuh): 64 query rows per CTA
        CAR_LAUNCH(attn_prefill_mma_kernel, dim3((s->T + 63) / 64, d.n_head, s->b_eff), 128, 0, st, (const bf16*)s->qP, (const bf16*)s->kc[l],
                   (const bf16*)s->vc[l], (const int*)s->emb_mask, s->T, d.n_head, s->S, s->T, s->T, (bf16*)s->attnP);
    } else {
        CAR_TRY(launch_attn_prefill<bf16>(s, l, st));
    }
    CAR_TRY(dense_linear(st, s->attnP, dim, m->wo[l], rows, dim, dim, ACT_NONE, s->hP, dim, s->hP, dim));
    CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), rows, 256, 0, st, (const bf16*)s->hP, (const bf16*)m->ffn_norm[l], (bf16*)s->t1, dim, d.norm_eps);
    CAR_TRY(dense_linear(st, s->t1, dim, m->w1[l], rows, F, dim, ACT_NONE, nullptr, 0, s->gP, F));
    CAR_TRY(dense_linear(st, s->t1, dim, m->w3[l], rows, F, dim, ACT_NONE, nullptr, 0, s->uP, F));
    CAR_LAUNCH(swiglu_kernel, sm_count() * 8, 256, 0, st, (const bf16*)s->gP, (const bf16*)s->uP, (bf16*)s->actP, (long long)rows * F);
    CAR_TRY(dense_linear(st, s->actP, F, m->w2[l], rows, dim, F, ACT_NONE, s->hP, dim, s->hP, dim));
    return CAR_OK;
}

// one transformer block on `rows` rows; decode (rpb = 1, pos from device scalar) or prefill (rpb = T, pos = t)
static int enqueue_block(CarState* s, int l, bool decode, cudaStream_t st) {
    CarModel* m = s->m;
    const CarModelDesc& d = m->d;
    const int dt = d.dtype, dim = d.dim, F = d.ffn_dim;
    const int rows = decode ? s->b_eff : s->b_eff * s->T;
    void* h = decode ? s->h : s->hP;
    void* q = decode ? s->q : s->qP;
    void* attn = decode ? s->attn : s->attnP;
    void* act = decode ? s->act : s->actP;
    const int rpb = decode ? 1 : s->T;
    const int* posp = decode ? s->pos : nullptr;

    EpiParams e1 = epi_base(EPI_QKV);
    e1.rpb = rpb; e1.pos_ptr = posp; e1.rope = s->rope; e1.kc = s->kc[l]; e1.vc = s->vc[l]; e1.q = q; e1.S = s->S;
    e1.H = d.n_head; e1.d = dim;
    const int skip = decode ? tune().skip : 0;
    if (tune().dbg && decode && l == 3) {
        if (!g_dbg) { cudaMalloc(&g_dbg, 5 * 8 * 8); cudaMemset(g_dbg, 0, 5 * 8 * 8); }
        e1.dbg = g_dbg;
    }
    if (!(skip & 1)) CAR_TRY(launch_skinny(st, dt, h, dim, m->g_wqkv[l], m->attention_norm[l], d.norm_eps, rows, 3 * dim, dim, e1, true));

    if (decode) { if (!(skip & 2)) { if (dt == CAR_BF16) CAR_TRY(launch_attn_decode<bf16>(s, l, st)); else CAR_TRY(launch_attn_decode<float>(s, l, st)); } }
    else { if (dt == CAR_BF16) CAR_TRY(launch_attn_prefill<bf16>(s, l, st)); else CAR_TRY(launch_attn_prefill<float>(s, l, st)); }

    EpiParams e2 = epi_base(EPI_RESID);
    e2.rpb = rpb; e2.pos_ptr = posp; e2.h = h; e2.ldh = dim;
    if (tune().dbg && decode && l == 3) e2.dbg = g_dbg + 8;
    if (!(skip & 4)) CAR_TRY(launch_skinny(st, dt, attn, dim, m->g_wo[l], nullptr, 0.f, rows, dim, dim, e2, false));

    EpiParams e3 = epi_base(EPI_SWIGLU);
    e3.rpb = rpb; e3.out = act; e3.ldo = F;
    if (tune().dbg && decode && l == 3) e3.dbg = g_dbg + 16;
    if (!(skip & 8)) CAR_TRY(launch_skinny(st, dt, h, dim, m->g_w13[l], m->ffn_norm[l], d.norm_eps, rows, 2 * F, dim, e3, true));

    EpiParams e4 = epi_base(EPI_RESID);
    e4.rpb = rpb; e4.pos_ptr = posp; e4.h = h; e4.ldh = dim;
    const int step3 = d.n_layer / 3;
    if (decode && s->has_ctrl && (l + 1) < d.n_layer && (l + 1) % step3 == 0) {
        // control add of the NEXT layer group fused here (gpt_t2i.py:466)
        e4.ctrl = s->ctrl[(l + 1) / step3]; e4.n_img = s->N; e4.T = s->T; e4.cs = s->cs;
    }
    if (tune().dbg && decode && l == 3) e4.dbg = g_dbg + 24;
    if (!(skip & 16)) CAR_TRY(launch_skinny(st, dt, act, F, m->g_w2[l], nullptr, 0.f, rows, dim, F, e4, false));
    return CAR_OK;
}

static int enqueue_head(CarState* s, const void* hrows, int rows, float* logits, cudaStream_t st) {
    CarModel* m = s->m;
    const CarModelDesc& d = m->d;
    EpiParams e = epi_base(EPI_LOGITS);
    e.logits = logits; e.ldl = d.vocab_size;
    return launch_skinny(st, d.dtype, hrows, d.dim, m->g_output, m->norm, d.norm_eps, rows, d.vocab_size, d.dim, e, true);
}

static int enqueue_decode_layers(CarState* s, float* logits, cudaStream_t st) {
    for (int l = 0; l < s->m->d.n_layer; ++l) CAR_TRY(enqueue_block(s, l, true, st));
    return enqueue_head(s, s->h, s->b_eff, logits, st);
}

static int enqueue_mlp(CarState* s, const void* x, int rows, int K, const void* fc1, const void* fc2, const void* fc1_plain,
                       const void* fc2_plain, void* tmp, void* out, cudaStream_t st) {
    const CarModelDesc& d = s->m->d;
    if (use_dense(s, rows) && K % 8 == 0) {      // MLP.forward gpt_t2i.py:165-181: fc2(gelu_tanh(fc1 x)), bias-free
        CAR_TRY(dense_linear(st, x, K, fc1_plain, rows, d.dim, K, ACT_GELU_TANH, nullptr, 0, tmp, d.dim));
        return dense_linear(st, tmp, d.dim, fc2_plain, rows, d.dim, d.dim, ACT_NONE, nullptr, 0, out, d.dim);
    }
    EpiParams a = epi_base(EPI_STORE);
    a.out = tmp; a.ldo = d.dim; a.act = 1;
    CAR_TRY(launch_skinny(st, d.dtype, x, K, fc1, nullptr, 0.f, rows, d.dim, K, a, false));
    EpiParams b = epi_base(EPI_STORE);
    b.out = out; b.ldo = d.dim; b.act = 0;
    return launch_skinny(st, d.dtype, tmp, d.dim, fc2, nullptr, 0.f, rows, d.dim, d.dim, b, false);
}

template <typename T>
static int prefill_small_kernels(CarState* s, int l, cudaStream_t st) {
    const CarModelDesc& d = s->m->d;
    const int step3 = d.n_layer / 3;
    if (s->has_ctrl && l % step3 == 0)
        CAR_LAUNCH((prefill_ctrl_add_kernel<T>), s->b_eff, 256, 0, st, (T*)s->hP, (const T*)s->ctrl[l / step3], s->T, s->N, d.dim, s->cs);
    return CAR_OK;
}

extern "C" int car_prefill(CarState* s, const void* cond, const void* condition, float control_strength, float* logits_out,
                           int32_t all_rows, void* stream) {
    if (!s || !cond) CAR_FAIL(CAR_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CarModel* m = s->m;
    const CarModelDesc& d = m->d;
    const int rows = s->b_eff * s->T;
    s->cs = control_strength;
    s->has_ctrl = condition != nullptr;
    s->graph_ok = false;
    // 1. prefix embeddings: CaptionEmbedder MLP (gpt_t2i.py:156-162) or LabelEmbedder gather (:89-97)
    if (d.model_type == 1) CAR_TRY(enqueue_mlp(s, cond, rows, d.caption_dim, m->g_cap_fc1, m->g_cap_fc2, m->cap_fc1, m->cap_fc2, s->t1, s->hP, st));
    else {
        if (d.dtype == CAR_BF16) CAR_LAUNCH((gather_rows_kernel<bf16>), rows, 256, 0, st, (const bf16*)m->label_table, (const int*)cond, (bf16*)s->hP, d.dim, (const bf16*)nullptr, 0, 0, 0.f);
        else CAR_LAUNCH((gather_rows_kernel<float>), rows, 256, 0, st, (const float*)m->label_table, (const int*)cond, (float*)s->hP, d.dim, (const float*)nullptr, 0, 0, 0.f);
    }
    // 2. control tokens: condition_mlp then the three condition_layers MLPs (gpt_t2i.py:438-442)
    if (condition) {
        const int crow = s->b_eff * s->N;
        CAR_TRY(enqueue_mlp(s, condition, crow, d.dim, m->g_cond_fc1, m->g_cond_fc2, m->cond_fc1, m->cond_fc2, s->t1, s->t2, st));
        for (int j = 0; j < 3; ++j)
            CAR_TRY(enqueue_mlp(s, s->t2, crow, d.dim, m->g_ctl_fc1[j], m->g_ctl_fc2[j], m->ctl_fc1[j], m->ctl_fc2[j], s->t1, s->ctrl[j], st));
    }
    // 3. blocks
    for (int l = 0; l < d.n_layer; ++l) {
        if (d.dtype == CAR_BF16) CAR_TRY(prefill_small_kernels<bf16>(s, l, st)); else CAR_TRY(prefill_small_kernels<float>(s, l, st));
        if (use_dense(s, rows)) CAR_TRY(enqueue_block_dense(s, l, st));
        else CAR_TRY(enqueue_block(s, l, false, st));
    }
    // 4. head: last prefix row always (feeds car_generate); all rows on request (forward() parity)
    if (d.dtype == CAR_BF16) CAR_LAUNCH((take_last_row_kernel<bf16>), s->b_eff, 256, 0, st, (const bf16*)s->hP, (bf16*)s->h, s->T, d.dim);
    else CAR_LAUNCH((take_last_row_kernel<float>), s->b_eff, 256, 0, st, (const float*)s->hP, (float*)s->h, s->T, d.dim);
    CAR_TRY(enqueue_head(s, s->h, s->b_eff, s->logits, st));
    if (logits_out) {
        if (all_rows) CAR_TRY(enqueue_head(s, s->hP, rows, logits_out, st));
        else CAR_CUDA(cudaMemcpyAsync(logits_out, s->logits, (size_t)s->b_eff * d.vocab_size * 4, cudaMemcpyDeviceToDevice, st));
    }
    CAR_LAUNCH(set_int_kernel, 1, 1, 0, st, s->pos, s->T - 1);
    s->prefilled = true;
    return CAR_OK;
}

extern "C" int car_decode_step(CarState* s, const int32_t* tok, int32_t pos, float* logits_out, void* stream) {
    if (!s || !tok || !logits_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (pos < s->T || pos >= s->S) CAR_FAIL(CAR_ERR_ARG, "pos out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const CarModelDesc& d = s->m->d;
    CAR_LAUNCH(set_int_kernel, 1, 1, 0, st, s->pos, pos);
    const int p = pos - s->T + 1;
    if (d.dtype == CAR_BF16)
        CAR_LAUNCH((gather_rows_kernel<bf16>), s->b_eff, 256, 0, st, (const bf16*)s->m->tok_emb, (const int*)tok, (bf16*)s->h, d.dim,
                   (const bf16*)(s->has_ctrl ? s->ctrl[0] : nullptr), s->N, p, s->cs);
    else
        CAR_LAUNCH((gather_rows_kernel<float>), s->b_eff, 256, 0, st, (const float*)s->m->tok_emb, (const int*)tok, (float*)s->h, d.dim,
                   (const float*)(s->has_ctrl ? s->ctrl[0] : nullptr), s->N, p, s->cs);
    return enqueue_decode_layers(s, logits_out, st);
}

// ---------------------------------------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------------------------------------
static int fill_sample_args(SampleArgs& a, const CarSampling* sp, int b_eff, int V) {
    memset(&a, 0, sizeof(a));
    a.V = V;
    a.use_cfg = sp->cfg_scale > 1.0f ? 1 : 0;
    if (a.use_cfg && (b_eff % 2)) CAR_FAIL(CAR_ERR_ARG, "cfg_scale > 1 needs an even number of rows");
    a.B = a.use_cfg ? b_eff / 2 : b_eff;
    a.cfg_on = 1; a.cfg_scale = sp->cfg_scale; a.cfg_interval = sp->cfg_interval;
    a.inv_temp = 1.0f / fmaxf(sp->temperature, 1e-5f);
    a.top_k = sp->top_k; a.top_p = sp->top_p; a.sample_logits = sp->sample_logits;
    a.seed_lo = (uint32_t)(sp->seed & 0xffffffffu); a.seed_hi = (uint32_t)(sp->seed >> 32);
    return CAR_OK;
}

static int launch_sampler(const SampleArgs& a, cudaStream_t st) {
    if (a.V % 4 != 0 || a.V < 4) CAR_FAIL(CAR_ERR_UNSUPPORTED, "the fused sampler needs a vocabulary size that is a multiple of 4");
    CAR_LAUNCH(sample_kernel, a.B, SMP_THREADS, 0, st, a);
    return CAR_OK;
}

extern "C" int car_sample(const float* logits, int32_t b_eff, int32_t V, const CarSampling* sp, int32_t cfg_on, int32_t step,
                          const float* noise, int32_t* idx_out, float* probs_out, void* stream) {
    if (!logits || !sp || !idx_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    SampleArgs a;
    CAR_TRY(fill_sample_args(a, sp, b_eff, V));
    a.logits = logits; a.cfg_on = cfg_on; a.cfg_interval = -1; a.step = step; a.noise = noise;
    a.idx_out = idx_out; a.tokens_ld = 0; a.probs_out = probs_out;
    return launch_sampler(a, (cudaStream_t)stream);
}

static bool same_sampling(const CarSampling& x, const CarSampling& y) {
    return x.temperature == y.temperature && x.top_k == y.top_k && x.top_p == y.top_p && x.sample_logits == y.sample_logits &&
           x.cfg_scale == y.cfg_scale && x.cfg_interval == y.cfg_interval && x.seed == y.seed;
}

static int loop_sample_args(CarState* s, const CarSampling* sp, const float* noise, SampleArgs& a) {
    const CarModelDesc& d = s->m->d;
    CAR_TRY(fill_sample_args(a, sp, s->b_eff, d.vocab_size));
    a.logits = s->logits; a.noise = noise; a.noise_per_step = noise ? 1 : 0;
    a.idx_out = s->tokens; a.tokens_ld = s->N; a.probs_out = nullptr;
    a.h_out = s->h; a.tok_emb = s->m->tok_emb; a.ctrl0 = s->has_ctrl ? s->ctrl[0] : nullptr; a.d = d.dim; a.n_img = s->N;
    a.T = s->T; a.cs = s->cs; a.dtype = d.dtype; a.tok_buf = s->tok; a.pos_ptr = s->pos; a.done_ctr = s->done_ctr;
    return CAR_OK;
}

// the whole decode loop as one persistent cooperative kernel (decode_persistent.cuh)
static int launch_pk(CarState* s, const SampleArgs& a, int n_tokens, cudaStream_t st, const int32_t* forced = nullptr, float* trace = nullptr) {
    const CarModelDesc& d = s->m->d;
    const int L = d.n_layer;
    if (s->pk_ptrs_gen != s->m->pack_gen) {   // car_model_repack since the table was uploaded: the borrowed norm-weight pointers may have moved
        const std::vector<const void*> hp = pk_pointer_table(s);
        CAR_CUDA(cudaMemcpyAsync(s->pk_ptrs, hp.data(), hp.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
        CAR_CUDA(cudaStreamSynchronize(st));                      // hp is a host temporary
        s->pk_ptrs_gen = s->m->pack_gen;
    }
    unsigned int tag_gen = 0;
    const unsigned int tag_base = pk_alloc_tags((unsigned int)n_tokens * (unsigned int)(L + 1) + 8u, &tag_gen);
    if (tag_gen != s->pk_tag_gen) {       // the process-wide tag counter wrapped since this state's packets were last zeroed
        CAR_TRY(fill_u32(s->pk_pkt_base, 0u, s->pk_pkt_bytes, st));
        s->pk_tag_gen = tag_gen;
    }
    PkParams P;
    memset(&P, 0, sizeof(P));
    P.dim = d.dim; P.F = d.ffn_dim; P.V = d.vocab_size; P.L = L; P.H = d.n_head; P.T = s->T; P.S = s->S; P.n_img = s->N;
    P.b_eff = s->b_eff; P.B = a.B; P.eps = d.norm_eps; P.cs = s->cs;
    P.tok_emb = (const bf16*)s->m->tok_emb; P.norm_w = (const bf16*)s->m->norm; P.w_out = (const uint4*)s->m->g_output;
    void** pp = s->pk_ptrs;
    P.wqkv = (const uint4* const*)(pp + 0 * L); P.wo = (const uint4* const*)(pp + 1 * L); P.w13 = (const uint4* const*)(pp + 2 * L);
    P.w2 = (const uint4* const*)(pp + 3 * L); P.attn_norm = (const bf16* const*)(pp + 4 * L); P.ffn_norm = (const bf16* const*)(pp + 5 * L);
    P.kc = (bf16* const*)(pp + 6 * L); P.vc = (bf16* const*)(pp + 7 * L);
    for (int j = 0; j < 3; ++j) P.ctrl[j] = (const bf16*)s->ctrl[j];
    P.has_ctrl = s->has_ctrl ? 1 : 0;
    P.rope = s->rope; P.emb_mask = s->emb_mask; P.logits = s->logits; P.part = s->pk_part;
    for (int par = 0; par < 2; ++par) {
        P.h2[par] = s->pk_h2[par]; P.h1[par] = s->pk_h1[par]; P.att[par] = s->pk_att[par]; P.act[par] = s->pk_act[par];
        P.qkv[par] = s->pk_qkv[par]; P.partial[par] = s->pk_partial[par];
    }
    P.part_slots = s->pk_part_slots; P.tag_base = tag_base; P.bar = s->pk_bar; P.bar_base = s->pk_bar_count;
    P.smp = a; P.n_steps = n_tokens;
    P.forced = forced; P.forced_ld = n_tokens; P.trace = trace; P.step_ts = s->pk_step_ts;
    if (trace) CAR_CUDA(cudaMemcpyAsync(trace, s->logits, (size_t)s->b_eff * d.vocab_size * 4, cudaMemcpyDeviceToDevice, st));
    { const char* e = getenv("CAR_EXP"); P.exp_flags = e ? atoi(e) : 0; }
    static DevOnce once;
    if (once.first()) {
        CAR_CUDA(cudaFuncSetAttribute(pk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_TOTAL));
    }
    int occ = 0;
    CAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pk_decode_kernel, PK_THREADS, PK_SMEM_TOTAL));
    if (occ < 1) CAR_FAIL(CAR_ERR_UNSUPPORTED, "persistent decode kernel does not fit on an SM");
    static long long* mdbg = nullptr;
    const size_t dbg_n = (size_t)s->pk_grid * 64 + 5 * 256;
    if (tune().dbg) {
        if (!mdbg) { cudaMalloc(&mdbg, dbg_n * 8); }
        cudaMemsetAsync(mdbg, 0, dbg_n * 8, st);
        P.dbg = mdbg; P.dbg_step = std::max(0, std::min(n_tokens - 2, tune().dbg));
    }
    void* args[] = {&P};
    CAR_CUDA(cudaLaunchCooperativeKernel((const void*)pk_decode_kernel, dim3(s->pk_grid), dim3(PK_THREADS), args, PK_SMEM_TOTAL, st));
    s->pk_bar_count += (unsigned int)(n_tokens - 1) * (unsigned int)s->pk_grid;        // one grid barrier per decoded token
    g_car_launches.fetch_add(1, std::memory_order_relaxed);
    if (tune().dbg) {
        cudaStreamSynchronize(st);
        std::vector<long long> t(dbg_n);
        cudaMemcpy(t.data(), mdbg, dbg_n * 8, cudaMemcpyDeviceToHost);
        const int G = s->pk_grid;
        long long t0 = t[0];
        for (int c = 0; c < G; ++c) if (t[(size_t)c * 64]) t0 = std::min(t0, t[(size_t)c * 64]);
        auto stat = [&](int slot, const char* name) {
            std::vector<long long> v;
            for (int c = 0; c < G; ++c) if (t[(size_t)c * 64 + slot]) v.push_back(t[(size_t)c * 64 + slot] - t0);
            if (v.empty()) return;
            std::sort(v.begin(), v.end());
            int amax = 0;
            for (int c = 0; c < G; ++c) if (t[(size_t)c * 64 + slot] - t0 == v.back()) amax = c;
            fprintf(stderr, "[pk] %-22s n=%3zu  min %8.2f  med %8.2f  max %8.2f us (CTA %d)\n", name, v.size(), v.front() * 1e-3, v[v.size() / 2] * 1e-3,
                    v.back() * 1e-3, amax);
        };
        if (!t[0] && !t[64]) fprintf(stderr, "[pk] no stamps: rebuild with CAR_PK_TRACE=1 (python -m controlar_b200.build --force)\n");
        fprintf(stderr, "[pk] step %d, times relative to the first CTA entering the sampler; layer 3 phases\n", P.dbg_step);
        stat(0, "step start"); stat(1, "sampler done");
        if (t[48]) fprintf(stderr, "[pk] sampler top-k of CTA 0: histogram built %.2f | boundary bin found %.2f | candidates gathered %.2f\n",
                           (t[53] - t[48]) * 1e-3, (t[54] - t[48]) * 1e-3, (t[55] - t[48]) * 1e-3);
        if (t[48]) fprintf(stderr, "[pk] sampler of CTA 0 (us after its start): loads issued %.2f | row in registers + CFG %.2f | top-k done %.2f | soft-max done %.2f | race done %.2f | CTA done %.2f\n",
                           0.0, (t[49] - t[48]) * 1e-3, (t[50] - t[48]) * 1e-3, (t[51] - t[48]) * 1e-3, (t[52] - t[48]) * 1e-3, (t[1] - t[48]) * 1e-3);
        const char* nm[5] = {"qkv", "attn", "wo", "w13", "w2"};
        for (int k = 0; k < 5; ++k) {
            char buf[64];
            const char* sub[5] = {"start", k == 1 ? "q polled" : "A polled", k == 1 ? "keys done" : "weights+norm", k == 1 ? "end" : "mma done", "end"};
            for (int j = 0; j < (k == 1 ? 4 : 5); ++j) { snprintf(buf, sizeof buf, "L3 %s %s", nm[k], sub[j]); stat(8 + 8 * k + j, buf); }
        }
        {
            const char* sub[5] = {"head start", "head A polled", "head weights+norm", "head mma done (batch 0)", "head end (batch 0)"};
            for (int j = 0; j < 5; ++j) stat(56 + j, sub[j]);
        }
        stat(3, "head done"); stat(4, "barrier passed");
        {   // per-warp stamps of one CTA (77): min / max over the 16 warps, relative to the phase's first stamp
            const char* wn[11] = {"start", "prepoll", "A loaded", "ssq out", "sync1", "normed", "mma", "red out", "sync2", "reduced", "end"};
            const char* pn[5] = {"qkv", "attn", "wo", "w13", "w2"};
            for (int ph = 0; ph < 5; ++ph) {
                if (ph == 1) continue;
                const long long* w = t.data() + (size_t)G * 64 + (size_t)ph * 256;
                long long base = 0;
                for (int k = 0; k < 16; ++k) if (w[k * 16] && (!base || w[k * 16] < base)) base = w[k * 16];
                if (!base) continue;
                fprintf(stderr, "[pk warp] %-3s", pn[ph]);
                for (int sI = 0; sI < 11; ++sI) {
                    long long mn = 0, mx = 0;
                    for (int k = 0; k < 16; ++k) { const long long v = w[k * 16 + sI]; if (!v) continue; if (!mn || v < mn) mn = v; if (v > mx) mx = v; }
                    if (mn) fprintf(stderr, " | %s %.2f-%.2f", wn[sI], (mn - base) * 1e-3, (mx - base) * 1e-3);
                }
                fprintf(stderr, "\n");
            }
        }
    }
    return CAR_OK;
}

extern "C" int car_generate(CarState* s, const CarSampling* sp, int32_t n_tokens, const float* noise, int32_t* tokens_out,
                            void* stream) {
    if (!s || !sp || !tokens_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (!s->prefilled) CAR_FAIL(CAR_ERR_STATE, "car_generate must follow car_prefill on the same state");
    if (n_tokens < 1 || n_tokens > s->N) CAR_FAIL(CAR_ERR_ARG, "n_tokens must be in [1, N]");
    cudaStream_t st = (cudaStream_t)stream;
    SampleArgs a;
    CAR_TRY(loop_sample_args(s, sp, noise, a));
    if (s->m->d.dtype == CAR_BF16 && tune().mega && s->pk_ok) {
        CAR_TRY(launch_pk(s, a, n_tokens, st));
        CAR_CUDA(cudaMemcpy2DAsync(tokens_out, (size_t)n_tokens * 4, s->tokens, (size_t)s->N * 4, (size_t)n_tokens * 4, a.B,
                                   cudaMemcpyDeviceToDevice, st));
        CAR_LAUNCH(set_int_kernel, 1, 1, 0, st, s->pos, s->T - 1 + n_tokens);
        s->prefilled = false;
        return CAR_OK;
    }
    // token 0 from the prefill logits (generate.py:198); its fused tail writes h for position T and bumps pos
    CAR_TRY(launch_sampler(a, st));
    if (n_tokens > 1) {
        if (!s->graph_ok || !same_sampling(s->gsp, *sp) || s->gnoise != noise || s->graph_pack_gen != s->m->pack_gen) {
            if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
            cudaGraph_t g = nullptr;
            const long long launched_before = g_car_launches.load();   // captured nodes are not launches yet
            CAR_CUDA(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeRelaxed));
            int r = enqueue_decode_layers(s, s->logits, s->cap_stream);
            if (r == CAR_OK) r = launch_sampler(a, s->cap_stream);
            cudaError_t ce = cudaStreamEndCapture(s->cap_stream, &g);
            g_car_launches.store(launched_before);
            if (r != CAR_OK) { if (g) cudaGraphDestroy(g); return r; }
            if (ce != cudaSuccess) CAR_FAIL(CAR_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&s->gexec, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) CAR_FAIL(CAR_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
            s->graph_ok = true; s->gsp = *sp; s->gnoise = noise; s->graph_pack_gen = s->m->pack_gen;
        }
        const int per_step = s->m->d.n_layer * 5 + 2;
        for (int i = 1; i < n_tokens; ++i) CAR_CUDA(cudaGraphLaunch(s->gexec, st));
        g_car_launches.fetch_add((long long)per_step * (n_tokens - 1), std::memory_order_relaxed);
    }
    const int B = a.B;
    CAR_CUDA(cudaMemcpy2DAsync(tokens_out, (size_t)n_tokens * 4, s->tokens, (size_t)s->N * 4, (size_t)n_tokens * 4, B,
                               cudaMemcpyDeviceToDevice, st));
    s->prefilled = false;
    if (tune().dbg && g_dbg) {
        cudaStreamSynchronize(st);
        long long hbuf[40];
        cudaMemcpy(hbuf, g_dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const char* names[4] = {"qkv", "wo", "w13", "w2"};
        for (int k = 0; k < 4; ++k) {
            long long* t = hbuf + 8 * k;
            fprintf(stderr, "[dbg] %-4s wait %6lld | stage %6lld | main %6lld | reduce %6lld | epi %6lld | total %6lld cycles\n", names[k],
                    t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[5] - t[0]);
        }
    }
    return CAR_OK;
}

// teacher-forced run of the SAME device-side loop (parity tests): the token fed to step i + 1 is forced[b][i]; the sampler still
// runs and tokens_out holds what it would have chosen at every step given the forced prefix; logits_trace (optional) receives the
// raw model logits of every step.
extern "C" int car_generate_forced(CarState* s, const CarSampling* sp, int32_t n_tokens, const float* noise, const int32_t* forced_tokens,
                                   float* logits_trace, int32_t* tokens_out, void* stream) {
    if (!s || !sp || !tokens_out || !forced_tokens) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (!s->prefilled) CAR_FAIL(CAR_ERR_STATE, "car_generate_forced must follow car_prefill on the same state");
    if (n_tokens < 1 || n_tokens > s->N) CAR_FAIL(CAR_ERR_ARG, "n_tokens must be in [1, N]");
    if (!(s->m->d.dtype == CAR_BF16 && s->pk_ok))
        CAR_FAIL(CAR_ERR_UNSUPPORTED, "teacher forcing through the device-side loop needs the persistent decode kernel (bf16); use car_decode_step");
    cudaStream_t st = (cudaStream_t)stream;
    SampleArgs a;
    CAR_TRY(loop_sample_args(s, sp, noise, a));
    CAR_TRY(launch_pk(s, a, n_tokens, st, forced_tokens, logits_trace));
    CAR_CUDA(cudaMemcpy2DAsync(tokens_out, (size_t)n_tokens * 4, s->tokens, (size_t)s->N * 4, (size_t)n_tokens * 4, a.B,
                               cudaMemcpyDeviceToDevice, st));
    CAR_LAUNCH(set_int_kernel, 1, 1, 0, st, s->pos, s->T - 1 + n_tokens);
    s->prefilled = false;
    return CAR_OK;
}

extern "C" int64_t car_decode_step_bytes(const CarState* s, int32_t n_context) {
    if (!s) return -1;
    const CarModelDesc& d = s->m->d;
    const int64_t es = d.dtype == CAR_BF16 ? 2 : 4;
    const int64_t P = (int64_t)d.n_layer * (4LL * d.dim * d.dim + 3LL * d.dim * d.ffn_dim) + (int64_t)d.vocab_size * d.dim;
    const int64_t kappa = 2LL * d.n_layer * d.dim * es;
    return es * P + (int64_t)s->b_eff * kappa * n_context + (int64_t)s->b_eff * kappa + 4LL * s->b_eff * d.vocab_size;
}

// ---------------------------------------------------------------------------------------------------------
// building-block ops for unit tests
// ---------------------------------------------------------------------------------------------------------
extern "C" int car_op_linear(int32_t dtype, const void* x, const void* w, const void* bias, void* y, int32_t M, int32_t N,
                             int32_t K, int32_t act, void* stream) {
    if (!x || !w || !y) CAR_FAIL(CAR_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    EpiParams e = epi_base(EPI_STORE);
    e.out = y; e.ldo = N; e.act = act; e.bias = bias;
    if (dtype == CAR_F32) return launch_skinny(st, dtype, x, K, w, nullptr, 0.f, M, N, K, e, false);
    void* packed = nullptr;
    CAR_CUDA(cudaMallocAsync(&packed, (size_t)N * K * 2, st));
    const int nblk = N / 8;
    const long long total = (long long)nblk * (K / 32) * 32;
    CAR_LAUNCH(pack_weight_bf16_kernel, (int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st, (const bf16*)w,
               (const bf16*)nullptr, (uint4*)packed, nblk, K, 0);
    int r = launch_skinny(st, dtype, x, K, packed, nullptr, 0.f, M, N, K, e, false);
    cudaFreeAsync(packed, st);
    return r;
}

// the dense (M >= 64 rows) tensor-core linear of the prefill / MLP path, exposed for unit tests and micro-benchmarks:
// y[M,N] = act(x[M,K] · w[N,K]^T) (+ resid), bf16, fp32 accumulate (gemm_tc5.cuh; CAR_TC5=0 selects the mma.sync kernel)
extern "C" int car_op_dense_linear(const void* x, const void* w, const void* resid, void* y, int32_t M, int32_t N, int32_t K, int32_t act,
                                   void* stream) {
    if (!x || !w || !y) CAR_FAIL(CAR_ERR_ARG, "null argument");
    return dense_linear((cudaStream_t)stream, x, K, w, M, N, K, act ? ACT_GELU_TANH : ACT_NONE, resid, N, y, N);
}

extern "C" int car_op_rmsnorm(int32_t dtype, const void* x, const void* w, void* y, int32_t M, int32_t K, float eps, void* stream) {
    if (!x || !w || !y) CAR_FAIL(CAR_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CAR_BF16) CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), M, 256, 0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, K, eps);
    else CAR_LAUNCH((rmsnorm_rows_kernel<float>), M, 256, 0, st, (const float*)x, (const float*)w, (float*)y, K, eps);
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// training forward: Transformer.forward(idx, cond_idx, targets, mask, valid, condition) in train mode,
// fp32 parameters under bf16 autocast — gpt_t2i.py:420-431,451-484.  First correct path: prefill GEMM kernels + train.cuh glue.
// ---------------------------------------------------------------------------------------------------------
struct CarTrain {
    CarModelDesc d;
    CarTrainWeights w;
    std::vector<const void*> attention_norm, wqkv, wo, ffn_norm, w1, w3, w2;   // borrowed fp32
    std::vector<bf16*> b_wqkv, b_wo, b_w1, b_w3, b_w2;                          // owned bf16 casts, refreshed every forward
    bf16 *b_out, *b_cap1, *b_cap2, *b_cond1, *b_cond2, *b_ctl1[3], *b_ctl2[3], *b_ad1, *b_ad2;
    int maxB, maxN, maxS;
    const float* rope;
    float *h, *nll;
    bf16 *x, *qkv, *q, *kc, *vc, *att, *g, *u, *act, *o, *cin, *ctmp, *ctok, *cadd, *lg;
    // ---- backward (car_train_backward): the stream saved at every block input, gradient / transpose workspaces ----
    float *hs, *dh, *h0, *scr, *part, *lse, *dsum;
    bf16 *capx, *x2, *db, *dact, *dg, *du, *dx, *datt, *dq, *dk, *dv, *dqkv, *dlg, *wT, *yT, *xT, *dWb;
    bf16 *m_t, *m_a, *m_da, *m_dt, *dctok, *dcin, *dadd;
    size_t Rp_max;
    // arguments of the last car_train_forward (borrowed until car_train_backward returns)
    int fB = 0, fN = 0;
    const int32_t *f_idx = nullptr, *f_targets = nullptr;
    const void *f_cond = nullptr, *f_feat = nullptr;
    const uint8_t *f_drop = nullptr, *f_mask = nullptr;
    const float* f_valid = nullptr;
    bool fwd_ok = false;
    std::vector<void*> owned;
};

static int tr_cast(cudaStream_t st, const void* src, bf16* dst, long long n) {
    CAR_LAUNCH(tr_cast_bf16_kernel, (int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st, (const float*)src, dst, n);
    return CAR_OK;
}
static int tr_grid(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }
// MLP.forward gpt_t2i.py:177-181 on bf16 operands: out = fc2(gelu_tanh(fc1 x))
static int tr_mlp(cudaStream_t st, const bf16* x, int rows, int K, const bf16* fc1, const bf16* fc2, int d, bf16* tmp, bf16* out) {
    CAR_TRY(dense_linear(st, x, K, fc1, rows, d, K, ACT_GELU_TANH, nullptr, 0, tmp, d));
    return dense_linear(st, tmp, d, fc2, rows, d, d, ACT_NONE, nullptr, 0, out, d);
}

extern "C" int car_train_create(const CarModelDesc* desc, const CarTrainWeights* w, int32_t max_batch, int32_t max_img_tokens,
                                const float* rope_table, void* stream, CarTrain** out) {
    if (!desc || !w || !out || !rope_table) CAR_FAIL(CAR_ERR_ARG, "null argument");
    const CarModelDesc& d = *desc;
    if (d.dtype != CAR_F32) CAR_FAIL(CAR_ERR_UNSUPPORTED, "training forward takes the fp32 master weights (bf16 autocast is applied inside)");
    if (d.dim % 64 != 0 || d.dim / d.n_head != 64 || d.n_layer % 3 != 0 || d.ffn_dim % 8 != 0 || d.vocab_size % 8 != 0 || w->adapter_dim % 8 != 0 ||
        (d.model_type == 1 && d.caption_dim % 8 != 0))
        CAR_FAIL(CAR_ERR_UNSUPPORTED, "shape not supported (head_dim 64, dims multiple of 8, n_layer multiple of 3)");
    if (max_batch <= 0 || max_img_tokens <= 0) CAR_FAIL(CAR_ERR_ARG, "bad capacity");
    (void)stream;
    CarTrain* t = new CarTrain();
    t->d = d; t->w = *w; t->rope = rope_table;
    t->maxB = max_batch; t->maxN = max_img_tokens; t->maxS = d.cls_token_num + max_img_tokens - 1;
    const int L = d.n_layer, dim = d.dim, F = d.ffn_dim, V = d.vocab_size;
    auto copyp = [&](std::vector<const void*>& v, const void* const* src) { v.assign(src, src + L); };
    copyp(t->attention_norm, w->w.attention_norm); copyp(t->wqkv, w->w.wqkv); copyp(t->wo, w->w.wo); copyp(t->ffn_norm, w->w.ffn_norm);
    copyp(t->w1, w->w.w1); copyp(t->w3, w->w.w3); copyp(t->w2, w->w.w2);
    int rc = CAR_OK;
    auto A = [&](bf16** p, size_t elems) { if (rc == CAR_OK) rc = alloc_dev(t->owned, (void**)p, elems * 2); };
    t->b_wqkv.resize(L); t->b_wo.resize(L); t->b_w1.resize(L); t->b_w3.resize(L); t->b_w2.resize(L);
    for (int l = 0; l < L; ++l) {
        A(&t->b_wqkv[l], (size_t)3 * dim * dim); A(&t->b_wo[l], (size_t)dim * dim);
        A(&t->b_w1[l], (size_t)F * dim); A(&t->b_w3[l], (size_t)F * dim); A(&t->b_w2[l], (size_t)dim * F);
    }
    A(&t->b_out, (size_t)V * dim);
    t->b_cap1 = t->b_cap2 = nullptr;
    if (d.model_type == 1) { A(&t->b_cap1, (size_t)dim * d.caption_dim); A(&t->b_cap2, (size_t)dim * dim); }
    A(&t->b_cond1, (size_t)dim * dim); A(&t->b_cond2, (size_t)dim * dim);
    for (int j = 0; j < 3; ++j) { A(&t->b_ctl1[j], (size_t)dim * dim); A(&t->b_ctl2[j], (size_t)dim * dim); }
    A(&t->b_ad1, (size_t)dim * w->adapter_dim); A(&t->b_ad2, (size_t)dim * dim);
    const size_t R = (size_t)t->maxB * t->maxS, RC = (size_t)t->maxB * t->maxN;
    if (rc == CAR_OK) rc = alloc_dev(t->owned, (void**)&t->h, R * dim * 4);
    if (rc == CAR_OK) rc = alloc_dev(t->owned, (void**)&t->nll, RC * 4);
    A(&t->x, std::max(R * dim, (size_t)t->maxB * d.cls_token_num * std::max(d.caption_dim, dim)));
    A(&t->qkv, R * 3 * dim); A(&t->q, R * dim); A(&t->kc, R * dim); A(&t->vc, R * dim); A(&t->att, R * dim);
    A(&t->g, R * F); A(&t->u, R * F); A(&t->act, R * F); A(&t->o, R * dim);
    A(&t->cin, RC * dim); A(&t->ctmp, std::max(RC, (size_t)t->maxB * d.cls_token_num) * dim); A(&t->ctok, RC * dim); A(&t->cadd, RC * dim);
    A(&t->lg, RC * V);
    // backward workspaces
    {
        auto AF = [&](float** p, size_t elems) { if (rc == CAR_OK) rc = alloc_dev(t->owned, (void**)p, elems * 4); };
        const size_t cap = (size_t)(d.model_type == 1 ? d.caption_dim : 8), ad = (size_t)w->adapter_dim;
        const size_t rows_mlp = std::max(RC, (size_t)t->maxB * d.cls_token_num);
        const size_t Rp = (std::max(R, rows_mlp) + 63) / 64 * 64;
        const size_t maxN = std::max({(size_t)3 * dim, (size_t)F, (size_t)V}), maxK = std::max({(size_t)F, (size_t)dim, cap, ad});
        const size_t maxW = std::max({(size_t)3 * dim * dim, (size_t)F * dim, (size_t)V * dim, (size_t)dim * cap, (size_t)dim * ad});
        t->Rp_max = Rp;
        AF(&t->hs, (size_t)L * R * dim); AF(&t->dh, R * dim); AF(&t->h0, R * dim); AF(&t->scr, R * dim);
        AF(&t->part, (size_t)TR_COLSUM_CHUNKS * dim); AF(&t->lse, (size_t)t->maxB * d.n_head * t->maxS); AF(&t->dsum, (size_t)t->maxB * d.n_head * t->maxS);
        A(&t->capx, (size_t)t->maxB * d.cls_token_num * cap);
        A(&t->x2, R * dim); A(&t->db, R * dim); A(&t->dact, R * F); A(&t->dg, R * F); A(&t->du, R * F); A(&t->dx, R * dim);
        A(&t->datt, R * dim); A(&t->dq, R * dim); A(&t->dk, R * dim); A(&t->dv, R * dim); A(&t->dqkv, R * 3 * dim); A(&t->dlg, RC * V);
        A(&t->wT, maxW); A(&t->dWb, maxW); A(&t->yT, maxN * Rp); A(&t->xT, maxK * Rp);
        A(&t->m_t, rows_mlp * dim); A(&t->m_a, rows_mlp * dim); A(&t->m_da, rows_mlp * dim); A(&t->m_dt, rows_mlp * dim);
        A(&t->dctok, RC * dim); A(&t->dcin, RC * dim); A(&t->dadd, rows_mlp * dim);
    }
    if (rc != CAR_OK) { for (void* p : t->owned) cudaFree(p); delete t; return rc; }
    *out = t;
    return CAR_OK;
}

extern "C" int car_train_destroy(CarTrain* t) {
    if (!t) return CAR_OK;
    for (void* p : t->owned) cudaFree(p);
    delete t;
    return CAR_OK;
}

// dynamic shared memory of the plain attention kernels (TRA_WARPS warps x floats_per_warp); above 48 KB the opt-in attribute is set on
// every call (cheap, and correct on every device of the process — no process-wide "already set" flag)
static int tr_attn_smem(size_t floats_per_warp, const void* fn, size_t* bytes) {
    *bytes = (size_t)TRA_WARPS * floats_per_warp * 4;
    if (*bytes > 200 * 1024) CAR_FAIL(CAR_ERR_UNSUPPORTED, "sequence too long for the plain attention kernels");
    if (*bytes > 48 * 1024) CAR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return CAR_OK;
}

// One TransformerBlock (gpt_t2i.py:303-307) on the fp32 stream t->h, preceded by the control add of gpt_t2i.py:458-460 when the
// block opens a third of the stack.  for_bwd: the recompute of car_train_backward — keeps the block input (after the control
// add) in t->h0, the attention-side norm output in t->x, the feed-forward-side one in t->x2, and stops before w2 (t->h then
// holds the stream between the two halves).
static int tr_block_fwd(CarTrain* t, cudaStream_t st, int l, int B, int n_img, const uint8_t* mask, bool has_feat, bool for_bwd) {
    const CarModelDesc& d = t->d;
    const int L = d.n_layer, dim = d.dim, F = d.ffn_dim, T = d.cls_token_num, H = d.n_head;
    const int n = n_img - 1, S = T + n, R = B * S, RC = B * n_img, step3 = L / 3;
    if (has_feat && l % step3 == 0) {
        CAR_TRY(tr_mlp(st, t->ctok, RC, dim, t->b_ctl1[l / step3], t->b_ctl2[l / step3], dim, t->ctmp, t->cadd));
        CAR_LAUNCH(tr_add_rows_kernel, tr_grid((long long)RC * dim), 256, 0, st, t->h, (const bf16*)t->cadd, B, n_img, S, T - 1, dim);
    }
    if (for_bwd) CAR_CUDA(cudaMemcpyAsync(t->h0, t->h, (size_t)R * dim * 4, cudaMemcpyDeviceToDevice, st));
    size_t att_smem = 0;
    CAR_TRY(tr_attn_smem((size_t)S, (const void*)tr_attention_kernel, &att_smem));
    CAR_LAUNCH(tr_rmsnorm_kernel, R, 256, 0, st, (const float*)t->h, (const float*)t->attention_norm[l], t->x, dim, d.norm_eps, S, S, 0);
    CAR_TRY(dense_linear(st, t->x, dim, t->b_wqkv[l], R, 3 * dim, dim, ACT_NONE, nullptr, 0, t->qkv, 3 * dim));
    CAR_LAUNCH(rope_kv_write_kernel, 148 * 8, 256, 0, st, (const bf16*)t->qkv, t->rope, t->q, t->kc, t->vc, R, S, dim, H, S);
    CAR_LAUNCH(tr_attention_kernel, (unsigned)(((long long)B * H * S + TRA_WARPS - 1) / TRA_WARPS), TRA_WARPS * 32, att_smem, st, (const bf16*)t->q,
               (const bf16*)t->kc, (const bf16*)t->vc, mask, B, H, S, t->att);
    CAR_TRY(dense_linear(st, t->att, dim, t->b_wo[l], R, dim, dim, ACT_NONE, nullptr, 0, t->o, dim));
    CAR_LAUNCH(tr_add_rows_kernel, tr_grid((long long)R * dim), 256, 0, st, t->h, (const bf16*)t->o, B, S, S, 0, dim);
    bf16* xn = for_bwd ? t->x2 : t->x;
    CAR_LAUNCH(tr_rmsnorm_kernel, R, 256, 0, st, (const float*)t->h, (const float*)t->ffn_norm[l], xn, dim, d.norm_eps, S, S, 0);
    CAR_TRY(dense_linear(st, xn, dim, t->b_w1[l], R, F, dim, ACT_NONE, nullptr, 0, t->g, F));
    CAR_TRY(dense_linear(st, xn, dim, t->b_w3[l], R, F, dim, ACT_NONE, nullptr, 0, t->u, F));
    CAR_LAUNCH(swiglu_kernel, 148 * 8, 256, 0, st, (const bf16*)t->g, (const bf16*)t->u, t->act, (long long)R * F);
    if (for_bwd) return CAR_OK;
    CAR_TRY(dense_linear(st, t->act, F, t->b_w2[l], R, dim, F, ACT_NONE, nullptr, 0, t->o, dim));
    CAR_LAUNCH(tr_add_rows_kernel, tr_grid((long long)R * dim), 256, 0, st, t->h, (const bf16*)t->o, B, S, S, 0, dim);
    return CAR_OK;
}

extern "C" int car_train_forward(CarTrain* t, int32_t B, int32_t n_img, const int32_t* idx, const void* cond, const void* feat,
                                 const uint8_t* drop_ids, const uint8_t* mask, const int32_t* targets, const float* valid,
                                 float* logits_out, float* loss_out, void* stream) {
    if (!t || !idx || !cond || !drop_ids) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (B <= 0 || B > t->maxB || n_img < 2 || n_img > t->maxN) CAR_FAIL(CAR_ERR_ARG, "batch / token count beyond the capacity given to car_train_create");
    if ((loss_out != nullptr) != (targets != nullptr)) CAR_FAIL(CAR_ERR_ARG, "loss_out and targets go together");
    cudaStream_t st = (cudaStream_t)stream;
    const CarModelDesc& d = t->d;
    const int L = d.n_layer, dim = d.dim, F = d.ffn_dim, V = d.vocab_size, T = d.cls_token_num;
    const int n = n_img - 1, S = T + n, R = B * S, RC = B * n_img;
    if (S > T + d.block_size) CAR_FAIL(CAR_ERR_ARG, "sequence longer than the RoPE table");
    t->fwd_ok = false;
    // 0. autocast: bf16 copies of every nn.Linear weight, re-cast each forward (the fp32 masters may have been stepped)
    for (int l = 0; l < L; ++l) {
        CAR_TRY(tr_cast(st, t->wqkv[l], t->b_wqkv[l], (long long)3 * dim * dim)); CAR_TRY(tr_cast(st, t->wo[l], t->b_wo[l], (long long)dim * dim));
        CAR_TRY(tr_cast(st, t->w1[l], t->b_w1[l], (long long)F * dim)); CAR_TRY(tr_cast(st, t->w3[l], t->b_w3[l], (long long)F * dim));
        CAR_TRY(tr_cast(st, t->w2[l], t->b_w2[l], (long long)dim * F));
    }
    CAR_TRY(tr_cast(st, t->w.w.output, t->b_out, (long long)V * dim));
    if (d.model_type == 1) { CAR_TRY(tr_cast(st, t->w.w.cap_fc1, t->b_cap1, (long long)dim * d.caption_dim)); CAR_TRY(tr_cast(st, t->w.w.cap_fc2, t->b_cap2, (long long)dim * dim)); }
    if (feat) {
        CAR_TRY(tr_cast(st, t->w.w.cond_fc1, t->b_cond1, (long long)dim * dim)); CAR_TRY(tr_cast(st, t->w.w.cond_fc2, t->b_cond2, (long long)dim * dim));
        for (int j = 0; j < 3; ++j) { CAR_TRY(tr_cast(st, t->w.w.ctl_fc1[j], t->b_ctl1[j], (long long)dim * dim)); CAR_TRY(tr_cast(st, t->w.w.ctl_fc2[j], t->b_ctl2[j], (long long)dim * dim)); }
        CAR_TRY(tr_cast(st, t->w.adapter_fc1, t->b_ad1, (long long)dim * t->w.adapter_dim)); CAR_TRY(tr_cast(st, t->w.adapter_fc2, t->b_ad2, (long long)dim * dim));
    }
    // 1. prefix rows: CaptionEmbedder (token_drop, cap_proj) gpt_t2i.py:145-162 or LabelEmbedder :78-97; image-token rows :423
    if (d.model_type == 1) {
        CAR_LAUNCH(tr_caption_select_kernel, tr_grid((long long)B * T * d.caption_dim), 256, 0, st, (const float*)cond, (const float*)t->w.cap_uncond,
                   drop_ids, t->capx, B, T, d.caption_dim);
        CAR_TRY(tr_mlp(st, t->capx, B * T, d.caption_dim, t->b_cap1, t->b_cap2, dim, t->ctmp, t->o));
        CAR_LAUNCH(tr_put_rows_bf16_kernel, tr_grid((long long)B * T * dim), 256, 0, st, (const bf16*)t->o, t->h, B, T, S, 0, dim);
    } else {
        CAR_LAUNCH(tr_embed_rows_kernel, B, 256, 0, st, (const float*)t->w.w.label_table, (const int*)cond, 1, drop_ids, t->w.num_classes, t->h, B, 1, S, 0, dim);
    }
    CAR_LAUNCH(tr_embed_rows_kernel, B * n, 256, 0, st, (const float*)t->w.w.tok_embeddings, (const int*)idx, n, (const unsigned char*)nullptr, 0, t->h, B, n, S, T, dim);
    // 2. control tokens: adapter_mlp -> token_drop -> condition_mlp  gpt_t2i.py:424-427 (feat = the control encoder's output tokens)
    if (feat) {
        CAR_TRY(tr_mlp(st, (const bf16*)feat, RC, t->w.adapter_dim, t->b_ad1, t->b_ad2, dim, t->ctmp, t->cin));
        CAR_LAUNCH(tr_select_uncond_kernel, tr_grid((long long)RC * dim), 256, 0, st, t->cin, (const float*)t->w.cond_uncond, drop_ids, B, (long long)n_img * dim);
        CAR_TRY(tr_mlp(st, t->cin, RC, dim, t->b_cond1, t->b_cond2, dim, t->ctmp, t->ctok));
    }
    // 3. blocks  gpt_t2i.py:456-468; the stream at every block input is kept for the backward's recompute
    for (int l = 0; l < L; ++l) {
        CAR_CUDA(cudaMemcpyAsync(t->hs + (size_t)l * R * dim, t->h, (size_t)R * dim * 4, cudaMemcpyDeviceToDevice, st));
        CAR_TRY(tr_block_fwd(t, st, l, B, n_img, mask, feat != nullptr, false));
    }
    // 4. head on rows T-1 .. S-1 of every sample (gpt_t2i.py:469-473), loss :474-481
    CAR_LAUNCH(tr_rmsnorm_kernel, RC, 256, 0, st, (const float*)t->h, (const float*)t->w.w.norm, t->x, dim, d.norm_eps, n_img, S, T - 1);
    CAR_TRY(dense_linear(st, t->x, dim, t->b_out, RC, V, dim, ACT_NONE, nullptr, 0, t->lg, V));
    if (targets) {
        CAR_LAUNCH(tr_ce_rows_kernel, RC, 256, 0, st, (const bf16*)t->lg, (const int*)targets, logits_out, t->nll, V);
        CAR_LAUNCH(tr_ce_reduce_kernel, 1, 1024, 0, st, (const float*)t->nll, valid, B, n_img, loss_out);
    } else if (logits_out) {
        CAR_LAUNCH(tr_put_rows_bf16_kernel, tr_grid((long long)RC * V), 256, 0, st, (const bf16*)t->lg, logits_out, 1, RC, RC, 0, V);
    }
    t->fB = B; t->fN = n_img; t->f_idx = idx; t->f_cond = cond; t->f_feat = feat; t->f_drop = drop_ids; t->f_mask = mask; t->f_targets = targets;
    t->f_valid = valid;
    t->fwd_ok = targets != nullptr;
    return CAR_OK;
}

// ---- backward helpers: the two GEMMs of a linear layer's backward on the [N][K] x [M][K]^T tensor-core kernels ----------------
// dX [rows][K] = bf16(dY [rows][N] . W [N][K] (+ resid)): needs W^T as the K-major operand
static int tr_dgrad(CarTrain* t, cudaStream_t st, const bf16* dY, const bf16* Wb, int rows, int N, int K, const bf16* resid, bf16* dX) {
    CAR_LAUNCH(tr_transpose_pad_kernel, dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, st, Wb, t->wT, N, K, N);
    return dense_linear(st, dY, N, t->wT, rows, K, N, ACT_NONE, resid, K, dX, K);
}
// grad [N][K] fp32 = float(bf16(dY^T . X)), dY [rows][N], X [rows][K]: both operands transposed, the row extent zero-padded to 64
static int tr_wgrad(CarTrain* t, cudaStream_t st, const bf16* dY, const bf16* X, int rows, int N, int K, float* grad) {
    if (!grad) return CAR_OK;
    const int Rp = (rows + 63) / 64 * 64;
    if ((size_t)Rp > t->Rp_max) CAR_FAIL(CAR_ERR_ARG, "row count beyond the transpose workspace");
    CAR_LAUNCH(tr_transpose_pad_kernel, dim3(Rp / 32, (N + 31) / 32), dim3(32, 8), 0, st, dY, t->yT, rows, N, Rp);
    CAR_LAUNCH(tr_transpose_pad_kernel, dim3(Rp / 32, (K + 31) / 32), dim3(32, 8), 0, st, X, t->xT, rows, K, Rp);
    CAR_TRY(dense_linear(st, t->yT, Rp, t->xT, N, K, Rp, ACT_NONE, nullptr, 0, t->dWb, K));
    CAR_LAUNCH(tr_bf16_to_f32_kernel, tr_grid((long long)N * K), 256, 0, st, (const bf16*)t->dWb, grad, (long long)N * K);
    return CAR_OK;
}
// RMSNorm backward on `rows` output rows (row map like tr_rmsnorm_kernel) + the weight gradient
static int tr_norm_bwd(CarTrain* t, cudaStream_t st, const float* h, const void* w, const bf16* dy, int rows, int nrows, int S, int row0, float* gw) {
    const int dim = t->d.dim;
    CAR_LAUNCH(tr_rmsnorm_bwd_kernel, rows, 256, 0, st, h, (const float*)w, dy, t->dh, t->scr, dim, t->d.norm_eps, nrows, S, row0);
    if (gw) {
        CAR_LAUNCH(tr_colsum_part_kernel, dim3((dim + 31) / 32, TR_COLSUM_CHUNKS), dim3(32, 8), 0, st, (const float*)t->scr, t->part, rows, dim);
        CAR_LAUNCH(tr_colsum_final_kernel, (dim + 255) / 256, 256, 0, st, (const float*)t->part, gw, dim);
    }
    return CAR_OK;
}
// MLP backward (gpt_t2i.py:177-181: fc2(gelu_tanh(fc1 x)), no bias), recomputing the two intermediates.  dX (optional) = bf16(dT . fc1 (+ resid))
static int tr_mlp_bwd(CarTrain* t, cudaStream_t st, const bf16* x, int rows, int K, const bf16* fc1, const bf16* fc2, const bf16* dY,
                      const bf16* resid, bf16* dX, float* g1, float* g2) {
    const int dim = t->d.dim;
    CAR_TRY(dense_linear(st, x, K, fc1, rows, dim, K, ACT_NONE, nullptr, 0, t->m_t, dim));
    CAR_LAUNCH(tr_gelu_kernel, tr_grid((long long)rows * dim), 256, 0, st, (const bf16*)t->m_t, t->m_a, (long long)rows * dim);
    CAR_TRY(tr_wgrad(t, st, dY, t->m_a, rows, dim, dim, g2));
    CAR_TRY(tr_dgrad(t, st, dY, fc2, rows, dim, dim, nullptr, t->m_da));
    CAR_LAUNCH(tr_gelu_bwd_kernel, tr_grid((long long)rows * dim), 256, 0, st, (const bf16*)t->m_t, (const bf16*)t->m_da, t->m_dt, (long long)rows * dim);
    CAR_TRY(tr_wgrad(t, st, t->m_dt, x, rows, dim, K, g1));
    if (dX) CAR_TRY(tr_dgrad(t, st, t->m_dt, fc1, rows, dim, K, resid, dX));
    return CAR_OK;
}

// Backward of the last car_train_forward(targets != NULL) on this handle: writes d loss / d parameter (fp32, OVERWRITTEN, scaled
// by *loss_grad when given) through the non-NULL pointers of `g` (a CarTrainWeights whose fields point at gradient buffers of the
// parameters' shapes) and d loss / d feat (bf16 [B, n_img, adapter_dim]) when d_feat is given.
extern "C" int car_train_backward(CarTrain* t, const CarTrainWeights* g, void* d_feat, const float* loss_grad, void* stream) {
    if (!t || !g) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (!t->fwd_ok) CAR_FAIL(CAR_ERR_ARG, "car_train_backward needs a preceding car_train_forward with targets on the same handle");
    cudaStream_t st = (cudaStream_t)stream;
    const CarModelDesc& d = t->d;
    const int L = d.n_layer, dim = d.dim, F = d.ffn_dim, V = d.vocab_size, T = d.cls_token_num, H = d.n_head;
    if (dim % 64 != 0 || F % 64 != 0 || V % 64 != 0) CAR_FAIL(CAR_ERR_UNSUPPORTED, "backward: dim, ffn_dim and vocab_size must be multiples of 64");
    const int B = t->fB, n_img = t->fN, n = n_img - 1, S = T + n, R = B * S, RC = B * n_img, step3 = L / 3;
    const bool has_feat = t->f_feat != nullptr;
    const uint8_t* mask = t->f_mask;
    size_t smem_q = 0, smem_kv = 0;
    CAR_TRY(tr_attn_smem((size_t)2 * S + 128, (const void*)tr_attn_bwd_q_kernel, &smem_q));
    CAR_TRY(tr_attn_smem((size_t)2 * S + 128, (const void*)tr_attn_bwd_kv_kernel, &smem_kv));
    const unsigned att_grid = (unsigned)(((long long)B * H * S + TRA_WARPS - 1) / TRA_WARPS);
    t->fwd_ok = false;                                         // the recompute below overwrites the forward's buffers
    // ---- head: loss -> logits -> output projection -> final norm (gpt_t2i.py:469-481) ----
    CAR_LAUNCH(tr_rmsnorm_kernel, RC, 256, 0, st, (const float*)t->h, (const float*)t->w.w.norm, t->x, dim, d.norm_eps, n_img, S, T - 1);
    CAR_LAUNCH(tr_ce_grad_kernel, RC, 256, 0, st, (const bf16*)t->lg, (const int*)t->f_targets, t->f_valid, loss_grad, B, n_img, t->dlg, V);
    CAR_TRY(tr_wgrad(t, st, t->dlg, t->x, RC, V, dim, (float*)g->w.output));
    CAR_TRY(tr_dgrad(t, st, t->dlg, t->b_out, RC, V, dim, nullptr, t->dx));
    CAR_CUDA(cudaMemsetAsync(t->dh, 0, (size_t)R * dim * 4, st));
    CAR_TRY(tr_norm_bwd(t, st, t->h, t->w.w.norm, t->dx, RC, n_img, S, T - 1, (float*)g->w.norm));
    // ---- blocks, last to first: recompute from the saved input stream, then feed-forward half, attention half, control add ----
    bool first_ctl = true;
    for (int l = L - 1; l >= 0; --l) {
        CAR_CUDA(cudaMemcpyAsync(t->h, t->hs + (size_t)l * R * dim, (size_t)R * dim * 4, cudaMemcpyDeviceToDevice, st));
        CAR_TRY(tr_block_fwd(t, st, l, B, n_img, mask, has_feat, true));
        // feed-forward: h_out = h_mid + w2(silu(w1 x2) * w3 x2)
        CAR_LAUNCH(tr_take_rows_bf16_kernel, tr_grid((long long)R * dim), 256, 0, st, (const float*)t->dh, t->db, B, S, S, 0, dim);
        CAR_TRY(tr_wgrad(t, st, t->db, t->act, R, dim, F, g->w.w2 ? (float*)g->w.w2[l] : nullptr));
        CAR_TRY(tr_dgrad(t, st, t->db, t->b_w2[l], R, dim, F, nullptr, t->dact));
        CAR_LAUNCH(tr_swiglu_bwd_kernel, 148 * 8, 256, 0, st, (const bf16*)t->g, (const bf16*)t->u, (const bf16*)t->dact, t->dg, t->du, (long long)R * F);
        CAR_TRY(tr_wgrad(t, st, t->dg, t->x2, R, F, dim, g->w.w1 ? (float*)g->w.w1[l] : nullptr));
        CAR_TRY(tr_wgrad(t, st, t->du, t->x2, R, F, dim, g->w.w3 ? (float*)g->w.w3[l] : nullptr));
        CAR_TRY(tr_dgrad(t, st, t->dg, t->b_w1[l], R, F, dim, nullptr, t->dx));
        CAR_TRY(tr_dgrad(t, st, t->du, t->b_w3[l], R, F, dim, t->dx, t->dx));
        CAR_TRY(tr_norm_bwd(t, st, t->h, t->ffn_norm[l], t->dx, R, S, S, 0, g->w.ffn_norm ? (float*)g->w.ffn_norm[l] : nullptr));
        // attention: h_mid = h0 + wo(sdpa(rope(wqkv x1)))
        CAR_LAUNCH(tr_take_rows_bf16_kernel, tr_grid((long long)R * dim), 256, 0, st, (const float*)t->dh, t->db, B, S, S, 0, dim);
        CAR_TRY(tr_wgrad(t, st, t->db, t->att, R, dim, dim, g->w.wo ? (float*)g->w.wo[l] : nullptr));
        CAR_TRY(tr_dgrad(t, st, t->db, t->b_wo[l], R, dim, dim, nullptr, t->datt));
        CAR_LAUNCH(tr_attn_bwd_q_kernel, att_grid, TRA_WARPS * 32, smem_q, st, (const bf16*)t->q, (const bf16*)t->kc, (const bf16*)t->vc, mask,
                   (const bf16*)t->datt, B, H, S, t->lse, t->dsum, t->dq);
        CAR_LAUNCH(tr_attn_bwd_kv_kernel, att_grid, TRA_WARPS * 32, smem_kv, st, (const bf16*)t->q, (const bf16*)t->kc, (const bf16*)t->vc, mask,
                   (const bf16*)t->datt, (const float*)t->lse, (const float*)t->dsum, B, H, S, t->dk, t->dv);
        CAR_LAUNCH(tr_rope_bwd_kernel, 148 * 8, 256, 0, st, (const bf16*)t->dq, (const bf16*)t->dk, (const bf16*)t->dv, t->rope, t->dqkv, R, S, dim, H, S);
        CAR_TRY(tr_wgrad(t, st, t->dqkv, t->x, R, 3 * dim, dim, g->w.wqkv ? (float*)g->w.wqkv[l] : nullptr));
        CAR_TRY(tr_dgrad(t, st, t->dqkv, t->b_wqkv[l], R, 3 * dim, dim, nullptr, t->dx));
        CAR_TRY(tr_norm_bwd(t, st, t->h0, t->attention_norm[l], t->dx, R, S, S, 0, g->w.attention_norm ? (float*)g->w.attention_norm[l] : nullptr));
        // control add h[:, T-1:] += condition_layers[j](condition_token)
        if (has_feat && l % step3 == 0) {
            const int j = l / step3;
            CAR_LAUNCH(tr_take_rows_bf16_kernel, tr_grid((long long)RC * dim), 256, 0, st, (const float*)t->dh, t->dadd, B, n_img, S, T - 1, dim);
            CAR_TRY(tr_mlp_bwd(t, st, t->ctok, RC, dim, t->b_ctl1[j], t->b_ctl2[j], t->dadd, first_ctl ? nullptr : t->dctok, t->dctok,
                               (float*)g->w.ctl_fc1[j], (float*)g->w.ctl_fc2[j]));
            first_ctl = false;
        }
    }
    // ---- embeddings and the prefix / control front ends ----
    if (g->w.tok_embeddings) {
        CAR_CUDA(cudaMemsetAsync((void*)g->w.tok_embeddings, 0, (size_t)V * dim * 4, st));
        CAR_LAUNCH(tr_embed_grad_kernel, B * n, 256, 0, st, (const float*)t->dh, (const int*)t->f_idx, n, (const unsigned char*)nullptr, 0,
                   (float*)g->w.tok_embeddings, B, n, S, T, dim);
    }
    if (d.model_type == 1) {
        CAR_LAUNCH(tr_take_rows_bf16_kernel, tr_grid((long long)B * T * dim), 256, 0, st, (const float*)t->dh, t->dadd, B, T, S, 0, dim);
        CAR_TRY(tr_mlp_bwd(t, st, t->capx, B * T, d.caption_dim, t->b_cap1, t->b_cap2, t->dadd, nullptr, nullptr, (float*)g->w.cap_fc1, (float*)g->w.cap_fc2));
    } else if (g->w.label_table) {
        CAR_CUDA(cudaMemsetAsync((void*)g->w.label_table, 0, (size_t)(t->w.num_classes + 1) * dim * 4, st));
        CAR_LAUNCH(tr_embed_grad_kernel, B, 256, 0, st, (const float*)t->dh, (const int*)t->f_cond, 1, t->f_drop, t->w.num_classes, (float*)g->w.label_table,
                   B, 1, S, 0, dim);
    }
    if (has_feat) {
        CAR_TRY(tr_mlp_bwd(t, st, t->cin, RC, dim, t->b_cond1, t->b_cond2, t->dctok, nullptr, t->dcin, (float*)g->w.cond_fc1, (float*)g->w.cond_fc2));
        CAR_LAUNCH(tr_zero_dropped_kernel, tr_grid((long long)RC * dim), 256, 0, st, t->dcin, t->f_drop, B, (long long)n_img * dim);
        CAR_TRY(tr_mlp_bwd(t, st, (const bf16*)t->f_feat, RC, t->w.adapter_dim, t->b_ad1, t->b_ad2, t->dcin, nullptr, (bf16*)d_feat,
                           (float*)g->adapter_fc1, (float*)g->adapter_fc2));
    }
    return CAR_OK;
}

// fused AdamW step over a device-resident tensor table (train.cuh); bias corrections from the step count (1-based)
extern "C" int car_adamw_step(const void* tensors_dev, const void* chunks_dev, int32_t n_chunks, float lr, float beta1, float beta2, float eps,
                              int32_t step, void* stream) {
    if (!tensors_dev || !chunks_dev) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (n_chunks <= 0 || step < 1) CAR_FAIL(CAR_ERR_ARG, "n_chunks must be positive and step 1-based");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    CAR_LAUNCH(adamw_multi_kernel, n_chunks, 256, 0, (cudaStream_t)stream, (const CarAdamWTensorDev*)tensors_dev, (const int2*)chunks_dev, lr, beta1, beta2,
               eps, bc1, sqrtf(bc2));
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// T5 text encoder forward: language/t5.py:58-79 -> HF T5EncoderModel(...).last_hidden_state, bf16.
// v1.1 / flan architecture: gated gelu_new feed-forward, no biases, RMS layer norm (eps 1e-6), relative position bias of block 0
// shared by every block, no 1/sqrt(d) scaling.  GEMMs: dense_linear (tcgen05); glue: t5.cuh.
// ---------------------------------------------------------------------------------------------------------
struct CarT5 {
    CarT5Desc d;
    const void *embed, *rel_bias, *final_norm;
    std::vector<const void*> ln1, wq, wk, wv, wo, ln2, wi0, wi1, wo2;
    int max_rows;
    bf16 *h, *x, *q, *k, *v, *att, *g, *u, *act;
    std::vector<void*> owned;
};

extern "C" int car_t5_create(const CarT5Desc* desc, const CarT5Weights* w, int32_t max_rows, void* stream, CarT5** out) {
    if (!desc || !w || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    (void)stream;
    const CarT5Desc& d = *desc;
    if (d.dtype != CAR_BF16) CAR_FAIL(CAR_ERR_UNSUPPORTED, "the T5 encoder is built for bf16 checkpoints (the reference's default torch_dtype, language/t5.py:22)");
    if (d.d_kv != 64 || d.d_model % 8 || d.d_ff % 8 || d.n_heads <= 0 || d.n_layers <= 0 || max_rows <= 0 || d.num_buckets % 2)
        CAR_FAIL(CAR_ERR_UNSUPPORTED, "shape not supported (d_kv 64, dims multiple of 8)");
    CarT5* t = new CarT5();
    t->d = d; t->embed = w->embed; t->rel_bias = w->rel_bias; t->final_norm = w->final_norm; t->max_rows = max_rows;
    const int L = d.n_layers;
    auto cp = [&](std::vector<const void*>& v, const void* const* src) { v.assign(src, src + L); };
    cp(t->ln1, w->ln1); cp(t->wq, w->q); cp(t->wk, w->k); cp(t->wv, w->v); cp(t->wo, w->o); cp(t->ln2, w->ln2);
    cp(t->wi0, w->wi_0); cp(t->wi1, w->wi_1); cp(t->wo2, w->wo);
    const size_t R = (size_t)max_rows, inner = (size_t)d.n_heads * 64;
    int rc = CAR_OK;
    auto A = [&](bf16** p, size_t elems) { if (rc == CAR_OK) rc = alloc_dev(t->owned, (void**)p, elems * 2); };
    A(&t->h, R * d.d_model); A(&t->x, R * d.d_model); A(&t->q, R * inner); A(&t->k, R * inner); A(&t->v, R * inner); A(&t->att, R * inner);
    A(&t->g, R * d.d_ff); A(&t->u, R * d.d_ff); A(&t->act, R * d.d_ff);
    if (rc != CAR_OK) { for (void* p : t->owned) cudaFree(p); delete t; return rc; }
    *out = t;
    return CAR_OK;
}
extern "C" int car_t5_destroy(CarT5* t) {
    if (!t) return CAR_OK;
    for (void* p : t->owned) cudaFree(p);
    delete t;
    return CAR_OK;
}
// ids int32 [B][L], mask int32 [B][L] (1 = token, 0 = padding) -> last_hidden_state bf16 [B][L][d_model]
extern "C" int car_t5_forward(CarT5* t, const int32_t* ids, const int32_t* mask, int32_t B, int32_t L, void* out, void* stream) {
    if (!t || !ids || !mask || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (B <= 0 || L <= 0 || (long long)B * L > t->max_rows) CAR_FAIL(CAR_ERR_ARG, "batch x length beyond the capacity given to car_t5_create");
    cudaStream_t st = (cudaStream_t)stream;
    const CarT5Desc& d = t->d;
    const int R = B * L, dm = d.d_model, inner = d.n_heads * 64, F = d.d_ff;
    const size_t att_smem = (size_t)T5A_WARPS * L * 4;
    if (att_smem > 48 * 1024) CAR_FAIL(CAR_ERR_UNSUPPORTED, "sequence too long for the T5 attention kernel (L <= 3072)");
    CAR_LAUNCH(t5_embed_kernel, R, 128, 0, st, (const int*)ids, (const bf16*)t->embed, t->h, R, dm);
    for (int l = 0; l < d.n_layers; ++l) {
        // T5LayerSelfAttention (modeling_t5.py: layer_norm -> SelfAttention -> residual)
        CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), R, 256, 0, st, (const bf16*)t->h, (const bf16*)t->ln1[l], t->x, dm, d.eps);
        CAR_TRY(dense_linear(st, t->x, dm, t->wq[l], R, inner, dm, ACT_NONE, nullptr, 0, t->q, inner));
        CAR_TRY(dense_linear(st, t->x, dm, t->wk[l], R, inner, dm, ACT_NONE, nullptr, 0, t->k, inner));
        CAR_TRY(dense_linear(st, t->x, dm, t->wv[l], R, inner, dm, ACT_NONE, nullptr, 0, t->v, inner));
        CAR_LAUNCH(t5_attention_kernel, (unsigned)(((long long)B * d.n_heads * L + T5A_WARPS - 1) / T5A_WARPS), T5A_WARPS * 32, att_smem, st,
                   (const bf16*)t->q, (const bf16*)t->k, (const bf16*)t->v, (const bf16*)t->rel_bias, (const int*)mask, B, d.n_heads, L, d.num_buckets,
                   d.max_distance, t->att);
        CAR_TRY(dense_linear(st, t->att, inner, t->wo[l], R, dm, inner, ACT_NONE, t->h, dm, t->h, dm));          // hidden + attention_output
        // T5LayerFF: layer_norm -> wi_0 / wi_1 -> gelu_new(.) * . -> wo -> residual
        CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), R, 256, 0, st, (const bf16*)t->h, (const bf16*)t->ln2[l], t->x, dm, d.eps);
        CAR_TRY(dense_linear(st, t->x, dm, t->wi0[l], R, F, dm, ACT_NONE, nullptr, 0, t->g, F));
        CAR_TRY(dense_linear(st, t->x, dm, t->wi1[l], R, F, dm, ACT_NONE, nullptr, 0, t->u, F));
        CAR_LAUNCH(t5_geglu_kernel, (int)std::min<long long>(((long long)R * F + 255) / 256, 148 * 16), 256, 0, st, (const bf16*)t->g, (const bf16*)t->u, t->act,
                   (long long)R * F);
        CAR_TRY(dense_linear(st, t->act, F, t->wo2[l], R, dm, F, ACT_NONE, t->h, dm, t->h, dm));
    }
    CAR_LAUNCH((rmsnorm_rows_kernel<bf16>), R, 256, 0, st, (const bf16*)t->h, (const bf16*)t->final_norm, (bf16*)out, dm, d.eps);
    return CAR_OK;
}
