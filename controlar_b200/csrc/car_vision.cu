// car_vision.cu — C-ABI entry points for the control encoder (DINOv2) and the VQGAN tokenizer.
#include <vector>
#include <cstdio>
#include <algorithm>

#include "common.cuh"
#include "gemm_dense.cuh"
#include "gemm_tc5.cuh"
#include "vision.cuh"
#include "frontend.cuh"

static int vis_sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}
// tcgen05 path of the vision stages (gemm_tc5.cuh): TMA tensor-map loads, accumulator in TMEM, persistent warp-specialised CTAs.
// CAR_TC5=0 sends everything back to the mma.sync kernel (dev A/B).
static bool tc5_on() {
    static const bool on = [] { const char* e = getenv("CAR_TC5"); return e ? atoi(e) != 0 : true; }();
    return on && t5_encoder() != nullptr;
}
static int tc5_launch(cudaStream_t st, const CUtensorMap& mapA, const CUtensorMap& mapB, const Tc5P& q, int tiles_m) {
    static DevOnce once5;
    if (once5.first()) CAR_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
    const int ntiles = tiles_m * ((q.N + T5_BN - 1) / T5_BN);
    CAR_LAUNCH(gemm_tc5_kernel, std::min(ntiles, vis_sm_count()), T5_THREADS, T5_SMEM, st, mapA, mapB, q);
    return CAR_OK;
}

static int dense(cudaStream_t st, DenseP p, int batch = 1) {
    static DevOnce once;
    if (once.first()) CAR_CUDA(cudaFuncSetAttribute(dense_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
    if (p.M <= 0 || p.N <= 0) return CAR_OK;
    if (p.alpha == 0.f) p.alpha = 1.f;
    // plain bf16 -> bf16 GEMMs (DINOv2 linears, 1x1 convolutions) go to the tcgen05 kernel; same epilogue order as below
    if (batch == 1 && p.amode == A_PLAIN && p.out_mode == 0 && !p.bias_along_m && p.alpha == 1.f && !p.bias_f && !p.resid_f && tc5_on() &&
        p.K % 8 == 0 && p.N % 8 == 0 && p.lda % 8 == 0 && p.ldb % 8 == 0 && p.ldc % 8 == 0 && (!p.resid || p.ldr % 8 == 0) &&
        ((uintptr_t)p.A % 16) == 0 && ((uintptr_t)p.B % 16) == 0 && ((uintptr_t)p.C % 16) == 0 && (!p.resid || ((uintptr_t)p.resid % 16) == 0)) {
        alignas(64) CUtensorMap mapA, mapB;
        if (!t5_make_map(&mapA, p.A, p.M, p.K, p.lda) || !t5_make_map(&mapB, p.B, p.N, p.K, p.ldb)) CAR_FAIL(CAR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        Tc5P q;
        memset(&q, 0, sizeof(q));
        q.M = p.M; q.N = p.N; q.K = p.K; q.resid = p.resid; q.ldr = p.ldr; q.C = (bf16*)p.C; q.ldc = p.ldc;
        q.act = p.act == ACT_GELU_TANH ? 1 : (p.act == ACT_GELU_ERF ? 2 : 0); q.bias = p.bias; q.scale = p.scale;
        return tc5_launch(st, mapA, mapB, q, (p.M + T5_BM - 1) / T5_BM);
    }
    dim3 grid((p.N + DG_BN - 1) / DG_BN, (p.M + DG_BM - 1) / DG_BM, batch);
    CAR_LAUNCH(dense_gemm_kernel, grid, DG_THREADS, DG_SMEM, st, p);
    return CAR_OK;
}
static DenseP dp_plain(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, void* C, int ldc) {
    DenseP p;
    memset(&p, 0, sizeof(p));
    p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.C = C; p.ldc = ldc; p.alpha = 1.f;
    return p;
}
static inline int gsz(long long total, int block = 256) { return (int)std::min<long long>((total + block - 1) / block, 148 * 16); }

struct Arena {   // grow-only device workspace, re-used across calls (no allocation in steady state)
    char* base = nullptr; size_t cap = 0, off = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return CAR_OK;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        CAR_CUDA(cudaMalloc(&base, bytes));
        cap = bytes;
        return CAR_OK;
    }
    void reset() { off = 0; }
    void* take(size_t bytes) { void* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; }
    void release() { if (base) cudaFree(base); base = nullptr; cap = 0; }
};

// =========================================================================================================
// DINOv2 control encoder
// =========================================================================================================
struct CarDino {
    CarDinoDesc d;
    std::vector<void*> owned;
    // bf16 GEMM-ready weights
    bf16* w_patch;                      // [C][kpad]  (k = 3 * patch^2 = 588 -> 608 for DINOv2, 768 for ViT-S/16)
    int kpatch, kpad;
    const void *b_patch, *cls, *pos, *ln_w, *ln_b;
    struct Layer { bf16 *w_qk, *b_qk, *w_v; const void *b_v, *w_o, *b_o, *ls1, *ls2, *n1w, *n1b, *n2w, *n2b, *w_fc1, *b_fc1, *w_fc2, *b_fc2; };
    std::vector<Layer> L;
    bf16 *ad_fc1, *ad_fc2;              // adapter_mlp (bias-free)
    int ad_dim;
    Arena ws;
};

template <typename TI>
static int to_bf16(cudaStream_t st, std::vector<void*>& owned, const void* src, long long n, bf16** dst) {
    CAR_CUDA(cudaMalloc((void**)dst, (size_t)n * 2));
    owned.push_back(*dst);
    CAR_LAUNCH((cast_to_bf16_kernel<TI>), gsz(n), 256, 0, st, (const TI*)src, *dst, n);
    return CAR_OK;
}
static int to_bf16_any(cudaStream_t st, int dtype, std::vector<void*>& owned, const void* src, long long n, bf16** dst) {
    return dtype == CAR_BF16 ? to_bf16<bf16>(st, owned, src, n, dst) : to_bf16<float>(st, owned, src, n, dst);
}

extern "C" int car_dino_create(const CarDinoDesc* desc, const CarDinoWeights* w, void* stream, CarDino** out) {
    if (!desc || !w || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    const CarDinoDesc& d = *desc;
    if (d.hidden % 64 || d.heads * 64 != d.hidden) CAR_FAIL(CAR_ERR_UNSUPPORTED, "DINOv2 head_dim must be 64");
    if (d.patch != 14 && d.patch != 16) CAR_FAIL(CAR_ERR_UNSUPPORTED, "patch size must be 14 (DINOv2) or 16 (ViT-S/16)");
    cudaStream_t st = (cudaStream_t)stream;
    CarDino* m = new CarDino();
    m->d = d;
    const int C = d.hidden, dt = d.dtype;
    int r = CAR_OK;
    auto T = [&](int rc) { if (r == CAR_OK) r = rc; };
    // patch projection [C][3*P*P] -> [C][kpad] zero padded (k-tiles of 32)
    m->kpatch = 3 * d.patch * d.patch;
    m->kpad = (m->kpatch + 31) & ~31;
    {
        bf16* tmp = nullptr;
        T(to_bf16_any(st, dt, m->owned, w->patch_w, (long long)C * m->kpatch, &tmp));
        if (r == CAR_OK && cudaMalloc((void**)&m->w_patch, (size_t)C * m->kpad * 2) != cudaSuccess) r = CAR_ERR_CUDA;
        if (r == CAR_OK) {
            m->owned.push_back(m->w_patch);
            cudaMemsetAsync(m->w_patch, 0, (size_t)C * m->kpad * 2, st);
            cudaMemcpy2DAsync(m->w_patch, (size_t)m->kpad * 2, tmp, (size_t)m->kpatch * 2, (size_t)m->kpatch * 2, C, cudaMemcpyDeviceToDevice, st);
        }
    }
    auto cv = [&](const void* src, long long n) -> const void* {   // bf16 view of a (possibly fp32) vector / matrix
        bf16* p = nullptr;
        T(to_bf16_any(st, dt, m->owned, src, n, &p));
        return p;
    };
    m->b_patch = cv(w->patch_b, C); m->cls = w->cls_token; m->pos = w->pos_emb;
    m->ln_w = cv(w->ln_w, C); m->ln_b = cv(w->ln_b, C);
    m->L.resize(d.layers);
    for (int l = 0; l < d.layers && r == CAR_OK; ++l) {
        CarDino::Layer& Ly = m->L[l];
        // q and k fused into one [2C][C] weight (+bias); v kept separate (computed transposed)
        if (cudaMalloc((void**)&Ly.w_qk, (size_t)2 * C * C * 2) != cudaSuccess || cudaMalloc((void**)&Ly.b_qk, (size_t)2 * C * 2) != cudaSuccess) { r = CAR_ERR_CUDA; break; }
        m->owned.push_back(Ly.w_qk); m->owned.push_back(Ly.b_qk);
        const bf16 *wq = (const bf16*)cv(w->q_w[l], (long long)C * C), *wk = (const bf16*)cv(w->k_w[l], (long long)C * C);
        const bf16 *bq = (const bf16*)cv(w->q_b[l], C), *bk = (const bf16*)cv(w->k_b[l], C);
        if (r != CAR_OK) break;
        cudaMemcpyAsync(Ly.w_qk, wq, (size_t)C * C * 2, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(Ly.w_qk + (size_t)C * C, wk, (size_t)C * C * 2, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(Ly.b_qk, bq, (size_t)C * 2, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(Ly.b_qk + C, bk, (size_t)C * 2, cudaMemcpyDeviceToDevice, st);
        Ly.w_v = (bf16*)cv(w->v_w[l], (long long)C * C); Ly.b_v = cv(w->v_b[l], C);
        Ly.w_o = cv(w->o_w[l], (long long)C * C); Ly.b_o = cv(w->o_b[l], C);
        Ly.ls1 = cv(w->ls1[l], C); Ly.ls2 = cv(w->ls2[l], C);
        Ly.n1w = cv(w->n1_w[l], C); Ly.n1b = cv(w->n1_b[l], C); Ly.n2w = cv(w->n2_w[l], C); Ly.n2b = cv(w->n2_b[l], C);
        Ly.w_fc1 = cv(w->fc1_w[l], (long long)4 * C * C); Ly.b_fc1 = cv(w->fc1_b[l], 4 * C);
        Ly.w_fc2 = cv(w->fc2_w[l], (long long)4 * C * C); Ly.b_fc2 = cv(w->fc2_b[l], C);
    }
    m->ad_fc1 = m->ad_fc2 = nullptr; m->ad_dim = d.adapter_out_dim;
    if (w->adapter_fc1 && d.adapter_out_dim > 0) {
        m->ad_fc1 = (bf16*)cv(w->adapter_fc1, (long long)d.adapter_out_dim * C);
        m->ad_fc2 = (bf16*)cv(w->adapter_fc2, (long long)d.adapter_out_dim * d.adapter_out_dim);
    }
    if (r != CAR_OK) { for (void* p : m->owned) cudaFree(p); delete m; return r; }
    *out = m;
    return CAR_OK;
}

extern "C" int car_dino_destroy(CarDino* m) {
    if (!m) return CAR_OK;
    for (void* p : m->owned) cudaFree(p);
    m->ws.release();
    delete m;
    return CAR_OK;
}

template <typename TI>
static int dino_forward_t(CarDino* m, const TI* image, int B, int H, int W, void* out, int apply_mlp, cudaStream_t st) {
    const CarDinoDesc& d = m->d;
    const int C = d.hidden, h = H / 16, w = W / 16, hw = h * w, Tn = hw + 1, heads = d.heads;
    const int Tp = (Tn + 31) & ~31;                      // key axis padded for the P·V GEMM
    const long long rows = (long long)B * Tn;
    // workspace
    size_t need = 0;
    auto sz = [&](size_t b) { need += (b + 255) & ~(size_t)255; };
    const int KP = m->kpad;
    sz((size_t)B * hw * KP * 2); sz((size_t)B * hw * C * 2); sz((size_t)hw * C * 2);
    sz(rows * C * 2); sz(rows * C * 2); sz(rows * 2 * C * 2); sz((size_t)B * C * Tp * 2);
    sz((size_t)B * heads * Tn * Tp * 4); sz((size_t)B * heads * Tn * Tp * 2); sz(rows * C * 2); sz(rows * 4 * C * 2);
    sz((size_t)B * hw * C * 2); sz((size_t)B * hw * std::max(m->ad_dim, 1) * 2);
    CAR_TRY(m->ws.reserve(need));
    m->ws.reset();
    bf16* patches = (bf16*)m->ws.take((size_t)B * hw * KP * 2);
    bf16* ptok = (bf16*)m->ws.take((size_t)B * hw * C * 2);
    bf16* posi = (bf16*)m->ws.take((size_t)hw * C * 2);
    bf16* x = (bf16*)m->ws.take(rows * C * 2);
    bf16* xn = (bf16*)m->ws.take(rows * C * 2);
    bf16* qk = (bf16*)m->ws.take(rows * 2 * C * 2);
    bf16* vT = (bf16*)m->ws.take((size_t)B * C * Tp * 2);
    float* S = (float*)m->ws.take((size_t)B * heads * Tn * Tp * 4);
    bf16* P = (bf16*)m->ws.take((size_t)B * heads * Tn * Tp * 2);
    bf16* ctx = (bf16*)m->ws.take(rows * C * 2);
    bf16* hid = (bf16*)m->ws.take(rows * 4 * C * 2);
    bf16* feat = (bf16*)m->ws.take((size_t)B * hw * C * 2);
    bf16* mlp_h = (bf16*)m->ws.take((size_t)B * hw * std::max(m->ad_dim, 1) * 2);

    CAR_CUDA(cudaMemsetAsync(vT, 0, (size_t)B * C * Tp * 2, st));   // padded key columns must be finite (x 0 prob)
    // 1. resize to (h*P, w*P) + patchify (dinov2_adapter.py:16-24; ViT: P = 16, no resize), patch projection + bias
    CAR_LAUNCH((resize_patchify_kernel<TI>), gsz((long long)B * hw * KP), 256, 0, st, image, patches, B, H, W, h, w, KP, d.resize_mode, d.patch);
    {
        DenseP p = dp_plain(patches, KP, m->w_patch, KP, B * hw, C, KP, ptok, C);
        p.bias = (const bf16*)m->b_patch;
        CAR_TRY(dense(st, p));
    }
    // 2. CLS + interpolated position embeddings
    CAR_LAUNCH((pos_embed_interp_kernel<TI>), gsz((long long)hw * C), 256, 0, st, (const TI*)m->pos, posi, d.pos_grid, h, w, C);
    CAR_LAUNCH((dino_assemble_kernel<TI>), gsz(rows * C), 256, 0, st, ptok, (const TI*)m->cls, (const TI*)m->pos, posi, x, B, hw, C);
    // 3. encoder layers (modeling_dinov2.py Dinov2Layer.forward)
    const float scale = 0.125f;
    for (int l = 0; l < d.layers; ++l) {
        const CarDino::Layer& Ly = m->L[l];
        CAR_LAUNCH(layernorm_kernel, (unsigned)rows, 128, 0, st, x, (const bf16*)Ly.n1w, (const bf16*)Ly.n1b, xn, C, d.eps, (long long)C, (long long)C);
        {   // q | k  : [rows][2C]
            DenseP p = dp_plain(xn, C, Ly.w_qk, C, (int)rows, 2 * C, C, qk, 2 * C);
            p.bias = Ly.b_qk;
            CAR_TRY(dense(st, p));
        }
        {   // V^T per image: [C][Tp] = Wv[C][C] · xn_b[Tn][C]^T  (+ bias along M); padded key columns stay 0-weighted
            DenseP p = dp_plain(Ly.w_v, C, xn, C, C, Tn, C, vT, Tp);
            p.sB = (long long)Tn * C; p.sC = (long long)C * Tp; p.bias = (const bf16*)Ly.b_v; p.bias_along_m = 1;
            CAR_TRY(dense(st, p, B));
        }
        static const bool fused_attn = [] { const char* e = getenv("CAR_VIT_FA"); return e ? atoi(e) != 0 : true; }();
        if (fused_attn && C % 64 == 0 && ((uintptr_t)qk % 16) == 0 && ((uintptr_t)vT % 16) == 0) {
            // fused attention (vision.cuh): scores / probabilities never leave the SM
            CAR_LAUNCH(vit_attention_kernel, dim3((Tn + 63) / 64, heads, B), 128, 0, st, (const bf16*)qk, (const bf16*)vT, ctx, Tn, Tp, C, scale);
        } else {
        // scores S[b,hd] = q k^T * 1/8  (fp32), soft-max -> P (bf16, zero padded), ctx = P V
        for (int hd = 0; hd < heads; ++hd) {
            DenseP p = dp_plain(qk + hd * 64, 2 * C, qk + C + hd * 64, 2 * C, Tn, Tn, 64, S + (size_t)hd * Tn * Tp, Tp);
            p.sA = (long long)Tn * 2 * C; p.sB = (long long)Tn * 2 * C; p.sC = (long long)heads * Tn * Tp; p.alpha = scale; p.out_mode = 1;
            CAR_TRY(dense(st, p, B));
        }
        CAR_LAUNCH(softmax_rows_kernel, (unsigned)((long long)B * heads * Tn), 128, 0, st, S, P, Tn, Tp, Tp);
        for (int hd = 0; hd < heads; ++hd) {
            DenseP p = dp_plain(P + (size_t)hd * Tn * Tp, Tp, vT + (size_t)hd * 64 * Tp, Tp, Tn, 64, Tp, ctx + hd * 64, C);
            p.sA = (long long)heads * Tn * Tp; p.sB = (long long)C * Tp; p.sC = (long long)Tn * C;
            CAR_TRY(dense(st, p, B));
        }
        }
        {   // x = x + ls1 * (dense(ctx) + b)
            DenseP p = dp_plain(ctx, C, (const bf16*)Ly.w_o, C, (int)rows, C, C, x, C);
            p.bias = (const bf16*)Ly.b_o; p.scale = (const bf16*)Ly.ls1; p.resid = x; p.ldr = C;
            CAR_TRY(dense(st, p));
        }
        CAR_LAUNCH(layernorm_kernel, (unsigned)rows, 128, 0, st, x, (const bf16*)Ly.n2w, (const bf16*)Ly.n2b, xn, C, d.eps, (long long)C, (long long)C);
        {
            DenseP p = dp_plain(xn, C, (const bf16*)Ly.w_fc1, C, (int)rows, 4 * C, C, hid, 4 * C);
            p.bias = (const bf16*)Ly.b_fc1; p.act = ACT_GELU_ERF;
            CAR_TRY(dense(st, p));
        }
        {
            DenseP p = dp_plain(hid, 4 * C, (const bf16*)Ly.w_fc2, 4 * C, (int)rows, C, 4 * C, x, C);
            p.bias = (const bf16*)Ly.b_fc2; p.scale = (const bf16*)Ly.ls2; p.resid = x; p.ldr = C;
            CAR_TRY(dense(st, p));
        }
    }
    // 4. final LayerNorm, drop CLS (dinov2_adapter.py:29): rows of image b start at token 1
    bf16* fdst = apply_mlp ? feat : (bf16*)out;
    for (int b = 0; b < B; ++b)
        CAR_LAUNCH(layernorm_kernel, (unsigned)hw, 128, 0, st, x + ((size_t)b * Tn + 1) * C, (const bf16*)m->ln_w, (const bf16*)m->ln_b,
                   fdst + (size_t)b * hw * C, C, d.eps, (long long)C, (long long)C);
    // 5. optional adapter_mlp (generate.py:138): fc2(gelu_tanh(fc1 x)), bias-free
    if (apply_mlp) {
        if (!m->ad_fc1) CAR_FAIL(CAR_ERR_STATE, "adapter_mlp weights were not registered");
        DenseP p1 = dp_plain(feat, C, m->ad_fc1, C, B * hw, m->ad_dim, C, mlp_h, m->ad_dim);
        p1.act = ACT_GELU_TANH;
        CAR_TRY(dense(st, p1));
        DenseP p2 = dp_plain(mlp_h, m->ad_dim, m->ad_fc2, m->ad_dim, B * hw, m->ad_dim, m->ad_dim, out, m->ad_dim);
        CAR_TRY(dense(st, p2));
    }
    return CAR_OK;
}

extern "C" int car_dino_forward(CarDino* m, const void* image, int32_t B, int32_t H, int32_t W, void* out_bf16, int32_t apply_mlp,
                                void* stream) {
    if (!m || !image || !out_bf16) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (H % 16 || W % 16 || B <= 0) CAR_FAIL(CAR_ERR_ARG, "H and W must be multiples of 16");
    if (m->d.dtype == CAR_BF16) return dino_forward_t<bf16>(m, (const bf16*)image, B, H, W, out_bf16, apply_mlp, (cudaStream_t)stream);
    return dino_forward_t<float>(m, (const float*)image, B, H, W, out_bf16, apply_mlp, (cudaStream_t)stream);
}

// =========================================================================================================
// VQGAN tokenizer
// =========================================================================================================
struct ConvW { bf16* w; bf16* b; int cin, cin_pad, cout, k; bf16* w3; float* bf; };     // w3 / bf: split-bf16 weights + fp32 bias (encoder)
struct NormW { bf16 *w, *b; int c; float *wf, *bff; };                                   // wf / bff: fp32 copies (encoder)
struct ResW { NormW n1, n2; ConvW c1, c2, nin; bool has_nin; };
struct AttnW { NormW n; ConvW q, k, v, o; };

struct CarVQ {
    CarVQDesc d;
    std::vector<void*> owned;
    float* codebook_n;                 // l2-normalised fp32 [n_codes][e_dim]
    // decoder
    ConvW post_quant, d_conv_in, d_conv_out; NormW d_norm_out;
    ResW d_mid0, d_mid2; AttnW d_mid1;
    std::vector<std::vector<ResW>> d_res; std::vector<std::vector<AttnW>> d_attn; std::vector<ConvW> d_up; std::vector<bool> d_has_up;
    // encoder
    ConvW quant_conv, e_conv_in, e_conv_out; NormW e_norm_out;
    ResW e_mid0, e_mid2; AttnW e_mid1;
    std::vector<std::vector<ResW>> e_res; std::vector<std::vector<AttnW>> e_attn; std::vector<ConvW> e_down; std::vector<bool> e_has_down;
    Arena ws;
};

struct TensorCursor { const void* const* t; int n; int i; };

static int keep_f32(CarVQ* m, cudaStream_t st, const void* src, long long n, float** dst) {
    CAR_CUDA(cudaMalloc((void**)dst, (size_t)n * 4));
    m->owned.push_back(*dst);
    CAR_CUDA(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return CAR_OK;
}
static int take_conv(CarVQ* m, cudaStream_t st, TensorCursor& tc, int cout, int cin, int k, ConvW* c, bool x3 = false) {
    if (tc.i + 2 > tc.n) CAR_FAIL(CAR_ERR_ARG, "tensor list too short");
    c->cin = cin; c->cout = cout; c->k = k; c->cin_pad = (cin + 31) & ~31; c->w3 = nullptr; c->bf = nullptr;
    if (x3) {   // fp32-grade encoder path (vision.cuh "x3"): [w_hi | w_hi | w_lo] per tap + the fp32 bias
        const long long n3 = (long long)cout * k * k * c->cin_pad;
        CAR_CUDA(cudaMalloc((void**)&c->w3, (size_t)n3 * 3 * 2));
        m->owned.push_back(c->w3);
        CAR_LAUNCH(conv_weight_pack_x3_kernel, gsz(n3), 256, 0, st, (const float*)tc.t[tc.i], c->w3, cout, cin, k, k, c->cin_pad);
        CAR_TRY(keep_f32(m, st, tc.t[tc.i + 1], cout, &c->bf));
    }
    const long long n = (long long)cout * k * k * c->cin_pad;
    CAR_CUDA(cudaMalloc((void**)&c->w, (size_t)n * 2));
    m->owned.push_back(c->w);
    CAR_LAUNCH((conv_weight_pack_kernel<float>), gsz(n), 256, 0, st, (const float*)tc.t[tc.i], c->w, cout, cin, k, k, c->cin_pad);
    CAR_TRY(to_bf16<float>(st, m->owned, tc.t[tc.i + 1], cout, &c->b));
    tc.i += 2;
    return CAR_OK;
}
static int take_norm(CarVQ* m, cudaStream_t st, TensorCursor& tc, int c, NormW* nw, bool x3 = false) {
    if (tc.i + 2 > tc.n) CAR_FAIL(CAR_ERR_ARG, "tensor list too short");
    nw->c = c; nw->wf = nullptr; nw->bff = nullptr;
    if (x3) { CAR_TRY(keep_f32(m, st, tc.t[tc.i], c, &nw->wf)); CAR_TRY(keep_f32(m, st, tc.t[tc.i + 1], c, &nw->bff)); }
    CAR_TRY(to_bf16<float>(st, m->owned, tc.t[tc.i], c, &nw->w));
    CAR_TRY(to_bf16<float>(st, m->owned, tc.t[tc.i + 1], c, &nw->b));
    tc.i += 2;
    return CAR_OK;
}
// canonical order inside a ResnetBlock: norm1.{w,b} conv1.{w,b} norm2.{w,b} conv2.{w,b} [nin_shortcut.{w,b}]
static int take_res(CarVQ* m, cudaStream_t st, TensorCursor& tc, int cin, int cout, ResW* r, bool x3 = false) {
    CAR_TRY(take_norm(m, st, tc, cin, &r->n1, x3)); CAR_TRY(take_conv(m, st, tc, cout, cin, 3, &r->c1, x3));
    CAR_TRY(take_norm(m, st, tc, cout, &r->n2, x3)); CAR_TRY(take_conv(m, st, tc, cout, cout, 3, &r->c2, x3));
    r->has_nin = cin != cout;
    if (r->has_nin) CAR_TRY(take_conv(m, st, tc, cout, cin, 1, &r->nin, x3));
    return CAR_OK;
}
// AttnBlock: norm.{w,b} q.{w,b} k.{w,b} v.{w,b} proj_out.{w,b}
static int take_attn(CarVQ* m, cudaStream_t st, TensorCursor& tc, int c, AttnW* a, bool x3 = false) {
    CAR_TRY(take_norm(m, st, tc, c, &a->n, x3)); CAR_TRY(take_conv(m, st, tc, c, c, 1, &a->q, x3)); CAR_TRY(take_conv(m, st, tc, c, c, 1, &a->k, x3));
    CAR_TRY(take_conv(m, st, tc, c, c, 1, &a->v, x3)); CAR_TRY(take_conv(m, st, tc, c, c, 1, &a->o, x3));
    return CAR_OK;
}

static int vq_build(CarVQ* m, const void* const* tensors, int n, cudaStream_t st) {
    const CarVQDesc& d = m->d;
    TensorCursor tc{tensors, n, 0};
    const int ch = d.ch, nres = d.n_levels, nrb = d.num_res_blocks;
    // ---- encoder (vq_model.py:65-125)
    CAR_TRY(take_conv(m, st, tc, ch, 3, 3, &m->e_conv_in, true));
    m->e_res.resize(nres); m->e_attn.resize(nres); m->e_down.resize(nres); m->e_has_down.assign(nres, false);
    int block_in = ch;
    for (int lvl = 0; lvl < nres; ++lvl) {
        block_in = ch * (lvl == 0 ? 1 : d.ch_mult[lvl - 1]);
        const int block_out = ch * d.ch_mult[lvl];
        for (int b = 0; b < nrb; ++b) {
            ResW r; CAR_TRY(take_res(m, st, tc, block_in, block_out, &r, true)); m->e_res[lvl].push_back(r);
            block_in = block_out;
            if (lvl == nres - 1) { AttnW a; CAR_TRY(take_attn(m, st, tc, block_in, &a, true)); m->e_attn[lvl].push_back(a); }
        }
        if (lvl != nres - 1) { CAR_TRY(take_conv(m, st, tc, block_in, block_in, 3, &m->e_down[lvl], true)); m->e_has_down[lvl] = true; }
    }
    CAR_TRY(take_res(m, st, tc, block_in, block_in, &m->e_mid0, true)); CAR_TRY(take_attn(m, st, tc, block_in, &m->e_mid1, true));
    CAR_TRY(take_res(m, st, tc, block_in, block_in, &m->e_mid2, true));
    CAR_TRY(take_norm(m, st, tc, block_in, &m->e_norm_out, true)); CAR_TRY(take_conv(m, st, tc, d.z_channels, block_in, 3, &m->e_conv_out, true));
    // ---- decoder (vq_model.py:129-195)
    block_in = ch * d.ch_mult[nres - 1];
    CAR_TRY(take_conv(m, st, tc, block_in, d.z_channels, 3, &m->d_conv_in));
    CAR_TRY(take_res(m, st, tc, block_in, block_in, &m->d_mid0)); CAR_TRY(take_attn(m, st, tc, block_in, &m->d_mid1));
    CAR_TRY(take_res(m, st, tc, block_in, block_in, &m->d_mid2));
    m->d_res.resize(nres); m->d_attn.resize(nres); m->d_up.resize(nres); m->d_has_up.assign(nres, false);
    for (int idx = 0; idx < nres; ++idx) {
        const int lvl = nres - 1 - idx;
        const int block_out = ch * d.ch_mult[lvl];
        for (int b = 0; b < nrb + 1; ++b) {
            ResW r; CAR_TRY(take_res(m, st, tc, block_in, block_out, &r)); m->d_res[idx].push_back(r);
            block_in = block_out;
            if (lvl == nres - 1) { AttnW a; CAR_TRY(take_attn(m, st, tc, block_in, &a)); m->d_attn[idx].push_back(a); }
        }
        if (lvl != 0) { CAR_TRY(take_conv(m, st, tc, block_in, block_in, 3, &m->d_up[idx])); m->d_has_up[idx] = true; }
    }
    CAR_TRY(take_norm(m, st, tc, block_in, &m->d_norm_out)); CAR_TRY(take_conv(m, st, tc, 3, block_in, 3, &m->d_conv_out));
    // ---- quantiser + 1x1 convs
    if (tc.i + 1 > tc.n) CAR_FAIL(CAR_ERR_ARG, "tensor list too short");
    CAR_CUDA(cudaMalloc((void**)&m->codebook_n, (size_t)d.codebook_size * d.embed_dim * 4));
    m->owned.push_back(m->codebook_n);
    CAR_LAUNCH(codebook_normalize_kernel, (d.codebook_size + 255) / 256, 256, 0, st, (const float*)tc.t[tc.i], m->codebook_n, d.codebook_size, d.embed_dim);
    tc.i += 1;
    CAR_TRY(take_conv(m, st, tc, d.embed_dim, d.z_channels, 1, &m->quant_conv, true));
    CAR_TRY(take_conv(m, st, tc, d.z_channels, d.embed_dim, 1, &m->post_quant));
    if (tc.i != tc.n) CAR_FAIL(CAR_ERR_ARG, "tensor list length does not match the VQ architecture");
    return CAR_OK;
}

extern "C" int car_vq_create(const CarVQDesc* desc, const void* const* tensors, int32_t n_tensors, void* stream, CarVQ** out) {
    if (!desc || !tensors || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (desc->embed_dim > 8 || desc->n_levels > 8 || desc->ch % 32) CAR_FAIL(CAR_ERR_UNSUPPORTED, "unsupported VQ shape");
    CarVQ* m = new CarVQ();
    m->d = *desc;
    int r = vq_build(m, tensors, n_tensors, (cudaStream_t)stream);
    if (r != CAR_OK) { for (void* p : m->owned) cudaFree(p); delete m; return r; }
    *out = m;
    return CAR_OK;
}
extern "C" int car_vq_destroy(CarVQ* m) {
    if (!m) return CAR_OK;
    for (void* p : m->owned) cudaFree(p);
    m->ws.release();
    delete m;
    return CAR_OK;
}

// ---- layer helpers on NHWC bf16 activations ----
struct Act { bf16* p; int B, H, W, C; long long n() const { return (long long)B * H * W * C; } };

static int conv_fwd(cudaStream_t st, const ConvW& c, const Act& x, int ups, int stride2, bf16* out, const bf16* resid, int Ho, int Wo,
                    float* out_nchw_f32 = nullptr) {
    if (x.C != c.cin_pad && !(c.k == 1 && x.C == c.cin_pad)) CAR_FAIL(CAR_ERR_STATE, "conv input channel mismatch");
    if (c.k == 3 && !stride2 && !ups && !out_nchw_f32 && c.cin_pad % T5_BK == 0 && c.cout % 8 == 0 && Ho == x.H && Wo == x.W && x.H >= T5_TH && x.W >= T5_TW && tc5_on() &&
        ((uintptr_t)x.p % 16) == 0 && ((uintptr_t)out % 16) == 0 && (!resid || ((uintptr_t)resid % 16) == 0)) {
        // 3x3 / pad 1 convolution on the tcgen05 kernel: one 4-D TMA box per (tap, 64-channel block), padding by TMA zero fill
        alignas(64) CUtensorMap mapA, mapB;
        if (!t5_make_map_nhwc(&mapA, x.p, x.B, x.H, x.W, c.cin_pad) || !t5_make_map(&mapB, c.w, c.cout, 9 * c.cin_pad, 9 * c.cin_pad))
            CAR_FAIL(CAR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        Tc5P q;
        memset(&q, 0, sizeof(q));
        q.M = x.B * x.H * x.W; q.N = c.cout; q.K = 9 * c.cin_pad; q.resid = resid; q.ldr = c.cout; q.C = out; q.ldc = c.cout; q.bias = c.b;
        q.conv = 1; q.H = x.H; q.W = x.W; q.tiles_x = (x.W + T5_TW - 1) / T5_TW; q.tiles_y = (x.H + T5_TH - 1) / T5_TH; q.cblks = c.cin_pad / T5_BK;
        return tc5_launch(st, mapA, mapB, q, x.B * q.tiles_x * q.tiles_y);
    }
    DenseP p;
    memset(&p, 0, sizeof(p));
    p.A = x.p; p.B = c.w; p.M = x.B * Ho * Wo; p.N = c.cout; p.K = c.k * c.k * c.cin_pad; p.ldb = p.K; p.alpha = 1.f;
    if (c.k == 1) { p.amode = A_PLAIN; p.lda = x.C; }
    else { p.amode = stride2 ? A_CONV3x3S2 : A_CONV3x3; p.Hs = x.H; p.Ws = x.W; p.Cin = c.cin_pad; p.ups = ups; }
    p.Ho = Ho; p.Wo = Wo;
    p.bias = c.b;
    if (out_nchw_f32) { p.C = out_nchw_f32; p.out_mode = 2; }
    else { p.C = out; p.ldc = c.cout; p.resid = resid; p.ldr = c.cout; }
    return dense(st, p);
}
static int gn_fwd(cudaStream_t st, const NormW& nw, const Act& x, bf16* y, int swish, float* stats) {
    const int G = 32;
    const int C8 = x.C / 8;
    if (x.C % 64 == 0 && 256 % C8 == 0 && x.H * x.W >= 4 * GN_CHUNKS && ((uintptr_t)x.p % 16) == 0 && ((uintptr_t)y % 16) == 0 &&
        ((uintptr_t)nw.w % 16) == 0 && ((uintptr_t)nw.b % 16) == 0) {
        // coalesced two-stage statistics + 16-byte apply (vision.cuh); `stats` has room for the partial sums behind the [B*32][2] block
        float* part = stats + (size_t)x.B * G * 2;
        CAR_LAUNCH(groupnorm_partial_kernel, dim3(GN_CHUNKS, x.B), 256, 0, st, (const bf16*)x.p, part, x.H * x.W, x.C);
        CAR_LAUNCH(groupnorm_finish_kernel, x.B, 32, 0, st, (const float*)part, stats, x.H * x.W, x.C);
        CAR_LAUNCH(groupnorm_apply8_kernel, gsz(x.n() / 8), 256, 0, st, (const bf16*)x.p, (const float*)stats, (const bf16*)nw.w, (const bf16*)nw.b, y, x.n() / 8,
                   x.H * x.W, x.C, swish);
        return CAR_OK;
    }
    CAR_LAUNCH(groupnorm_stats_kernel, x.B * G, 512, 0, st, x.p, stats, x.H * x.W, x.C, G);
    CAR_LAUNCH(groupnorm_apply_kernel, gsz(x.n()), 256, 0, st, x.p, stats, nw.w, nw.b, y, x.n(), x.H * x.W, x.C, G, swish);
    return CAR_OK;
}

struct VqScratch { bf16 *t0, *t1, *t2; float* stats; float* S; bf16* P; bf16* vT; };

// ResnetBlock.forward (vq_model.py:300-315): x + conv2(swish(GN(conv1(swish(GN(x))))))  [+ 1x1 shortcut]
static int res_fwd(cudaStream_t st, const ResW& r, Act& x, bf16* out, VqScratch& s) {
    Act a = x;
    CAR_TRY(gn_fwd(st, r.n1, x, s.t0, 1, s.stats));
    a.p = s.t0;
    CAR_TRY(conv_fwd(st, r.c1, a, 0, 0, s.t1, nullptr, x.H, x.W));
    Act b{s.t1, x.B, x.H, x.W, r.c1.cout};
    CAR_TRY(gn_fwd(st, r.n2, b, s.t0, 1, s.stats));
    b.p = s.t0;
    const bf16* sc = x.p;
    if (r.has_nin) { CAR_TRY(conv_fwd(st, r.nin, x, 0, 0, s.t2, nullptr, x.H, x.W)); sc = s.t2; }
    CAR_TRY(conv_fwd(st, r.c2, b, 0, 0, out, sc, x.H, x.W));
    x.p = out; x.C = r.c2.cout;
    return CAR_OK;
}
// AttnBlock.forward (vq_model.py:328-352): single head over H*W tokens, scale C^-0.5
static int attn_fwd(cudaStream_t st, const AttnW& a, Act& x, bf16* out, VqScratch& s) {
    const int C = x.C, hw = x.H * x.W, B = x.B;
    const int hwp = (hw + 31) & ~31;
    CAR_TRY(gn_fwd(st, a.n, x, s.t0, 0, s.stats));
    Act xn{s.t0, B, x.H, x.W, C};
    bf16* q = s.t1;
    bf16* k = s.t2;
    CAR_TRY(conv_fwd(st, a.q, xn, 0, 0, q, nullptr, x.H, x.W));
    CAR_TRY(conv_fwd(st, a.k, xn, 0, 0, k, nullptr, x.H, x.W));
    {   // V^T [B][C][hwp]
        DenseP p = dp_plain(a.v.w, C, xn.p, C, C, hw, C, s.vT, hwp);
        p.sB = (long long)hw * C; p.sC = (long long)C * hwp; p.bias = a.v.b; p.bias_along_m = 1;
        CAR_TRY(dense(st, p, B));
    }
    {
        DenseP p = dp_plain(q, C, k, C, hw, hw, C, s.S, hwp);
        p.sA = (long long)hw * C; p.sB = (long long)hw * C; p.sC = (long long)hw * hwp; p.alpha = 1.0f / sqrtf((float)C); p.out_mode = 1;
        CAR_TRY(dense(st, p, B));
    }
    CAR_LAUNCH(softmax_rows_kernel, (unsigned)((long long)B * hw), 256, 0, st, s.S, s.P, hw, hwp, hwp);
    {
        DenseP p = dp_plain(s.P, hwp, s.vT, hwp, hw, C, hwp, q, C);     // ctx overwrites q
        p.sA = (long long)hw * hwp; p.sB = (long long)C * hwp; p.sC = (long long)hw * C;
        CAR_TRY(dense(st, p, B));
    }
    Act ctx{q, B, x.H, x.W, C};
    CAR_TRY(conv_fwd(st, a.o, ctx, 0, 0, out, x.p, x.H, x.W));
    x.p = out;
    return CAR_OK;
}

static int vq_scratch(cudaStream_t st, CarVQ* m, int B, int Hmax, int Wmax, int h16, int w16, int Cfull, VqScratch& s, bf16** bufA, bf16** bufB,
                      size_t extra, void** extra_p) {
    const CarVQDesc& d = m->d;
    const size_t act = (size_t)B * Hmax * Wmax * Cfull * 2;            // largest activation (ch channels at full res)
    const int hw = h16 * w16, hwp = (hw + 31) & ~31, Cmax = d.ch * d.ch_mult[d.n_levels - 1];
    size_t need = 5 * (act + 256) + (size_t)B * 32 * 2 * 4 * (1 + GN_CHUNKS) + 256 + (size_t)B * hw * hwp * 6 + 512 + (size_t)B * Cmax * hwp * 2 + 256 + extra + 256;
    CAR_TRY(m->ws.reserve(need));
    m->ws.reset();
    *bufA = (bf16*)m->ws.take(act); *bufB = (bf16*)m->ws.take(act);
    s.t0 = (bf16*)m->ws.take(act); s.t1 = (bf16*)m->ws.take(act); s.t2 = (bf16*)m->ws.take(act);
    s.stats = (float*)m->ws.take((size_t)B * 32 * 2 * 4 * (1 + GN_CHUNKS));
    s.S = (float*)m->ws.take((size_t)B * hw * hwp * 4); s.P = (bf16*)m->ws.take((size_t)B * hw * hwp * 2);
    s.vT = (bf16*)m->ws.take((size_t)B * Cmax * hwp * 2);
    if (extra_p) *extra_p = m->ws.take(extra);
    CAR_CUDA(cudaMemsetAsync(s.vT, 0, (size_t)B * Cmax * hwp * 2, st));
    return CAR_OK;
}

// VQModel.decode_code (vq_model.py:53-56): codes int32 [B][h*w] -> image fp32 NCHW [B][3][16h][16w] (VQ-16)
// VQModel.decode (vq_model.py:48-51): quant fp32 NCHW [B][e][h][w] -> image
static int vq_decode_impl(CarVQ* m, const int32_t* codes, const float* quant, int32_t B, int32_t h, int32_t w, float* out, void* stream) {
    if (!m || !(codes || quant) || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const CarVQDesc& d = m->d;
    const int up = 1 << (d.n_levels - 1);
    VqScratch s; bf16 *bA, *bB; void* zbuf;
    CAR_TRY(vq_scratch(st, m, B, h * up, w * up, h, w, d.ch, s, &bA, &bB, (size_t)B * h * w * 32 * 2, &zbuf));
    // get_codebook_entry (vq_model.py:262-277) + post_quant_conv
    if (codes) CAR_LAUNCH(codebook_lookup_kernel, gsz((long long)B * h * w * 32), 256, 0, st, m->codebook_n, codes, (bf16*)zbuf, (long long)B * h * w, d.embed_dim, 32, d.codebook_size);
    else CAR_LAUNCH((nchw_to_nhwc_bf16_kernel<float>), gsz((long long)B * h * w * 32), 256, 0, st, quant, (bf16*)zbuf, B, d.embed_dim, h * w, 32);
    Act x{(bf16*)zbuf, B, h, w, 32};
    CAR_TRY(conv_fwd(st, m->post_quant, x, 0, 0, bA, nullptr, h, w));
    x = Act{bA, B, h, w, d.z_channels};
    bf16* cur = bB;
    auto flip = [&](bf16* used) { return used == bA ? bB : bA; };
    CAR_TRY(conv_fwd(st, m->d_conv_in, x, 0, 0, cur, nullptr, h, w));
    x = Act{cur, B, h, w, m->d_conv_in.cout};
    CAR_TRY(res_fwd(st, m->d_mid0, x, flip(x.p), s));
    CAR_TRY(attn_fwd(st, m->d_mid1, x, flip(x.p), s));
    CAR_TRY(res_fwd(st, m->d_mid2, x, flip(x.p), s));
    for (int idx = 0; idx < d.n_levels; ++idx) {
        for (size_t b = 0; b < m->d_res[idx].size(); ++b) {
            CAR_TRY(res_fwd(st, m->d_res[idx][b], x, flip(x.p), s));
            if (!m->d_attn[idx].empty()) CAR_TRY(attn_fwd(st, m->d_attn[idx][b], x, flip(x.p), s));
        }
        if (m->d_has_up[idx]) {   // Upsample (vq_model.py:368-379): nearest x2, then the 3x3 convolution
            bf16* o = flip(x.p);
            if (tc5_on() && x.C % T5_BK == 0) {   // materialise the up-sampled tensor (bandwidth-trivial) so the convolution is a plain TMA box walk
                CAR_LAUNCH(upsample2x_nhwc_kernel, gsz((long long)B * x.H * 2 * x.W * 2 * (x.C / 8)), 256, 0, st, (const bf16*)x.p, s.t2, B, x.H, x.W, x.C);
                Act u{s.t2, B, x.H * 2, x.W * 2, x.C};
                CAR_TRY(conv_fwd(st, m->d_up[idx], u, 0, 0, o, nullptr, u.H, u.W));
            } else {                              // nearest x2 folded into the conv's addressing (mma.sync kernel)
                CAR_TRY(conv_fwd(st, m->d_up[idx], x, 1, 0, o, nullptr, x.H * 2, x.W * 2));
            }
            x = Act{o, B, x.H * 2, x.W * 2, m->d_up[idx].cout};
        }
    }
    CAR_TRY(gn_fwd(st, m->d_norm_out, x, s.t0, 1, s.stats));
    Act y{s.t0, B, x.H, x.W, x.C};
    if (m->d_conv_out.cout == 3 && m->d_conv_out.cin_pad == y.C && y.C % 8 == 0 && 27 * y.C * 2 <= 48 * 1024) {
        CAR_LAUNCH(conv3x3_to3_kernel, gsz((long long)B * y.H * y.W), 256, (size_t)27 * y.C * 2, st, (const bf16*)y.p, (const bf16*)m->d_conv_out.w,
                   (const bf16*)m->d_conv_out.b, out, B, y.H, y.W, y.C);
        return CAR_OK;
    }
    return conv_fwd(st, m->d_conv_out, y, 0, 0, nullptr, nullptr, x.H, x.W, out);
}

extern "C" int car_vq_decode_code(CarVQ* m, const int32_t* codes, int32_t B, int32_t h, int32_t w, float* out, void* stream) {
    if (!codes) CAR_FAIL(CAR_ERR_ARG, "null codes");
    return vq_decode_impl(m, codes, nullptr, B, h, w, out, stream);
}
extern "C" int car_vq_decode(CarVQ* m, const float* quant, int32_t B, int32_t h, int32_t w, float* out, void* stream) {
    if (!quant) CAR_FAIL(CAR_ERR_ARG, "null quant");
    return vq_decode_impl(m, nullptr, quant, B, h, w, out, stream);
}

// ---- VQModel.encode (vq_model.py:41-46) at fp32 grade: the "x3" split-bf16 path of vision.cuh ----
// fp32 NHWC activations; every convolution = one launch of the bf16 implicit-GEMM kernel over tripled K, fp32 output / bias / residual.
struct ActF { float* p; int B, H, W, C; long long npix() const { return (long long)B * H * W; } long long n() const { return npix() * C; } };
struct EncScratch { bf16 *t3, *x3; float *h1, *sc; float* stats; float *qf, *kf, *vf, *S, *P, *ctx; bf16 *q3, *k3, *P3, *vT3; };

static int split3(cudaStream_t st, const float* x, bf16* y, long long npix, int C, int bside = 0) {
    CAR_LAUNCH(split3_kernel, gsz(npix * C), 256, 0, st, x, y, npix, C, bside);
    return CAR_OK;
}
// a3: S3 activations [B][Hs][Ws][3 cin_pad]; out fp32 [B][Ho][Wo][cout] (+ fp32 residual of the same shape)
static int conv_x3(cudaStream_t st, const ConvW& c, const bf16* a3, int B, int Hs, int Ws, int stride2, float* out, const float* resid, int Ho, int Wo,
                   int act = ACT_NONE) {
    if (!c.w3 || !c.bf) CAR_FAIL(CAR_ERR_STATE, "convolution has no split-bf16 weights (encoder layers only)");
    DenseP p;
    memset(&p, 0, sizeof(p));
    p.A = a3; p.B = c.w3; p.M = B * Ho * Wo; p.N = c.cout; p.K = c.k * c.k * 3 * c.cin_pad; p.ldb = p.K; p.alpha = 1.f;
    if (c.k == 1) { p.amode = A_PLAIN; p.lda = 3 * c.cin_pad; }
    else { p.amode = stride2 ? A_CONV3x3S2 : A_CONV3x3; p.Hs = Hs; p.Ws = Ws; p.Cin = 3 * c.cin_pad; p.ups = 0; }
    p.Ho = Ho; p.Wo = Wo;
    p.bias_f = c.bf; p.C = out; p.ldc = c.cout; p.out_mode = 1; p.resid_f = resid; p.ldr = c.cout; p.act = act;
    return dense(st, p);
}
static int gn_x3(cudaStream_t st, const NormW& nw, const ActF& x, bf16* y3, int swish, float* stats) {
    const int G = 32;
    CAR_LAUNCH(groupnorm_stats_f32_kernel, x.B * G, 512, 0, st, (const float*)x.p, stats, x.H * x.W, x.C, G);
    CAR_LAUNCH(groupnorm_apply_split3_kernel, gsz(x.n()), 256, 0, st, (const float*)x.p, (const float*)stats, (const float*)nw.wf, (const float*)nw.bff, y3,
               x.n(), x.H * x.W, x.C, G, swish);
    return CAR_OK;
}
// ResnetBlock.forward (vq_model.py:300-315)
static int res_x3(cudaStream_t st, const ResW& r, ActF& x, float* out, EncScratch& s) {
    CAR_TRY(gn_x3(st, r.n1, x, s.t3, 1, s.stats));
    CAR_TRY(conv_x3(st, r.c1, s.t3, x.B, x.H, x.W, 0, s.h1, nullptr, x.H, x.W));
    ActF h{s.h1, x.B, x.H, x.W, r.c1.cout};
    CAR_TRY(gn_x3(st, r.n2, h, s.t3, 1, s.stats));
    const float* sc = x.p;
    if (r.has_nin) {
        CAR_TRY(split3(st, x.p, s.x3, x.npix(), x.C));
        CAR_TRY(conv_x3(st, r.nin, s.x3, x.B, x.H, x.W, 0, s.sc, nullptr, x.H, x.W));
        sc = s.sc;
    }
    CAR_TRY(conv_x3(st, r.c2, s.t3, x.B, x.H, x.W, 0, out, sc, x.H, x.W));
    x.p = out; x.C = r.c2.cout;
    return CAR_OK;
}
// AttnBlock.forward (vq_model.py:328-352): single head over H*W tokens, scale C^-0.5, everything fp32-grade
static int attn_x3(cudaStream_t st, const AttnW& a, ActF& x, float* out, EncScratch& s) {
    const int C = x.C, hw = x.H * x.W, B = x.B;
    const int hwp = (hw + 31) & ~31;
    const long long rows = (long long)B * hw;
    CAR_TRY(gn_x3(st, a.n, x, s.t3, 0, s.stats));
    CAR_TRY(conv_x3(st, a.q, s.t3, B, x.H, x.W, 0, s.qf, nullptr, x.H, x.W));
    CAR_TRY(conv_x3(st, a.k, s.t3, B, x.H, x.W, 0, s.kf, nullptr, x.H, x.W));
    CAR_TRY(conv_x3(st, a.v, s.t3, B, x.H, x.W, 0, s.vf, nullptr, x.H, x.W));
    CAR_TRY(split3(st, s.qf, s.q3, rows, C, 0));
    CAR_TRY(split3(st, s.kf, s.k3, rows, C, 1));
    {   // scores [B][hw][hwp] fp32
        DenseP p = dp_plain(s.q3, 3 * C, s.k3, 3 * C, hw, hw, 3 * C, s.S, hwp);
        p.sA = (long long)hw * 3 * C; p.sB = (long long)hw * 3 * C; p.sC = (long long)hw * hwp; p.alpha = 1.0f / sqrtf((float)C); p.out_mode = 1;
        CAR_TRY(dense(st, p, B));
    }
    CAR_LAUNCH(softmax_rows_f32_kernel, (unsigned)rows, 256, 0, st, (const float*)s.S, s.P, hw, hwp);
    CAR_TRY(split3(st, s.P, s.P3, rows, hwp, 0));
    CAR_LAUNCH(transpose_split3b_kernel, gsz((long long)B * C * hwp), 256, 0, st, (const float*)s.vf, s.vT3, B, hw, hwp, C);
    {   // context [B][hw][C] fp32
        DenseP p = dp_plain(s.P3, 3 * hwp, s.vT3, 3 * hwp, hw, C, 3 * hwp, s.ctx, C);
        p.sA = (long long)hw * 3 * hwp; p.sB = (long long)C * 3 * hwp; p.sC = (long long)hw * C; p.out_mode = 1;
        CAR_TRY(dense(st, p, B));
    }
    CAR_TRY(split3(st, s.ctx, s.q3, rows, C, 0));              // (q3 is free again)
    CAR_TRY(conv_x3(st, a.o, s.q3, B, x.H, x.W, 0, out, x.p, x.H, x.W));
    x.p = out;
    return CAR_OK;
}

// VQModel.encode (vq_model.py:41-46): image fp32 NCHW [B][3][H][W] -> indices int32 [B*h*w] (+ quant fp32 [B][e][h][w])
extern "C" int car_vq_encode(CarVQ* m, const float* img, int32_t B, int32_t H, int32_t W, int32_t* idx_out, float* quant_out,
                             void* stream) {
    if (!m || !img || !idx_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const CarVQDesc& d = m->d;
    const int down = 1 << (d.n_levels - 1);
    if (H % down || W % down) CAR_FAIL(CAR_ERR_ARG, "image size must be a multiple of the down-sampling factor");
    const int h = H / down, w = W / down;
    const size_t npix = (size_t)B * h * w;
    const int hw = h * w, hwp = (hw + 31) & ~31, Cmax = d.ch * d.ch_mult[d.n_levels - 1];
    // workspace: fp32 activations (largest: ch channels at full resolution) and their S3 forms
    const size_t actf = (size_t)B * H * W * d.ch * 4, act3 = (size_t)B * H * W * d.ch * 3 * 2;
    const size_t x0 = (size_t)B * H * W * 32 * 3 * 2;
    const size_t af = npix * Cmax * 4, a3 = npix * Cmax * 3 * 2;
    const size_t need = 4 * (actf + 256) + 2 * (act3 + 256) + x0 + 256 + (size_t)B * 32 * 2 * 4 + 256 + 4 * (af + 256) + 2 * (a3 + 256) +
                        2 * ((size_t)B * hw * hwp * 4 + 256) + (size_t)B * hw * hwp * 3 * 2 + 256 + (size_t)B * Cmax * hwp * 3 * 2 + 256 + npix * 8 * 4 * 2 + 1024;
    CAR_TRY(m->ws.reserve(need));
    m->ws.reset();
    float* fA = (float*)m->ws.take(actf); float* fB = (float*)m->ws.take(actf);
    EncScratch s;
    s.h1 = (float*)m->ws.take(actf); s.sc = (float*)m->ws.take(actf);
    s.t3 = (bf16*)m->ws.take(act3); s.x3 = (bf16*)m->ws.take(act3);
    bf16* img3 = (bf16*)m->ws.take(x0);
    s.stats = (float*)m->ws.take((size_t)B * 32 * 2 * 4);
    s.qf = (float*)m->ws.take(af); s.kf = (float*)m->ws.take(af); s.vf = (float*)m->ws.take(af); s.ctx = (float*)m->ws.take(af);
    s.q3 = (bf16*)m->ws.take(a3); s.k3 = (bf16*)m->ws.take(a3);
    s.S = (float*)m->ws.take((size_t)B * hw * hwp * 4); s.P = (float*)m->ws.take((size_t)B * hw * hwp * 4);
    s.P3 = (bf16*)m->ws.take((size_t)B * hw * hwp * 3 * 2); s.vT3 = (bf16*)m->ws.take((size_t)B * Cmax * hwp * 3 * 2);
    float* zf = (float*)m->ws.take(npix * 8 * 4); float* zq = (float*)m->ws.take(npix * 8 * 4);

    CAR_LAUNCH(nchw_to_nhwc_split3_kernel, gsz((long long)B * H * W * 32), 256, 0, st, img, img3, B, 3, H * W, 32);
    auto flip = [&](float* used) { return used == fA ? fB : fA; };
    CAR_TRY(conv_x3(st, m->e_conv_in, img3, B, H, W, 0, fA, nullptr, H, W));
    ActF x{fA, B, H, W, m->e_conv_in.cout};
    for (int lvl = 0; lvl < d.n_levels; ++lvl) {
        for (size_t b = 0; b < m->e_res[lvl].size(); ++b) {
            CAR_TRY(res_x3(st, m->e_res[lvl][b], x, flip(x.p), s));
            if (!m->e_attn[lvl].empty()) CAR_TRY(attn_x3(st, m->e_attn[lvl][b], x, flip(x.p), s));
        }
        if (m->e_has_down[lvl]) {   // Downsample (vq_model.py:382-397): pad (0,1,0,1), 3x3 stride 2
            float* o = flip(x.p);
            CAR_TRY(split3(st, x.p, s.x3, x.npix(), x.C));
            CAR_TRY(conv_x3(st, m->e_down[lvl], s.x3, B, x.H, x.W, 1, o, nullptr, x.H / 2, x.W / 2));
            x = ActF{o, B, x.H / 2, x.W / 2, m->e_down[lvl].cout};
        }
    }
    CAR_TRY(res_x3(st, m->e_mid0, x, flip(x.p), s));
    CAR_TRY(attn_x3(st, m->e_mid1, x, flip(x.p), s));
    CAR_TRY(res_x3(st, m->e_mid2, x, flip(x.p), s));
    CAR_TRY(gn_x3(st, m->e_norm_out, x, s.t3, 1, s.stats));
    float* zc = flip(x.p);
    CAR_TRY(conv_x3(st, m->e_conv_out, s.t3, B, x.H, x.W, 0, zc, nullptr, x.H, x.W));              // [npix][z_channels]
    CAR_TRY(split3(st, zc, s.x3, (long long)npix, d.z_channels));
    CAR_TRY(conv_x3(st, m->quant_conv, s.x3, B, x.H, x.W, 0, zf, nullptr, x.H, x.W));              // [npix][embed_dim] fp32
    CAR_LAUNCH(vq_argmin_kernel, (unsigned)((npix + 127) / 128), 128, 0, st, (const float*)zf, m->codebook_n, idx_out, quant_out ? zq : nullptr, (long long)npix, d.embed_dim, d.codebook_size);
    if (quant_out) CAR_LAUNCH(nhwc_to_nchw_f32_kernel, gsz((long long)npix * d.embed_dim), 256, 0, st, zq, quant_out, B, h * w, d.embed_dim);
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// antialiased bilinear resize (row f2): width pass into tmp, height pass into out
// ---------------------------------------------------------------------------------------------------------
extern "C" int car_resize_bilinear_aa(const float* in, int32_t B, int32_t Cc, int32_t H, int32_t W, float* out, int32_t OH, int32_t OW,
                                      float* tmp, void* stream) {
    if (!in || !out || !tmp) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (B <= 0 || Cc <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) CAR_FAIL(CAR_ERR_ARG, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n1 = (long long)B * Cc * H * OW, n2 = (long long)B * Cc * OH * OW;
    CAR_LAUNCH(resize_aa_axis_kernel, (int)std::min<long long>((n1 + 255) / 256, 148 * 16), 256, 0, st, in, tmp, (long long)B * Cc * H, W, OW, 1);
    CAR_LAUNCH(resize_aa_axis_kernel, (int)std::min<long long>((n2 + 255) / 256, 148 * 16), 256, 0, st, (const float*)tmp, out, (long long)B * Cc, H, OH, OW);
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// control-map / prompt front-end (row f3): Canny edges (condition/canny.py:14) and caption left-padding (sample_t2i.py:146-156)
// ---------------------------------------------------------------------------------------------------------
extern "C" int64_t car_canny_workspace_bytes(int32_t H, int32_t W) {
    const int64_t n = (int64_t)H * W;
    return ((n * 2 + 255) / 256 * 256) * 3 + (n + 255) / 256 * 256 + 256;     // mag u16, xs i16, ys i16, map u8, changed flag
}
// img uint8 [H][W][C] (device) -> edges uint8 [H][W] in {0, 255}.  restart != 0: gradients + non-maximum suppression + thresholds,
// then `sweeps` hysteresis sweeps; restart == 0: `sweeps` more sweeps on the map left in `work`.  *changed_dev (int32 in device
// memory) is 1 afterwards iff the LAST sweep still grew an edge — the caller repeats with restart = 0 until it reads 0
// (the reference's cv2 call is synchronous host code; here only that convergence check needs the host).
extern "C" int car_canny_u8(const uint8_t* img, int32_t H, int32_t W, int32_t C, int32_t low, int32_t high, uint8_t* edges_out, void* work,
                            int32_t sweeps, int32_t restart, int32_t* changed_dev, void* stream) {
    if (!img || !edges_out || !work || !changed_dev) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (H <= 0 || W <= 0 || C <= 0 || C > 4 || sweeps < 1) CAR_FAIL(CAR_ERR_ARG, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)H * W;
    const size_t a2 = ((size_t)n * 2 + 255) / 256 * 256;
    unsigned short* mag = (unsigned short*)work;
    short* xs = (short*)((char*)work + a2);
    short* ys = (short*)((char*)work + 2 * a2);
    unsigned char* map = (unsigned char*)work + 3 * a2;
    if (low > high) std::swap(low, high);
    if (restart) {
        CAR_LAUNCH(canny_grad_kernel, gsz(n), 256, 0, st, img, H, W, C, mag, xs, ys);
        CAR_LAUNCH(canny_nms_kernel, gsz(n), 256, 0, st, (const unsigned short*)mag, (const short*)xs, (const short*)ys, H, W, low, high, map);
    }
    dim3 grid((W + CH_T - 1) / CH_T, (H + CH_T - 1) / CH_T);
    for (int s = 0; s < sweeps; ++s) {
        CAR_CUDA(cudaMemsetAsync(changed_dev, 0, 4, st));
        CAR_LAUNCH(canny_hyst_kernel, grid, CH_T * CH_T / 4, 0, st, map, H, W, changed_dev);
    }
    CAR_LAUNCH(canny_finish_kernel, gsz(n), 256, 0, st, (const unsigned char*)map, edges_out, n);
    return CAR_OK;
}

// caption_embs [B][L][row_bytes] (any dtype, row_bytes % 16 == 0), emb_masks int64 [B][L] -> rotated embeddings + flipped masks
extern "C" int car_left_pad_captions(const void* embs, const int64_t* masks, int32_t B, int32_t L, int32_t row_bytes, void* embs_out,
                                     int64_t* masks_out, void* stream) {
    if (!embs || !masks || !embs_out || !masks_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (B <= 0 || L <= 0 || row_bytes <= 0 || row_bytes % 16) CAR_FAIL(CAR_ERR_ARG, "row size must be a positive multiple of 16 bytes");
    if (((uintptr_t)embs % 16) || ((uintptr_t)embs_out % 16)) CAR_FAIL(CAR_ERR_ARG, "embeddings must be 16-byte aligned");
    CAR_LAUNCH(left_pad_pack_kernel, B * L, 128, 0, (cudaStream_t)stream, (const uint4*)embs, (const long long*)masks, (uint4*)embs_out,
               (long long*)masks_out, L, row_bytes / 16);
    return CAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// HED soft-edge detector (row f3): condition/hed.py:17-84 — 13 ReLU 3x3 convolutions in five blocks with 2x2 max-pooling between
// them, a 1x1 projection per block, bilinear resize of the five maps, mean, sigmoid.  fp32 in the reference => fp32-grade here.
// ---------------------------------------------------------------------------------------------------------
struct CarHED {
    std::vector<void*> owned;
    std::vector<ConvW> conv;           // 13, in forward order
    float* norm;                       // [3]
    float* pw[5]; float* pb[5];        // projection weights [C] / bias [1]
    Arena ws;
};
static const int HED_BLK[5][3] = {{3, 64, 2}, {64, 128, 2}, {128, 256, 3}, {256, 512, 3}, {512, 512, 3}};

// tensors (fp32, device), in order: norm [3]; per block: conv0.weight, conv0.bias, conv1.weight, ... , projection.weight [C], projection.bias [1]
extern "C" int car_hed_create(const void* const* tensors, int32_t n_tensors, void* stream, CarHED** out) {
    if (!tensors || !out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (n_tensors != 1 + 2 * 13 + 2 * 5) CAR_FAIL(CAR_ERR_ARG, "HED expects 37 tensors (norm, 13 x (weight, bias), 5 x (projection weight, bias))");
    cudaStream_t st = (cudaStream_t)stream;
    CarHED* m = new CarHED();
    int rc = CAR_OK, ti = 0;
    auto keep = [&](const void* src, long long n, float** dst) {
        if (rc != CAR_OK) return;
        if (cudaMalloc((void**)dst, (size_t)n * 4) != cudaSuccess) { rc = CAR_ERR_CUDA; g_car_err = "car_hed_create: cudaMalloc failed"; return; }
        m->owned.push_back(*dst);
        if (cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = CAR_ERR_CUDA; g_car_err = "car_hed_create: copy failed"; }
    };
    keep(tensors[ti++], 3, &m->norm);
    for (int b = 0; b < 5 && rc == CAR_OK; ++b) {
        int cin = HED_BLK[b][0];
        const int cout = HED_BLK[b][1];
        for (int i = 0; i < HED_BLK[b][2] && rc == CAR_OK; ++i) {
            ConvW c;
            memset(&c, 0, sizeof(c));
            c.cin = cin; c.cout = cout; c.k = 3; c.cin_pad = (cin + 31) & ~31;
            const long long n3 = (long long)cout * 9 * c.cin_pad;
            if (cudaMalloc((void**)&c.w3, (size_t)n3 * 3 * 2) != cudaSuccess) { rc = CAR_ERR_CUDA; g_car_err = "car_hed_create: cudaMalloc failed"; break; }
            m->owned.push_back(c.w3);
            conv_weight_pack_x3_kernel<<<gsz(n3), 256, 0, st>>>((const float*)tensors[ti], c.w3, cout, cin, 3, 3, c.cin_pad);
            keep(tensors[ti + 1], cout, &c.bf);
            ti += 2;
            m->conv.push_back(c);
            cin = cout;
        }
        keep(tensors[ti], cout, &m->pw[b]); keep(tensors[ti + 1], 1, &m->pb[b]);
        ti += 2;
    }
    if (rc == CAR_OK && cudaGetLastError() != cudaSuccess) { rc = CAR_ERR_CUDA; g_car_err = "car_hed_create: weight packing failed"; }
    if (rc != CAR_OK) { for (void* p : m->owned) cudaFree(p); delete m; return rc; }
    *out = m;
    return CAR_OK;
}
extern "C" int car_hed_destroy(CarHED* m) {
    if (!m) return CAR_OK;
    for (void* p : m->owned) cudaFree(p);
    m->ws.release();
    delete m;
    return CAR_OK;
}
// image fp32 NCHW [B][3][H][W] (values 0..255) -> edge fp32 [B][H][W] in [0, 255]; proj_out (optional): the five projection maps
// back to back, map k = [B][hk][wk] with hk = H >> k (floor), wk = W >> k
extern "C" int car_hed_forward(CarHED* m, const float* img, int32_t B, int32_t H, int32_t W, float* edge_out, float* proj_out, void* stream) {
    if (!m || !img || !edge_out) CAR_FAIL(CAR_ERR_ARG, "null argument");
    if (B <= 0 || H < 16 || W < 16) CAR_FAIL(CAR_ERR_ARG, "image must be at least 16 x 16");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t full = (size_t)B * H * W;
    size_t proj_elems = 0;
    { int h = H, w = W; for (int k = 0; k < 5; ++k) { proj_elems += (size_t)B * h * w; h /= 2; w /= 2; } }
    const size_t actf = full * 64 * 4, act3 = full * 64 * 3 * 2;                 // largest activation: 64 channels at full resolution
    CAR_TRY(m->ws.reserve(2 * (actf + 256) + act3 + 256 + full * 32 * 3 * 2 + 256 + proj_elems * 4 + 256));
    m->ws.reset();
    float* fA = (float*)m->ws.take(actf); float* fB = (float*)m->ws.take(actf);
    bf16* s3 = (bf16*)m->ws.take(act3);
    bf16* img3 = (bf16*)m->ws.take(full * 32 * 3 * 2);
    float* maps = (float*)m->ws.take(proj_elems * 4);
    CAR_LAUNCH(hed_input_split3_kernel, gsz((long long)full * 32), 256, 0, st, img, (const float*)m->norm, img3, B, 3, H * W, 32);
    HedMaps hm;
    int h = H, w = W, ci = 0, curC = 3;
    float* cur = nullptr;                                                       // the one live fp32 activation (NHWC); ping-pong fA / fB
    auto other = [&](float* p) { return p == fA ? fB : fA; };
    size_t moff = 0;
    for (int b = 0; b < 5; ++b) {
        if (b > 0) {                                                            // down_sampling=True (:28-30)
            float* o = other(cur);
            CAR_LAUNCH(maxpool2_nhwc_f32_kernel, gsz((long long)B * (h / 2) * (w / 2) * curC), 256, 0, st, (const float*)cur, o, B, h, w, curC);
            h /= 2; w /= 2; cur = o;
        }
        for (int i = 0; i < HED_BLK[b][2]; ++i, ++ci) {
            const ConvW& c = m->conv[ci];
            const bf16* a3 = img3;
            if (ci > 0) { CAR_TRY(split3(st, cur, s3, (long long)B * h * w, curC)); a3 = s3; }
            float* o = other(cur);
            CAR_TRY(conv_x3(st, c, a3, B, h, w, 0, o, nullptr, h, w, ACT_RELU));  // conv + bias, ReLU (:31-33)
            cur = o; curC = c.cout;
        }
        float* mp = maps + moff;
        CAR_LAUNCH(hed_proj_kernel, (unsigned)(((long long)B * h * w * 32 + 255) / 256), 256, 0, st, (const float*)cur, (const float*)m->pw[b], (const float*)m->pb[b], mp,
                   (long long)B * h * w, curC);
        hm.p[b] = mp; hm.h[b] = h; hm.w[b] = w;
        moff += (size_t)B * h * w;
    }
    CAR_LAUNCH(hed_merge_kernel, gsz((long long)full), 256, 0, st, hm, B, H, W, edge_out);
    if (proj_out) CAR_CUDA(cudaMemcpyAsync(proj_out, maps, proj_elems * 4, cudaMemcpyDeviceToDevice, st));
    return CAR_OK;
}
