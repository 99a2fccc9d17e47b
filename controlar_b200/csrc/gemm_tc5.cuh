// gemm_tc5.cuh — dense bf16 GEMM on the 5th-generation tensor cores: TMA tensor-map loads (128-byte swizzle) -> tcgen05.mma with
// the accumulator in TMEM -> tcgen05.ld epilogue, warp-specialised and persistent.  C[M,N] = epi(A[M,K] · B[N,K]^T), fp32 accumulate.
// Used by the prefill (autoregressive/models/gpt_t2i.py:433-470: wqkv / wo / w1 / w3 / w2 over B_eff·T rows) and the control-token
// MLPs (gpt_t2i.py:165-181 over B_eff·N rows).  lda, ldb % 8 == 0 (16-byte row pitch for the tensor map); M, N, K tails are
// zero-filled by TMA on load and masked on store (N % 8 == 0).
//
// Roles (192 threads, one CTA per SM, static tile schedule t = blockIdx.x + i · gridDim.x over 128 x 128 output tiles):
//   warp 0, one lane : TMA producer — per 64-wide k-block two cp.async.bulk.tensor.2d (A 128 x 64, B 128 x 64, SWIZZLE_128B) into
//                      a 6-stage ring (32 KB per stage), completion on full[stage] (expect_tx), slot reuse on empty[stage]
//   warp 1, one lane : MMA issuer — 4 x tcgen05.mma.cta_group::1.kind::f16 (m128 n128 k16) per k-block through shared-memory
//                      descriptors (K-major, 128-byte swizzle, SBO = 1024 B), tcgen05.commit -> empty[stage]; after the last
//                      k-block of a tile tcgen05.commit -> acc_full[a].  Two accumulator stages (2 x 128 TMEM columns): the
//                      epilogue of tile i overlaps the main loop of tile i + 1.
//   warps 2..5       : epilogue — warp w reads TMEM lanes [32 (w % 4), +32) with tcgen05.ld.32x32b.x32, applies
//                      bf16 round / GELU-tanh / residual (same rounding points as gemm_dense.cuh), stores 64 contiguous bytes per
//                      thread and chunk, then releases the accumulator stage (acc_empty[a]).
// Descriptor encodings: instruction descriptor and version bit validated by scripts/tc5_probe.cu (r1); the swizzled
// shared-memory descriptor follows the sm_100 layout (start >> 4 | LBO | SBO >> 4 << 32 | version 1 << 46 | layout 2 << 61).
#pragma once
#include "common.cuh"
#include <cuda.h>

constexpr int T5_BM = 128, T5_BN = 128, T5_BK = 64, T5_STAGES = 6, T5_THREADS = 192, T5_ACC = 2;
constexpr int T5_TILE_BYTES = T5_BM * T5_BK * 2;                               // 16 KB per operand and stage
constexpr int T5_STAGE_BYTES = 2 * T5_TILE_BYTES;
constexpr int T5_SMEM = T5_STAGES * T5_STAGE_BYTES + 1024;                     // + 1024-byte alignment slack (swizzle atoms)

struct Tc5P {
    int M, N, K;
    const bf16* resid; int ldr;
    bf16* C; int ldc;
    int act;            // 0 none, 1 GELU-tanh, 2 exact (erf) GELU
    const bf16* bias;   // per output column, added to the fp32 accumulator before the bf16 rounding (nn.Linear / Conv2d bias)
    const bf16* scale;  // per output column, applied after the activation: r(r(v) * scale) (DINOv2 LayerScale)
    // 3x3 / pad 1 / stride 1 convolution over an NHWC tensor as an implicit GEMM (conv = 1): the A tile of output-pixel block
    // (image n, rows 8 ty .., columns 16 tx ..) and k-block (tap, 64-channel block) is ONE 4-D TMA box {64 ch, 16 x, 8 y, 1 n} at
    // (c0, 16 tx + kx - 1, 8 ty + ky - 1, n): out-of-bounds pixels (the padding) are zero-filled by TMA, and the box lands in shared
    // memory as 128 rows x 128 bytes — exactly the K-major SWIZZLE_128B operand tile of the plain GEMM.
    int conv, H, W, tiles_x, tiles_y, cblks;
};
constexpr int T5_TW = 16, T5_TH = 8;                                            // output-pixel block of a conv tile (T5_TW * T5_TH = T5_BM)

__device__ __forceinline__ uint32_t t5_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t5_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void t5_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();            // never hang the box
    }
}
__device__ __forceinline__ void t5_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t5_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void t5_tma_2d(uint32_t sdst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sdst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void t5_tma_4d(uint32_t sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(sdst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// K-major operand tile [128 rows][64 bf16] written by TMA with SWIZZLE_128B: 8-row groups are 1024-byte atoms
__device__ __forceinline__ uint64_t t5_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

static __global__ void __launch_bounds__(T5_THREADS, 1) gemm_tc5_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                        const Tc5P p) {
    extern __shared__ unsigned char t5_raw[];
    __shared__ __align__(8) uint64_t bar_full[T5_STAGES], bar_empty[T5_STAGES], bar_acc_full[T5_ACC], bar_acc_empty[T5_ACC];
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (t5_smem(t5_raw) + 1023u) & ~1023u;                 // stage s: A at smem0 + s * 32 KB, B 16 KB after
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_xy = p.tiles_x * p.tiles_y;                                 // (conv) pixel blocks per image
    const int tiles_m = p.conv ? (p.M / (p.H * p.W)) * tiles_xy : (p.M + T5_BM - 1) / T5_BM, tiles_n = (p.N + T5_BN - 1) / T5_BN;
    const int ntiles = tiles_m * tiles_n;
    const int nkb = (p.K + T5_BK - 1) / T5_BK;

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) { t5_mbar_init(t5_smem(&bar_full[s]), 1); t5_mbar_init(t5_smem(&bar_empty[s]), 1); }
        for (int a = 0; a < T5_ACC; ++a) { t5_mbar_init(t5_smem(&bar_acc_full[a]), 1); t5_mbar_init(t5_smem(&bar_acc_empty[a]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(t5_smem(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // tile t -> (m-tile, n-tile): n fastest inside groups of 8 m-tiles... plain n-major walk keeps the B (weight) tile hot in L2
    // for the CTAs that run the same n-tile at the same time
    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tm = t % tiles_m, tn = t / tiles_m;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % T5_STAGES, use = it / T5_STAGES;
                    if (use > 0) t5_mbar_wait(t5_smem(&bar_empty[s]), (use - 1) & 1);
                    const uint32_t sA = smem0 + s * T5_STAGE_BYTES, sB = sA + T5_TILE_BYTES, fb = t5_smem(&bar_full[s]);
                    t5_mbar_expect(fb, T5_STAGE_BYTES);
                    if (p.conv) {
                        const int n_img = tm / tiles_xy, r = tm - n_img * tiles_xy, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                        const int tap = kb / p.cblks, cb = kb - tap * p.cblks, ky = tap / 3, kx = tap - 3 * ky;
                        t5_tma_4d(sA, &mapA, cb * T5_BK, tx * T5_TW + kx - 1, ty * T5_TH + ky - 1, n_img, fb);
                    } else {
                        t5_tma_2d(sA, &mapA, kb * T5_BK, tm * T5_BM, fb);
                    }
                    t5_tma_2d(sB, &mapB, kb * T5_BK, tn * T5_BN, fb);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T5_BN >> 3) << 17) | ((uint32_t)(T5_BM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
                const uint32_t a = ti % T5_ACC, ause = ti / T5_ACC;
                if (ause > 0) t5_mbar_wait(t5_smem(&bar_acc_empty[a]), (ause - 1) & 1);        // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + a * T5_BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % T5_STAGES, use = it / T5_STAGES;
                    t5_mbar_wait(t5_smem(&bar_full[s]), use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sA = smem0 + s * T5_STAGE_BYTES, sB = sA + T5_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < T5_BK / 16; ++kk) {
                        const uint64_t da = t5_desc_sw128(sA + kk * 32), db = t5_desc_sw128(sB + kk * 32);
                        const uint32_t accf = (kb > 0 || kk > 0) ? 1u : 0u;
                        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                     ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
                    }
                    // commit: arrives on the barrier when the MMAs issued so far have finished reading shared memory / writing TMEM
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t5_smem(&bar_empty[s])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t5_smem(&bar_acc_full[a])) : "memory");
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
            const int tm = t % tiles_m, tn = t / tiles_m;
            const uint32_t a = ti % T5_ACC, ause = ti / T5_ACC;
            t5_mbar_wait(t5_smem(&bar_acc_full[a]), ause & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int row = tm * T5_BM + quad * 32 + lane;                            // output row (plain) / NHWC pixel index (conv)
            bool row_ok = row < p.M;
            if (p.conv) {
                const int n_img = tm / tiles_xy, r = tm - n_img * tiles_xy, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                const int rr = quad * 32 + lane, py = ty * T5_TH + (rr >> 4), px = tx * T5_TW + (rr & 15);
                row_ok = py < p.H && px < p.W;
                row = (n_img * p.H + py) * p.W + px;
            }
#pragma unroll 1
            for (int cc = 0; cc < T5_BN / 32; ++cc) {
                const int c0 = cc * 32;
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * T5_BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                    "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                      "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == T5_BN / 32 - 1) {                       // every column of this accumulator stage is in registers: release it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) t5_mbar_arrive(t5_smem(&bar_acc_empty[a]));
                }
                if (row_ok) {
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const int n = tn * T5_BN + c0 + j8 * 8;
                        if (n < p.N) {                                     // N % 8 == 0 (host-checked): whole 16-byte groups
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float a = __uint_as_float(v[j8 * 8 + j]);
                                if (p.bias) a += tof(p.bias[n + j]);
                                f[j] = rnd<bf16>(a);
                                if (p.act == 1) f[j] = rnd<bf16>(gelu_tanh_f(f[j]));
                                else if (p.act == 2) f[j] = rnd<bf16>(gelu_erf_f(f[j]));
                                if (p.scale) f[j] = rnd<bf16>(f[j] * tof(p.scale[n + j]));
                            }
                            if (p.resid) {
                                const uint4 rv = *reinterpret_cast<const uint4*>(p.resid + (size_t)row * p.ldr + n);
                                const uint32_t ri[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float x, y;
                                    unpack_bf16x2(ri[q], x, y);
                                    f[2 * q] = rnd<bf16>(f[2 * q] + x); f[2 * q + 1] = rnd<bf16>(f[2 * q + 1] + y);
                                }
                            }
                            uint4 o;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                            o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
                            o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
                            *reinterpret_cast<uint4*>(p.C + (size_t)row * p.ldc + n) = o;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// =========================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one TPC) computes a 256 x 256 output tile.  Each CTA loads ITS half of both
// operands per 64-wide k-block — A rows [m0 + 128 rank, +128), B rows [n0 + 128 rank, +128) — and the leader CTA issues
// tcgen05.mma.cta_group::2 (UMMA M = 256, N = 256): every tensor core reads its own A half and BOTH B halves, so each SM does
// twice the math of the 1-CTA tile per byte it pulls from L2 (128 instead of 64 FLOP per operand byte: the 1-CTA kernel is bounded
// by L2 -> SM operand bandwidth, measured 543 TFLOP/s on M1920 N3584 K1280 against cuBLAS' 1019).  Accumulators: 128 lanes x 256
// fp32 columns per CTA and stage, two stages = all 512 TMEM columns.  Barriers live in the leader's shared memory for the
// TMA -> MMA direction (both CTAs' TMA loads complete_tx on the leader's full[s], address with the peer bit cleared) and are
// multicast to both CTAs for the MMA -> producer / MMA -> epilogue direction (tcgen05.commit ... multicast::cluster).
// =========================================================================================================
constexpr int T2_BM = 256, T2_BN = 256;
__device__ __forceinline__ void t5_tma_2d_2cta(uint32_t sdst, const CUtensorMap* map, int c0, int c1, uint32_t bar_leader) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sdst), "l"(map), "r"(c0), "r"(c1), "r"(bar_leader) : "memory");
}
__device__ __forceinline__ void t5_commit_2cta(uint32_t bar) {       // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void t5_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T5_THREADS, 1)
gemm_tc5x2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Tc5P p) {
    extern __shared__ unsigned char t5_raw[];
    __shared__ __align__(8) uint64_t bar_full[T5_STAGES], bar_empty[T5_STAGES], bar_acc_full[T5_ACC], bar_acc_empty[T5_ACC];
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (t5_smem(t5_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int tiles_m = (p.M + T2_BM - 1) / T2_BM, tiles_n = (p.N + T2_BN - 1) / T2_BN;
    const int ntiles = tiles_m * tiles_n;
    const int nkb = (p.K + T5_BK - 1) / T5_BK;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) { t5_mbar_init(t5_smem(&bar_full[s]), 1); t5_mbar_init(t5_smem(&bar_empty[s]), 1); }
        for (int a = 0; a < T5_ACC; ++a) { t5_mbar_init(t5_smem(&bar_acc_full[a]), 1); t5_mbar_init(t5_smem(&bar_acc_empty[a]), 8); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(t5_smem(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    t5_cluster_sync();                                   // both CTAs' barriers are initialised before any remote arrive / TMA signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): own halves of A and B, completion on the LEADER's full barrier =====
            uint32_t it = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                const int tm = t % tiles_m, tn = t / tiles_m;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % T5_STAGES, use = it / T5_STAGES;
                    if (use > 0) t5_mbar_wait(t5_smem(&bar_empty[s]), (use - 1) & 1);
                    const uint32_t sA = smem0 + s * T5_STAGE_BYTES, sB = sA + T5_TILE_BYTES;
                    const uint32_t fb = t5_smem(&bar_full[s]);
                    if (leader) t5_mbar_expect(fb, 2 * T5_STAGE_BYTES);            // 4 tiles of 16 KB: two from each CTA
                    uint32_t fb_leader;                                            // the same barrier in the leader CTA (rank 0) of the pair
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(fb_leader) : "r"(fb), "r"(0));
                    t5_tma_2d_2cta(sA, &mapA, kb * T5_BK, tm * T2_BM + (int)rank * 128, fb_leader);
                    t5_tma_2d_2cta(sB, &mapB, kb * T5_BK, tn * T2_BN + (int)rank * 128, fb_leader);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===== MMA issuer (leader CTA only) =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters, ++ti) {
                const uint32_t a = ti % T5_ACC, ause = ti / T5_ACC;
                if (ause > 0) t5_mbar_wait(t5_smem(&bar_acc_empty[a]), (ause - 1) & 1);        // both CTAs' epilogues drained this stage
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + a * T2_BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % T5_STAGES, use = it / T5_STAGES;
                    t5_mbar_wait(t5_smem(&bar_full[s]), use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sA = smem0 + s * T5_STAGE_BYTES, sB = sA + T5_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < T5_BK / 16; ++kk) {
                        const uint64_t da = t5_desc_sw128(sA + kk * 32), db = t5_desc_sw128(sB + kk * 32);
                        const uint32_t accf = (kb > 0 || kk > 0) ? 1u : 0u;
                        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                                     ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
                    }
                    t5_commit_2cta(t5_smem(&bar_empty[s]));                        // frees stage s in BOTH CTAs
                }
                t5_commit_2cta(t5_smem(&bar_acc_full[a]));                         // wakes both CTAs' epilogue warps
            }
        }
    } else {
        // ===== epilogue warps 2..5 of both CTAs: this CTA's 128 rows x 256 columns =====
        const int quad = warp & 3;
        uint32_t acc_empty_leader[T5_ACC];
#pragma unroll
        for (int a = 0; a < T5_ACC; ++a)
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(acc_empty_leader[a]) : "r"(t5_smem(&bar_acc_empty[a])), "r"(0));
        uint32_t ti = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters, ++ti) {
            const int tm = t % tiles_m, tn = t / tiles_m;
            const uint32_t a = ti % T5_ACC, ause = ti / T5_ACC;
            t5_mbar_wait(t5_smem(&bar_acc_full[a]), ause & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = tm * T2_BM + (int)rank * 128 + quad * 32 + lane;
#pragma unroll 1
            for (int cc = 0; cc < T2_BN / 32; ++cc) {
                const int c0 = cc * 32;
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * T2_BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                    "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                      "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == T2_BN / 32 - 1) {                        // the whole accumulator stage is in registers: release it (to the leader)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_leader[a]) : "memory");
                }
                if (row < p.M) {
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const int n = tn * T2_BN + c0 + j8 * 8;
                        if (n < p.N) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float x = __uint_as_float(v[j8 * 8 + j]);
                                if (p.bias) x += tof(p.bias[n + j]);
                                f[j] = rnd<bf16>(x);
                                if (p.act == 1) f[j] = rnd<bf16>(gelu_tanh_f(f[j]));
                                else if (p.act == 2) f[j] = rnd<bf16>(gelu_erf_f(f[j]));
                                if (p.scale) f[j] = rnd<bf16>(f[j] * tof(p.scale[n + j]));
                            }
                            if (p.resid) {
                                const uint4 rv = *reinterpret_cast<const uint4*>(p.resid + (size_t)row * p.ldr + n);
                                const uint32_t ri[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float x, y;
                                    unpack_bf16x2(ri[q], x, y);
                                    f[2 * q] = rnd<bf16>(f[2 * q] + x); f[2 * q + 1] = rnd<bf16>(f[2 * q + 1] + y);
                                }
                            }
                            uint4 o;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                            o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
                            o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
                            *reinterpret_cast<uint4*>(p.C + (size_t)row * p.ldc + n) = o;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    t5_cluster_sync();                                   // no CTA frees TMEM / exits while its partner still reads its shared memory
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- host: tensor maps (driver entry point fetched through the runtime: the library does not link libcuda) ----
typedef CUresult (*t5_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static t5_encode_fn t5_encoder() {
    static t5_encode_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (t5_encode_fn)f;
    }();
    return fn;
}
// row-major bf16 matrix [rows][cols] with row pitch ld elements; box = 64 columns x 128 rows, 128-byte swizzle, zero fill out of bounds
static bool t5_make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld) {
    t5_encode_fn enc = t5_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {T5_BK, T5_BM};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// NHWC bf16 tensor [N][H][W][C] as a 4-D map {C, W, H, N}; box = {64 channels, 16 x, 8 y, 1 image}, 128-byte swizzle, zero fill outside
static bool t5_make_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C) {
    t5_encode_fn enc = t5_encoder();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {T5_BK, T5_TW, T5_TH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
