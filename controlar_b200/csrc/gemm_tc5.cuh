// gemm_tc5.cuh — dense bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM) for the plain
// (non-convolution) dense stages: C[M,N] = epi(A[M,K] · B[N,K]^T), fp32 accumulate.  Used by the prefill
// (autoregressive/models/gpt_t2i.py:433-470: wqkv / wo / w1 / w3 / w2 over B_eff·T rows) and the control-token MLPs
// (gpt_t2i.py:165-181 over B_eff·N rows).  K % 64 == 0; M and N tails are zero-filled on load and masked on store.
//
//   * 128 x 128 output tile per CTA, 64-wide k-blocks, 4-stage shared-memory ring (128 KB).
//   * Operands are copied with 16-byte cp.async into the canonical no-swizzle K-major core-matrix layout the tensor core
//     reads through shared-memory descriptors: core matrix = 8 rows x 16 B (128 B contiguous); core (r8, kc) of a
//     [128][64] tile at (kc·16 + r8)·128 B  =>  LBO (K direction) = 2048 B, SBO (row direction) = 128 B.
//     (Encodings validated against the host by scripts/tc5_probe.cu on B200.)
//   * ONE thread issues the four m128n128k16 MMAs of a k-block and commits them to the stage's mbarrier; a stage is
//     refilled when that mbarrier flips.  The accumulator (128 lanes x 128 fp32 columns) is read back with tcgen05.ld
//     (32 lanes x 32 columns per warp and instruction) for the fused epilogue.
// Epilogue (same rounding points as gemm_dense.cuh): v = bf16(acc); act: v = bf16(gelu_tanh(v)); resid: v = bf16(v + resid).
#pragma once
#include "common.cuh"

constexpr int T5_BM = 128, T5_BN = 128, T5_BK = 64, T5_STAGES = 4, T5_THREADS = 256;
constexpr int T5_TILE_BYTES = T5_BM * T5_BK * 2;                               // 16 KB per operand and stage
constexpr int T5_SMEM = T5_STAGES * 2 * T5_TILE_BYTES + 1024;                  // + alignment slack

struct Tc5P {
    const bf16* A; const bf16* B; int M, N, K, lda, ldb;
    const bf16* resid; int ldr;
    bf16* C; int ldc;
    int act;            // 0 none, 1 GELU-tanh
};

__device__ __forceinline__ uint32_t t5_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48), no swizzle
__device__ __forceinline__ uint64_t t5_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void t5_cp16(uint32_t sdst, const void* gsrc, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void t5_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();            // never hang the box
    }
}

static __global__ void __launch_bounds__(T5_THREADS, 1) gemm_tc5_kernel(Tc5P p) {
    extern __shared__ unsigned char t5_raw[];
    __shared__ __align__(8) uint64_t mma_done[T5_STAGES];
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (t5_smem(t5_raw) + 1023u) & ~1023u;                 // stage s: A at smem0 + s*32 KB, B 16 KB after
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * T5_BM, n0 = blockIdx.x * T5_BN;
    const int nkb = p.K / T5_BK;

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t5_smem(&mma_done[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(t5_smem(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T5_BN >> 3) << 17) | ((uint32_t)(T5_BM >> 4) << 24);

    // loader: a warp-instruction covers 8 rows x 4 k-chunks (4 x 128 B contiguous in shared memory, 8 x 64 B in global)
    const int r_lo = lane & 7, kcl = lane >> 3;
    auto load_stage = [&](int slot, int kb) {
        const uint32_t sA = smem0 + (uint32_t)slot * 2u * T5_TILE_BYTES, sB = sA + T5_TILE_BYTES;
        const int k0 = kb * T5_BK;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int wi = warp + it * 8;                     // 32 warp-iterations per operand: r8 = wi >> 1, k-half = wi & 1
            const int r = (wi >> 1) * 8 + r_lo, kc = (wi & 1) * 4 + kcl;
            const uint32_t off = (uint32_t)((kc * 16 + (r >> 3)) * 128 + (r & 7) * 16);
            const bool va = m0 + r < p.M, vb = n0 + r < p.N;
            t5_cp16(sA + off, p.A + (size_t)(va ? m0 + r : 0) * p.lda + k0 + kc * 8, va);
            t5_cp16(sB + off, p.B + (size_t)(vb ? n0 + r : 0) * p.ldb + k0 + kc * 8, vb);
        }
    };
#pragma unroll
    for (int s = 0; s < T5_STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nkb; ++kb) {
        asm volatile("cp.async.wait_group %0;" ::"n"(T5_STAGES - 2) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // this thread's copies -> visible to the tensor core
        __syncthreads();
        const int slot = kb % T5_STAGES;
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sA = smem0 + (uint32_t)slot * 2u * T5_TILE_BYTES, sB = sA + T5_TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < T5_BK / 16; ++kk) {
                const uint64_t da = t5_desc(sA + kk * 2 * 2048, 2048, 128), db = t5_desc(sB + kk * 2 * 2048, 2048, 128);
                const uint32_t accf = (kb > 0 || kk > 0) ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t5_smem(&mma_done[slot])) : "memory");
        }
        // refill the slot of k-block kb - 1 (its MMAs were committed one iteration ago) with k-block kb + STAGES - 1
        if (kb + T5_STAGES - 1 < nkb) {
            const int ps = (kb + T5_STAGES - 1) % T5_STAGES;        // == (kb - 1) % STAGES; never used before when kb == 0
            if (kb >= 1) t5_mbar_wait(t5_smem(&mma_done[ps]), (uint32_t)(((kb - 1) / T5_STAGES) & 1));
            load_stage(ps, kb + T5_STAGES - 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // the commit of the last k-block covers every MMA issued before it
    t5_mbar_wait(t5_smem(&mma_done[(nkb - 1) % T5_STAGES]), (uint32_t)(((nkb - 1) / T5_STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) (= rows), columns [64 (w >> 2), +64)
    const int row = m0 + (warp & 3) * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (warp >> 2) * 64 + cc * 32;
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.M) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                const int n = n0 + c0 + j8 * 8;
                if (n < p.N) {                                     // N % 8 == 0 (host-checked): whole 16-byte groups
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[j] = rnd<bf16>(__uint_as_float(v[j8 * 8 + j]));
                        if (p.act == 1) f[j] = rnd<bf16>(gelu_tanh_f(f[j]));
                    }
                    if (p.resid) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(p.resid + (size_t)row * p.ldr + n);
                        const uint32_t ri[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float a, b;
                            unpack_bf16x2(ri[q], a, b);
                            f[2 * q] = rnd<bf16>(f[2 * q] + a); f[2 * q + 1] = rnd<bf16>(f[2 * q + 1] + b);
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
