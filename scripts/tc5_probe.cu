// tc5_probe.cu — dev probe: one CTA computes C[128 x 128] = A[128 x K] · B[128 x K]^T (bf16 -> fp32) with tcgen05.mma
// (accumulator in TMEM, operands in shared memory in the canonical no-swizzle K-major core-matrix layout), and checks it
// against the host.  Purpose: validate the instruction / shared-memory descriptor encodings before they go into the
// dense GEMM of the prefill / DINOv2 / VQGAN path.  Every wait has a watchdog (trap), the box can never hang.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tc5_probe scripts/tc5_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>

typedef __nv_bfloat16 bf16;
constexpr int BM = 128, BN = 128, BK = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (tcgen05): start address >> 4 [0,14), leading-dimension byte offset >> 4 [16,30),
// stride-dimension byte offset >> 4 [32,46), version = 1 at [46,48), swizzle mode [61,64) (0 = none)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// canonical K-major layout of a [rows][BK] tile: core matrix (8 rows x 8 elements = 128 B, row-contiguous 16 B each);
// core (r8, kc) at ((kc * rows/8) + r8) * 128 B  =>  LBO (K direction) = rows/8 * 128, SBO (row direction) = 128
__device__ __forceinline__ uint32_t tile_off(int r, int k, int rows) {
    return (uint32_t)((((k >> 3) * (rows >> 3)) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(128, 1) tc5_gemm(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ C, int K,
                                                    int swap_lbo_sbo) {
    __shared__ __align__(1024) unsigned char sA[BM * BK * 2];
    __shared__ __align__(1024) unsigned char sB[BN * BK * 2];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // one warp allocates 128 TMEM columns (fp32 accumulator 128 lanes x 128 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor, kind::f16: D = F32 (1 << 4), A = BF16 (1 << 7), B = BF16 (1 << 10), both K-major,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // generic-proxy stores of the tile into the canonical layout (a production kernel would use bulk copies of
        // pre-arranged data)
        for (int i = tid; i < BM * BK / 8; i += 128) {
            const int r = i / (BK / 8), kc = i % (BK / 8);
            *reinterpret_cast<uint4*>(sA + tile_off(r, kc * 8, BM)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + k0 + kc * 8);
            *reinterpret_cast<uint4*>(sB + tile_off(r, kc * 8, BN)) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + k0 + kc * 8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lboA = swap_lbo_sbo ? 128u : (BM / 8) * 128u, sboA = swap_lbo_sbo ? (BM / 8) * 128u : 128u;
            const uint32_t lboB = swap_lbo_sbo ? 128u : (BN / 8) * 128u, sboB = swap_lbo_sbo ? (BN / 8) * 128u : 128u;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                // one MMA consumes K = 16 = two core-matrix columns; the next one starts 2 columns further
                const uint64_t da = make_desc(smem_u32(sA) + kk * 2 * (BM / 8) * 128, lboA, sboA);
                const uint64_t db = make_desc(smem_u32(sB) + kk * 2 * (BN / 8) * 128, lboB, sboB);
                const uint32_t acc = (k0 > 0 || kk > 0) ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // everybody waits until the MMAs of this k-block have read the tiles (single-buffered probe)
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            if (!ok && ++spins > (1u << 22)) __trap();
        }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // epilogue: warp q reads TMEM lanes [32 q, 32 q + 32): thread = row, 32 columns per load
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + (tid & 31);
#pragma unroll
        for (int j = 0; j < 32; ++j) C[(size_t)row * BN + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int K = 256;
    const int swap = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<bf16> hA(BM * K), hB(BN * K);
    std::vector<float> fA(BM * K), fB(BN * K), ref(BM * BN), out(BM * BN);
    srand(1);
    for (int i = 0; i < BM * K; ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < BN * K; ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < BM; ++m)
        for (int n = 0; n < BN; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k]; ref[m * BN + n] = s; }
    bf16 *dA, *dB; float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, out.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, out.size() * 4);
    tc5_gemm<<<1, 128>>>(dA, dB, dC, K, swap);
    cudaError_t e = cudaDeviceSynchronize();
    printf("swap=%d launch: %s\n", swap, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < BM * BN; ++i) { const double d = fabs(out[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("swap=%d max abs err %.6f, mismatches %d of %d; C[0][0..3] = %.3f %.3f %.3f %.3f (ref %.3f %.3f %.3f %.3f); C[5][7] %.3f ref %.3f\n", swap,
           maxerr, bad, BM * BN, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3], out[5 * BN + 7], ref[5 * BN + 7]);
    return bad ? 2 : 0;
}
