// Probe: tcgen05.mma kind::f16 (bf16 in, fp32 out) with SMALL tiles fed from shared memory in the canonical
// NO-SWIZZLE K-major layout (8-row x 16-byte core matrices) — the operand shape of the IIC joint contraction:
//   D[M, N] += A[M, K16] * B[N, K16]^T,   M in {64, 128}, N in {16 .. 128}
// It (1) validates the descriptor / layout conventions against a host reference (and prints where the rows of an M=64
// accumulator live in TMEM) and (2) measures cycles per MMA, optionally while 8 other warps hammer shared memory
// (LDS.128 + STS.128), to see whether UMMA operand fetches and LSU traffic share the 128 B/clk/SM port.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma_small probe_umma_small.cu
// Run:   ./probe_umma_small M N [hammer=0|1]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// no-swizzle descriptor: layout type 0
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int KTOT = 64;          // K extent held in smem (4 MMAs of K16)
constexpr int THREADS = 32 * 13;  // warps 0-3: issue / readout, warps 4-12: shared-memory hammer (hammer mode)

// smem operand layout: offset(r, c) = (r/8)*SBO + (c/8)*128 + (r%8)*16 + (c%8)*2,  SBO = (KTOT/8)*128
__global__ void __launch_bounds__(THREADS, 1)
umma_probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int M, int N,
           int loops, int hammer, long long* stats, int ts = 0, int nacc = 1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t SBO = (KTOT / 8) * 128;
    uint8_t* sa = smem;                          // 128 rows max
    uint8_t* sb = smem + 16 * SBO;               // 128 rows max... (N/8 groups)
    uint8_t* scratch = sb + 16 * SBO;            // 64 KB hammer area
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + 65536);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    volatile int* stop = reinterpret_cast<volatile int*>(tmem_slot + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < 128 * KTOT; i += THREADS) {
        const int r = i / KTOT, c = i % KTOT;
        const uint32_t off = (r / 8) * SBO + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sa + off) = r < M ? A[r * KTOT + c] : __float2bfloat16(0.f);
        *reinterpret_cast<__nv_bfloat16*>(sb + off) = r < N ? B[r * KTOT + c] : __float2bfloat16(0.f);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = idesc_bf16(M, N);

    long long t0 = 0, t1 = 0;
    unsigned long long hbytes = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int l = 0; l < loops; ++l) {
#pragma unroll
            for (int j = 0; j < KTOT / 16; ++j) {
                const uint64_t da = desc_noswz(smem_u32(sa) + j * 256, 128, SBO);
                const uint64_t db = desc_noswz(smem_u32(sb) + j * 256, 128, SBO);
                const uint32_t acc = (l | j) ? 1u : 0u;
                const uint32_t dcol = tmem + (nacc > 1 ? (uint32_t)(j % nacc) * 32u : 0u);      // nacc > 1: independent accumulators (N <= 32)
                if (ts)     // A operand from tensor memory (columns 96 + 8 j; contents are whatever is there: timing only)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(dcol), "r"(tmem + 96u + 8u * j), "l"(db), "r"(idesc), "r"(acc) : "memory");
                else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        while (!mbar_try_wait(bar, 0)) {}
        t1 = clock64();
        *stop = 1;
    } else if (warp >= 4 && hammer) {
        // each thread: LDS.128 + STS.128 on its own 16-byte slot pattern (conflict-free)
        float4* p = reinterpret_cast<float4*>(scratch) + (threadIdx.x - 128);
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        while (!*stop) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float4 w = p[u * 288];
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                p[u * 288] = v;
            }
            hbytes += 8 * 32;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4 && blockIdx.x == 0) {
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 128 + c0 + i] = __uint_as_float(r[i]);
        }
    }
    if (threadIdx.x == 0) stats[blockIdx.x * 2] = t1 - t0;
    if (hammer && warp >= 4) atomicAdd(reinterpret_cast<unsigned long long*>(stats + blockIdx.x * 2 + 1), hbytes);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 72, hammer = argc > 3 ? atoi(argv[3]) : 0;
    std::vector<__nv_bfloat16> hA(128 * KTOT), hB(128 * KTOT);
    std::vector<float> fA(128 * KTOT), fB(128 * KTOT);
    srand(1);
    for (int i = 0; i < 128 * KTOT; ++i) {
        hA[i] = __float2bfloat16((float)(rand() % 17 - 8) / 8.f);
        hB[i] = __float2bfloat16((float)(rand() % 13 - 6) / 4.f);
        fA[i] = __bfloat162float(hA[i]);
        fB[i] = __bfloat162float(hB[i]);
    }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    long long* dS;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dD, 128 * 128 * 4);
    cudaMalloc(&dS, 148 * 2 * 8);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 128 * 4);
    cudaMemset(dS, 0, 148 * 2 * 8);
    const size_t smem = 32 * (KTOT / 8) * 128 + 65536 + 64;
    cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // correctness: one pass over K = 64
    umma_probe<<<1, THREADS, smem>>>(dA, dB, dD, M, N, 1, 0, dS);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d N=%d correctness launch: %s\n", M, N, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hD(128 * 128);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    // find the TMEM lane of each row by matching against the reference
    int bad = 0;
    for (int r = 0; r < M; ++r) {
        std::vector<float> ref(N);
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < KTOT; ++k) s += fA[r * KTOT + k] * fB[n * KTOT + k];
            ref[n] = s;
        }
        int found = -1;
        for (int l = 0; l < 128 && found < 0; ++l) {
            bool ok = true;
            for (int n = 0; n < N && ok; ++n) ok = fabsf(hD[l * 128 + n] - ref[n]) < 1e-3f;
            if (ok) found = l;
        }
        if (found < 0) ++bad;
        if (r < 4 || r == 15 || r == 16 || r == 31 || r == 32 || r == 63 || r == M - 1) printf("  row %3d -> TMEM lane %d\n", r, found);
    }
    printf("rows not found: %d of %d\n", bad, M);
    // throughput on all SMs
    for (int ts = 0; ts <= 1; ++ts)
        for (int nacc = 1; nacc <= (N <= 32 ? 3 : 1); ++nacc) {
            const int loops = 2000;
            cudaMemset(dS, 0, 148 * 2 * 8);
            umma_probe<<<148, THREADS, smem>>>(dA, dB, dD, M, N, loops, 0, dS, ts, nacc);
            e = cudaDeviceSynchronize();
            long long st[4];
            cudaMemcpy(st, dS, 32, cudaMemcpyDeviceToHost);
            printf("%s, %d accumulator(s): %s  cycles/MMA %.2f\n", ts ? "A from TMEM" : "A from smem", nacc, cudaGetErrorString(e),
                   st[0] / ((double)loops * (KTOT / 16)));
        }
    for (int h = 0; h <= hammer; ++h) {
        const int loops = 2000;
        cudaMemset(dS, 0, 148 * 2 * 8);
        umma_probe<<<148, THREADS, smem>>>(dA, dB, dD, M, N, loops, h, dS);
        e = cudaDeviceSynchronize();
        long long st[4];
        cudaMemcpy(st, dS, 32, cudaMemcpyDeviceToHost);
        const double n_mma = (double)loops * (KTOT / 16);
        printf("hammer=%d: %s  cycles/MMA %.2f  (operand bytes/MMA %d -> %.1f B/clk)  LSU traffic %.1f B/clk\n", h, cudaGetErrorString(e),
               st[0] / n_mma, (M + N) * 32, (M + N) * 32.0 * n_mma / st[0], (double)st[1] / st[0]);
    }
    return 0;
}
