// Probe: how fast can small tcgen05.mma instructions (M = 128, N = 16 / 32, K = 16, A operand in tensor memory, B from shared
// memory) be issued and executed when the issuing code is warp-uniform (`if (elect_one())` region: ptxas keeps the
// descriptors in uniform registers and emits back-to-back UTCHMMA), with 1-3 issuer warps, and what tcgen05.st / tcgen05.ld
// cost per warp.  This is the shape of the IIC adjoint (csrc/iic_bwd_tc.cu): 12-18 tiny MMAs per 128-pixel output row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma_issue2 probe_umma_issue2.cu
// Run:   ./probe_umma_issue2
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int ACC>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int THREADS = 32 * 8;     // warps 0-2: issuers, warp 3: TMEM allocator, warps 4-7: st / ld tests

// mode 0: NI issuer warps, each issues `loops` batches of NB MMAs (N columns) into its own accumulator, commit per batch, the
//         batch ring is 4 deep (the issuer waits for batch l-4 before issuing batch l)
// mode 1: warps 4-7 each issue `loops` x (tcgen05.st x16 twice + wait::st)
// mode 2: warps 4-7 each issue `loops` x (tcgen05.ld x16 + wait::ld)
template <int N, int NB>
__global__ void __launch_bounds__(THREADS, 1)
probe(int mode, int NI, int loops, long long* stats) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sb = smem;                                     // weights: 12 tiles of [N rows][16 k] bf16 (N * 32 B each)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384 / 4; i += THREADS) reinterpret_cast<uint32_t*>(sb)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t idesc = idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    if (mode == 0 && warp < NI) {
        uint64_t* bar = bars + warp * 4;
        const uint32_t d = tmem + warp * 32;                // accumulator columns of this issuer
        const uint32_t a0 = tmem + 128;                     // A tiles: 6 slots x 32 columns (contents arbitrary: timing only)
        const uint32_t sb_addr = smem_u32(sb);
        t0 = clock64();
        if (elect_one()) {
            for (int l = 0; l < loops; ++l) {
                if (l >= 4) { while (!mbar_try_wait(bar + (l & 3), ((l >> 2) - 1) & 1)) {} }
                const uint32_t a = a0 + (uint32_t)(l % 3) * 32;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const uint64_t db = desc_noswz(sb_addr + (j % 12) * (N * 32), 128, 256);
                    if (j == 0) umma_ts<0>(d, a + (j % 12) * 8, db, idesc);
                    else umma_ts<1>(d, a + (j % 12) * 8, db, idesc);
                }
                umma_commit(bar + (l & 3));
            }
            const int last = loops - 1;
            while (!mbar_try_wait(bar + (last & 3), (last >> 2) & 1)) {}
        }
        __syncwarp();
        t1 = clock64();
        if (lane == 0) stats[blockIdx.x * 4 + warp] = t1 - t0;
    } else if (mode == 1 && warp >= 4) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = lane * 16 + i;
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128;
        t0 = clock64();
        for (int l = 0; l < loops; ++l) {
            const uint32_t t = ta + (uint32_t)(l % 6) * 32;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                         ::"r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                           "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                         ::"r"(t + 16), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                           "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] += 1;
        }
        t1 = clock64();
        if (lane == 0) stats[blockIdx.x * 4 + (warp & 3)] = t1 - t0;
    } else if (mode == 2 && warp >= 4) {
        uint32_t r[16], acc = 0;
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        t0 = clock64();
        for (int l = 0; l < loops; ++l) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(ta + (uint32_t)(l & 3) * 16) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += r[i];
        }
        t1 = clock64();
        if (lane == 0) stats[blockIdx.x * 4 + (warp & 3)] = (t1 - t0) + (acc == 0x12345 ? 1 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, int NB>
void run(long long* dS) {
    const size_t smem = 16384 + 256;
    cudaFuncSetAttribute(probe<N, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int loops = 4000;
    for (int ni = 1; ni <= 3; ++ni) {
        cudaMemset(dS, 0, 148 * 4 * 8);
        probe<N, NB><<<148, THREADS, smem>>>(0, ni, loops, dS);
        cudaError_t e = cudaDeviceSynchronize();
        long long st[4];
        cudaMemcpy(st, dS, 32, cudaMemcpyDeviceToHost);
        printf("N=%2d batch=%2d issuers=%d: %s  cycles/batch/issuer %.1f  cycles/MMA (CTA-wide) %.2f\n", N, NB, ni, cudaGetErrorString(e),
               (double)st[0] / loops, (double)st[0] / loops / NB / ni);
        if (e != cudaSuccess) exit(1);
    }
}

int main() {
    long long* dS;
    cudaMalloc(&dS, 148 * 4 * 8);
    run<16, 18>(dS);
    run<16, 12>(dS);
    run<32, 12>(dS);
    run<32, 6>(dS);
    run<48, 9>(dS);
    run<64, 6>(dS);
    const size_t smem = 16384 + 256;
    for (int mode = 1; mode <= 2; ++mode) {
        const int loops = 4000;
        cudaMemset(dS, 0, 148 * 4 * 8);
        probe<16, 12><<<148, THREADS, smem>>>(mode, 0, loops, dS);
        cudaError_t e = cudaDeviceSynchronize();
        long long st[4];
        cudaMemcpy(st, dS, 32, cudaMemcpyDeviceToHost);
        printf("%s, 4 warps: %s  cycles/iteration %.1f\n", mode == 1 ? "2 x tcgen05.st.x16 + wait::st" : "tcgen05.ld.x16 + wait::ld",
               cudaGetErrorString(e), (double)st[0] / loops);
    }
    return 0;
}
