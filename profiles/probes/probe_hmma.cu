// Probe: legacy warp-level mma.sync (HMMA) bf16 m16n8k16 throughput per SM on sm_100a, as a function of resident warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_hmma probe_hmma.cu ; run: ./probe_hmma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void hmma_loop(float* out, int iters, long long* cycles) {
    float acc[NACC][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + threadIdx.x + i;
    b[0] = 0x3f803f80u + threadIdx.x;
    b[1] = 0x3f803f80u - threadIdx.x;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma_bf16(acc[j], a, b);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// operands vary between consecutive HMMAs (NA distinct A fragments, NB distinct B fragments), as in a real kernel
template <int NACC, int NA, int NB>
__global__ void hmma_loop_var(float* out, int iters, long long* cycles) {
    float acc[NACC][4];
    uint32_t a[NA][4], b[NB][2];
#pragma unroll
    for (int k = 0; k < NA; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[k][i] = 0x3f803f80u + threadIdx.x + i + 7 * k;
#pragma unroll
    for (int k = 0; k < NB; ++k) { b[k][0] = 0x3f803f80u + threadIdx.x + k; b[k][1] = 0x3f803f80u - threadIdx.x - k; }
#pragma unroll
    for (int j = 0; j < NACC; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma_bf16(acc[j], a[j % NA], b[(j / NA) % NB]);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// does an HMMA hold the issue port?  Per iteration: 2 independent HMMAs + NALU independent integer adds (8 accumulators).
// If CUDA-core instructions issue in the shadow of the HMMA pipe the time per iteration is max(16, 2 + NALU) per warp-slot,
// if not it is 16 + NALU.
template <int NALU>
__global__ void hmma_alu_mix(float* out, int iters, long long* cycles) {
    float acc[2][4] = {};
    uint32_t a[4], b[2], x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + threadIdx.x + i;
    b[0] = 0x3f803f80u + threadIdx.x;
    b[1] = 0x3f803f80u - threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 7 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        mma_bf16(acc[0], a, b);
        mma_bf16(acc[1], a, b);
#pragma unroll
        for (int j = 0; j < NALU; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j % 8]) : "r"(x[(j + 3) % 8] | 1u));
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t sx = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sx ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0][0] + acc[1][1] + __uint_as_float(sx & 0xff);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * 4);
    cudaMalloc(&cyc, 148 * 8 * 8);
    const int iters = 4096;
    // dependent-issue latency: one warp per SMSP, NACC independent accumulator chains
    {
        long long c;
        hmma_loop<1><<<148, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 warp/SMSP, 1 chain : %.1f cycles per dependent HMMA\n", (double)c / iters);
        hmma_loop<2><<<148, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 warp/SMSP, 2 chains: %.1f cycles per HMMA\n", (double)c / iters / 2);
        hmma_loop<4><<<148, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 warp/SMSP, 4 chains: %.1f cycles per HMMA\n", (double)c / iters / 4);
        hmma_loop<2><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 2 chains: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 2 / 4);
        hmma_loop<1><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 1 chain : %.1f cycles per HMMA per SMSP\n", (double)c / iters / 4);
    }
    {
        long long c;
        hmma_loop_var<8, 4, 2><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 8 chains, 4 A x 2 B fragments: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 8 / 4);
        hmma_loop_var<8, 2, 4><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 8 chains, 2 A x 4 B fragments: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 8 / 4);
        hmma_loop_var<8, 1, 8><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 8 chains, 1 A x 8 B fragments: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 8 / 4);
        hmma_loop_var<8, 8, 1><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("4 warps/SMSP, 8 chains, 8 A x 1 B fragments: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 8 / 4);
        hmma_loop_var<8, 4, 2><<<148, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 warp/SMSP,  8 chains, 4 A x 2 B fragments: %.1f cycles per HMMA per SMSP\n", (double)c / iters / 8);
    }
    {
        long long c;
#define MIX(N) hmma_alu_mix<N><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("4 warps/SMSP, 2 HMMA + %2d int adds per iteration: %.1f cycles per iteration per SMSP (%.1f per warp)\n", N, (double)c / iters / 4, (double)c / iters);
        MIX(0) MIX(4) MIX(8) MIX(16) MIX(24) MIX(32) MIX(48)
#undef MIX
#define MIX(N) hmma_alu_mix<N><<<148, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("1 warp/SMSP,  2 HMMA + %2d int adds per iteration: %.1f cycles per iteration\n", N, (double)c / iters);
        MIX(0) MIX(8) MIX(16) MIX(32)
#undef MIX
    }
    const int NACC = 12;
    for (int warps = 4; warps <= 32; warps *= 2) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        hmma_loop<NACC><<<148, warps * 32>>>(out, 16, cyc);
        cudaEventRecord(e0);
        hmma_loop<NACC><<<148, warps * 32>>>(out, iters, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double macs = (double)warps * iters * NACC * 16 * 8 * 16;
        printf("warps/SM %2d: %s  %.3f ms  cycles %lld  -> %.1f MAC/clk/SM  (%.1f TFLOP/s chip)\n", warps, cudaGetErrorString(e), ms, c,
               macs / (double)c, 2.0 * macs * 148 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
