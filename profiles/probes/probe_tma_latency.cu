// Probe: latency / throughput of cp.async.bulk.tensor.3d (fp32, no swizzle) box loads as a function of the box shape,
// as used by the IIC kernels (tensor (W, H, B*K), boxes of (bw x bh x bk) with short rows).
// One CTA per SM, one thread issues `depth` loads back to back into a ring and waits for each in order.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tma_latency probe_tma_latency.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm, int box_bytes, int depth, int loads, int tiles_w, int tiles_h, int bw, int bh, int bk,
      int planes, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint8_t* buf = smem + 128;
    const int stride = (box_bytes + 127) & ~127;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + i)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto issue = [&](int i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, b = (tile / (tiles_w * tiles_h)) % (planes / bk);
        const int s = i % depth;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(buf + (size_t)s * stride)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(bar + s)),
                       "r"(tw * bw), "r"(th * bh), "r"(b * bk) : "memory");
    };
    const long long t0 = clock64();
    for (int i = 0; i < depth && i < loads; ++i) issue(i);
    for (int i = 0; i < loads; ++i) {
        const int s = i % depth;
        const uint32_t ph = (i / depth) & 1;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar + s)), "r"(ph) : "memory");
        if (i + depth < loads) issue(i + depth);
    }
    out[blockIdx.x] = clock64() - t0;
}

int main() {
    const int W = 224, H = 224, planes = 640;      // two maps of config 3 = 2 x 320 planes; here one tensor of 640 planes
    float* d;
    cudaMalloc(&d, (size_t)W * H * planes * 4);
    cudaMemset(d, 0, (size_t)W * H * planes * 4);
    long long* out;
    cudaMalloc(&out, 148 * 8);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int shapes[][3] = {{32, 11, 10}, {36, 11, 10}, {40, 9, 10}, {44, 9, 10}, {64, 11, 10}, {68, 11, 10}, {132, 11, 10}, {228, 11, 10}, {228, 3, 10}, {224, 16, 4}, {64, 64, 1}};
    for (auto& sh : shapes) {
        const int bw = sh[0], bh = sh[1], bk = sh[2];
        CUtensorMap tm;
        cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
        cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bk};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d for %dx%dx%d\n", (int)r, bw, bh, bk); continue; }
        const int box_bytes = bw * bh * bk * 4;
        const int tiles_w = W / bw > 0 ? W / bw : 1, tiles_h = H / bh;
        for (int depth : {1, 2, 4}) {
            const size_t smem = 128 + (size_t)depth * ((box_bytes + 127) & ~127);
            if (smem > 220 * 1024) continue;
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const int loads = 64;
            probe<<<148, 128, smem>>>(tm, box_bytes, depth, loads, tiles_w, tiles_h, bw, bh, bk, planes, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[148];
            cudaMemcpy(c, out, sizeof(c), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += c[i];
            avg /= 148;
            printf("box %3d x %3d x %2d (%6d B, %4d rows) depth %d: %s  %.0f clk per load, %.1f B/clk/SM, chip %.0f GB/s @1.9GHz\n", bw, bh, bk,
                   box_bytes, bh * bk, depth, cudaGetErrorString(e), avg / loads, box_bytes * (double)loads / avg,
                   box_bytes * (double)loads / avg * 148 * 1.9);
        }
    }
    return 0;
}
