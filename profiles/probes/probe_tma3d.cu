#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int nfl, int c0, int c1, int c2, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* dst = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((nfl * 4 + 127) & ~127));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nfl * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < nfl; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
    int W = 56, H = 36, BK = 20, bw = argc > 1 ? atoi(argv[1]) : 60, bh = argc > 2 ? atoi(argv[2]) : 11, bk = argc > 3 ? atoi(argv[3]) : 10;
    int c0 = argc > 4 ? atoi(argv[4]) : -3, c1 = argc > 5 ? atoi(argv[5]) : -1, c2 = argc > 6 ? atoi(argv[6]) : 10; int l2 = 1;
    std::vector<float> h(W * H * BK);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int nfl = bw * bh * bk;
    cudaMalloc(&o, nfl * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BK};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bk};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d box=(%d,%d,%d)\n", (int)r, bw, bh, bk);
    size_t smem = ((nfl * 4 + 127) & ~127) + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(tm, o, nfl, c0, c1, c2, 0); printf("coords (%d,%d,%d)\n", c0, c1, c2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> res(nfl);
        cudaMemcpy(res.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
        // element (k=0,row=1,col=3) of the box = global (plane 10, h 0, w 0)
        printf("box[0][1][3]=%.0f expect %.0f ; box[0][0][0]=%.0f expect 0 ; box[2][5][10]=%.0f expect %.0f\n", res[1 * bw + 3],
               (float)(10 * W * H), res[0], res[(2 * bh + 5) * bw + 10], (float)((12 * H + 4) * W + 7));
    }
    return 0;
}
