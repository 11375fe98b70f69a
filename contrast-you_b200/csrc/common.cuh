// Shared helpers for the sm_100a kernels of libcontrastyou_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/contrastyou_b200.h"

namespace cy {

// thread-local error text behind cy_last_error()
void set_error(const char* fmt, ...);

// Per-DEVICE host-side caches (a process may drive several GPUs: the reference works on any device).
constexpr int CY_MAX_DEVICES = 64;
int current_device();       // cudaGetDevice(); 0 when the runtime cannot tell
int device_sm_count();      // SM count of the current device, cached per device
// dynamic shared memory opt-in of one kernel, remembered per device: need(smem) is true when the attribute has to be raised
struct SmemAttrCache {
    size_t granted[CY_MAX_DEVICES] = {};
    bool need(size_t smem) const { return smem > granted[current_device() % CY_MAX_DEVICES]; }
    void set(size_t smem) { granted[current_device() % CY_MAX_DEVICES] = smem; }
};

#define CY_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            cy::set_error(__VA_ARGS__);    \
            return CY_ERR_ARG;             \
        }                                  \
    } while (0)

// every kernel launch of the library passes through CY_CHECK_LAUNCH: it also feeds cy_launch_count() (bench.py's
// `gpu_launches` is this counter's difference over the timed region, not an estimate)
void count_launch();

#define CY_CHECK_LAUNCH(what)                                                    \
    do {                                                                         \
        cy::count_launch();                                                      \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            cy::set_error("%s: %s", what, cudaGetErrorString(e__));              \
            return (int)e__;                                                     \
        }                                                                        \
    } while (0)

__device__ __forceinline__ float ld_as_float(const void* p, int dtype, size_t idx) {
    if (dtype == CY_F32) return __ldg(reinterpret_cast<const float*>(p) + idx);
    if (dtype == CY_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
    return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}

__device__ __forceinline__ void st_from_float(void* p, int dtype, size_t idx, float v) {
    if (dtype == CY_F32) reinterpret_cast<float*>(p)[idx] = v;
    else if (dtype == CY_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(p)[idx] = __float2half_rn(v);
}

__host__ __device__ __forceinline__ int dtype_size(int dtype) { return dtype == CY_F32 ? 4 : 2; }

// ---- pair masks: the ONE definition every kernel uses (contrastive.py:31-48, :62-71) ----
// returns bit0 = positive, bit1 = negative for the ordered pair (i, j) of the 2n x 2n problem.
struct PairMask {
    const int32_t* labels;  // [N] tiled labels (label path) or nullptr
    const uint8_t* codes;   // [n, n] mask codes (mask= path) or nullptr
    int64_t n;              // rows per view
};

__device__ __forceinline__ int pair_bits_labels(int32_t li, int32_t lj, bool offdiag) {
    const bool same = (li == lj);
    return offdiag ? (same ? 1 : 2) : 0;
}

__device__ __forceinline__ int pair_bits_codes(const uint8_t* codes, int64_t n, int64_t i, int64_t j) {
    if (i == j) return 0;
    const int64_t ii = i >= n ? i - n : i, jj = j >= n ? j - n : j;
    const uint8_t c = __ldg(codes + ii * n + jj);
    return c == 1 ? 1 : (c == 0 ? 2 : 0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cy
