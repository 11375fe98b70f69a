// Feeders of the InfoNCE kernels: one pass that stacks the two views, applies the label-sort permutation and runs the
// reference's `is_normalized` assertion on the device; and its adjoint (scatter of dZ back to the two views).
// Reference: contrastyou/losses/contrastive.py:9-11 (is_normalized: allclose(norm, 1) evaluated in the input dtype),
// :15 (torch.cat of the two views), :58 (the assertion).  The permutation is this implementation's own (rows sorted by
// label so that the tensor-core kernels take their mask-free inner loop; the loss is permutation invariant).
#include "common.cuh"

namespace cy {

namespace {

__device__ __forceinline__ float round_to_dtype(float v, int dtype) {
    if (dtype == CY_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (dtype == CY_F16) return __half2float(__float2half_rn(v));
    return v;
}

// one warp per output row; 16-byte accesses when the row pitch allows it
template <int ESIZE>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t* __restrict__ f1, const uint8_t* __restrict__ f2, int dtype, int64_t n, int64_t d,
                 int64_t ld1, int64_t ld2, const int64_t* __restrict__ order, uint8_t* __restrict__ z, int* __restrict__ bad_rows,
                 float* __restrict__ inv_norm, const int64_t* __restrict__ pix_off, int64_t chan_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= 2 * n) return;
    const int64_t src = order ? order[row] : row;
    uint8_t* o = z + row * d * ESIZE;
    float ss = 0.f;
    if (pix_off) {
        // dense-feature gather fused into the pack (semi_seg/hooks/infonce.py:31-46): row = the C-vector of one sampled pixel of
        // a [B, C, h, w] map, element c at pix_off[i] + c * chan_stride (the same pixel list serves both views)
        const int64_t i = src < n ? src : src - n;
        const uint8_t* base = (src < n ? f1 : f2) + pix_off[i] * ESIZE;
        for (int64_t c = lane; c < d; c += 32) {
            const float a = ld_as_float(base, dtype, c * chan_stride);
            st_from_float(o, dtype, c, a);
            ss = fmaf(a, a, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0 && bad_rows) {
            const float nrm = round_to_dtype(sqrtf(ss), dtype);
            if (!(fabsf(nrm - 1.f) <= 1e-8f + 1e-5f)) atomicAdd(bad_rows, 1);
        }
        return;
    }
    const uint8_t* s = src < n ? f1 + src * ld1 * ESIZE : f2 + (src - n) * ld2 * ESIZE;
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(o)) & 15) == 0 && (d * ESIZE) % 16 == 0;
    if (inv_norm) {
        // fused F.normalize (contrastyou/projectors/nn.py:47-54, heads.py:20): z = f / max(||f||_2, 1e-12).  Two sweeps over
        // the row (the second one hits L1); the reciprocal norm is kept for the adjoint.
        for (int64_t c = lane; c < d; c += 32) {
            const float a = ld_as_float(s, dtype, c);
            ss = fmaf(a, a, ss);
        }
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        for (int64_t c = lane; c < d; c += 32) st_from_float(o, dtype, c, ld_as_float(s, dtype, c) * inv);
        if (lane == 0) inv_norm[row] = inv;
        return;
    }
    if (vec) {
        constexpr int PER = 16 / ESIZE;
        for (int64_t c = lane; c < d / PER; c += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(s + c * 16);
            *reinterpret_cast<uint4*>(o + c * 16) = v;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ESIZE == 4) {
                    const float a = __uint_as_float(w[k]);
                    ss = fmaf(a, a, ss);
                } else if (dtype == CY_BF16) {
                    const float a = __uint_as_float(w[k] << 16), b = __uint_as_float(w[k] & 0xffff0000u);
                    ss = fmaf(a, a, ss);
                    ss = fmaf(b, b, ss);
                } else {
                    const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
                    const float2 f = __half22float2(h);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
        }
    } else {
        for (int64_t c = lane; c < d; c += 32) {
            const float a = ld_as_float(s, dtype, c);
            st_from_float(o, dtype, c, a);
            ss = fmaf(a, a, ss);
        }
    }
    ss = warp_sum(ss);
    if (lane == 0 && bad_rows) {
        // torch: norm accumulated in fp32, rounded to the tensor dtype; allclose(norm, 1): |norm - 1| <= 1e-8 + 1e-5 * 1
        const float nrm = round_to_dtype(sqrtf(ss), dtype);
        if (!(fabsf(nrm - 1.f) <= 1e-8f + 1e-5f)) atomicAdd(bad_rows, 1);
    }
}

// 16 bytes of gradient (4 fp32 or 8 half-precision elements) times a scalar
__device__ __forceinline__ uint4 scale16(uint4 v, int dtype, float g) {
    if (dtype == CY_F32) {
        return make_uint4(__float_as_uint(__uint_as_float(v.x) * g), __float_as_uint(__uint_as_float(v.y) * g),
                          __float_as_uint(__uint_as_float(v.z) * g), __float_as_uint(__uint_as_float(v.w) * g));
    }
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (dtype == CY_BF16) {
            float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            const __nv_bfloat162 r = __floats2bfloat162_rn(f.x * g, f.y * g);
            w[k] = *reinterpret_cast<const uint32_t*>(&r);
        } else {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            const __half2 r = __floats2half2_rn(f.x * g, f.y * g);
            w[k] = *reinterpret_cast<const uint32_t*>(&r);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int ESIZE>
__global__ void __launch_bounds__(256)
unpack_rows_kernel(const uint8_t* __restrict__ dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* __restrict__ order,
                   uint8_t* __restrict__ g1, uint8_t* __restrict__ g2, const uint8_t* __restrict__ z, const float* __restrict__ inv_norm,
                   const float* __restrict__ gscale, const int64_t* __restrict__ pix_off, int64_t chan_stride) {
    const int lane = threadIdx.x & 31;
    const float gs = gscale ? gscale[0] : 1.f;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= 2 * n) return;
    const int64_t dst = order ? order[row] : row;
    const uint8_t* s = dz + row * lddz * ESIZE;
    if (pix_off) {      // adjoint of the fused gather: the sampled pixels of one image are distinct, so plain stores into zeroed maps
        const int64_t i = dst < n ? dst : dst - n;
        uint8_t* base = (dst < n ? g1 : g2) + pix_off[i] * ESIZE;
        for (int64_t c = lane; c < d; c += 32) st_from_float(base, dtype, c * chan_stride, ld_as_float(s, dtype, c) * gs);
        return;
    }
    uint8_t* o = dst < n ? g1 + dst * d * ESIZE : g2 + (dst - n) * d * ESIZE;
    if (inv_norm) {
        // adjoint of z = f / ||f||:  g = (dz - z (z . dz)) / ||f||
        const uint8_t* zr = z + row * d * ESIZE;
        float dot = 0.f;
        for (int64_t c = lane; c < d; c += 32) dot = fmaf(ld_as_float(s, dtype, c), ld_as_float(zr, dtype, c), dot);
        dot = warp_sum(dot);
        const float inv = inv_norm[row];
        for (int64_t c = lane; c < d; c += 32)
            st_from_float(o, dtype, c, (ld_as_float(s, dtype, c) - ld_as_float(zr, dtype, c) * dot) * (inv * gs));
        return;
    }
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(o)) & 15) == 0 && (d * ESIZE) % 16 == 0;
    if (vec) {
        if (gscale) {
            for (int64_t c = lane; c < d * ESIZE / 16; c += 32)
                *reinterpret_cast<uint4*>(o + c * 16) = scale16(*reinterpret_cast<const uint4*>(s + c * 16), dtype, gs);
        } else {
            for (int64_t c = lane; c < d * ESIZE / 16; c += 32) *reinterpret_cast<uint4*>(o + c * 16) = *reinterpret_cast<const uint4*>(s + c * 16);
        }
    } else if (gscale) {
        for (int64_t c = lane; c < d; c += 32) st_from_float(o, dtype, c, ld_as_float(s, dtype, c) * gs);
    } else {
        for (int64_t c = lane; c < d * ESIZE; c += 32) o[c] = s[c];
    }
}

// fp32 rows -> [hi | lo] bf16 halves (CY_F32_SPLIT): hi = bf16(v) (round to nearest), lo = bf16(v - hi); one warp per row
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                  const int64_t* __restrict__ order, __nv_bfloat16* __restrict__ zs, int* __restrict__ bad_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= 2 * n) return;
    const int64_t src = order ? order[row] : row;
    const float* s = src < n ? f1 + src * ld1 : f2 + (src - n) * ld2;
    __nv_bfloat16* o = zs + row * 2 * d;
    float ss = 0.f;
    for (int64_t c = lane; c < d; c += 32) {
        const float v = s[c];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        o[c] = hi;
        o[d + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0 && bad_rows) {
        if (!(fabsf(sqrtf(ss) - 1.f) <= 1e-8f + 1e-5f)) atomicAdd(bad_rows, 1);      // is_normalized on the fp32 values
    }
}

}  // namespace

int infonce_pack_split(const void* f1, const void* f2, int64_t n, int64_t d, int64_t ld1, int64_t ld2, const int64_t* order, void* zs,
                       int* bad_rows, cudaStream_t st) {
    const unsigned grid = (unsigned)((2 * n + 7) / 8);
    pack_split_kernel<<<grid, 256, 0, st>>>((const float*)f1, (const float*)f2, n, d, ld1, ld2, order, (__nv_bfloat16*)zs, bad_rows);
    CY_CHECK_LAUNCH("infonce_pack_split");
    return CY_OK;
}

int infonce_pack(const void* f1, const void* f2, int dtype, int64_t n, int64_t d, int64_t ld1, int64_t ld2, const int64_t* order,
                 void* z, int* bad_rows, float* inv_norm, const int64_t* pix_off, int64_t chan_stride, cudaStream_t st) {
    const unsigned grid = (unsigned)((2 * n + 7) / 8);
    if (dtype == CY_F32)
        pack_rows_kernel<4><<<grid, 256, 0, st>>>((const uint8_t*)f1, (const uint8_t*)f2, dtype, n, d, ld1, ld2, order, (uint8_t*)z, bad_rows, inv_norm, pix_off, chan_stride);
    else
        pack_rows_kernel<2><<<grid, 256, 0, st>>>((const uint8_t*)f1, (const uint8_t*)f2, dtype, n, d, ld1, ld2, order, (uint8_t*)z, bad_rows, inv_norm, pix_off, chan_stride);
    CY_CHECK_LAUNCH("infonce_pack");
    return CY_OK;
}

int infonce_unpack(const void* dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* order, void* g1, void* g2,
                   const void* z, const float* inv_norm, const float* gscale, const int64_t* pix_off, int64_t chan_stride, cudaStream_t st) {
    const unsigned grid = (unsigned)((2 * n + 7) / 8);
    if (dtype == CY_F32)
        unpack_rows_kernel<4><<<grid, 256, 0, st>>>((const uint8_t*)dz, dtype, n, d, lddz, order, (uint8_t*)g1, (uint8_t*)g2, (const uint8_t*)z, inv_norm, gscale, pix_off, chan_stride);
    else
        unpack_rows_kernel<2><<<grid, 256, 0, st>>>((const uint8_t*)dz, dtype, n, d, lddz, order, (uint8_t*)g1, (uint8_t*)g2, (const uint8_t*)z, inv_norm, gscale, pix_off, chan_stride);
    CY_CHECK_LAUNCH("infonce_unpack");
    return CY_OK;
}

}  // namespace cy
