// SoftmaxWithT forward for the IIC feeders (reference: contrastyou/projectors/nn.py:36-44, applied per sub-head by
// DenseClusterHead / ClusterHead, heads.py:127-172):  p[b, :, h, w] = softmax(logits[b, :, h, w] / T) over the K planes.
//
// One streaming pass: every map is read once and written once (HBM-bound; 2 * B*K*H*W * 4 B per map), all maps of a call —
// both views of every sub-head — in ONE launch (blockIdx.y = map).  A thread owns 4 consecutive pixels (float4 per plane,
// coalesced along W) and keeps their K logits in registers; K > 32, 16-bit inputs or unaligned planes take a scalar kernel that
// re-reads the logits from L1/L2.  The backward half lives in the IIC adjoint's epilogue (iic_bwd_tc.cu).
#include "common.cuh"

namespace cy {
namespace {

constexpr int SM_MAXMAPS = 16;
struct SmPtrs {
    const void* in[SM_MAXMAPS];
    void* out[SM_MAXMAPS];
};

template <int KT>
__global__ void __launch_bounds__(256)
softmax_t_fwd_vec4_kernel(const __grid_constant__ SmPtrs ptrs, int K, long long plane, long long n_groups, float c) {
    const float* in = reinterpret_cast<const float*>(ptrs.in[blockIdx.y]);      // (out may alias in: no __restrict__)
    float* out = reinterpret_cast<float*>(ptrs.out[blockIdx.y]);
    for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < n_groups; gi += (long long)gridDim.x * blockDim.x) {
        const long long pix = gi * 4;
        const size_t base = (size_t)(pix / plane) * K * plane + (size_t)(pix % plane);
        float4 v[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) v[k] = __ldcs(reinterpret_cast<const float4*>(in + base + (size_t)k * plane));      // streamed: no reuse
        float4 m = v[0];
#pragma unroll
        for (int k = 1; k < KT; ++k)
            if (k < K) { m.x = fmaxf(m.x, v[k].x); m.y = fmaxf(m.y, v[k].y); m.z = fmaxf(m.z, v[k].z); m.w = fmaxf(m.w, v[k].w); }
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                v[k].x = exp2f((v[k].x - m.x) * c); v[k].y = exp2f((v[k].y - m.y) * c);
                v[k].z = exp2f((v[k].z - m.z) * c); v[k].w = exp2f((v[k].w - m.w) * c);
                s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w;
            }
        const float4 r = make_float4(1.f / s.x, 1.f / s.y, 1.f / s.z, 1.f / s.w);
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const float4 o = make_float4(v[k].x * r.x, v[k].y * r.y, v[k].z * r.z, v[k].w * r.w);
                *reinterpret_cast<float4*>(out + base + (size_t)k * plane) = o;      // (the joint kernel reads it next: keep it in L2)
            }
    }
}

// any K / dtype / alignment: one pixel per thread, three passes over its K logits (the second and third hit L1 / L2)
__global__ void __launch_bounds__(256)
softmax_t_fwd_scalar_kernel(const __grid_constant__ SmPtrs ptrs, int dtype, int K, long long plane, long long n_pix, float c) {
    const void* in = ptrs.in[blockIdx.y];
    void* out = ptrs.out[blockIdx.y];
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pix; pix += (long long)gridDim.x * blockDim.x) {
        const size_t base = (size_t)(pix / plane) * K * plane + (size_t)(pix % plane);
        float m = ld_as_float(in, dtype, base);
        for (int k = 1; k < K; ++k) m = fmaxf(m, ld_as_float(in, dtype, base + (size_t)k * plane));
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += exp2f((ld_as_float(in, dtype, base + (size_t)k * plane) - m) * c);
        const float r = 1.f / s;
        for (int k = 0; k < K; ++k)
            st_from_float(out, dtype, base + (size_t)k * plane, exp2f((ld_as_float(in, dtype, base + (size_t)k * plane) - m) * c) * r);
    }
}

template <int KT>
int launch_vec4(const SmPtrs& p, int n, int K, long long plane, long long n_groups, float c, cudaStream_t st) {
    const long long want = (n_groups + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8 / n + 1;
    const unsigned gx = (unsigned)(want < cap ? want : cap);
    softmax_t_fwd_vec4_kernel<KT><<<dim3(gx, (unsigned)n), 256, 0, st>>>(p, K, plane, n_groups, c);
    CY_CHECK_LAUNCH("softmax_t_fwd");
    return CY_OK;
}

}  // namespace

int softmax_t_fwd(const void* const* logits, void* const* probs, int n_maps, int dtype, int B, int K, int H, int W, float T,
                  cudaStream_t st) {
    const long long plane = (long long)H * W, n_pix = (long long)B * plane;
    const float c = 1.4426950408889634f / T;
    for (int m0 = 0; m0 < n_maps; m0 += SM_MAXMAPS) {
        const int n = n_maps - m0 < SM_MAXMAPS ? n_maps - m0 : SM_MAXMAPS;
        SmPtrs p;
        bool vec = dtype == CY_F32 && (plane % 4) == 0 && K <= 32;
        for (int i = 0; i < SM_MAXMAPS; ++i) {
            const int j = m0 + (i < n ? i : 0);
            p.in[i] = logits[j];
            p.out[i] = probs[j];
            vec = vec && (reinterpret_cast<uintptr_t>(logits[j]) & 15) == 0 && (reinterpret_cast<uintptr_t>(probs[j]) & 15) == 0;
        }
        int rc;
        if (vec) {
            const long long ng = n_pix / 4;
            if (K <= 4) rc = launch_vec4<4>(p, n, K, plane, ng, c, st);
            else if (K <= 8) rc = launch_vec4<8>(p, n, K, plane, ng, c, st);
            else if (K <= 12) rc = launch_vec4<12>(p, n, K, plane, ng, c, st);
            else if (K <= 16) rc = launch_vec4<16>(p, n, K, plane, ng, c, st);
            else if (K <= 24) rc = launch_vec4<24>(p, n, K, plane, ng, c, st);
            else rc = launch_vec4<32>(p, n, K, plane, ng, c, st);
        } else {
            const long long want = (n_pix + 255) / 256;
            const long long cap = (long long)device_sm_count() * 8 / n + 1;
            softmax_t_fwd_scalar_kernel<<<dim3((unsigned)(want < cap ? want : cap), (unsigned)n), 256, 0, st>>>(p, dtype, K, plane, n_pix, c);
            CY_CHECK_LAUNCH("softmax_t_fwd");
            rc = CY_OK;
        }
        if (rc != CY_OK) return rc;
    }
    return CY_OK;
}

}  // namespace cy
