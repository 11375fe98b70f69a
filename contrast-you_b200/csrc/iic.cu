// IIC discrete-MI segmentation loss: joint accumulation, 900-float epilogue, adjoint.
// Reference: contrastyou/losses/discreteMI.py:139-165 (IIDSegmentationLoss.forward), :225-243 (compute_joint_2D: the
// F.conv2d with a B x H x W "filter"), :246-261 (padding == 0).  SURVEY.md Appendix A5/A6 for the closed forms.
//
// Data layout.  x, y: [B, K, H, W] contiguous.  A CTA stages one (TH x TW) pixel tile of one image in shared memory as
// fp32 rows indexed [(hh * K + k) * XW + col] (hh = tile row incl. halo): consecutive (dy, k1) "roles" then sit on
// consecutive shared-memory rows and, with XW/4 odd, 8 consecutive roles hit 8 distinct 16-byte bank groups.
//
// Forward (iic_joint_kernel): lane = role (dy, k1) [x k2-chunk], warp-level strips walk along w four pixels at a
// time keeping a (4 + 2*PAD)-wide register window of x and broadcast float4 loads of y; every lane owns the
// T x KC accumulators J[k1, k2chunk, dy, :].  Accumulators live in registers across all tiles of a persistent CTA
// and are reduced once (shared-memory atomics, then one [K,K,T,T] partial per CTA); cy_iic_epilogue sums the partials
// in a fixed order (deterministic).
// Backward (iic_bwd_kernel): thread = 4 consecutive pixels x all output channels; dL/dJ (and its flipped transpose)
// sit in shared memory and are read as warp-uniform float4 broadcasts.
#include <stdlib.h>

#include "common.cuh"

namespace cy {

constexpr int IIC_NT = 256;

struct IICGeom {
    int B, K, H, W, pad, T;
    int TH, TW;          // tile (pixels)
    int XW;              // shared row pitch (floats) of a halo tile row
    int CO;              // smem column of the tile's first pixel column (so that float4 window loads are aligned)
    int tiles_h, tiles_w, n_tiles;
    int HH;              // TH + 2*pad
};

__host__ __device__ inline int iic_col_origin(int pad) {
    // smallest CO >= pad with (CO + pad) % 4 == 0
    int co = pad;
    while ((co + pad) % 4) ++co;
    return co;
}

// stage one halo tile of `src` (image b, tile origin h0,w0) into sm[(hh*K + k)*XW + c]; zero outside the image
__device__ __forceinline__ void stage_tile(float* __restrict__ sm, const void* __restrict__ src, int dtype, const IICGeom& g,
                                           int b, int h0, int w0, int halo, int origin) {
    const int rows = g.TH + 2 * halo;
    const int c_lo = origin - halo, c_hi = origin + g.TW + halo;   // [c_lo, c_hi): tile column 0 sits at `origin`
    const int ncol = c_hi - c_lo;
    const int total = rows * g.K * ncol;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int c = idx % ncol;
        const int rk = idx / ncol;
        const int k = rk % g.K, hh = rk / g.K;
        const int h = h0 + hh - halo, w = w0 + c - halo;
        float v = 0.f;
        if (h >= 0 && h < g.H && w >= 0 && w < g.W)
            v = ld_as_float(src, dtype, (((size_t)b * g.K + k) * g.H + h) * g.W + w);
        sm[(hh * g.K + k) * g.XW + c_lo + c] = v;
    }
}

// ---------------------------------------------------------------------------------------------- forward (fast)
template <int PAD, int KC>
__global__ void __launch_bounds__(IIC_NT)
iic_joint_kernel(const void* __restrict__ x, const void* __restrict__ y, int dtype, IICGeom g, float* __restrict__ partials) {
    constexpr int T = 2 * PAD + 1;
    extern __shared__ __align__(16) float smem[];
    const int K = g.K;
    const int nchunk = (K + KC - 1) / KC;
    const int R = T * K * nchunk;                 // roles
    const int nslot = IIC_NT / R;
    float* xs = smem;                                              // [(HH*K) * XW]
    float* ys = xs + (size_t)g.HH * K * g.XW;                      // [(TH*K + KC) * YW], YW == XW
    float* jsm = ys + (size_t)(g.TH * K + KC) * g.XW;              // [K*K*T*T]
    const int nj = K * K * T * T;

    const int tid = threadIdx.x;
    const int slot = tid / R, role = tid % R;
    const bool active = slot < nslot;
    const int rr = role % (T * K), c2 = role / (T * K);            // rr = dy*K + k1
    const int k2base = c2 * KC;

    float acc[T][KC];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int c = 0; c < KC; ++c) acc[a][c] = 0.f;

    for (int i = tid; i < nj; i += IIC_NT) jsm[i] = 0.f;
    // the KC padding rows behind the y tile are read (and their products discarded) when KC does not divide K
    for (int i = tid; i < KC * g.XW; i += IIC_NT) ys[(size_t)g.TH * K * g.XW + i] = 0.f;

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int b = tile / (g.tiles_h * g.tiles_w);
        const int trem = tile % (g.tiles_h * g.tiles_w);
        const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * g.TW;
        __syncthreads();
        stage_tile(xs, x, dtype, g, b, h0, w0, PAD, g.CO);
        stage_tile(ys, y, dtype, g, b, h0, w0, 0, 0);    // y tile: rows (h*K + k2), same pitch, columns at 0..
        __syncthreads();
        if (active) {
            for (int h = slot; h < g.TH; h += nslot) {
                const float* xr = xs + (size_t)(h * K + rr) * g.XW + (g.CO - PAD);
                const float* yr = ys + (size_t)(h * K + k2base) * g.XW;
                float xw[4 + 2 * PAD];
#pragma unroll
                for (int e = 0; e < 2 * PAD; ++e) xw[4 + e] = xr[e];
                for (int w = 0; w < g.TW; w += 4) {
#pragma unroll
                    for (int e = 0; e < 2 * PAD; ++e) xw[e] = xw[4 + e];
                    const float4 nx = *reinterpret_cast<const float4*>(xr + 2 * PAD + w);
                    xw[2 * PAD + 0] = nx.x; xw[2 * PAD + 1] = nx.y; xw[2 * PAD + 2] = nx.z; xw[2 * PAD + 3] = nx.w;
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) {
                        const float4 yv = *reinterpret_cast<const float4*>(yr + (size_t)kk * g.XW + w);
#pragma unroll
                        for (int dx = 0; dx < T; ++dx) {
                            float a = acc[dx][kk];
                            a = fmaf(xw[dx + 0], yv.x, a);
                            a = fmaf(xw[dx + 1], yv.y, a);
                            a = fmaf(xw[dx + 2], yv.z, a);
                            a = fmaf(xw[dx + 3], yv.w, a);
                            acc[dx][kk] = a;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // per-CTA sum over the row slots in slot order (plain read-modify-writes, one slot per barrier round): bitwise reproducible,
    // unlike shared-memory atomics — the padding-0 loss is a 1e-3 residual of O(1) terms and shows one ulp of a partial joint
    for (int sv = 0; sv < nslot; ++sv) {
        if (active && slot == sv) {
            const int dy = rr / K, k1 = rr % K;
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const int k2 = k2base + kk;
                if (k2 < K) {
#pragma unroll
                    for (int dx = 0; dx < T; ++dx) jsm[((k1 * K + k2) * T + dy) * T + dx] += acc[dx][kk];
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < nj; i += IIC_NT) partials[(size_t)blockIdx.x * nj + i] = jsm[i];
}

// ---------------------------------------------------------------------------------------------- forward (generic)
// any K / pad: one J entry per thread iteration, two shared loads per FMA.  Correctness path for exotic shapes.
__global__ void __launch_bounds__(IIC_NT)
iic_joint_generic_kernel(const void* __restrict__ x, const void* __restrict__ y, int dtype, IICGeom g,
                         float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    const int K = g.K, T = g.T, pad = g.pad;
    float* xs = smem;
    float* ys = xs + (size_t)g.HH * K * g.XW;
    const int nj = K * K * T * T;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x, first = 1; tile < g.n_tiles; tile += gridDim.x, first = 0) {
        const int b = tile / (g.tiles_h * g.tiles_w);
        const int trem = tile % (g.tiles_h * g.tiles_w);
        const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * g.TW;
        __syncthreads();
        stage_tile(xs, x, dtype, g, b, h0, w0, pad, g.CO);
        stage_tile(ys, y, dtype, g, b, h0, w0, 0, 0);
        __syncthreads();
        for (int e = tid; e < nj; e += IIC_NT) {
            const int dx = e % T, dy = (e / T) % T, k2 = (e / (T * T)) % K, k1 = e / (T * T * K);
            float s = 0.f;
            for (int h = 0; h < g.TH; ++h) {
                const float* xr = xs + (size_t)((h + dy) * K + k1) * g.XW + g.CO - pad + dx;
                const float* yr = ys + (size_t)(h * K + k2) * g.XW;
                for (int w = 0; w < g.TW; ++w) s = fmaf(xr[w], yr[w], s);
            }
            float* dst = partials + (size_t)blockIdx.x * nj + e;
            *dst = first ? s : (*dst + s);
        }
    }
}

// ---------------------------------------------------------------------------------------------- epilogue
// One CTA.  Sums the per-CTA partial joints in a fixed order (fp64), then discreteMI.py:233-243 / :246-261 and
// :154-165, and the analytic dLoss/dJoint.  All in fp64: 900 elements.
// block = 32 joint entries (lanes) x 32 warps striding over the partials; fixed summation order: deterministic
constexpr int RED_WARPS = 32;
__global__ void __launch_bounds__(32 * RED_WARPS)
iic_reduce_partials_kernel(const float* __restrict__ partials, int n_partials, int nj, double* __restrict__ joint,
                           long long partials_stride, long long joint_stride) {
    __shared__ double acc[RED_WARPS][33];
    partials += (size_t)blockIdx.y * partials_stride;      // blockIdx.y = sub-head of a heads launch (strides 0 otherwise)
    joint += (size_t)blockIdx.y * joint_stride;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    // launched with programmatic stream serialization right behind the joint kernel: everything above overlaps its tail,
    // the partial joints are read only after the dependency has resolved; the epilogue behind THIS kernel is launched the same
    // way (its CTA may become resident now and waits in its own cudaGridDependencySynchronize until this grid is complete)
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;");
    cudaGridDependencySynchronize();
#endif
    double s = 0.0;
    if (i < nj) {
        // fixed order per thread; the loads of a batch of 8 are independent and in flight together (one L2 round trip per batch)
        int p = w;
        for (; p + 7 * RED_WARPS < n_partials; p += 8 * RED_WARPS) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(partials + (size_t)(p + u * RED_WARPS) * nj + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
        for (; p < n_partials; p += RED_WARPS) s += (double)__ldcg(partials + (size_t)p * nj + i);
    }
    acc[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < nj) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < RED_WARPS; ++k) t += acc[k][lane];
        joint[i] = t;      // kept in double: the min-shift of the epilogue amplifies a float32 rounding of J ~sqrt(pixels)-fold
    }
}
// launch helper: programmatic dependent launch (the kernel may start while its predecessor on the stream drains)
static int launch_reduce_partials(const float* partials, int n_partials, int nj, double* joint, cudaStream_t st, int n_heads = 1,
                                  long long partials_stride = 0, long long joint_stride = 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((nj + 31) / 32), (unsigned)n_heads);
    cfg.blockDim = dim3(32 * RED_WARPS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, iic_reduce_partials_kernel, partials, n_partials, nj, joint, partials_stride, joint_stride);
    count_launch();
    if (e != cudaSuccess) { set_error("iic_reduce_partials: %s", cudaGetErrorString(e)); return (int)e; }
    return CY_OK;
}

// Block-wide reductions of the one-CTA epilogue, two butterfly levels (lanes, then the <= 32 warp totals re-read by every warp): the
// kernel is a chain of ~15 dependent phases whose cost is instruction latency per warp, so the second level is 5 shuffle steps,
// not a 32-long serial chain of shared-memory adds.  Fixed order: bitwise reproducible.  red: 64 doubles.
__device__ __noinline__ double block_reduce_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                     // (the previous use of red is over)
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = lane < nw ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// two sums at once (same barriers)
__device__ __noinline__ void block_reduce_sum2(double& a, double& b, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if (lane == 0) { red[w] = a; red[32 + w] = b; }
    __syncthreads();
    a = lane < nw ? red[lane] : 0.0;
    b = lane < nw ? red[32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
}

// sum of v[0..n) by one warp (fixed order per lane, then a butterfly): every lane returns the total
__device__ __forceinline__ double warp_sum_n(const double* v, int n, int lane) {
    double s = 0.0;
    for (int e = lane; e < n; e += 32) s += v[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

__device__ __noinline__ double block_reduce_min(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = lane < nw ? red[lane] : red[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
    return t;
}

// One copy each of the long fp64 sequences (log ~100 instructions, division ~40): the epilogue is ONE CTA that runs straight-line
// code once, and ncu showed it waiting for instruction fetch as much as for anything else (58 k warp instructions in 20 us,
// "no instruction" the second largest stall) — inlining six logs and eight divisions made the kernel mostly cold code.
__device__ __noinline__ double ep_log(double v) { return log(v); }
__device__ __noinline__ double ep_div(double a, double b) { return a / b; }

// dynamic smem: double A[nj] (B matrix, [dd][k1][k2]), P[nj], G[nj], sdisp[TT], colsum[TT*K], rowsum[TT*K], then the marginals'
// log(m + eps) and m / (m + eps) for both ([4][TT*K]): computed once per marginal instead of once per joint entry
constexpr int EPI_THREADS = 1024;
__global__ void __launch_bounds__(EPI_THREADS)
iic_epilogue_kernel(const double* __restrict__ joint, int n_slots, int K, int pad, int symmetric, double lamda, double eps,
                    double n_pixels, float* __restrict__ loss, float* __restrict__ p00, float* __restrict__ p_ij,
                    float* __restrict__ djoint, double* __restrict__ gscratch, long long joint_stride, long long out_stride,
                    long long slot_stride) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[64];
    // programmatic dependent launch behind the partial-joint reduction (or any other producer of `joint`: a kernel that never
    // signals simply completes first): the launch latency and this prologue overlap the producer's tail
#if __CUDA_ARCH__ >= 900
    cudaGridDependencySynchronize();
#endif
    // one CTA per sub-head of a heads launch (gridDim.x = 1 and strides 0 otherwise); p_ij and gscratch are single-head only
    joint += (size_t)blockIdx.x * joint_stride;
    loss += (size_t)blockIdx.x * out_stride;
    p00 += (size_t)blockIdx.x * out_stride;
    if (djoint) djoint += (size_t)blockIdx.x * out_stride;
    const int T = 2 * pad + 1, TT = T * T, KK = K * K, nj = KK * TT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double* Bm = gscratch ? gscratch : sm;      // large K / padding: the arrays live in the caller's workspace
    double* P = Bm + nj;
    double* G = P + nj;
    double* sd = G + nj;
    double* asum = sd + TT;        // [TT][K]  sum over k1 -> function of k2 ("p_i_mat", dim 2)
    double* bsum = asum + TT * K;  // [TT][K]  sum over k2 -> function of k1 ("p_j_mat", dim 3)
    double* la = bsum + TT * K;    // log(asum + eps), asum / (asum + eps), and the same for bsum
    double* ra = la + TT * K;
    double* lb = ra + TT * K;
    double* rb = lb + TT * K;
    const int tid = threadIdx.x, nt = blockDim.x;
    // index splits by the runtime K / T: floor(n / d) = int((n + 0.5) * (1 / d)) is exact in fp32 for n < 2^22 (the host checks
    // K*K*T*T against it) and three instructions instead of the ~25 of an integer division, of which there were 20 per thread
    const float invK = 1.f / (float)K, invTT = 1.f / (float)TT, invKK = 1.f / (float)KK;
    auto qK = [&](int n) { return __float2int_rz(((float)n + 0.5f) * invK); };
    auto qTT = [&](int n) { return __float2int_rz(((float)n + 0.5f) * invTT); };
    auto qKK = [&](int n) { return __float2int_rz(((float)n + 0.5f) * invKK); };
    // i -> (k1, k2, dd) for the external layout [k1][k2][dd] and for the internal layout [dd][k1][k2]
    auto split_ext = [&](int i, int& k1, int& k2, int& dd) { const int t = qTT(i); dd = i - t * TT; k1 = qK(t); k2 = t - k1 * K; };
    auto split_int = [&](int i, int& k1, int& k2, int& dd) { dd = qKK(i); const int t = qK(i); k2 = i - t * K; k1 = t - dd * K; };
    // joint is [k1][k2][dd]; internal layout [dd][k1][k2]
    auto JIDX = [&](int k1, int k2, int dd) { return (k1 * K + k2) * TT + dd; };
    auto PIDX = [&](int dd, int k1, int k2) { return (dd * K + k1) * K + k2; };

    // batch-sharded multi-GPU form: `joint` holds one partial joint per rank ([n_slots][nj], written by the ranks themselves
    // through peer memory); they are summed here in slot order, so every rank forms the identical global joint
    auto JV = [&](int i) {
        double t = joint[i];
        #pragma unroll 1
        for (int sl = 1; sl < n_slots; ++sl) t += joint[(size_t)sl * slot_stride + i];
        return t;
    };
    double total = 1.0;
    if (pad > 0) {
        double mn = 1e300;
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) mn = fmin(mn, JV(i));
        mn = block_reduce_min(mn, red);
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) {
            int k1, k2, dd;
            split_ext(i, k1, k2, dd);
            Bm[PIDX(dd, k1, k2)] = JV(i) - mn + 1e-8;
        }
        __syncthreads();
        #pragma unroll 1
        for (int dd = warp; dd < TT; dd += nwarp) {          // one warp per displacement
            const double s = warp_sum_n(Bm + dd * KK, KK, lane);
            if (lane == 0) sd[dd] = s;
        }
        __syncthreads();
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) Bm[i] = ep_div(Bm[i], sd[qKK(i)]);
        __syncthreads();
        double part = 0.0;
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) {
            int k1, k2, dd;
            split_int(i, k1, k2, dd);
            const double c = symmetric ? 0.5 * (Bm[i] + Bm[PIDX(dd, k2, k1)]) : Bm[i];
            P[i] = c;
            part += c;
        }
        total = block_reduce_sum(part, red);
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) P[i] = ep_div(P[i], total);
    } else {
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) Bm[i] = ep_div(JV(i), n_pixels);   // TT == 1: layouts coincide
        __syncthreads();
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) {
            const int k1 = qK(i), k2 = i - k1 * K;
            P[i] = symmetric ? 0.5 * (Bm[i] + Bm[k2 * K + k1]) : Bm[i];
        }
    }
    __syncthreads();
    // marginals: threads [0, TT*K) sum over k1 (asum), threads [TT*K, 2*TT*K) over k2 (bsum); each also takes the logarithm and
    // the ratio its marginal contributes to every entry of its row / column
    #pragma unroll 1
    for (int i = tid; i < 2 * TT * K; i += nt) {
        const bool second = i >= TT * K;
        const int r = second ? i - TT * K : i, dd = qK(r), k = r - dd * K;
        double m = 0.0;
        #pragma unroll 1
        for (int q = 0; q < K; ++q) m += second ? P[PIDX(dd, k, q)] : P[PIDX(dd, q, k)];
        (second ? bsum : asum)[r] = m;
        (second ? lb : la)[r] = ep_log(m + eps);
        (second ? rb : ra)[r] = ep_div(m, m + eps);
    }
    __syncthreads();
    double lpart = 0.0, gpart = 0.0;
    #pragma unroll 1
    for (int i = tid; i < nj; i += nt) {
        int k1, k2, dd;
        split_int(i, k1, k2, dd);
        const double p = P[i];
        const double lp = ep_log(p + eps), la_ = la[dd * K + k2], lb_ = lb[dd * K + k1];
        lpart += -p * (lp - lamda * la_ - lamda * lb_);
        const double gq = -(lp + ep_div(p, p + eps) - lamda * (la_ + ra[dd * K + k2]) - lamda * (lb_ + rb[dd * K + k1])) / (double)TT;
        G[i] = gq;
        gpart += gq * p;
    }
    block_reduce_sum2(lpart, gpart, red);
    const double L = lpart, gdotP = gpart;
    if (tid == 0) loss[0] = (float)(L / (double)TT);
    #pragma unroll 1
    for (int i = tid; i < KK; i += nt) p00[i] = (float)P[i];
    if (p_ij)
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) p_ij[i] = (float)P[i];
    if (!djoint) return;
    if (pad > 0) {
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) G[i] = ep_div(G[i] - gdotP, total);
        __syncthreads();
        // per displacement: dot = sum gB * B
        #pragma unroll 1
        for (int dd = warp; dd < TT; dd += nwarp) {          // one warp per displacement
            double dot = 0.0;
            #pragma unroll 1
            for (int e = lane; e < KK; e += 32) {
                const int k1 = qK(e), k2 = e - k1 * K;
                const double gb = symmetric ? 0.5 * (G[PIDX(dd, k1, k2)] + G[PIDX(dd, k2, k1)]) : G[PIDX(dd, k1, k2)];
                dot += gb * Bm[PIDX(dd, k1, k2)];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (lane == 0) asum[dd] = dot;   // reuse
        }
        __syncthreads();
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) {
            int k1, k2, dd;
            split_int(i, k1, k2, dd);
            const double gb = symmetric ? 0.5 * (G[i] + G[PIDX(dd, k2, k1)]) : G[i];
            djoint[JIDX(k1, k2, dd)] = (float)ep_div(gb - asum[dd], sd[dd]);
        }
    } else {
        #pragma unroll 1
        for (int i = tid; i < nj; i += nt) {
            const int k1 = qK(i), k2 = i - k1 * K;
            const double gb = symmetric ? 0.5 * (G[i] + G[k2 * K + k1]) : G[i];
            djoint[i] = (float)ep_div(gb, n_pixels);
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
// g tables in shared memory: gA[((k_in*T + dyy)*KP + k_out)*TP + dxx] for phase A (k_in = k1 over x, k_out = k2),
// and the flipped transpose for phase B (k_in = k2 over y, k_out = k1, dy -> 2p-dy, dx -> 2p-dx).
template <int PAD, int KC>
__device__ __forceinline__ void bwd_phase(const float* __restrict__ tile, const float* __restrict__ gtab, const IICGeom& g,
                                          int KP, int TP, int r, int q, int k_out_base, float (&acc)[KC][4]) {
    constexpr int T = 2 * PAD + 1;
    const int K = g.K;
#pragma unroll
    for (int c = 0; c < KC; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    for (int kin = 0; kin < K; ++kin) {
#pragma unroll
        for (int dyy = 0; dyy < T; ++dyy) {
            const float* row = tile + (size_t)((r + dyy) * K + kin) * g.XW + (g.CO - PAD) + 4 * q;
            float xw[4 + 2 * PAD];
#pragma unroll
            for (int e = 0; e < 2 * PAD; ++e) xw[e] = row[e];
            const float4 nx = *reinterpret_cast<const float4*>(row + 2 * PAD);
            xw[2 * PAD + 0] = nx.x; xw[2 * PAD + 1] = nx.y; xw[2 * PAD + 2] = nx.z; xw[2 * PAD + 3] = nx.w;
            const float* gp = gtab + (size_t)((kin * T + dyy) * KP + k_out_base) * TP;
#pragma unroll
            for (int c = 0; c < KC; ++c) {
#pragma unroll
                for (int dxx = 0; dxx < T; ++dxx) {
                    const float gv = gp[c * TP + dxx];
                    acc[c][0] = fmaf(gv, xw[dxx + 0], acc[c][0]);
                    acc[c][1] = fmaf(gv, xw[dxx + 1], acc[c][1]);
                    acc[c][2] = fmaf(gv, xw[dxx + 2], acc[c][2]);
                    acc[c][3] = fmaf(gv, xw[dxx + 3], acc[c][3]);
                }
            }
        }
    }
}

template <int PAD, int KC>
__global__ void __launch_bounds__(IIC_NT)
iic_bwd_kernel(const void* __restrict__ x, const void* __restrict__ y, int dtype, IICGeom g, const float* __restrict__ djoint,
               const float* __restrict__ gscale, void* __restrict__ dx_out, void* __restrict__ dy_out) {
    constexpr int T = 2 * PAD + 1;
    constexpr int TP = (T == 1) ? 1 : ((T <= 4) ? 4 : 8);
    extern __shared__ __align__(16) float smem[];
    const int K = g.K;
    const int nchunk = (K + KC - 1) / KC;
    const int KP = nchunk * KC;
    float* xs = smem;
    float* ys = xs + (size_t)g.HH * K * g.XW;
    float* gA = ys + (size_t)g.HH * K * g.XW;
    float* gB = gA + (size_t)K * T * KP * TP;
    const int tid = threadIdx.x;
    const float scale = gscale[0];

    for (int i = tid; i < K * T * KP * TP; i += blockDim.x) {
        const int dxx = i % TP, ko = (i / TP) % KP, dyy = (i / (TP * KP)) % T, kin = i / (TP * KP * T);
        float a = 0.f, b = 0.f;
        if (dxx < T && ko < K) {
            a = djoint[((kin * K + ko) * T + dyy) * T + dxx] * scale;                       // g[k1=kin, k2=ko, dy, dx]
            b = djoint[((ko * K + kin) * T + (T - 1 - dyy)) * T + (T - 1 - dxx)] * scale;   // g[k1=ko, k2=kin, flipped]
        }
        gA[i] = a;
        gB[i] = b;
    }

    const int qpr = g.TW / 4;                 // 4-pixel groups per tile row
    const int r = tid / qpr, q = tid % qpr;   // blockDim.x == TH * qpr

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int b = tile / (g.tiles_h * g.tiles_w);
        const int trem = tile % (g.tiles_h * g.tiles_w);
        const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * g.TW;
        __syncthreads();
        stage_tile(xs, x, dtype, g, b, h0, w0, PAD, g.CO);
        stage_tile(ys, y, dtype, g, b, h0, w0, PAD, g.CO);
        __syncthreads();
        const int h = h0 + r, w = w0 + 4 * q;
        if (r < g.TH && h < g.H && w < g.W) {
            float acc[KC][4];
            for (int c2 = 0; c2 < nchunk; ++c2) {
                // dL/dy[b, k2, h, w..w+3] = sum_{k1,dy,dx} g[k1,k2,dy,dx] x[b,k1,h+dy-p,w+dx-p]
                bwd_phase<PAD, KC>(xs, gA, g, KP, TP, r, q, c2 * KC, acc);
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const int ko = c2 * KC + c;
                    if (ko < K) {
                        const size_t base = (((size_t)b * K + ko) * g.H + h) * g.W + w;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (w + e < g.W) st_from_float(dy_out, dtype, base + e, acc[c][e]);
                    }
                }
                // dL/dx[b, k1, h, w..w+3] = sum_{k2,dy,dx} g[k1,k2,dy,dx] y[b,k2,h-dy+p,w-dx+p]
                bwd_phase<PAD, KC>(ys, gB, g, KP, TP, r, q, c2 * KC, acc);
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const int ko = c2 * KC + c;
                    if (ko < K) {
                        const size_t base = (((size_t)b * K + ko) * g.H + h) * g.W + w;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (w + e < g.W) st_from_float(dx_out, dtype, base + e, acc[c][e]);
                    }
                }
            }
        }
    }
}

// generic backward: any K / pad; one output element per thread, djoint read from global (L1-resident)
__global__ void iic_bwd_generic_kernel(const void* __restrict__ x, const void* __restrict__ y, int dtype, int B, int K, int H,
                                       int W, int pad, const float* __restrict__ djoint, const float* __restrict__ gscale,
                                       void* __restrict__ dx_out, void* __restrict__ dy_out) {
    const int T = 2 * pad + 1;
    const size_t total = (size_t)B * K * H * W;
    const float scale = gscale[0];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int w = idx % W, h = (idx / W) % H, k = (idx / ((size_t)W * H)) % K, b = idx / ((size_t)W * H * K);
        float sy = 0.f, sx = 0.f;
        for (int ko = 0; ko < K; ++ko)
            for (int dy = 0; dy < T; ++dy)
                for (int dx = 0; dx < T; ++dx) {
                    // dy-map: this element is y[b,k2=k]; partner x[b,k1=ko] at (h+dy-p, w+dx-p)
                    int hh = h + dy - pad, ww = w + dx - pad;
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                        sy = fmaf(djoint[((ko * K + k) * T + dy) * T + dx],
                                  ld_as_float(x, dtype, (((size_t)b * K + ko) * H + hh) * W + ww), sy);
                    // dx-map: this element is x[b,k1=k]; partner y[b,k2=ko] at (h-dy+p, w-dx+p)
                    hh = h - dy + pad; ww = w - dx + pad;
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                        sx = fmaf(djoint[((k * K + ko) * T + dy) * T + dx],
                                  ld_as_float(y, dtype, (((size_t)b * K + ko) * H + hh) * W + ww), sx);
                }
        st_from_float(dy_out, dtype, idx, sy * scale);
        st_from_float(dx_out, dtype, idx, sx * scale);
    }
}

// ---------------------------------------------------------------------------------------------- host side
static int pick_kc(int K) {
    const int opts[] = {4, 5, 8, 10, 16, 20};
    for (int o : opts)
        if (K <= o) return o;
    return 8;   // chunked
}

static int sm_count_cached() { return device_sm_count(); }

static int pick_tw(int W, int maxtw) {
    // multiple of 4, <= maxtw, minimising padded width then number of tiles
    int best = 4, best_cost = 1 << 30;
    for (int tw = 4; tw <= maxtw; tw += 4) {
        const int tiles = (W + tw - 1) / tw;
        const int cost = tiles * tw * 64 + tiles;   // padded pixels dominate
        if (cost <= best_cost) { best_cost = cost; best = tw; }
    }
    return best;
}

static IICGeom make_geom(int B, int K, int H, int W, int pad, int TH, int TW) {
    IICGeom g;
    g.B = B; g.K = K; g.H = H; g.W = W; g.pad = pad; g.T = 2 * pad + 1;
    g.TH = TH; g.TW = TW;
    g.CO = iic_col_origin(pad);
    int xw = g.CO + TW + pad;
    xw = (xw + 3) / 4 * 4;
    if (((xw / 4) & 1) == 0) xw += 4;       // XW/4 odd: conflict-free role rows
    g.XW = xw;
    g.HH = TH + 2 * pad;
    g.tiles_h = (H + TH - 1) / TH;
    g.tiles_w = (W + TW - 1) / TW;
    g.n_tiles = B * g.tiles_h * g.tiles_w;
    return g;
}

struct JointPlan {
    IICGeom g;
    int kc;          // 0 -> generic kernel
    size_t smem;
    int grid;
};

static JointPlan plan_joint(int B, int K, int H, int W, int pad) {
    JointPlan p;
    const int T = 2 * pad + 1;
    const int kc = pick_kc(K);
    const int nchunk = (K + kc - 1) / kc;
    const bool fast = pad <= 3 && T * K * nchunk <= IIC_NT && T * kc <= 100;
    const int TW = pick_tw(W, 64);
    int TH = 8;
    if (fast) {
        const int nslot = IIC_NT / (T * K * nchunk);
        TH = nslot < 4 ? 4 : (nslot > 16 ? 16 : nslot);
    }
    if (TH > H) TH = H;
    const size_t budget = 100 * 1024;
    for (;;) {
        p.g = make_geom(B, K, H, W, pad, TH, TW);
        p.smem = ((size_t)p.g.HH * K * p.g.XW + (size_t)(TH * K + kc) * p.g.XW + (size_t)K * K * T * T) * sizeof(float);
        if (p.smem <= budget || TH <= 1) break;
        TH = TH / 2;
    }
    p.kc = fast ? kc : 0;
    const int sms = sm_count_cached();
    p.grid = p.g.n_tiles < 2 * sms ? p.g.n_tiles : 2 * sms;
    if (p.grid < 1) p.grid = 1;
    return p;
}

// iic_tma.cu (TMA + packed-FMA fast path; returns CY_ERR_UNSUPPORTED for shapes it does not take)
int iic_joint_tma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, float* partials, int* n_partials,
                  cudaStream_t st);
int iic_bwd_tma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
                const float* gscale, void* dx, void* dy, cudaStream_t st);

// iic_mma.cu (tensor-pipe path: bf16 hi/lo split mma.sync fed from TMA boxes; CY_ERR_UNSUPPORTED for shapes it does not take)
int iic_joint_mma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, float* partials, int* n_partials,
                  cudaStream_t st);
int iic_bwd_mma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
                const float* gscale, void* dx, void* dy, cudaStream_t st);

// iic_bwd_tc.cu (tcgen05 adjoint for padding = 1, K <= 16, fp32: pixels on the TMEM lanes, A slots written by converter warps,
// 18 small MMAs per output row; CY_ERR_UNSUPPORTED for shapes it does not take).  The default adjoint.
int iic_joint_mma_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                        float* partials, int* n_partials, cudaStream_t st);
int iic_bwd_tc_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                     const float* djoint, long long dj_stride, const float* gscale, void* const* dxs, void* const* dys,
                     float softmax_inv_T, cudaStream_t st);
int iic_bwd_tc(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
               const float* gscale, void* dx, void* dy, cudaStream_t st);

static bool tc_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("CY_IIC_TC");      // CY_IIC_TC=0 pins the mma.sync adjoint (A/B measurements)
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

static bool mma_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("CY_IIC_MMA");     // CY_IIC_MMA=0 pins the CUDA-core kernels (A/B measurements)
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

static bool tma_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("CY_IIC_TMA");     // CY_IIC_TMA=0 pins the SIMT kernels (A/B measurements)
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

size_t iic_workspace_bytes(int B, int K, int H, int W, int pad) {
    (void)B; (void)H; (void)W;
    const int T = 2 * pad + 1;
    return (size_t)2 * sm_count_cached() * K * K * T * T * sizeof(float);     // one partial joint per persistent CTA
}

template <int PAD, int KC>
static int launch_joint(const void* x, const void* y, int dtype, const JointPlan& p, float* partials, cudaStream_t st) {
    auto k = iic_joint_kernel<PAD, KC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) { set_error("iic_joint smem attr (%zu B): %s", p.smem, cudaGetErrorString(e)); return (int)e; }
    k<<<p.grid, IIC_NT, p.smem, st>>>(x, y, dtype, p.g, partials);
    CY_CHECK_LAUNCH("iic_joint");
    return CY_OK;
}

#define IIC_DISPATCH_KC(FN, PADV, ...)                                   \
    switch (kc) {                                                        \
        case 4: return FN<PADV, 4>(__VA_ARGS__);                         \
        case 5: return FN<PADV, 5>(__VA_ARGS__);                         \
        case 8: return FN<PADV, 8>(__VA_ARGS__);                         \
        case 10: return FN<PADV, 10>(__VA_ARGS__);                       \
        case 16: return FN<PADV, 16>(__VA_ARGS__);                       \
        case 20: return FN<PADV, 20>(__VA_ARGS__);                       \
    }

#define IIC_DISPATCH(FN, ...)                                            \
    switch (pad) {                                                       \
        case 0: IIC_DISPATCH_KC(FN, 0, __VA_ARGS__) break;               \
        case 1: IIC_DISPATCH_KC(FN, 1, __VA_ARGS__) break;               \
        case 2: IIC_DISPATCH_KC(FN, 2, __VA_ARGS__) break;               \
        case 3: IIC_DISPATCH_KC(FN, 3, __VA_ARGS__) break;               \
    }

static int dispatch_joint(int pad, int kc, const void* x, const void* y, int dtype, const JointPlan& p, float* partials,
                          cudaStream_t st) {
    IIC_DISPATCH(launch_joint, x, y, dtype, p, partials, st)
    set_error("iic_joint: no instantiation for pad=%d kc=%d", pad, kc);
    return CY_ERR_UNSUPPORTED;
}

int iic_joint(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, double* joint, void* workspace,
              size_t workspace_bytes, cudaStream_t st) {
    const JointPlan p = plan_joint(B, K, H, W, pad);
    const int T = 2 * pad + 1, nj = K * K * T * T;
    const size_t need = (size_t)p.grid * nj * sizeof(float);
    CY_CHECK_ARG(workspace && workspace_bytes >= need, "iic_joint: workspace %zu < %zu", workspace_bytes, need);
    float* partials = reinterpret_cast<float*>(workspace);
    int rc;
    if (mma_enabled()) {
        int np = 0;
        rc = iic_joint_mma(x, y, dtype, B, K, H, W, pad, partials, &np, st);
        if (rc == CY_OK) {
            return launch_reduce_partials(partials, np, nj, joint, st);
        }
        if (rc != CY_ERR_UNSUPPORTED) return rc;
    }
    if (tma_enabled()) {
        int np = 0;
        rc = iic_joint_tma(x, y, dtype, B, K, H, W, pad, partials, &np, st);
        if (rc == CY_OK) {
            return launch_reduce_partials(partials, np, nj, joint, st);
        }
        if (rc != CY_ERR_UNSUPPORTED) return rc;
    }
    if (p.kc) {
        rc = dispatch_joint(pad, p.kc, x, y, dtype, p, partials, st);
    } else {
        cudaError_t e = cudaFuncSetAttribute(iic_joint_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) { set_error("iic_joint_generic smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        iic_joint_generic_kernel<<<p.grid, IIC_NT, p.smem, st>>>(x, y, dtype, p.g, partials);
        CY_CHECK_LAUNCH("iic_joint_generic");
        rc = CY_OK;
    }
    if (rc != CY_OK) return rc;
    return launch_reduce_partials(partials, p.grid, nj, joint, st);
}

// n_heads (x, y) pairs of one shape: head s writes joint + s * joint_stride (doubles).  workspace = n_heads *
// iic_workspace_bytes().  One joint launch + one reduction launch when the tensor-core kernel takes the shape (heads in
// chunks of 8), head by head otherwise.
int iic_joint_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                    double* joint, long long joint_stride, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t per = iic_workspace_bytes(B, K, H, W, pad);
    CY_CHECK_ARG(n_heads >= 1 && workspace && workspace_bytes >= per * (size_t)n_heads, "iic_joint_heads: workspace %zu < %zu",
                 workspace_bytes, per * (size_t)n_heads);
    const int T = 2 * pad + 1, nj = K * K * T * T;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    for (int s0 = 0; s0 < n_heads;) {
        const int n = n_heads - s0 < 8 ? n_heads - s0 : 8;
        int rc = CY_ERR_UNSUPPORTED;
        if (mma_enabled() && n > 1) {
            int np = 0;
            float* partials = reinterpret_cast<float*>(ws + per * s0);
            rc = iic_joint_mma_heads(xs + s0, ys + s0, n, dtype, B, K, H, W, pad, partials, &np, st);
            if (rc == CY_OK)
                rc = launch_reduce_partials(partials, np, nj, joint + (size_t)s0 * joint_stride, st, n, (long long)np * nj, joint_stride);
            if (rc != CY_OK && rc != CY_ERR_UNSUPPORTED) return rc;
        }
        if (rc == CY_OK) { s0 += n; continue; }
        for (int s = s0; s < s0 + n; ++s) {
            rc = iic_joint(xs[s], ys[s], dtype, B, K, H, W, pad, joint + (size_t)s * joint_stride, ws + per * s, per, st);
            if (rc != CY_OK) return rc;
        }
        s0 += n;
    }
    return CY_OK;
}

static size_t epilogue_scratch_doubles(int K, int pad) {
    const int T = 2 * pad + 1, TT = T * T, nj = K * K * TT;
    return (size_t)3 * nj + TT + 6 * TT * K;
}

size_t iic_epilogue_workspace_bytes(int K, int pad) {
    const size_t b = epilogue_scratch_doubles(K, pad) * sizeof(double);
    return b > 160 * 1024 ? b : 0;      // small problems keep everything in shared memory
}

static int iic_epilogue_impl(const double* joint, int n_slots, int K, int pad, int symmetric, float lamda, float eps, double n_pixels,
                             float* loss, float* p00, float* p_ij, float* djoint, void* workspace, size_t workspace_bytes,
                             cudaStream_t st, int n_heads, long long joint_stride, long long out_stride, long long slot_stride = 0) {
    // (the kernel splits indices through fp32 reciprocals: exact below 2^22 entries)
    CY_CHECK_ARG((long long)K * K * (2 * pad + 1) * (2 * pad + 1) < (1LL << 22), "iic_epilogue: K=%d pad=%d: joint too large", K, pad);
    size_t smem = epilogue_scratch_doubles(K, pad) * sizeof(double);
    double* gscratch = nullptr;
    if (iic_epilogue_workspace_bytes(K, pad)) {
        CY_CHECK_ARG(workspace && workspace_bytes >= smem, "iic_epilogue: K=%d pad=%d needs a %zu B workspace", K, pad, smem);
        gscratch = reinterpret_cast<double*>(workspace);
        smem = 0;
    }
    static SmemAttrCache attr;
    if (attr.need(smem)) {
        cudaError_t e = cudaFuncSetAttribute(iic_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("iic_epilogue smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set(smem);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_heads);
    cfg.blockDim = dim3(EPI_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute lattr[1];
    lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    lattr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = lattr;
    cfg.numAttrs = 1;
    const long long sstride = slot_stride > 0 ? slot_stride : (long long)K * K * (2 * pad + 1) * (2 * pad + 1);
    cudaError_t le = cudaLaunchKernelEx(&cfg, iic_epilogue_kernel, joint, n_slots, K, pad, symmetric, (double)lamda, (double)eps, n_pixels,
                                        loss, p00, p_ij, djoint, gscratch, joint_stride, out_stride, sstride);
    count_launch();
    if (le != cudaSuccess) { set_error("iic_epilogue: %s", cudaGetErrorString(le)); return (int)le; }
    return CY_OK;
}

int iic_epilogue(const double* joint, int n_slots, int K, int pad, int symmetric, float lamda, float eps, double n_pixels, float* loss,
                 float* p00, float* p_ij, float* djoint, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    return iic_epilogue_impl(joint, n_slots, K, pad, symmetric, lamda, eps, n_pixels, loss, p00, p_ij, djoint, workspace,
                             workspace_bytes, st, 1, 0, 0);
}

// n_heads epilogues in one launch (one CTA each): head s reads joint + s * joint_stride (doubles; [n_slots][nj] inside) and writes
// loss / p00 / djoint + s * out_stride (floats); slot sl of a head lies slot_stride doubles further (0: K*K*T*T, slots back to
// back).  Shapes whose arrays need the global scratch run head by head.
int iic_epilogue_heads(const double* joint, long long joint_stride, long long slot_stride, int n_heads, int n_slots, int K, int pad,
                       int symmetric, float lamda, float eps, double n_pixels, float* loss, float* p00, float* djoint, long long out_stride,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
    CY_CHECK_ARG(n_heads >= 1, "iic_epilogue_heads: n_heads=%d", n_heads);
    if (iic_epilogue_workspace_bytes(K, pad) == 0)
        return iic_epilogue_impl(joint, n_slots, K, pad, symmetric, lamda, eps, n_pixels, loss, p00, nullptr, djoint, nullptr, 0, st,
                                 n_heads, joint_stride, out_stride, slot_stride);
    for (int s = 0; s < n_heads; ++s) {
        const int rc = iic_epilogue_impl(joint + s * joint_stride, n_slots, K, pad, symmetric, lamda, eps, n_pixels, loss + s * out_stride,
                                         p00 + s * out_stride, nullptr, djoint ? djoint + s * out_stride : nullptr, workspace,
                                         workspace_bytes, st, 1, 0, 0, slot_stride);
        if (rc != CY_OK) return rc;
    }
    return CY_OK;
}

struct BwdPlan {
    IICGeom g;
    int kc;
    size_t smem;
    int grid, threads;
};

template <int PAD, int KC>
static int launch_bwd(const void* x, const void* y, int dtype, const BwdPlan& p, const float* djoint, const float* gscale,
                      void* dx, void* dy, cudaStream_t st) {
    auto k = iic_bwd_kernel<PAD, KC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) { set_error("iic_bwd smem attr (%zu B): %s", p.smem, cudaGetErrorString(e)); return (int)e; }
    k<<<p.grid, p.threads, p.smem, st>>>(x, y, dtype, p.g, djoint, gscale, dx, dy);
    CY_CHECK_LAUNCH("iic_bwd");
    return CY_OK;
}

static int dispatch_bwd(int pad, int kc, const void* x, const void* y, int dtype, const BwdPlan& p, const float* djoint,
                        const float* gscale, void* dx, void* dy, cudaStream_t st) {
    IIC_DISPATCH(launch_bwd, x, y, dtype, p, djoint, gscale, dx, dy, st)
    set_error("iic_bwd: no instantiation for pad=%d kc=%d", pad, kc);
    return CY_ERR_UNSUPPORTED;
}

int iic_bwd(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
            const float* gscale, void* dx, void* dy, cudaStream_t st) {
    const int T = 2 * pad + 1;
    if (mma_enabled() && tc_enabled()) {
        const int rc = iic_bwd_tc(x, y, dtype, B, K, H, W, pad, djoint, gscale, dx, dy, st);
        if (rc != CY_ERR_UNSUPPORTED) return rc;
    }
    if (mma_enabled()) {
        const int rc = iic_bwd_mma(x, y, dtype, B, K, H, W, pad, djoint, gscale, dx, dy, st);
        if (rc != CY_ERR_UNSUPPORTED) return rc;
    }
    if (tma_enabled()) {
        const int rc = iic_bwd_tma(x, y, dtype, B, K, H, W, pad, djoint, gscale, dx, dy, st);
        if (rc != CY_ERR_UNSUPPORTED) return rc;
    }
    if (pad <= 3) {
        BwdPlan p;
        p.kc = pick_kc(K);
        const int nchunk = (K + p.kc - 1) / p.kc, KP = nchunk * p.kc;
        const int TP = (T == 1) ? 1 : ((T <= 4) ? 4 : 8);
        const int TW = W >= 32 ? 32 : (W + 3) / 4 * 4;
        int TH = IIC_NT / (TW / 4);
        if (TH > H) TH = H;
        const size_t budget = 110 * 1024;
        for (;;) {
            p.g = make_geom(B, K, H, W, pad, TH, TW);
            p.smem = ((size_t)2 * p.g.HH * K * p.g.XW + (size_t)2 * K * T * KP * TP) * sizeof(float);
            if (p.smem <= budget || TH <= 1) break;
            TH /= 2;
        }
        if (p.smem <= 200 * 1024) {
            p.threads = TH * (TW / 4);
            if (p.threads < 32) p.threads = 32;
            const int sms = sm_count_cached();
            p.grid = p.g.n_tiles < 4 * sms ? p.g.n_tiles : 4 * sms;
            return dispatch_bwd(pad, p.kc, x, y, dtype, p, djoint, gscale, dx, dy, st);
        }
    }
    const size_t total = (size_t)B * K * H * W;
    const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    iic_bwd_generic_kernel<<<grid, 256, 0, st>>>(x, y, dtype, B, K, H, W, pad, djoint, gscale, dx, dy);
    CY_CHECK_LAUNCH("iic_bwd_generic");
    return CY_OK;
}

// in place: g <- p * (g - sum_k p_k g_k) * inv_T per pixel  (the backward of p = softmax(logits / T) over the K planes);
// one thread per pixel of both maps, coalesced along W.  The fused form lives in the tcgen05 adjoint's epilogue; this is the
// route of the shapes that kernel does not take.
__global__ void __launch_bounds__(256)
iic_softmax_bwd_kernel(const void* __restrict__ px, const void* __restrict__ py, void* gx, void* gy, int dtype,
                       int K, long long plane, long long n_pix, float inv_T) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * n_pix; idx += (long long)gridDim.x * blockDim.x) {
        const bool second = idx >= n_pix;
        const long long t = second ? idx - n_pix : idx;
        const void* p = second ? py : px;
        void* g = second ? gy : gx;
        const size_t base = (size_t)(t / plane) * K * plane + (size_t)(t % plane);
        // (g is read AND written here: plain loads, not the read-only path of ld_as_float)
        auto ld_g = [&](size_t e) {
            if (dtype == CY_F32) return reinterpret_cast<const float*>(g)[e];
            if (dtype == CY_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[e]);
            return __half2float(reinterpret_cast<const __half*>(g)[e]);
        };
        float dot = 0.f;
        for (int k = 0; k < K; ++k) dot = fmaf(ld_as_float(p, dtype, base + (size_t)k * plane), ld_g(base + (size_t)k * plane), dot);
        for (int k = 0; k < K; ++k) {
            const size_t e = base + (size_t)k * plane;
            st_from_float(g, dtype, e, ld_as_float(p, dtype, e) * (ld_g(e) - dot) * inv_T);
        }
    }
}

// n_heads adjoints of one shape: ONE tcgen05 launch per chunk of 8 heads when the shape is eligible, head by head otherwise.
// softmax_inv_T != 0: xs / ys are softmax(logits / T) over the K planes and dxs / dys receive dL/dlogits (softmax backward fused
// into the adjoint's epilogue, or applied in place behind the other adjoint kernels).
int iic_bwd_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                  const float* djoint, long long dj_stride, const float* gscale, void* const* dxs, void* const* dys, float softmax_inv_T,
                  cudaStream_t st) {
    CY_CHECK_ARG(n_heads >= 1, "iic_bwd_heads: n_heads=%d", n_heads);
    for (int s0 = 0; s0 < n_heads;) {
        const int n = n_heads - s0 < 8 ? n_heads - s0 : 8;
        int rc = CY_ERR_UNSUPPORTED;
        if ((n > 1 || softmax_inv_T != 0.f) && mma_enabled() && tc_enabled()) {
            rc = iic_bwd_tc_heads(xs + s0, ys + s0, n, dtype, B, K, H, W, pad, djoint + (size_t)s0 * dj_stride, dj_stride, gscale,
                                  dxs + s0, dys + s0, softmax_inv_T, st);
            if (rc != CY_OK && rc != CY_ERR_UNSUPPORTED) return rc;
        }
        if (rc != CY_OK)
            for (int s = s0; s < s0 + n; ++s) {
                rc = iic_bwd(xs[s], ys[s], dtype, B, K, H, W, pad, djoint + (size_t)s * dj_stride, gscale, dxs[s], dys[s], st);
                if (rc != CY_OK) return rc;
                if (softmax_inv_T != 0.f) {
                    const long long plane = (long long)H * W, n_pix = (long long)B * plane;
                    const long long want = (2 * n_pix + 255) / 256;
                    const int grid = (int)(want < 8LL * sm_count_cached() ? want : 8LL * sm_count_cached());
                    iic_softmax_bwd_kernel<<<grid, 256, 0, st>>>(xs[s], ys[s], dxs[s], dys[s], dtype, K, plane, n_pix, softmax_inv_T);
                    CY_CHECK_LAUNCH("iic_softmax_bwd");
                }
            }
        s0 += n;
    }
    return CY_OK;
}

}  // namespace cy
