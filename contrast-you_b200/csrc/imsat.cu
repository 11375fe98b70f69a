// IMSAT entropies in one streaming pass (reference: contrastyou/losses/discreteMI.py:275-297 imsat_loss / the
// marginal + conditional entropy pair, entropy_criterion = Entropy(reduction="none", eps=1e-8) of semi_seg/hooks/midl.py:13).
//
//   prediction [N, K, S] (S = product of the trailing dims, 1 for classification) — simplex over K per (n, s)
//   conditional = mean_{n,s} ( -sum_k p log(p + eps) )                         "mean entropy of the predictions"
//   marginal    = -sum_k q_k log(q_k + eps),  q_k = mean_{n,s} p[n,k,s]        "entropy of the mean prediction"
//
// HBM-bound: the forward reads the map once (per-block partial sums -> fixed-order reduction: deterministic), the backward
// reads it once and writes the gradient once; the reference's eager graph makes ~10 passes (moveaxis copies, log, mul, sum,
// mean, and their autograd).
#include "common.cuh"

namespace cy {

constexpr int IMSAT_THREADS = 256;
constexpr int IMSAT_MAXK = 64;      // per-class partial sums live in shared memory / registers up to this many classes

// partials [grid][K + 1]: [0..K) class sums of the block's pixels, [K] its entropy sum
__global__ void __launch_bounds__(IMSAT_THREADS)
imsat_partial_kernel(const void* __restrict__ pred, int dtype, int64_t N, int K, int64_t S, float eps, double* __restrict__ partials) {
    __shared__ double sh[IMSAT_MAXK + 1][IMSAT_THREADS / 32];
    const int64_t M = N * S;
    double cls[IMSAT_MAXK];
#pragma unroll 1
    for (int k = 0; k < K; ++k) cls[k] = 0.0;
    double ent = 0.0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = m / S, s = m % S;
        const size_t base = (size_t)n * K * S + s;
        float e = 0.f;
        for (int k = 0; k < K; ++k) {
            const float p = ld_as_float(pred, dtype, base + (size_t)k * S);
            e -= p * logf(p + eps);
            cls[k] += (double)p;
        }
        ent += (double)e;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k <= K; ++k) {
        double v = k < K ? cls[k] : ent;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][w] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= K; k += blockDim.x) {
        double t = 0.0;
        for (int i = 0; i < IMSAT_THREADS / 32; ++i) t += sh[k][i];
        partials[(size_t)blockIdx.x * (K + 1) + k] = t;
    }
}

// out2 = {marginal, conditional}; q [K] = class means (kept for the backward)
__global__ void imsat_final_kernel(const double* __restrict__ partials, int nblk, int K, double M, float eps, float* __restrict__ out2,
                                   float* __restrict__ q) {
    __shared__ double tot[IMSAT_MAXK + 1];
    for (int k = threadIdx.x; k <= K; k += blockDim.x) {
        double t = 0.0;
        for (int b = 0; b < nblk; ++b) t += partials[(size_t)b * (K + 1) + k];
        tot[k] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double marg = 0.0;
        for (int k = 0; k < K; ++k) {
            const double qk = tot[k] / M;
            q[k] = (float)qk;
            marg -= qk * log(qk + (double)eps);
        }
        out2[0] = (float)marg;
        out2[1] = (float)(tot[K] / M);
    }
}

// grad[n,k,s] = g_cond * (-(log(p+eps) + p/(p+eps))) / M  +  g_marg * (-(log(q_k+eps) + q_k/(q_k+eps))) / M
__global__ void __launch_bounds__(IMSAT_THREADS)
imsat_bwd_kernel(const void* __restrict__ pred, int dtype, int64_t N, int K, int64_t S, float eps, const float* __restrict__ q,
                 const float* __restrict__ g2, void* __restrict__ grad) {
    __shared__ float dq[IMSAT_MAXK];
    const int64_t M = N * S;
    const float inv_m = 1.f / (float)M;
    if (threadIdx.x < K) {
        const float qk = q[threadIdx.x];
        dq[threadIdx.x] = -(logf(qk + eps) + qk / (qk + eps)) * g2[0] * inv_m;
    }
    __syncthreads();
    const float gc = g2[1] * inv_m;
    const int64_t total = N * (int64_t)K * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)((i / S) % K);
        const float p = ld_as_float(pred, dtype, (size_t)i);
        st_from_float(grad, dtype, (size_t)i, dq[k] - gc * (logf(p + eps) + p / (p + eps)));
    }
}

static int imsat_grid(int64_t work) {
    const int64_t want = (work + IMSAT_THREADS - 1) / IMSAT_THREADS;
    const int64_t cap = (int64_t)device_sm_count() * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

size_t imsat_workspace_bytes(int K) { return (size_t)device_sm_count() * 8 * (K + 1) * sizeof(double); }

int imsat_fwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, float* out2, float* q, void* workspace,
              size_t workspace_bytes, cudaStream_t st) {
    CY_CHECK_ARG(K >= 1 && K <= IMSAT_MAXK, "cy_imsat: K=%d outside [1, %d]", K, IMSAT_MAXK);
    const int grid = imsat_grid(N * S);
    CY_CHECK_ARG(workspace && workspace_bytes >= (size_t)grid * (K + 1) * sizeof(double), "cy_imsat_fwd: workspace too small");
    double* partials = reinterpret_cast<double*>(workspace);
    imsat_partial_kernel<<<grid, IMSAT_THREADS, 0, st>>>(pred, dtype, N, K, S, eps, partials);
    CY_CHECK_LAUNCH("imsat_partial");
    imsat_final_kernel<<<1, 128, 0, st>>>(partials, grid, K, (double)(N * S), eps, out2, q);
    CY_CHECK_LAUNCH("imsat_final");
    return CY_OK;
}

int imsat_bwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, const float* q, const float* g2, void* grad,
              cudaStream_t st) {
    CY_CHECK_ARG(K >= 1 && K <= IMSAT_MAXK, "cy_imsat: K=%d outside [1, %d]", K, IMSAT_MAXK);
    imsat_bwd_kernel<<<imsat_grid(N * K * S), IMSAT_THREADS, 0, st>>>(pred, dtype, N, K, S, eps, q, g2, grad);
    CY_CHECK_LAUNCH("imsat_bwd");
    return CY_OK;
}

}  // namespace cy
