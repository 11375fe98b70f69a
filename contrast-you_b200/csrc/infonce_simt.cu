// InfoNCE / SupCon family on CUDA cores (fp32 tiles).  This is the general path of libcontrastyou_b200: every
// variant (SupConLoss1 with / without exclude_other_pos, SelfPacedSupConLoss hard / soft), both mask sources
// (integer labels, explicit [n,n] codes), any input dtype, d <= 256.  The tcgen05 path (infonce_tc.cu) covers the
// large-N bf16 label case; both share the row-statistics layout (contrastyou_b200.h, CY_STAT_*) and finalize.
//
// Math (reference contrastyou/losses/contrastive.py:51-100, :136-204), with the fixed shift
// m = 1/t (rows are unit vectors, so max_ij S_ij = 1/t up to rounding; the shift cancels except inside the 1e-16
// guards where the difference is far below fp32 resolution):
//   L_ij = (z_i.z_j - 1)/t,  E_ij = exp(L_ij),  P / Neg from labels or codes, diagonal removed
//   pass 1: posE_i = sum P E, negE_i = sum Neg E, c_i = sum P, nc_i = sum Neg, posL_i = sum P L
//   SUPCON          loss = -(1/N) sum_i [ posL_i/c_i - log(posE_i+negE_i+1e-16) ]
//   EXCLUDE  pass 2 loss = -(1/N) sum_i (1/c_i) sum_j P_ij [ L_ij - log(E_ij + A_i + 1e-16) ],  A_i = negE_i/(r_i+1e-4)
//   SELFPACED pass 2 loss = -(1/N) sum_i (1/c_i) sum_j P_ij w_ij (L_ij - logden_i),  w from (L_ij - logden_i), gamma
//   backward: dZ_i = gscale (1/t)(1/N) sum_k (G_ik + G_ki) z_k
#include "common.cuh"

namespace cy {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;
constexpr int DMAX = 256;  // backward keeps whole rows in shared memory

__device__ __forceinline__ float sp_weight(int variant, float logp, float gamma) {
    // contrastive.py:197-204 (evaluated for positive pairs only; elsewhere max(w, 1-P) = 1 and is unused)
    if (variant == CY_SELFPACED_HARD) return (-logp <= gamma) ? 1.f : 0.f;
    return fmaxf(1.f + logp / gamma, 0.f);
}

// ------------------------------------------------------------------------------------------------- forward
template <int VARIANT, int PASS, bool CODES>
__global__ void __launch_bounds__(NT)
infonce_fwd_simt_kernel(const void* __restrict__ z, int dtype, int64_t N, int d, int64_t ldz,
                        const int32_t* __restrict__ labels, const uint8_t* __restrict__ codes, int64_t n,
                        int64_t row_begin, int64_t row_end, float inv_t, float gamma, float* __restrict__ stats) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ int32_t lab_i[BM], lab_j[BN];
    __shared__ float aux_i[BM];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = row_begin + (int64_t)blockIdx.x * BM;

    if (tid < BM) {
        const int64_t i = i0 + tid;
        lab_i[tid] = (!CODES && i < row_end) ? labels[i] : 0;
        if (PASS == 2) {
            float a = 0.f;
            if (i < row_end) a = stats[(VARIANT == CY_SUPCON_EXCLUDE ? CY_STAT_AUX : CY_STAT_LOGDEN) * N + i];
            aux_i[tid] = a;
        }
    }

    float r0[4] = {0.f, 0.f, 0.f, 0.f}, r1[4] = {0.f, 0.f, 0.f, 0.f}, r2[4] = {0.f, 0.f, 0.f, 0.f};
    float r3[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {0.f, 0.f, 0.f, 0.f};

    const int lr = tid >> 2, lk = (tid & 3) * 4;  // tile loader: row lr, 4 consecutive k starting at lk

    for (int64_t j0 = 0; j0 < N; j0 += BN) {
        __syncthreads();
        if (!CODES && tid < BN) lab_j[tid] = (j0 + tid < N) ? labels[j0 + tid] : 0;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

        for (int k0 = 0; k0 < d; k0 += BK) {
            {
                const int64_t ia = i0 + lr, jb = j0 + lr;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k0 + lk + e;
                    As[lk + e][lr] = (ia < row_end && k < d) ? ld_as_float(z, dtype, ia * ldz + k) : 0.f;
                    Bs[lk + e][lr] = (jb < N && k < d) ? ld_as_float(z, dtype, jb * ldz + k) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) av[a] = As[kk][ty + 16 * a];
#pragma unroll
                for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx + 16 * b];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            }
            __syncthreads();
        }

#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int il = ty + 16 * a;
            const int64_t i = i0 + il;
            if (i >= row_end) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int jl = tx + 16 * b;
                const int64_t j = j0 + jl;
                if (j >= N) continue;
                const int bits = CODES ? pair_bits_codes(codes, n, i, j) : pair_bits_labels(lab_i[il], lab_j[jl], i != j);
                if (bits == 0) continue;
                const float L = fmaf(acc[a][b], inv_t, -inv_t);
                const float E = __expf(L);
                if (PASS == 1) {
                    if (bits & 1) { r0[a] += E; r2[a] += 1.f; r4[a] += L; }
                    else          { r1[a] += E; r3[a] += 1.f; }
                } else if (bits & 1) {
                    if (VARIANT == CY_SUPCON_EXCLUDE) {
                        const float den = E + aux_i[il] + 1e-16f;
                        r0[a] += L - __logf(den);
                        r1[a] += 1.f / den;
                    } else {
                        const float logp = L - aux_i[il];
                        const float w = sp_weight(VARIANT, logp, gamma);
                        r0[a] += w * logp;
                        r1[a] += w;
                    }
                }
            }
        }
    }

    // reduce over the 16 lanes (tx) that share the rows
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            r0[a] += __shfl_xor_sync(0xffffffffu, r0[a], o);
            r1[a] += __shfl_xor_sync(0xffffffffu, r1[a], o);
            if (PASS == 1) {
                r2[a] += __shfl_xor_sync(0xffffffffu, r2[a], o);
                r3[a] += __shfl_xor_sync(0xffffffffu, r3[a], o);
                r4[a] += __shfl_xor_sync(0xffffffffu, r4[a], o);
            }
        }
        const int64_t i = i0 + ty + 16 * a;
        if (tx == 0 && i < row_end) {
            if (PASS == 1) {
                stats[CY_STAT_POSE * N + i] = r0[a];
                stats[CY_STAT_AUX * N + i] = r1[a];
                stats[CY_STAT_INVC * N + i] = r2[a];
                stats[CY_STAT_NEGC * N + i] = r3[a];
                stats[CY_STAT_POSL * N + i] = r4[a];
            } else {
                stats[CY_STAT_POSL * N + i] = r0[a];
                stats[CY_STAT_SW * N + i] = r1[a];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------- backward
struct BwdSmem {
    float zi[BM][DMAX + 1];
    float zj[BN][DMAX + 1];
    float w[BM][BN + 1];
    float st_j[4][BN];   // logden, invc, coef, aux of the column rows
    float st_i[4][BM];
    int32_t lab_i[BM], lab_j[BN];
};

template <int VARIANT, bool CODES>
__global__ void __launch_bounds__(NT)
infonce_bwd_simt_kernel(const void* __restrict__ z, int dtype, int64_t N, int d, int64_t ldz,
                        const int32_t* __restrict__ labels, const uint8_t* __restrict__ codes, int64_t n,
                        int64_t row_begin, int64_t row_end, float inv_t, float gamma, const float4* __restrict__ xstat,
                        const float* __restrict__ gscale, void* __restrict__ dz, int64_t lddz) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem& s = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = row_begin + (int64_t)blockIdx.x * BM;

    for (int idx = tid; idx < BM * DMAX; idx += NT) {
        const int r = idx / DMAX, k = idx % DMAX;
        const int64_t i = i0 + r;
        s.zi[r][k] = (i < row_end && k < d) ? ld_as_float(z, dtype, i * ldz + k) : 0.f;
    }
    if (tid < BM) {
        const int64_t i = i0 + tid;
        const bool ok = i < row_end;
        s.lab_i[tid] = (!CODES && ok) ? labels[i] : 0;
        // xstat row (contrastyou_b200.h CY_XS_*): x = log-denominator (self-paced), y = 1/c, z = coefficient, w = A_i (exclude)
        const float4 xs = ok ? xstat[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s.st_i[0][tid] = xs.x;
        s.st_i[1][tid] = xs.y;
        s.st_i[2][tid] = xs.z;
        s.st_i[3][tid] = xs.w;
    }

    float dacc[4][DMAX / 16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < DMAX / 16; ++c) dacc[a][c] = 0.f;

    for (int64_t j0 = 0; j0 < N; j0 += BN) {
        __syncthreads();  // previous tile fully consumed
        for (int idx = tid; idx < BN * DMAX; idx += NT) {
            const int r = idx / DMAX, k = idx % DMAX;
            const int64_t j = j0 + r;
            s.zj[r][k] = (j < N && k < d) ? ld_as_float(z, dtype, j * ldz + k) : 0.f;
        }
        if (tid < BN) {
            const int64_t j = j0 + tid;
            const bool ok = j < N;
            s.lab_j[tid] = (!CODES && ok) ? labels[j] : 0;
            const float4 xs = ok ? xstat[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            s.st_j[0][tid] = xs.x;
            s.st_j[1][tid] = xs.y;
            s.st_j[2][tid] = xs.z;
            s.st_j[3][tid] = xs.w;
        }
        __syncthreads();

        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
        for (int k = 0; k < DMAX; ++k) {
            if (k >= d) break;
            float av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = s.zi[ty + 16 * a][k];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = s.zj[tx + 16 * b][k];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }

#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int il = ty + 16 * a;
            const int64_t i = i0 + il;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int jl = tx + 16 * b;
                const int64_t j = j0 + jl;
                float wv = 0.f;
                if (i < row_end && j < N && i != j) {
                    int bij, bji;
                    if (CODES) {
                        bij = pair_bits_codes(codes, n, i, j);
                        bji = pair_bits_codes(codes, n, j, i);
                    } else {
                        bij = bji = pair_bits_labels(s.lab_i[il], s.lab_j[jl], true);
                    }
                    const float L = fmaf(acc[a][b], inv_t, -inv_t);
                    const float E = __expf(L);
                    // G_ij (row i, column j) and G_ji (row j, column i); both use the same S_ij = S_ji
                    float gij = 0.f, gji = 0.f;
                    if (VARIANT == CY_SUPCON_EXCLUDE) {
                        if (bij & 1) gij = s.st_i[1][il] * (E / (E + s.st_i[3][il] + 1e-16f) - 1.f);
                        else if (bij & 2) gij = s.st_i[2][il] * E;
                        if (bji & 1) gji = s.st_j[1][jl] * (E / (E + s.st_j[3][jl] + 1e-16f) - 1.f);
                        else if (bji & 2) gji = s.st_j[2][jl] * E;
                    } else {
                        if (bij) gij = s.st_i[2][il] * E;
                        if (bji) gji = s.st_j[2][jl] * E;
                        if (bij & 1) {
                            const float w = (VARIANT == CY_SUPCON) ? 1.f : sp_weight(VARIANT, L - s.st_i[0][il], gamma);
                            gij -= w * s.st_i[1][il];
                        }
                        if (bji & 1) {
                            const float w = (VARIANT == CY_SUPCON) ? 1.f : sp_weight(VARIANT, L - s.st_j[0][jl], gamma);
                            gji -= w * s.st_j[1][jl];
                        }
                    }
                    wv = gij + gji;
                }
                s.w[il][jl] = wv;
            }
        }
        __syncthreads();

        // dZ_i[64 x d] += W[64 x 64] Zj[64 x d]
#pragma unroll 2
        for (int k = 0; k < BN; ++k) {
            float wv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) wv[a] = s.w[ty + 16 * a][k];
#pragma unroll
            for (int c = 0; c < DMAX / 16; ++c) {
                const float zv = s.zj[k][tx + 16 * c];
#pragma unroll
                for (int a = 0; a < 4; ++a) dacc[a][c] = fmaf(wv[a], zv, dacc[a][c]);
            }
        }
    }

    const float scale = gscale[0] * inv_t / (float)N;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + ty + 16 * a;
        if (i >= row_end) continue;
#pragma unroll
        for (int c = 0; c < DMAX / 16; ++c) {
            const int k = tx + 16 * c;
            if (k < d) st_from_float(dz, dtype, i * lddz + k, dacc[a][c] * scale);
        }
    }
}

// ------------------------------------------------------------------------------------------------- masks / labels
__global__ void infonce_masks_kernel(int64_t N, int64_t n, const int32_t* __restrict__ labels,
                                     const uint8_t* __restrict__ codes, float* __restrict__ pos, float* __restrict__ neg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const int64_t i = idx / N, j = idx % N;
    const int bits = codes ? pair_bits_codes(codes, n, i, j) : pair_bits_labels(labels[i], labels[j], i != j);
    if (pos) pos[idx] = (bits & 1) ? 1.f : 0.f;
    if (neg) neg[idx] = (bits & 2) ? 1.f : 0.f;
}

__global__ void labels_canonicalize_kernel(const void* __restrict__ src, int kind, int64_t n, int32_t* __restrict__ dst,
                                           int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t a, b;
    if (kind == 2) {        // int64 labels (torch's default integer dtype): exact when they fit 32 bits, counted otherwise
        const int64_t v = reinterpret_cast<const int64_t*>(src)[i];
        if (v < (int64_t)INT32_MIN || v > (int64_t)INT32_MAX) atomicAdd(overflow, 1);
        a = b = (int32_t)v;
    } else if (kind == 0) {
        const float v = reinterpret_cast<const float*>(src)[i] + 0.0f;  // -0.0 -> +0.0
        if (v != v) {  // NaN never equals anything, not even its twin in the other view
            a = 0x7FC00000 | (int32_t)(i & 0x1FFFFF);
            b = 0x7FC00000 | 0x200000 | (int32_t)(i & 0x1FFFFF);
        } else {
            a = b = __float_as_int(v);
        }
    } else {
        a = b = reinterpret_cast<const int32_t*>(src)[i];
    }
    dst[i] = a;
    dst[i + n] = b;
}

// ------------------------------------------------------------------------------------------------- launchers
template <int VARIANT, int PASS>
static int launch_fwd(const void* z, int dtype, int64_t N, int d, int64_t ldz, const int32_t* labels, const uint8_t* codes,
                      int64_t row_begin, int64_t row_end, float inv_t, float gamma, float* stats, cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    const unsigned grid = (unsigned)((rows + BM - 1) / BM);
    if (codes)
        infonce_fwd_simt_kernel<VARIANT, PASS, true><<<grid, NT, 0, st>>>(z, dtype, N, d, ldz, labels, codes, N / 2,
                                                                       row_begin, row_end, inv_t, gamma, stats);
    else
        infonce_fwd_simt_kernel<VARIANT, PASS, false><<<grid, NT, 0, st>>>(z, dtype, N, d, ldz, labels, codes, N / 2,
                                                                        row_begin, row_end, inv_t, gamma, stats);
    CY_CHECK_LAUNCH("infonce_fwd_simt");
    return CY_OK;
}

int infonce_fwd_simt(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                     const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant, int pass,
                     float gamma, float* stats, cudaStream_t st) {
    if (row_end <= row_begin) return CY_OK;
    if (pass == 1) {
        // pass 1 is variant independent
        return launch_fwd<CY_SUPCON, 1>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, st);
    }
    switch (variant) {
        case CY_SUPCON_EXCLUDE:
            return launch_fwd<CY_SUPCON_EXCLUDE, 2>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, st);
        case CY_SELFPACED_HARD:
            return launch_fwd<CY_SELFPACED_HARD, 2>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, st);
        case CY_SELFPACED_SOFT:
            return launch_fwd<CY_SELFPACED_SOFT, 2>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, st);
        default:
            set_error("pass 2 is not defined for variant %d", variant);
            return CY_ERR_ARG;
    }
}

template <int VARIANT>
static int launch_bwd(const void* z, int dtype, int64_t N, int d, int64_t ldz, const int32_t* labels, const uint8_t* codes,
                      int64_t row_begin, int64_t row_end, float inv_t, float gamma, const float* stats_,
                      const float* gscale, void* dz, int64_t lddz, cudaStream_t st) {
    const float4* stats = reinterpret_cast<const float4*>(stats_);      // xstat [N][4]
    const int64_t rows = row_end - row_begin;
    const unsigned grid = (unsigned)((rows + BM - 1) / BM);
    const size_t smem = sizeof(BwdSmem);
    cudaError_t e;
    if (codes) {
        auto k = infonce_bwd_simt_kernel<VARIANT, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        k<<<grid, NT, smem, st>>>(z, dtype, N, d, ldz, labels, codes, N / 2, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz);
    } else {
        auto k = infonce_bwd_simt_kernel<VARIANT, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        k<<<grid, NT, smem, st>>>(z, dtype, N, d, ldz, labels, codes, N / 2, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz);
    }
    CY_CHECK_LAUNCH("infonce_bwd_simt");
    return CY_OK;
}

int infonce_bwd_simt(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                     const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant, float gamma,
                     const float* stats, const float* gscale, void* dz, int64_t lddz, cudaStream_t st) {
    if (row_end <= row_begin) return CY_OK;
    switch (variant) {
        case CY_SUPCON:
            return launch_bwd<CY_SUPCON>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz, st);
        case CY_SUPCON_EXCLUDE:
            return launch_bwd<CY_SUPCON_EXCLUDE>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz, st);
        case CY_SELFPACED_HARD:
            return launch_bwd<CY_SELFPACED_HARD>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz, st);
        case CY_SELFPACED_SOFT:
            return launch_bwd<CY_SELFPACED_SOFT>(z, dtype, N, (int)d, ldz, labels, codes, row_begin, row_end, inv_t, gamma, stats, gscale, dz, lddz, st);
    }
    set_error("unknown variant %d", variant);
    return CY_ERR_ARG;
}

int infonce_masks(int64_t N, const int32_t* labels, const uint8_t* codes, float* pos, float* neg, cudaStream_t st) {
    const int64_t total = N * N;
    infonce_masks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(N, N / 2, labels, codes, pos, neg);
    CY_CHECK_LAUNCH("infonce_masks");
    return CY_OK;
}

int labels_canonicalize(const void* src, int kind, int64_t n, int32_t* dst, int32_t* overflow, cudaStream_t st) {
    labels_canonicalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, kind, n, dst, overflow);
    CY_CHECK_LAUNCH("labels_canonicalize");
    return CY_OK;
}

}  // namespace cy
