// InfoNCE / SupCon on the 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM, flash-style.
//
// Scope: CY_SUPCON with label-derived masks, bf16 embeddings, d == 256, rows / N multiples of 128 (everything else
// runs on the SIMT path).  The N x N similarity never exists in HBM: each CTA owns a 128-row block of Z (TMA-loaded
// once, the MMA "A" operand), streams 128- or 64-row column tiles of Z through a TMA ring ("B" operand, K-major,
// 128-byte swizzle), accumulates S = Zi Zj^T in TMEM and lets eight epilogue warps read it back with tcgen05.ld.
//
//   forward  (infonce_fwd_tc_kernel):  per element E = 2^(s*c1 - c1), c1 = log2(e)/t; row sums D_i += E, positives
//            c_i += [lab_i == lab_j], posS_i += [lab_i == lab_j] s; diagonal excluded on diagonal tiles only.
//            Column range split over blockIdx.y; partial row sums go to a [slot][3][N] scratch, summed in a fixed
//            order by infonce_tc_reduce_kernel (deterministic) into the CY_STAT_* rows cy_infonce_finalize reads.
//   backward (infonce_bwd_tc_kernel):  S tile recomputed, W_ij = E_ij (coef_i + coef_j) - P_ij (invc_i + invc_j)
//            written as bf16 into a swizzled K-major shared tile, second MMA dZ_i[128x256] += W[128x64] Zj[64x256]
//            with the SAME Zj bytes read MN-major; dZ accumulates in TMEM across all column tiles of the row block
//            and is scaled by gscale/(t N) on the way out (dZ = (1/t) W Z, W = G + G^T).
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane each), warp 2 TMEM allocator,
// warps 4..11 epilogue (warp w reads TMEM lanes 32*(w%4).., column half (w-4)/4).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cy {

using namespace tc;

constexpr int TC_D = 256;            // embedding dim handled by this path
constexpr int TC_BM = 128;           // rows per CTA (UMMA M)
constexpr int TC_THREADS = 384;
constexpr int TC_KBLK = TC_D / 64;   // 64-element (128-byte) K blocks per row
constexpr float LOG2E = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------------------ forward
template <int BN>
struct FwdSmem {
    static constexpr int NSTAGE = 2;
    static constexpr int NACC = 512 / BN >= 4 ? 4 : 2;
    static constexpr uint32_t A_BYTES = TC_BM * TC_D * 2;
    static constexpr uint32_t B_BYTES = BN * TC_D * 2;
    static constexpr uint32_t OFF_B = A_BYTES;
    static constexpr uint32_t OFF_LAB = OFF_B + NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_LAB + 8 * (BN / 2) * 4;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;   // + barriers + 1024-alignment slack
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels, int N, int row_begin,
                      int tiles_per_split, float c1, float* __restrict__ part) {
    using S = FwdSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + S::OFF_B;
    int32_t* sLab = reinterpret_cast<int32_t*>(smem + S::OFF_LAB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + S::NSTAGE;
    uint64_t* acc_full = b_empty + S::NSTAGE;
    uint64_t* acc_empty = acc_full + S::NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + S::NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = row_begin + blockIdx.x * TC_BM;              // first global row of this CTA
    const int n_ctiles = N / BN;
    const int ct0 = blockIdx.y * tiles_per_split;
    const int ct1 = min(n_ctiles, ct0 + tiles_per_split);

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < S::NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < S::NACC; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(a_full, S::A_BYTES);
            for (int kb = 0; kb < TC_KBLK; ++kb)
                for (int hb = 0; hb < TC_BM / 64; ++hb)
                    tma_load_2d(sA + kb * (TC_BM * 128) + hb * 8192, &tmap, a_full, kb * 64, row0 + hb * 64);
            Ring<S::NSTAGE> ring;
            for (int ct = ct0; ct < ct1; ++ct, ring.next()) {
                const uint32_t s = ring.stage();
                mbar_wait(b_empty + s, ring.phase() ^ 1u);
                mbar_arrive_expect_tx(b_full + s, S::B_BYTES);
                uint8_t* dst = sB + s * S::B_BYTES;
                for (int kb = 0; kb < TC_KBLK; ++kb)
                    for (int hb = 0; hb < BN / 64; ++hb)
                        tma_load_2d(dst + kb * (BN * 128) + hb * 8192, &tmap, b_full + s, kb * 64, ct * BN + hb * 64);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_f32(TC_BM, BN, 0, 0);
            const uint32_t a_addr = smem_u32(sA);
            mbar_wait(a_full, 0);
            Ring<S::NSTAGE> ring;
            Ring<S::NACC> acc;
            for (int ct = ct0; ct < ct1; ++ct, ring.next(), acc.next()) {
                const uint32_t s = ring.stage(), a = acc.stage();
                mbar_wait(b_full + s, ring.phase());
                mbar_wait(acc_empty + a, acc.phase() ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * S::B_BYTES);
#pragma unroll
                for (int kb = 0; kb < TC_KBLK; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem_base + a * BN, smem_desc(a_addr + kb * (TC_BM * 128) + ks * 32, 16, 1024),
                                  smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                umma_commit(b_empty + s);
                umma_commit(acc_full + a);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int row = q * 32 + lane;
        const int gi = row0 + row;
        const int32_t my_lab = labels[gi];
        int32_t* wlab = sLab + ew * (BN / 2);
        float D = 0.f, posS = 0.f;
        int cnt = 0;
        Ring<S::NACC> acc;
        for (int ct = ct0; ct < ct1; ++ct, acc.next()) {
            const uint32_t a = acc.stage();
            const int jbase = ct * BN + h * (BN / 2);          // first global column of this warp's half
            __syncwarp();
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) wlab[c * 32 + lane] = labels[jbase + c * 32 + lane];
            __syncwarp();
            mbar_wait(acc_full + a, acc.phase());
            tc_fence_after();
            const bool diag_tile = (gi >= jbase) && (gi < jbase + BN / 2);   // warp-uniform up to the 32-row group
            const bool any_diag = __any_sync(0xffffffffu, diag_tile);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + a * BN + h * (BN / 2) + c * 32, r);
                tmem_ld_wait();
                if (!any_diag) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const int4 lj = *reinterpret_cast<const int4*>(wlab + c * 32 + e4 * 4);
                        const int32_t lv[4] = {lj.x, lj.y, lj.z, lj.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float s = __uint_as_float(r[e4 * 4 + u]);
                            D += ex2_approx(fmaf(s, c1, -c1));
                            if (lv[u] == my_lab) { cnt += 1; posS += s; }
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = jbase + c * 32 + e;
                        const float s = __uint_as_float(r[e]);
                        if (j != gi) {
                            D += ex2_approx(fmaf(s, c1, -c1));
                            if (wlab[c * 32 + e] == my_lab) { cnt += 1; posS += s; }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);
        }
        const int slot = blockIdx.y * 2 + h;
        float* p = part + (size_t)slot * 3 * N;
        p[gi] = D;
        p[(size_t)N + gi] = (float)cnt;
        p[2 * (size_t)N + gi] = posS;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// partial row sums -> the pass-1 raw statistics cy_infonce_finalize expects (fixed summation order)
__global__ void infonce_tc_reduce_kernel(const float* __restrict__ part, int nslot, int N, int row_begin, int row_end,
                                         float inv_t, float* __restrict__ stats) {
    const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    float D = 0.f, c = 0.f, ps = 0.f;
    for (int s = 0; s < nslot; ++s) {
        const float* p = part + (size_t)s * 3 * N;
        D += p[i];
        c += p[(size_t)N + i];
        ps += p[2 * (size_t)N + i];
    }
    stats[(size_t)CY_STAT_POSE * N + i] = D;                       // total sum_j E_ij (positives and negatives together)
    stats[(size_t)CY_STAT_AUX * N + i] = 0.f;
    stats[(size_t)CY_STAT_INVC * N + i] = c;
    stats[(size_t)CY_STAT_NEGC * N + i] = (float)(N - 1) - c;
    stats[(size_t)CY_STAT_POSL * N + i] = inv_t * ps - inv_t * c;  // sum_j P_ij (s_ij - 1)/t
}

// ------------------------------------------------------------------------------------------------------------ backward
struct BwdCfg {
    static constexpr int BN = 64;
    static constexpr int NSTAGE = 3;     // Zj ring (a stage lives from its MMA1 until its MMA2 retires)
    static constexpr int NS = 4;         // S accumulators in TMEM (64 columns each) at columns [256, 512)
    static constexpr int NW = 2;         // W tiles in shared memory
    static constexpr uint32_t A_BYTES = TC_BM * TC_D * 2;       // 64 KB
    static constexpr uint32_t B_BYTES = BN * TC_D * 2;          // 32 KB
    static constexpr uint32_t W_BYTES = TC_BM * BN * 2;         // 16 KB
    static constexpr uint32_t OFF_B = A_BYTES;
    static constexpr uint32_t OFF_W = OFF_B + NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_COL = OFF_W + NW * W_BYTES;   // per epilogue warp: lab[32], coef[32], invc[32]
    static constexpr uint32_t OFF_BAR = OFF_COL + 8 * 3 * 32 * 4;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels,
                      const float* __restrict__ stats, int N, int row_begin, float c1, const float* __restrict__ gscale,
                      float out_scale, __nv_bfloat16* __restrict__ dz, int64_t lddz) {
    using C = BwdCfg;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::OFF_B;
    uint8_t* sW = smem + C::OFF_W;
    float* sCol = reinterpret_cast<float*>(smem + C::OFF_COL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + C::NSTAGE;
    uint64_t* s_full = b_empty + C::NSTAGE;
    uint64_t* s_empty = s_full + C::NS;
    uint64_t* w_full = s_empty + C::NS;
    uint64_t* w_empty = w_full + C::NW;
    uint64_t* dz_full = w_empty + C::NW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dz_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = row_begin + blockIdx.x * TC_BM;
    const int nt = N / BN;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < C::NS; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 8); }
        for (int i = 0; i < C::NW; ++i) { mbar_init(w_full + i, 8); mbar_init(w_empty + i, 1); }
        mbar_init(dz_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_dz = tmem_base;             // columns [0, 256)
    const uint32_t tmem_s = tmem_base + 256;        // NS x 64 columns

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(a_full, C::A_BYTES);
            for (int kb = 0; kb < TC_KBLK; ++kb)
                for (int hb = 0; hb < TC_BM / 64; ++hb)
                    tma_load_2d(sA + kb * (TC_BM * 128) + hb * 8192, &tmap, a_full, kb * 64, row0 + hb * 64);
            Ring<C::NSTAGE> ring;
            for (int t = 0; t < nt; ++t, ring.next()) {
                const uint32_t s = ring.stage();
                mbar_wait(b_empty + s, ring.phase() ^ 1u);
                mbar_arrive_expect_tx(b_full + s, C::B_BYTES);
                uint8_t* dst = sB + s * C::B_BYTES;
                for (int kb = 0; kb < TC_KBLK; ++kb) tma_load_2d(dst + kb * (BN * 128), &tmap, b_full + s, kb * 64, t * BN);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc1 = idesc_bf16_f32(TC_BM, BN, 0, 0);      // S  = Zi (K-major) x Zj (K-major)
            constexpr uint32_t idesc2 = idesc_bf16_f32(TC_BM, TC_D, 0, 1);    // dZ += W (K-major) x Zj (MN-major)
            const uint32_t a_addr = smem_u32(sA);
            mbar_wait(a_full, 0);
            Ring<C::NSTAGE> ring1;      // stage / phase of the tile whose MMA1 is issued next
            Ring<C::NS> sacc;
            auto issue_mma1 = [&]() {
                const uint32_t s = ring1.stage(), a = sacc.stage();
                mbar_wait(b_full + s, ring1.phase());
                mbar_wait(s_empty + a, sacc.phase() ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
                for (int kb = 0; kb < TC_KBLK; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem_s + a * BN, smem_desc(a_addr + kb * (TC_BM * 128) + ks * 32, 16, 1024),
                                  smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc1, (kb | ks) != 0);
                umma_commit(s_full + a);
                ring1.next();
                sacc.next();
            };
            issue_mma1();
            Ring<C::NSTAGE> ring2;
            Ring<C::NW> wr;
            for (int t = 0; t < nt; ++t, ring2.next(), wr.next()) {
                if (t + 1 < nt) issue_mma1();                       // keep the tensor pipe busy while the epilogue works
                const uint32_t s = ring2.stage(), w = wr.stage();
                mbar_wait(w_full + w, wr.phase());
                tc_fence_after();
                const uint32_t w_addr = smem_u32(sW + w * C::W_BYTES);
                const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks)
                    // A: W [128 x 64] K-major, 32 B per k-step.  B: Zj read MN-major: N = d (4 groups of 64, LBO = BN*128),
                    // K = j (8-row groups, SBO = 1024); one k-step = 16 rows of j = 2048 B.
                    umma_bf16(tmem_dz, smem_desc(w_addr + ks * 32, 16, 1024), smem_desc(b_addr + ks * 2048, BN * 128, 1024),
                              idesc2, (t | ks) != 0);
                umma_commit(w_empty + w);
                umma_commit(b_empty + s);
            }
            umma_commit(dz_full);
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int row = q * 32 + lane;
        const int gi = row0 + row;
        const int32_t my_lab = labels[gi];
        const float coef_i = stats[(size_t)CY_STAT_COEF * N + gi];
        const float invc_i = stats[(size_t)CY_STAT_INVC * N + gi];
        float* wcol = sCol + ew * 96;                 // [0,32) labels (as int bits), [32,64) coef_j, [64,96) invc_j
        Ring<C::NS> sacc;
        Ring<C::NW> wr;
        for (int t = 0; t < nt; ++t, sacc.next(), wr.next()) {
            const uint32_t a = sacc.stage(), w = wr.stage();
            const int jbase = t * BN + h * 32;
            __syncwarp();
            wcol[lane] = __int_as_float(labels[jbase + lane]);
            wcol[32 + lane] = stats[(size_t)CY_STAT_COEF * N + jbase + lane];
            wcol[64 + lane] = stats[(size_t)CY_STAT_INVC * N + jbase + lane];
            __syncwarp();
            mbar_wait(s_full + a, sacc.phase());
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + a);      // S values are in registers: the accumulator can be reused
            uint32_t packed[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                float wv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float s = __uint_as_float(r[e + u]);
                    const float E = ex2_approx(fmaf(s, c1, -c1));
                    float v = E * (coef_i + wcol[32 + e + u]);
                    if (__float_as_int(wcol[e + u]) == my_lab) v -= invc_i + wcol[64 + e + u];
                    if (jbase + e + u == gi) v = 0.f;
                    wv[u] = v;
                }
                __nv_bfloat162 b2 = __floats2bfloat162_rn(wv[0], wv[1]);
                packed[e >> 1] = *reinterpret_cast<uint32_t*>(&b2);
            }
            mbar_wait(w_empty + w, wr.phase() ^ 1u);
            // row `row` of the [128 x 64] bf16 tile: 128 B, this warp's half = 16-byte chunks 4h..4h+3, swizzled by row % 8
            uint8_t* wrow = sW + w * C::W_BYTES + row * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int phys = ((4 * h + ch) ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(wrow + phys) =
                    make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_full + w);
        }
        // dZ rows of this CTA: TMEM -> registers -> bf16 -> global
        mbar_wait(dz_full, 0);
        tc_fence_after();
        const float scale = gscale[0] * out_scale;
        __nv_bfloat16* out = dz + (size_t)gi * lddz + h * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * 128 + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
                uint32_t pk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[e + 2 * u]) * scale,
                                                              __uint_as_float(r[e + 2 * u + 1]) * scale);
                    pk[u] = *reinterpret_cast<uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(out + c * 32 + e) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [N, 256] bf16, row pitch ldz elements; box = 64 columns (128 B, the swizzle span) x 64 rows
static int make_tmap(CUtensorMap* m, const void* z, int64_t N, int64_t ldz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CY_ERR_DEVICE; }
    cuuint64_t gdim[2] = {(cuuint64_t)TC_D, (cuuint64_t)N};
    cuuint64_t gstride[1] = {(cuuint64_t)ldz * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(z), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CY_ERR_ARG; }
    return CY_OK;
}

bool infonce_tc_supported(int dtype, int64_t N, int64_t d, int64_t ldz, const uint8_t* codes, int variant) {
    return dtype == CY_BF16 && d == TC_D && (N % 128) == 0 && N >= 256 && (ldz % 8) == 0 && codes == nullptr &&
           variant == CY_SUPCON && N < (int64_t(1) << 30);
}

constexpr int FWD_BN = 128;

static int fwd_splits(int64_t N, int64_t rows) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int64_t rb = rows / TC_BM, ctiles = N / FWD_BN;
    // enough CTAs for ~8 waves, but at least 8 column tiles per CTA so the A load and the prologue amortise
    int64_t want = (8LL * sms + rb - 1) / rb;
    int64_t maxs = ctiles / 8 > 0 ? ctiles / 8 : 1;
    if (maxs > 16) maxs = 16;
    int64_t s = want < 1 ? 1 : (want > maxs ? maxs : want);
    return (int)s;
}

size_t infonce_tc_workspace_bytes(int64_t N, int64_t d) {
    if (d != TC_D || (N % 128) != 0) return 0;
    // worst case over row ranges: a single 128-row block -> the most column splits
    const int smax = fwd_splits(N, TC_BM);
    return (size_t)smax * 2 * 3 * (size_t)N * sizeof(float);
}

int infonce_fwd_tc(const void* z, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                   float inv_t, float* stats, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    CY_CHECK_ARG((rows % TC_BM) == 0 && (row_begin % TC_BM) == 0, "tcgen05 path: row range must be 128-aligned");
    CY_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0, "tcgen05 path: z must be 16-byte aligned");
    const int splits = fwd_splits(N, rows);
    const int nslot = splits * 2;
    const size_t need = (size_t)nslot * 3 * (size_t)N * sizeof(float);
    CY_CHECK_ARG(workspace && workspace_bytes >= need, "infonce_fwd_tc: workspace %zu < %zu", workspace_bytes, need);
    CUtensorMap tmap;
    int rc = make_tmap(&tmap, z, N, ldz);
    if (rc) return rc;
    using S = FwdSmem<FWD_BN>;
    auto k = infonce_fwd_tc_kernel<FWD_BN>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
    if (e != cudaSuccess) { set_error("fwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int ctiles = (int)(N / FWD_BN);
    const int tps = (ctiles + splits - 1) / splits;
    dim3 grid((unsigned)(rows / TC_BM), (unsigned)splits);
    k<<<grid, TC_THREADS, S::TOTAL, st>>>(tmap, labels, (int)N, (int)row_begin, tps, inv_t * LOG2E,
                                          reinterpret_cast<float*>(workspace));
    CY_CHECK_LAUNCH("infonce_fwd_tc");
    infonce_tc_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(reinterpret_cast<float*>(workspace), nslot, (int)N,
                                                                            (int)row_begin, (int)row_end, inv_t, stats);
    CY_CHECK_LAUNCH("infonce_tc_reduce");
    return CY_OK;
}

int infonce_bwd_tc(const void* z, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                   float inv_t, const float* stats, const float* gscale, void* dz, int64_t lddz, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
    (void)workspace; (void)workspace_bytes; (void)d;
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    CY_CHECK_ARG((rows % TC_BM) == 0 && (row_begin % TC_BM) == 0, "tcgen05 path: row range must be 128-aligned");
    CY_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (lddz % 8) == 0,
                 "tcgen05 path: z / dz must be 16-byte aligned");
    CUtensorMap tmap;
    int rc = make_tmap(&tmap, z, N, ldz);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg::TOTAL);
    if (e != cudaSuccess) { set_error("bwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    infonce_bwd_tc_kernel<<<(unsigned)(rows / TC_BM), TC_THREADS, BwdCfg::TOTAL, st>>>(
        tmap, labels, stats, (int)N, (int)row_begin, inv_t * LOG2E, gscale, inv_t / (float)N,
        reinterpret_cast<__nv_bfloat16*>(dz), lddz);
    CY_CHECK_LAUNCH("infonce_bwd_tc");
    return CY_OK;
}

}  // namespace cy
