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
// -DCY_FWD_EX2_POLY=1 sends every other exponential of the forward epilogue to the FMA-pipe polynomial (tc_common.cuh:
// ex2_poly).  Measured on B200 at N = 65536: 1.757 ms vs 1.646 ms with all of them on MUFU — the epilogue is issue-bound,
// not MUFU-bound, so the default keeps MUFU.
#ifndef CY_FWD_EX2_POLY
#define CY_FWD_EX2_POLY 0
#endif
#if CY_FWD_EX2_POLY
#define EX2_ALT ex2_poly
#else
#define EX2_ALT ex2_approx
#endif
#ifndef CY_FWD_A_TMEM
#define CY_FWD_A_TMEM 1
#endif
// CY_FWD_A_TMEM: Zi in tensor memory (columns [384, 512)), three S accumulators instead of four, three Zj stages instead
// of two.  With both operands in shared memory an M128 N128 K16 MMA fetches 8 KB per 64 clk — the whole 128 B/clk port —
// while TMA refills the ring through the same port (64 KB per 1024-clk tile): the forward was shared-memory bound at
// ~2/3 tensor utilisation.  A in TMEM halves the operand fetch.
template <int BN>
struct FwdSmem {
    static constexpr bool A_TMEM = CY_FWD_A_TMEM != 0 && BN == 128;
    static constexpr int NSTAGE = A_TMEM ? 3 : 2;
    static constexpr int NACC = A_TMEM ? 3 : (512 / BN >= 4 ? 4 : 2);
    static constexpr uint32_t A_BYTES = A_TMEM ? 0 : TC_BM * TC_D * 2;
    static constexpr uint32_t B_BYTES = BN * TC_D * 2;
    static constexpr uint32_t OFF_B = A_BYTES;
    static constexpr uint32_t OFF_LAB = OFF_B + NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_LAB + 8 * (BN / 2) * 4;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;   // + barriers + 1024-alignment slack
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels, int N, int row_begin,
                      int tiles_per_split, float c1, float* __restrict__ part, uint32_t idesc, const uint16_t* __restrict__ zrows,
                      int64_t ldz) {
    using S = FwdSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // 1024-aligned, still a shared pointer
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
        mbar_init(a_full, S::A_TMEM ? 8 : 1);
        for (int i = 0; i < S::NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < S::NACC; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_a = tmem_base + 384;        // A_TMEM: Zi, columns [384, 512)

    if (warp == 0) {
        if (elect_one()) {
            if constexpr (!S::A_TMEM) {
                mbar_arrive_expect_tx(a_full, S::A_BYTES);
                for (int kb = 0; kb < TC_KBLK; ++kb)
                    for (int hb = 0; hb < TC_BM / 64; ++hb)
                        tma_load_2d(sA + kb * (TC_BM * 128) + hb * 8192, &tmap, a_full, kb * 64, row0 + hb * 64);
            }
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
            const uint32_t a_addr = smem_u32(sA);        // idesc: M128 x N(BN) x K16, both operands K-major, bf16 or fp16
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
                    for (int ks = 0; ks < 4; ++ks) {
                        if constexpr (S::A_TMEM)
                            umma_bf16_ts(tmem_base + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                         smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                        else
                            umma_bf16(tmem_base + a * BN, smem_desc(a_addr + kb * (TC_BM * 128) + ks * 32, 16, 1024),
                                      smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                    }
                umma_commit(b_empty + s);
                umma_commit(acc_full + a);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int row = q * 32 + lane;
        const int gi = row0 + row;
        const int32_t my_lab = labels[gi];
        if constexpr (S::A_TMEM) {
            // this thread's half row of Zi (128 elements = 64 packed columns) -> tensor memory lanes q*32.., columns h*64..
            const uint4* src = reinterpret_cast<const uint4*>(zrows + (size_t)gi * ldz + h * 128);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint4 v = __ldg(src + c * 8 + e);
                    r[4 * e] = v.x; r[4 * e + 1] = v.y; r[4 * e + 2] = v.z; r[4 * e + 3] = v.w;
                }
                tmem_st_32x32(tmem_a + (uint32_t(q * 32) << 16) + h * 64 + c * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        int32_t* wlab = sLab + ew * (BN / 2);
        float D0 = 0.f, D1 = 0.f, D2 = 0.f, D3 = 0.f, posS = 0.f;
        int cnt = 0;
        constexpr int NCH = BN / 64;                       // 32-column chunks per warp and tile
        // label range of this warp's 32 rows: a column chunk whose label range does not intersect it holds no positive
        // pair, whatever the order of the rows (callers that sort rows by label make this the common case)
        const int32_t row_lo = __reduce_min_sync(0xffffffffu, my_lab), row_hi = __reduce_max_sync(0xffffffffu, my_lab);
        int32_t lab_next[NCH];
        if (ct0 < ct1) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) lab_next[c] = labels[ct0 * BN + h * (BN / 2) + c * 32 + lane];
        }
        Ring<S::NACC> acc;
        for (int ct = ct0; ct < ct1; ++ct, acc.next()) {
            const uint32_t a = acc.stage();
            const int jbase = ct * BN + h * (BN / 2);          // first global column of this warp's half
            __syncwarp();
            int32_t cmin = lab_next[0], cmax = lab_next[0];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                wlab[c * 32 + lane] = lab_next[c];
                cmin = min(cmin, lab_next[c]);
                cmax = max(cmax, lab_next[c]);
            }
            const bool may_have_pos = __reduce_max_sync(0xffffffffu, cmax) >= row_lo && __reduce_min_sync(0xffffffffu, cmin) <= row_hi;
            __syncwarp();
            if (ct + 1 < ct1) {                                // next tile's labels travel while this tile is processed
#pragma unroll
                for (int c = 0; c < NCH; ++c) lab_next[c] = labels[jbase + BN + c * 32 + lane];
            }
            mbar_wait(acc_full + a, acc.phase());
            tc_fence_after();
            uint32_t r[NCH][32];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + a * BN + h * (BN / 2) + c * 32, r[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);         // values are in registers: hand the accumulator back now
            const bool diag_tile = (gi >= jbase) && (gi < jbase + BN / 2);   // uniform per warp (32-row groups)
            if (!may_have_pos) {                              // (a diagonal chunk always intersects: i is its own label)
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        D0 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 0]), c1, -c1));
                        D1 += EX2_ALT(fmaf(__uint_as_float(r[c][e4 * 4 + 1]), c1, -c1));
                        D2 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 2]), c1, -c1));
                        D3 += EX2_ALT(fmaf(__uint_as_float(r[c][e4 * 4 + 3]), c1, -c1));
                    }
                }
            } else if (!__any_sync(0xffffffffu, diag_tile)) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const int4 lj = *reinterpret_cast<const int4*>(wlab + c * 32 + e4 * 4);
                        const float s0 = __uint_as_float(r[c][e4 * 4 + 0]), s1 = __uint_as_float(r[c][e4 * 4 + 1]);
                        const float s2 = __uint_as_float(r[c][e4 * 4 + 2]), s3 = __uint_as_float(r[c][e4 * 4 + 3]);
                        D0 += ex2_approx(fmaf(s0, c1, -c1));
                        D1 += EX2_ALT(fmaf(s1, c1, -c1));
                        D2 += ex2_approx(fmaf(s2, c1, -c1));
                        D3 += EX2_ALT(fmaf(s3, c1, -c1));
                        if (lj.x == my_lab) { cnt += 1; posS += s0; }
                        if (lj.y == my_lab) { cnt += 1; posS += s1; }
                        if (lj.z == my_lab) { cnt += 1; posS += s2; }
                        if (lj.w == my_lab) { cnt += 1; posS += s3; }
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = jbase + c * 32 + e;
                        const float sv = __uint_as_float(r[c][e]);
                        if (j != gi) {
                            D0 += ex2_approx(fmaf(sv, c1, -c1));
                            if (wlab[c * 32 + e] == my_lab) { cnt += 1; posS += sv; }
                        }
                    }
                }
            }
        }
        const float D = (D0 + D1) + (D2 + D3);
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
#ifndef CY_BWD_A_TMEM
#define CY_BWD_A_TMEM 1
#endif
// CY_BWD_A_TMEM: the row block Zi (the A operand of S = Zi Zj^T) lives in TENSOR MEMORY (128 columns: two bf16 per column)
// instead of shared memory.  An M128 N64 K16 MMA with both operands in shared memory fetches 6 KB and is bound by the
// 128 B/clk shared-memory port (48 clk, profiles/probes/probe_umma_small.cu) instead of its 32 clk of math; with A in TMEM
// it fetches 2 KB.  The 64 KB of shared memory this frees deepen the Zj ring.
#ifndef CY_BWD_W_TMEM
#define CY_BWD_W_TMEM 1
#endif
// CY_BWD_W_TMEM (needs A_TMEM): the weight tile W (A operand of dZ += W Zj) is written by the epilogue straight back into
// the tensor-memory columns of the S accumulator it was computed from (bf16 pairs: 16 columns per 32-column half) and
// read from there by the second MMA — no shared-memory W tile, no swizzled stores, no proxy fence, 32 KB less
// shared-memory traffic per 128 x 64 tile.  The S slot is recycled by the in-order tensor pipe: MMA1(t+2) is issued
// after MMA2(t), which is the last reader of slot t % 2.
struct BwdCfg {
    static constexpr int BN = 64;
    static constexpr bool A_TMEM = CY_BWD_A_TMEM != 0;
    static constexpr bool W_TMEM = A_TMEM && CY_BWD_W_TMEM != 0;
    static constexpr int NSTAGE = W_TMEM ? 6 : (A_TMEM ? 5 : 3);     // Zj ring (a stage lives from its MMA1 until its MMA2 retires)
    static constexpr int NS = A_TMEM ? 2 : 4;         // S accumulators in TMEM (64 columns each), the last NS*64 columns
    static constexpr int NW = 2;         // W tiles in shared memory
    static constexpr uint32_t A_BYTES = A_TMEM ? 0 : TC_BM * TC_D * 2;       // 64 KB
    static constexpr uint32_t B_BYTES = BN * TC_D * 2;          // 32 KB
    static constexpr uint32_t W_BYTES = W_TMEM ? 0 : TC_BM * BN * 2;         // 16 KB
    static constexpr uint32_t OFF_B = A_BYTES;
    static constexpr uint32_t OFF_W = OFF_B + NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_COL = OFF_W + NW * W_BYTES;   // per epilogue warp: lab[32], coef[32], invc[32]
    static constexpr uint32_t OFF_BAR = OFF_COL + 8 * 3 * 32 * 4;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

// two 16-bit elements of the embedding dtype from two floats
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (F16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        // round-half-up on the bit patterns + one PRMT: F2FP.BF16.PACK_AB issues on the XU pipe (1/16 rate), which the
        // exponentials of the same epilogue already keep busy
        return __byte_perm(__float_as_uint(a) + 0x8000u, __float_as_uint(b) + 0x8000u, 0x7632);
    }
}

// F16: fp16 embeddings.  W is then stored as fp16 scaled by 2^10 (softmax-sized weights below 6e-8 would flush to zero in
// fp16; scaled, the flush threshold drops to 6e-11 while the largest weight, ~2, stays far from 65504); the scale is
// undone in out_scale by the host.
template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels,
                      const float* __restrict__ stats, int N, int row_begin, int tiles_per_split, float c1,
                      const float* __restrict__ gscale, float out_scale, void* __restrict__ dz_v, int64_t lddz,
                      float* __restrict__ dz32, uint32_t idesc1, uint32_t idesc2, const uint16_t* __restrict__ zrows, int64_t ldz) {
    uint16_t* dz = reinterpret_cast<uint16_t*>(dz_v);
    constexpr float WS = F16 ? 1024.f : 1.f;
    using C = BwdCfg;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // 1024-aligned, still a shared pointer
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
    // column tiles [t0, t1) of this CTA (blockIdx.y splits the columns so that small row ranges still fill the chip)
    const int t0 = blockIdx.y * tiles_per_split;
    const int t1 = min(N / BN, t0 + tiles_per_split);
    const int nt = t1 - t0;
    if (nt <= 0) return;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, C::A_TMEM ? 8 : 1);
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
    const uint32_t tmem_a = tmem_base + 256;        // A_TMEM: Zi, columns [256, 384)
    const uint32_t tmem_s = tmem_base + 512 - C::NS * BN;        // NS x 64 columns

    if (warp == 0) {
        if (elect_one()) {
            if constexpr (!C::A_TMEM) {
                mbar_arrive_expect_tx(a_full, C::A_BYTES);
                for (int kb = 0; kb < TC_KBLK; ++kb)
                    for (int hb = 0; hb < TC_BM / 64; ++hb)
                        tma_load_2d(sA + kb * (TC_BM * 128) + hb * 8192, &tmap, a_full, kb * 64, row0 + hb * 64);
            }
            Ring<C::NSTAGE> ring;
            for (int t = 0; t < nt; ++t, ring.next()) {
                const uint32_t s = ring.stage();
                mbar_wait(b_empty + s, ring.phase() ^ 1u);
                mbar_arrive_expect_tx(b_full + s, C::B_BYTES);
                uint8_t* dst = sB + s * C::B_BYTES;
                for (int kb = 0; kb < TC_KBLK; ++kb) tma_load_2d(dst + kb * (BN * 128), &tmap, b_full + s, kb * 64, (t0 + t) * BN);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // idesc1: S = Zi (K-major) x Zj (K-major), M128 N64;  idesc2: dZ += W (K-major) x Zj (MN-major), M128 N256
            const uint32_t a_addr = smem_u32(sA);
            mbar_wait(a_full, 0);
            Ring<C::NSTAGE> ring1;      // stage / phase of the tile whose MMA1 is issued next
            Ring<C::NS> sacc;
            auto issue_mma1 = [&]() {
                const uint32_t s = ring1.stage(), a = sacc.stage();
                mbar_wait(b_full + s, ring1.phase());
                if constexpr (!C::W_TMEM) mbar_wait(s_empty + a, sacc.phase() ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
                for (int kb = 0; kb < TC_KBLK; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if constexpr (C::A_TMEM)      // k-step (kb, ks) = elements 64*kb + 16*ks ..: 8 columns of Zi in TMEM
                            umma_bf16_ts(tmem_s + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                         smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc1, (kb | ks) != 0);
                        else
                            umma_bf16(tmem_s + a * BN, smem_desc(a_addr + kb * (TC_BM * 128) + ks * 32, 16, 1024),
                                      smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc1, (kb | ks) != 0);
                    }
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
                for (int ks = 0; ks < BN / 16; ++ks) {
                    // A: W [128 x 64] K-major, 32 B per k-step.  B: Zj read MN-major: N = d (4 groups of 64, LBO = BN*128),
                    // K = j (8-row groups, SBO = 1024); one k-step = 16 rows of j = 2048 B.
                    if constexpr (C::W_TMEM)      // W of tile t sits in S slot t % NS: half h at columns h*32 .. h*32+15, 8 per k-step
                        umma_bf16_ts(tmem_dz, tmem_s + (uint32_t)(t % C::NS) * BN + (ks >> 1) * 32 + (ks & 1) * 8,
                                     smem_desc(b_addr + ks * 2048, BN * 128, 1024), idesc2, (t | ks) != 0);
                    else
                        umma_bf16(tmem_dz, smem_desc(w_addr + ks * 32, 16, 1024), smem_desc(b_addr + ks * 2048, BN * 128, 1024),
                                  idesc2, (t | ks) != 0);      // t counts from this CTA's first tile
                }
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
        const float coef_i = stats[(size_t)CY_STAT_COEF * N + gi] * WS;
        const float invc_i = stats[(size_t)CY_STAT_INVC * N + gi] * WS;
        if constexpr (C::A_TMEM) {
            // this thread's half row of Zi (128 elements = 64 packed columns) -> tensor memory lanes q*32.., columns h*64..
            const uint4* src = reinterpret_cast<const uint4*>(zrows + (size_t)gi * ldz + h * 128);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint4 v = __ldg(src + c * 8 + e);
                    r[4 * e] = v.x; r[4 * e + 1] = v.y; r[4 * e + 2] = v.z; r[4 * e + 3] = v.w;
                }
                tmem_st_32x32(tmem_a + (uint32_t(q * 32) << 16) + h * 64 + c * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        float* wcol = sCol + ew * 96;                 // [0,32) labels (as int bits), [32,64) coef_j, [64,96) invc_j
        const int32_t row_lo = __reduce_min_sync(0xffffffffu, my_lab), row_hi = __reduce_max_sync(0xffffffffu, my_lab);
        float nx_lab = __int_as_float(labels[t0 * BN + h * 32 + lane]);
        float nx_coef = stats[(size_t)CY_STAT_COEF * N + t0 * BN + h * 32 + lane] * WS;
        float nx_invc = stats[(size_t)CY_STAT_INVC * N + t0 * BN + h * 32 + lane] * WS;
        Ring<C::NS> sacc;
        Ring<C::NW> wr;
        for (int t = 0; t < nt; ++t, sacc.next(), wr.next()) {
            const uint32_t a = sacc.stage(), w = wr.stage();
            const int jbase = (t0 + t) * BN + h * 32;
            __syncwarp();
            wcol[lane] = nx_lab;
            wcol[32 + lane] = nx_coef;
            wcol[64 + lane] = nx_invc;
            const bool may_have_pos = __reduce_max_sync(0xffffffffu, __float_as_int(nx_lab)) >= row_lo &&
                                      __reduce_min_sync(0xffffffffu, __float_as_int(nx_lab)) <= row_hi;
            __syncwarp();
            if (t + 1 < nt) {                             // next tile's column statistics travel during this tile
                nx_lab = __int_as_float(labels[jbase + BN + lane]);
                nx_coef = stats[(size_t)CY_STAT_COEF * N + jbase + BN + lane] * WS;
                nx_invc = stats[(size_t)CY_STAT_INVC * N + jbase + BN + lane] * WS;
            }
            mbar_wait(s_full + a, sacc.phase());
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            // S values are in registers: the accumulator can be reused (W_TMEM: the slot is handed back by MMA2 instead)
            if (!C::W_TMEM && lane == 0) mbar_arrive(s_empty + a);
            uint32_t packed[16];
            if (!may_have_pos) {                          // no positive pair in this 32 x 32 block: W = E (coef_i + coef_j)
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 cj = *reinterpret_cast<const float4*>(wcol + 32 + e4 * 4);
                    const float w0 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 0]), c1, -c1)) * (coef_i + cj.x);
                    const float w1 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 1]), c1, -c1)) * (coef_i + cj.y);
                    const float w2 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 2]), c1, -c1)) * (coef_i + cj.z);
                    const float w3 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 3]), c1, -c1)) * (coef_i + cj.w);
                    packed[e4 * 2] = pack2<F16>(w0, w1);
                    packed[e4 * 2 + 1] = pack2<F16>(w2, w3);
                }
            } else {
                const bool diag_tile = __any_sync(0xffffffffu, (gi >= jbase) && (gi < jbase + 32));
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const int4 lj = *reinterpret_cast<const int4*>(wcol + e4 * 4);
                    const float4 cj = *reinterpret_cast<const float4*>(wcol + 32 + e4 * 4);
                    const float4 ij = *reinterpret_cast<const float4*>(wcol + 64 + e4 * 4);
                    const int lv[4] = {lj.x, lj.y, lj.z, lj.w};
                    const float cv[4] = {cj.x, cj.y, cj.z, cj.w};
                    const float iv[4] = {ij.x, ij.y, ij.z, ij.w};
                    float wv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float sv = __uint_as_float(r[e4 * 4 + u]);
                        const float E = ex2_approx(fmaf(sv, c1, -c1));
                        float v = E * (coef_i + cv[u]);
                        if (lv[u] == my_lab) v -= invc_i + iv[u];
                        wv[u] = v;
                    }
                    if (diag_tile) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (jbase + e4 * 4 + u == gi) wv[u] = 0.f;
                    }
                    packed[e4 * 2] = pack2<F16>(wv[0], wv[1]);
                    packed[e4 * 2 + 1] = pack2<F16>(wv[2], wv[3]);
                }
            }
            if constexpr (C::W_TMEM) {
                // in place: this warp's 32 x 32 block of S (columns h*32 ..) becomes 32 x 32 bf16 weights in columns h*32 .. +15
                tc_fence_after();
                tmem_st_32x16(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, packed);
                tmem_st_wait();
                tc_fence_before();
            } else {
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(w_full + w);
        }
        // dZ rows of this CTA: TMEM -> registers -> bf16 -> global
        mbar_wait(dz_full, 0);
        tc_fence_after();
        if (dz32 == nullptr) {
            const float scale = gscale[0] * out_scale;
            uint16_t* out = dz + (size_t)gi * lddz + h * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * 128 + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        pk[u] = pack2<F16>(__uint_as_float(r[e + 2 * u]) * scale, __uint_as_float(r[e + 2 * u + 1]) * scale);
                    *reinterpret_cast<uint4*>(out + c * 32 + e) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        } else {
            // column-split launch: partial row-block gradients meet in an fp32 accumulator (vector reductions at L2)
            float* out = dz32 + (size_t)(gi - row_begin) * TC_D + h * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * 128 + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c * 32 + e),
                                 "f"(__uint_as_float(r[e])), "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])),
                                 "f"(__uint_as_float(r[e + 3]))
                                 : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// fp32 split accumulator -> bf16 gradient rows (scaled by gscale / (t N))
template <bool F16>
__global__ void infonce_tc_convert_kernel(const float* __restrict__ dz32, int64_t rows, int64_t row_begin,
                                          const float* __restrict__ gscale, float out_scale, uint16_t* __restrict__ dz,
                                          int64_t lddz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 8 consecutive columns
    if (idx >= rows * (TC_D / 8)) return;
    const int64_t r = idx / (TC_D / 8);
    const int c = (int)(idx % (TC_D / 8)) * 8;
    const float scale = gscale[0] * out_scale;
    const float4 a = *reinterpret_cast<const float4*>(dz32 + r * TC_D + c);
    const float4 b = *reinterpret_cast<const float4*>(dz32 + r * TC_D + c + 4);
    *reinterpret_cast<uint4*>(dz + (row_begin + r) * lddz + c) =
        make_uint4(pack2<F16>(a.x * scale, a.y * scale), pack2<F16>(a.z * scale, a.w * scale),
                   pack2<F16>(b.x * scale, b.y * scale), pack2<F16>(b.z * scale, b.w * scale));
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encode_fn() {
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
static int make_tmap(CUtensorMap* m, const void* z, int64_t N, int64_t ldz, int dtype) {
    EncodeTiledFn fn = tensor_map_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CY_ERR_DEVICE; }
    cuuint64_t gdim[2] = {(cuuint64_t)TC_D, (cuuint64_t)N};
    cuuint64_t gstride[1] = {(cuuint64_t)ldz * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dtype == CY_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(z), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CY_ERR_ARG; }
    return CY_OK;
}

bool infonce_tc_supported(int dtype, int64_t N, int64_t d, int64_t ldz, const uint8_t* codes, int variant) {
    return (dtype == CY_BF16 || dtype == CY_F16) && d == TC_D && (N % 128) == 0 && N >= 256 && (ldz % 8) == 0 && codes == nullptr &&
           variant == CY_SUPCON && N < (int64_t(1) << 30);
}

constexpr int FWD_BN = 128;

static int sm_count() { return device_sm_count(); }

// column splits of the backward: minimise waves x tiles-per-CTA (one CTA per SM), at least 16 tiles per CTA
static int bwd_splits(int64_t N, int64_t rows) {
    const int64_t sms = sm_count(), rb = rows / TC_BM, nt = N / BwdCfg::BN;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= 16; ++s) {
        if (s > 1 && nt / s < 16) break;
        const int64_t tps = (nt + s - 1) / s;
        const int64_t waves = (rb * s + sms - 1) / sms;
        const double cost = (double)waves * (double)(tps + 6);      // + fixed per-CTA cost (A load, drain) in tile units
        if (cost < best_cost * 0.97) { best_cost = cost; best = s; }
    }
    return best;
}

static int fwd_splits(int64_t N, int64_t rows) {
    const int sms = sm_count();
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
    const size_t fwd = (size_t)smax * 2 * 3 * (size_t)N * sizeof(float);
    const size_t bwd = (size_t)N * TC_D * sizeof(float);            // fp32 split accumulator, worst case rows == N
    return fwd > bwd ? fwd : bwd;
}

int infonce_fwd_tc(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
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
    int rc = make_tmap(&tmap, z, N, ldz, dtype);
    if (rc) return rc;
    const int fmt = dtype == CY_F16 ? 0 : 1;
    using S = FwdSmem<FWD_BN>;
    auto k = infonce_fwd_tc_kernel<FWD_BN>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
    if (e != cudaSuccess) { set_error("fwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int ctiles = (int)(N / FWD_BN);
    const int tps = (ctiles + splits - 1) / splits;
    dim3 grid((unsigned)(rows / TC_BM), (unsigned)splits);
    k<<<grid, TC_THREADS, S::TOTAL, st>>>(tmap, labels, (int)N, (int)row_begin, tps, inv_t * LOG2E,
                                          reinterpret_cast<float*>(workspace), idesc_f16kind_f32(TC_BM, FWD_BN, 0, 0, fmt),
                                          reinterpret_cast<const uint16_t*>(z), ldz);
    CY_CHECK_LAUNCH("infonce_fwd_tc");
    infonce_tc_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(reinterpret_cast<float*>(workspace), nslot, (int)N,
                                                                            (int)row_begin, (int)row_end, inv_t, stats);
    CY_CHECK_LAUNCH("infonce_tc_reduce");
    return CY_OK;
}

int infonce_bwd_tc(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                   float inv_t, const float* stats, const float* gscale, void* dz, int64_t lddz, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
    (void)d;
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    CY_CHECK_ARG((rows % TC_BM) == 0 && (row_begin % TC_BM) == 0, "tcgen05 path: row range must be 128-aligned");
    CY_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (lddz % 8) == 0,
                 "tcgen05 path: z / dz must be 16-byte aligned");
    CUtensorMap tmap;
    int rc = make_tmap(&tmap, z, N, ldz, dtype);
    if (rc) return rc;
    const bool f16 = dtype == CY_F16;
    const int fmt = f16 ? 0 : 1;
    auto kern = f16 ? infonce_bwd_tc_kernel<true> : infonce_bwd_tc_kernel<false>;
    const float out_scale = (inv_t / (float)N) * (f16 ? (1.f / 1024.f) : 1.f);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg::TOTAL);
    if (e != cudaSuccess) { set_error("bwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int splits = bwd_splits(N, rows);
    const int nt = (int)(N / BwdCfg::BN);
    const int tps = (nt + splits - 1) / splits;
    float* dz32 = nullptr;
    if (splits > 1) {
        const size_t need = (size_t)rows * TC_D * sizeof(float);
        CY_CHECK_ARG(workspace && workspace_bytes >= need, "infonce_bwd_tc: workspace %zu < %zu", workspace_bytes, need);
        dz32 = reinterpret_cast<float*>(workspace);
        e = cudaMemsetAsync(dz32, 0, need, st);
        if (e != cudaSuccess) { set_error("bwd_tc memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)(rows / TC_BM), (unsigned)splits);
    kern<<<grid, TC_THREADS, BwdCfg::TOTAL, st>>>(tmap, labels, stats, (int)N, (int)row_begin, tps, inv_t * LOG2E, gscale, out_scale,
                                                  dz, lddz, dz32, idesc_f16kind_f32(TC_BM, BwdCfg::BN, 0, 0, fmt),
                                                  idesc_f16kind_f32(TC_BM, TC_D, 0, 1, fmt), reinterpret_cast<const uint16_t*>(z), ldz);
    CY_CHECK_LAUNCH("infonce_bwd_tc");
    if (dz32) {
        const int64_t n8 = rows * (TC_D / 8);
        if (f16)
            infonce_tc_convert_kernel<true><<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(dz32, rows, row_begin, gscale, out_scale,
                                                                                        reinterpret_cast<uint16_t*>(dz), lddz);
        else
            infonce_tc_convert_kernel<false><<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(dz32, rows, row_begin, gscale, out_scale,
                                                                                         reinterpret_cast<uint16_t*>(dz), lddz);
        CY_CHECK_LAUNCH("infonce_tc_convert");
    }
    return CY_OK;
}

}  // namespace cy
