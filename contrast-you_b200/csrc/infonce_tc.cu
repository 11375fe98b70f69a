// InfoNCE / SupCon family on the 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM, flash-style.
//
// Scope: every variant of the family (SupConLoss1 with / without exclude_other_pos, SelfPacedSupConLoss hard / soft) with
// label-derived masks, bf16 / fp16 embeddings, d in {128, 256}, any N >= 256 (a ragged last tile is masked in the
// epilogue; TMA zero-fills the rows past N).  Explicit mask= codes, fp32 inputs and other d run on the SIMT path.
// The N x N similarity never exists in HBM: each CTA owns a 128-row block of Z (held in TENSOR MEMORY as the MMA "A"
// operand), streams column tiles of Z through a TMA ring ("B" operand, K-major, 128-byte swizzle), accumulates
// S = Zi Zj^T in TMEM and lets eight epilogue warps read it back with tcgen05.ld.
//
//   pass 1   (infonce_fwd_tc_kernel<D,1,*>):  per element E = 2^(s*c1 - c1), c1 = log2(e)/t; row sums D_i += E, positives
//            c_i += [lab_i == lab_j], posS_i += [..] s, posE_i += [..] E; diagonal / ragged columns excluded on the tiles
//            that hold them only.  Column range split over blockIdx.y; partial row sums go to a [slot][4][N] scratch and
//            are summed in a fixed order by infonce_rowstats_kernel (deterministic), which also forms the per-row
//            statistics (log-denominator, 1/c, coefficient, loss term) the backward and the loss reduction read.
//   pass 2   (infonce_fwd_tc_kernel<D,2,VARIANT>): exclude_other_pos / self-paced need per-element denominators / weights
//            that depend on the pass-1 row sums — but only for POSITIVE pairs.  A CTA therefore visits only the column
//            tiles whose label range intersects the label range of its rows (list built in shared memory from a per-tile
//            [min,max] table): with rows sorted by label (the modules do that) this is a handful of tiles per row block
//            instead of N/128, so the second sweep costs O(N), not O(N^2).  Unsorted callers get every tile: still exact.
//   backward (infonce_bwd_tc_kernel<D,F16,VARIANT>):  S tile recomputed, W_ij = G_ij + G_ji written as bf16 straight back
//            into the TMEM columns S was read from, second MMA dZ_i[128 x D] += W[128x64] Zj[64 x D] with the SAME Zj
//            bytes read MN-major; dZ accumulates in TMEM across all column tiles of the row block and is scaled by
//            gscale/(t N) on the way out (dZ = (1/t) W Z).  The variant only changes the
//            positive-pair branch of the epilogue.  Column splits write fp32 slabs that one conversion kernel sums in a
//            fixed order: bitwise reproducible gradients (the reference runs under use_deterministic_algorithms).
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane each), warp 2 TMEM allocator,
// warp 3 tile-list builder (pass 2), warps 4..11 epilogue (warp w reads TMEM lanes 32*(w%4).., column half (w-4)/4).
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cy {

using namespace tc;

constexpr int TC_BM = 128;           // rows per CTA (UMMA M)
constexpr int TC_THREADS = 384;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int FWD_BN = 128;          // column tile of the forward sweeps
constexpr int BWD_BN = 64;           // column tile of the backward
constexpr int P2_LISTCAP = 4096;     // pass-2 tile list capacity (uint16 entries) -> N <= 4096 * 128

// contrastive.py:197-204 (evaluated for positive pairs only; elsewhere max(w, 1-P) = 1 and is unused)
__device__ __forceinline__ float sp_weight_tc(int variant, float logp, float gamma) {
    if (variant == CY_SELFPACED_HARD) return (-logp <= gamma) ? 1.f : 0.f;
    return fmaxf(1.f + logp / gamma, 0.f);
}

// ------------------------------------------------------------------------------------------------------------ forward
// Zi lives in tensor memory (columns [384, 384 + D/2): two 16-bit elements per column), three S accumulators of 128 columns,
// three Zj stages.  With both operands in shared memory an M128 N128 K16 MMA fetches 8 KB per 64 clk — the whole 128 B/clk
// port, which TMA needs as well (round-1 profiles: 72 % tensor-active); with A in TMEM it is at 79 %.
// SPLIT (fp32 embeddings, CY_F32_SPLIT): every row arrives as [hi | lo] bf16 halves (z = hi + lo to 2^-17, cy_infonce_pack_split)
// and S = hi_i hi_j + hi_i lo_j + lo_i hi_j — three MMA terms into the same accumulator, products good to 2^-16: fp32 parity
// (1e-4) at a third of the bf16 rate instead of the CUDA-core path's 1 %.  Column tiles shrink to 64 so that three [hi | lo]
// stages still fit; Zi hi AND lo live in tensor memory (D columns), four 64-column accumulators.
template <int D, bool SPLIT>
struct FwdCfg {
    static constexpr int BN = SPLIT ? 64 : FWD_BN;
    static constexpr int KBLK = D / 64;                         // 64-element (128-byte) K blocks per term
    static constexpr int KBLK_B = SPLIT ? 2 * KBLK : KBLK;      // K blocks staged per column tile
    static constexpr int NSTAGE = 3;
    static constexpr int NACC = SPLIT ? 4 : 3;
    static constexpr int A_COLS = SPLIT ? D : D / 2;            // tensor-memory columns of Zi
    static constexpr int TMEM_A = 512 - A_COLS;
    static_assert(NACC * BN <= TMEM_A, "accumulators and Zi overlap in tensor memory");
    static constexpr uint32_t B_BYTES = BN * KBLK_B * 128;
    static constexpr uint32_t OFF_LAB = NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_LIST = OFF_LAB + 8 * (FWD_BN / 2) * 4;
    static constexpr uint32_t OFF_BAR = OFF_LIST + P2_LISTCAP * 2;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;   // + barriers / scalars + 1024-alignment slack
};

template <int D, int PASS, int VARIANT, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels, int N, int row_begin,
                      int ct_begin, int ct_end, int tiles_per_split, float c1, float inv_t, float gamma,
                      float* __restrict__ part, int slot_base, uint32_t idesc, const uint16_t* __restrict__ zrows, int64_t ldz,
                      const int2* __restrict__ tile_range, const float4* __restrict__ xstat) {
    using S = FwdCfg<D, SPLIT>;
    constexpr int BN = S::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // 1024-aligned, still a shared pointer
    uint8_t* sB = smem;
    int32_t* sLab = reinterpret_cast<int32_t*>(smem + S::OFF_LAB);
    uint16_t* sList = reinterpret_cast<uint16_t*>(smem + S::OFF_LIST);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + S::NSTAGE;
    uint64_t* acc_full = b_empty + S::NSTAGE;
    uint64_t* acc_empty = acc_full + S::NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + S::NACC);
    int32_t* sMisc = reinterpret_cast<int32_t*>(tmem_slot + 1);      // [0..7] row label min / max per warp, [8] list length

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = row_begin + blockIdx.x * TC_BM;              // first global row of this CTA

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 8);
        for (int i = 0; i < S::NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < S::NACC; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if constexpr (PASS == 2) {
        if (threadIdx.x < TC_BM) {                                // label range of this CTA's rows
            const int gi = row0 + (int)threadIdx.x;
            const int32_t lab = labels[min(gi, N - 1)];
            const int32_t lo = __reduce_min_sync(0xffffffffu, gi < N ? lab : INT_MAX);
            const int32_t hi = __reduce_max_sync(0xffffffffu, gi < N ? lab : INT_MIN);
            if (lane == 0) { sMisc[2 * warp] = lo; sMisc[2 * warp + 1] = hi; }
        }
        __syncthreads();
        if (warp == 3) {                                          // column tiles that can hold a positive pair of these rows
            const int32_t rlo = min(min(sMisc[0], sMisc[2]), min(sMisc[4], sMisc[6]));
            const int32_t rhi = max(max(sMisc[1], sMisc[3]), max(sMisc[5], sMisc[7]));
            int count = 0;
            for (int base = ct_begin; base < ct_end; base += 32) {
                const int ct = base + lane;
                bool hit = false;
                if (ct < ct_end) {
                    const int2 r = tile_range[ct];
                    hit = r.y >= rlo && r.x <= rhi;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (hit) sList[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)ct;
                count += __popc(m);
            }
            if (lane == 0) sMisc[8] = count;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_a = tmem_base + S::TMEM_A;  // Zi: hi in columns [0, D/2) of this range, lo (SPLIT) in [D/2, D)

    int ct0 = 0, ntile;
    if constexpr (PASS == 1) {
        ct0 = ct_begin + (int)blockIdx.y * tiles_per_split;
        ntile = max(0, min(ct_end, ct0 + tiles_per_split) - ct0);
    } else {
        ntile = sMisc[8];
    }
    auto tile_at = [&](int k) -> int { return PASS == 1 ? ct0 + k : (int)sList[k]; };

    if (warp == 0) {
        if (elect_one()) {
            Ring<S::NSTAGE> ring;
            for (int k = 0; k < ntile; ++k, ring.next()) {
                const int ct = tile_at(k);
                const uint32_t s = ring.stage();
                mbar_wait(b_empty + s, ring.phase() ^ 1u);
                mbar_arrive_expect_tx(b_full + s, S::B_BYTES);
                uint8_t* dst = sB + s * S::B_BYTES;
                for (int kb = 0; kb < S::KBLK_B; ++kb)      // (SPLIT: K blocks KBLK.. are the lo halves, columns D.. of the row)
                    for (int hb = 0; hb < BN / 64; ++hb)
                        tma_load_2d(dst + kb * (BN * 128) + hb * 8192, &tmap, b_full + s, kb * 64, ct * BN + hb * 64);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            mbar_wait(a_full, 0);                   // idesc: M128 x N(BN) x K16, A from TMEM, B K-major, bf16 or fp16
            Ring<S::NSTAGE> ring;
            Ring<S::NACC> acc;
            for (int k = 0; k < ntile; ++k, ring.next(), acc.next()) {
                const uint32_t s = ring.stage(), a = acc.stage();
                mbar_wait(b_full + s, ring.phase());
                mbar_wait(acc_empty + a, acc.phase() ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * S::B_BYTES);
#pragma unroll
                for (int kb = 0; kb < S::KBLK; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16_ts(tmem_base + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                     smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                if constexpr (SPLIT) {
#pragma unroll
                    for (int kb = 0; kb < S::KBLK; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // hi_i . lo_j
                            umma_bf16_ts(tmem_base + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                         smem_desc(b_addr + (S::KBLK + kb) * (BN * 128) + ks * 32, 16, 1024), idesc, 1u);
#pragma unroll
                    for (int kb = 0; kb < S::KBLK; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // lo_i . hi_j
                            umma_bf16_ts(tmem_base + a * BN, tmem_a + D / 2 + (kb * 4 + ks) * 8,
                                         smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc, 1u);
                }
                umma_commit(b_empty + s);
                umma_commit(acc_full + a);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int row = q * 32 + lane;
        const int gi = row0 + row;
        const int gic = min(gi, N - 1);                        // rows past N (ragged last block) compute on row N-1, store nothing
        const int32_t my_lab = labels[gic];
        {
            // this thread's half row of Zi (D/2 elements = D/4 packed columns) -> tensor memory lanes q*32.., columns h*D/4..
            // (SPLIT: the same for the lo half of the row, D elements further, into the columns D/2 further)
#pragma unroll
            for (int part_ = 0; part_ < (SPLIT ? 2 : 1); ++part_) {
                const uint4* src = reinterpret_cast<const uint4*>(zrows + (size_t)gic * ldz + part_ * D + h * (D / 2));
#pragma unroll
                for (int c = 0; c < D / 128; ++c) {
                    uint32_t r[32];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint4 v = __ldg(src + c * 8 + e);
                        r[4 * e] = v.x; r[4 * e + 1] = v.y; r[4 * e + 2] = v.z; r[4 * e + 3] = v.w;
                    }
                    tmem_st_32x32(tmem_a + (uint32_t(q * 32) << 16) + part_ * (D / 2) + h * (D / 4) + c * 32, r);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        int32_t* wlab = sLab + ew * (BN / 2);
        constexpr int NCH = BN / 64;                       // 32-column chunks per warp and tile (its half of the BN columns)
        // label range of this warp's 32 rows: a column chunk whose label range does not intersect it holds no positive
        // pair, whatever the order of the rows (callers that sort rows by label make this the common case)
        const int32_t row_lo = __reduce_min_sync(0xffffffffu, my_lab), row_hi = __reduce_max_sync(0xffffffffu, my_lab);
        int32_t lab_next[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) lab_next[c] = 0;
        if (ntile > 0) {
            const int j0 = tile_at(0) * BN + h * (BN / 2);
#pragma unroll
            for (int c = 0; c < NCH; ++c) lab_next[c] = labels[min(j0 + c * 32 + lane, N - 1)];
        }
        // pass 1 accumulators
        float D0 = 0.f, D1 = 0.f, D2 = 0.f, D3 = 0.f, posS = 0.f, posE = 0.f;
        int cnt = 0;
        // pass 2 accumulators and the row statistic they need (exclude: A_i; self-paced: log-denominator)
        float r0 = 0.f, r1 = 0.f, rstat = 0.f;
        if constexpr (PASS == 2) {
            const float4 xs = xstat[gic];
            rstat = (VARIANT == CY_SUPCON_EXCLUDE) ? xs.w : xs.x;
        }
        Ring<S::NACC> acc;
        for (int k = 0; k < ntile; ++k, acc.next()) {
            const uint32_t a = acc.stage();
            const int ct = tile_at(k);
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
            if (k + 1 < ntile) {                               // next tile's labels travel while this tile is processed
                const int jn = tile_at(k + 1) * BN + h * (BN / 2);
#pragma unroll
                for (int c = 0; c < NCH; ++c) lab_next[c] = labels[min(jn + c * 32 + lane, N - 1)];
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
            const bool ragged = (ct + 1) * BN > N;             // uniform per CTA: columns past N are TMA zero fill
            const bool diag_tile = __any_sync(0xffffffffu, (gi >= jbase) && (gi < jbase + BN / 2));
            if constexpr (PASS == 1) {
                if (!may_have_pos && !ragged) {               // (a diagonal chunk always intersects: i is its own label)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) {
                            D0 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 0]), c1, -c1));
                            D1 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 1]), c1, -c1));
                            D2 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 2]), c1, -c1));
                            D3 += ex2_approx(fmaf(__uint_as_float(r[c][e4 * 4 + 3]), c1, -c1));
                        }
                    }
                } else if (!diag_tile && !ragged) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) {
                            const int4 lj = *reinterpret_cast<const int4*>(wlab + c * 32 + e4 * 4);
                            const float s0 = __uint_as_float(r[c][e4 * 4 + 0]), s1 = __uint_as_float(r[c][e4 * 4 + 1]);
                            const float s2 = __uint_as_float(r[c][e4 * 4 + 2]), s3 = __uint_as_float(r[c][e4 * 4 + 3]);
                            const float e0 = ex2_approx(fmaf(s0, c1, -c1)), e1 = ex2_approx(fmaf(s1, c1, -c1));
                            const float e2 = ex2_approx(fmaf(s2, c1, -c1)), e3 = ex2_approx(fmaf(s3, c1, -c1));
                            D0 += e0; D1 += e1; D2 += e2; D3 += e3;
                            if (lj.x == my_lab) { cnt += 1; posS += s0; posE += e0; }
                            if (lj.y == my_lab) { cnt += 1; posS += s1; posE += e1; }
                            if (lj.z == my_lab) { cnt += 1; posS += s2; posE += e2; }
                            if (lj.w == my_lab) { cnt += 1; posS += s3; posE += e3; }
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int j = jbase + c * 32 + e;
                            const float sv = __uint_as_float(r[c][e]);
                            if (j != gi && j < N) {
                                const float ev = ex2_approx(fmaf(sv, c1, -c1));
                                D0 += ev;
                                if (wlab[c * 32 + e] == my_lab) { cnt += 1; posS += sv; posE += ev; }
                            }
                        }
                    }
                }
            } else {
                if (may_have_pos) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int j = jbase + c * 32 + e;
                            if (j != gi && j < N && wlab[c * 32 + e] == my_lab) {
                                const float L = fmaf(__uint_as_float(r[c][e]), inv_t, -inv_t);
                                if (VARIANT == CY_SUPCON_EXCLUDE) {
                                    const float den = __expf(L) + rstat + 1e-16f;
                                    r0 += L - __logf(den);
                                    r1 += 1.f / den;
                                } else {
                                    const float logp = L - rstat;
                                    const float w = sp_weight_tc(VARIANT, logp, gamma);
                                    r0 += w * logp;
                                    r1 += w;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (gi < N) {
            if constexpr (PASS == 1) {
                float* p = part + (size_t)(slot_base + (int)blockIdx.y * 2 + h) * 4 * N;
                p[gi] = (D0 + D1) + (D2 + D3);
                p[(size_t)N + gi] = (float)cnt;
                p[2 * (size_t)N + gi] = posS;
                p[3 * (size_t)N + gi] = posE;
            } else {
                float* p = part + (size_t)h * 2 * N;
                p[gi] = r0;
                p[(size_t)N + gi] = r1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// [min, max] label of every 128-column tile (pass-2 tile selection)
__global__ void infonce_tile_range_kernel(const int32_t* __restrict__ labels, int N, int bn, int2* __restrict__ tile_range) {
    const int ct = blockIdx.x, lane = threadIdx.x;          // one warp per tile of bn columns
    int32_t lo = INT_MAX, hi = INT_MIN;
    for (int j = ct * bn + lane; j < min(N, (ct + 1) * bn); j += 32) {
        const int32_t l = labels[j];
        lo = min(lo, l);
        hi = max(hi, l);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) tile_range[ct] = make_int2(lo, hi);
}

// ------------------------------------------------------------------------------------------------------------ row statistics
// Per-row epilogue of a forward sweep, for the rows [row_begin, row_end) this process owns.  Source of the raw sums: the
// tensor path's partial slots (`part` != nullptr: summed here in a fixed order) or the CY_STAT_* rows the SIMT sweep wrote.
// Output: `xstat` [N][4] — what the backward needs for every row it meets as a row OR as a column, plus the row's loss
// term, so that ONE all-gather of the owned rows of this array is the whole exchange of the row-sharded multi-GPU form
// (include/contrastyou_b200.h: CY_XS_*).
template <int PASS>
__global__ void infonce_rowstats_kernel(int variant, int N, int row_begin, int row_end, float inv_t, const float* __restrict__ part,
                                        int nslot, float* __restrict__ stats, float4* __restrict__ xstat) {
    const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    const size_t Ns = (size_t)N;
    if (PASS == 1) {
        float posE, negE, c, nc, posL;
        if (part) {
            float Dt = 0.f, ct = 0.f, ps = 0.f, pe = 0.f;
            for (int s = 0; s < nslot; ++s) {
                const float* p = part + (size_t)s * 4 * Ns;
                Dt += p[i];
                ct += p[Ns + i];
                ps += p[2 * Ns + i];
                pe += p[3 * Ns + i];
            }
            posE = pe;
            negE = Dt - pe;                                     // every off-diagonal pair is positive or negative (label masks)
            c = ct;
            nc = (float)(N - 1) - ct;
            posL = inv_t * ps - inv_t * ct;                     // sum_j P_ij (s_ij - 1)/t
            stats[CY_STAT_AUX * Ns + i] = negE;
            stats[CY_STAT_NEGC * Ns + i] = nc;
            stats[CY_STAT_POSL * Ns + i] = posL;
        } else {
            posE = stats[CY_STAT_POSE * Ns + i];
            negE = stats[CY_STAT_AUX * Ns + i];
            c = stats[CY_STAT_INVC * Ns + i];
            nc = stats[CY_STAT_NEGC * Ns + i];
            posL = stats[CY_STAT_POSL * Ns + i];
        }
        const float den = posE + negE + 1e-16f;
        const float logden = logf(den);
        const float invc = 1.f / c;
        stats[CY_STAT_LOGDEN * Ns + i] = logden;
        stats[CY_STAT_INVC * Ns + i] = invc;
        stats[CY_STAT_POSE * Ns + i] = c;                       // slot reused: positive count (read by pass 2)
        float4 xs = make_float4(logden, invc, 0.f, 0.f);
        if (variant == CY_SUPCON) {
            xs.z = 1.f / den;
            xs.w = -(posL * invc - logden);                     // c == 0: 0 * inf = NaN, like the reference's 0/0 (contrastive.py:95)
            stats[CY_STAT_COEF * Ns + i] = xs.z;
        } else if (variant == CY_SUPCON_EXCLUDE) {
            const float ratio = nc / (c + nc);                  // contrastive.py:88 (float32)
            xs.w = negE / (ratio + 1e-4f);                      // A_i
            stats[CY_STAT_AUX * Ns + i] = xs.w;
        }
        xstat[i] = xs;
    } else {
        float r0, r1;
        if (part) {
            r0 = part[i] + part[2 * Ns + i];
            r1 = part[Ns + i] + part[3 * Ns + i];
            stats[CY_STAT_POSL * Ns + i] = r0;
            stats[CY_STAT_SW * Ns + i] = r1;
        } else {
            r0 = stats[CY_STAT_POSL * Ns + i];
            r1 = stats[CY_STAT_SW * Ns + i];
        }
        const float invc = stats[CY_STAT_INVC * Ns + i], c = stats[CY_STAT_POSE * Ns + i];
        const float term = -r0 * invc;
        float4 xs = xstat[i];
        if (variant == CY_SUPCON_EXCLUDE) {
            const float nc = stats[CY_STAT_NEGC * Ns + i];
            const float ratio = nc / (c + nc);
            xs.z = r1 * invc / (ratio + 1e-4f);
            xs.x = term;                                        // the log-denominator is not needed by this variant's backward
        } else {
            xs.z = r1 * invc * expf(-xs.x);
            xs.w = term;
        }
        stats[CY_STAT_COEF * Ns + i] = xs.z;
        xstat[i] = xs;
    }
}

// ------------------------------------------------------------------------------------------------------------ loss reduction
// out4 = { sum_i term_i / N, sum_i sw_i, sum_i c_i, #non-finite terms } over ALL N rows of xstat, in a fixed order (block
// partials in index order): every rank of a sharded run computes the identical bits from the gathered array, so no
// scalar collective is needed.
constexpr int LOSS_THREADS = 1024;
__global__ void __launch_bounds__(LOSS_THREADS)
infonce_loss_partial_kernel(int variant, int N, const float4* __restrict__ xstat, float* __restrict__ partials) {
    const int i = blockIdx.x * LOSS_THREADS + threadIdx.x;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < N) {
        const float4 xs = xstat[i];
        const float term = (variant == CY_SUPCON_EXCLUDE) ? xs.x : xs.w;
        v[0] = term;
        v[3] = isfinite(term) ? 0.f : 1.f;
        if (variant == CY_SELFPACED_HARD || variant == CY_SELFPACED_SOFT) {
            // coef = sw * invc * exp(-logden)  ->  sw = coef * exp(logden) / invc ;  c = 1 / invc
            const bool ok = isfinite(xs.y) && xs.y > 0.f;
            v[1] = ok ? xs.z * expf(xs.x) / xs.y : 0.f;
            v[2] = ok ? 1.f / xs.y : 0.f;
        }
    }
    __shared__ float red[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = warp_sum(v[q]);
    if (lane == 0)
        for (int q = 0; q < 4; ++q) red[q][w] = v[q];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x = warp_sum(red[q][lane]);
            if (lane == 0) partials[blockIdx.x * 4 + q] = x;
        }
    }
}

// out8[0..3] as above; out8[4] = *bad_rows (un-normalised rows counted by cy_infonce_pack), out8[5] = *overflow (int64 labels
// outside int32, cy_labels_canonicalize), out8[6..7] = 0: everything the module's per-step host check needs in ONE 32-byte read
__global__ void infonce_loss_final_kernel(int nblk, int N, const float* __restrict__ partials, float* __restrict__ out8,
                                          int32_t* __restrict__ bad_rows, const int32_t* __restrict__ overflow) {
    const int q = threadIdx.x;       // 4 threads, each walks its component in block order
    if (q < 4) {
        float s = 0.f;
        for (int b = 0; b < nblk; ++b) s += partials[b * 4 + q];
        out8[q] = (q == 0) ? s / (float)N : s;
    } else if (q == 4) {
        out8[4] = bad_rows ? (float)bad_rows[0] : 0.f;
        if (bad_rows) bad_rows[0] = 0;      // the counter is consumed: a persistent one is ready for the next step's pack
    } else if (q == 5) {
        out8[5] = overflow ? (float)overflow[0] : 0.f;
    } else if (q < 8) {
        out8[q] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------ backward
// The row block Zi (the A operand of S = Zi Zj^T) lives in TENSOR MEMORY (D/2 columns: two bf16 per column): an M128 N64
// K16 MMA with both operands in shared memory fetches 6 KB and is bound by the 128 B/clk shared-memory port (48 clk,
// profiles/probes/probe_umma_small.cu) instead of its 32 clk of math; with A in TMEM it fetches 2 KB.  The weight tile W
// (A operand of dZ += W Zj) is written by the epilogue straight back into the tensor-memory columns of the S accumulator
// it was computed from (bf16 pairs: 16 columns per 32-column half) and read from there by the second MMA — no
// shared-memory W tile, no swizzled stores, no proxy fence.  The S slot is recycled by the in-order tensor pipe: MMA1(t+2)
// is issued after MMA2(t), which is the last reader of slot t % 2.
// SPLIT (fp32 embeddings as [hi | lo] bf16 halves): S = hi_i hi_j + hi_i lo_j + lo_i hi_j and dZ += W_hi Zj_hi + W_lo Zj_hi +
// W_hi Zj_lo — six MMA groups per tile instead of two.  Tensor memory is full with dZ (D fp32 columns), Zi_hi and the two S
// slots, so Zi_lo is a K-major shared-memory tile (the one SS-mode MMA of the kernel) and the Zj ring shrinks to two
// [hi | lo] stages; the gradient leaves in fp32.
template <int D, bool SPLIT>
struct BwdCfg {
    static constexpr int BN = BWD_BN;
    static constexpr int KBLK = D / 64;
    static constexpr int KBLK_B = SPLIT ? 2 * KBLK : KBLK;
    static constexpr int NSTAGE = SPLIT ? 2 : 6;     // Zj ring (a stage lives from its MMA1 until its MMA2 retires)
    static constexpr int NS = 2;         // S accumulators in TMEM (64 columns each), the last NS*64 columns
    static constexpr uint32_t B_BYTES = BN * KBLK_B * 128;
    static constexpr uint32_t ALO_BYTES = SPLIT ? TC_BM * D * 2 : 0;      // Zi_lo [KBLK][128 rows x 128 B]
    static constexpr uint32_t OFF_ALO = NSTAGE * B_BYTES;
    static constexpr uint32_t OFF_COL = OFF_ALO + ALO_BYTES;    // per epilogue warp: lab[32], coef[32], invc[32], aux[32]
    static constexpr uint32_t OFF_BAR = OFF_COL + 8 * 4 * 32 * 4;
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
template <int D, bool F16, int VARIANT, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ labels,
                      const float4* __restrict__ xstat, int N, int row_begin, int row_end, int tiles_per_split, float c1,
                      float inv_t, float gamma, const float* __restrict__ gscale, float out_scale, void* __restrict__ dz_v,
                      int64_t lddz, float* __restrict__ dz32, int rows_total, uint32_t idesc1, uint32_t idesc2,
                      const uint16_t* __restrict__ zrows, int64_t ldz) {
    uint16_t* dz = reinterpret_cast<uint16_t*>(dz_v);
    constexpr float WS = F16 ? 1024.f : 1.f;
    using C = BwdCfg<D, SPLIT>;
    constexpr int BN = C::BN;
    static_assert(!(SPLIT && F16), "the split path carries bf16 halves");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // 1024-aligned, still a shared pointer
    uint8_t* sB = smem;
    uint8_t* sAlo = smem + C::OFF_ALO;
    float* sCol = reinterpret_cast<float*>(smem + C::OFF_COL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + C::NSTAGE;
    uint64_t* s_full = b_empty + C::NSTAGE;
    uint64_t* w_full = s_full + C::NS;
    uint64_t* dz_full = w_full + C::NS;
    uint64_t* alo_full = dz_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(alo_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = row_begin + blockIdx.x * TC_BM;
    // column tiles [t0, t1) of this CTA (blockIdx.y splits the columns so that small row ranges still fill the chip)
    const int n_tiles = (N + BN - 1) / BN;
    const int t0 = blockIdx.y * tiles_per_split;
    const int t1 = min(n_tiles, t0 + tiles_per_split);
    const int nt = t1 - t0;
    if (nt <= 0) return;          // (the host sizes the splits so that this never happens; slabs would stay unwritten)

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 8);
        for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < C::NS; ++i) { mbar_init(s_full + i, 1); mbar_init(w_full + i, 8); }
        mbar_init(dz_full, 1);
        mbar_init(alo_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_dz = tmem_base;             // columns [0, D)
    const uint32_t tmem_a = tmem_base + 256;        // Zi, columns [256, 256 + D/2)
    const uint32_t tmem_s = tmem_base + 512 - C::NS * BN;        // NS x 64 columns

    if (warp == 0) {
        if (elect_one()) {
            if constexpr (SPLIT) {          // Zi_lo: columns D.. of the row block, K-major tile (rows past N: TMA zero fill)
                mbar_arrive_expect_tx(alo_full, C::ALO_BYTES);
                for (int kb = 0; kb < C::KBLK; ++kb)
                    for (int hb = 0; hb < TC_BM / 64; ++hb)
                        tma_load_2d(sAlo + kb * (TC_BM * 128) + hb * 8192, &tmap, alo_full, D + kb * 64, row0 + hb * 64);
            }
            Ring<C::NSTAGE> ring;
            for (int t = 0; t < nt; ++t, ring.next()) {
                const uint32_t s = ring.stage();
                mbar_wait(b_empty + s, ring.phase() ^ 1u);
                mbar_arrive_expect_tx(b_full + s, C::B_BYTES);
                uint8_t* dst = sB + s * C::B_BYTES;
                for (int kb = 0; kb < C::KBLK_B; ++kb) tma_load_2d(dst + kb * (BN * 128), &tmap, b_full + s, kb * 64, (t0 + t) * BN);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // idesc1: S = Zi (TMEM) x Zj (K-major), M128 N64;  idesc2: dZ += W (TMEM) x Zj (MN-major), M128 N(D)
            mbar_wait(a_full, 0);
            if constexpr (SPLIT) mbar_wait(alo_full, 0);
            const uint32_t alo_addr = smem_u32(sAlo);
            Ring<C::NSTAGE> ring1;      // stage / phase of the tile whose MMA1 is issued next
            Ring<C::NS> sacc;
            auto issue_mma1 = [&]() {
                const uint32_t s = ring1.stage(), a = sacc.stage();
                mbar_wait(b_full + s, ring1.phase());
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
                for (int kb = 0; kb < C::KBLK; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)      // k-step (kb, ks) = elements 64*kb + 16*ks ..: 8 columns of Zi in TMEM
                        umma_bf16_ts(tmem_s + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                     smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc1, (kb | ks) != 0);
                if constexpr (SPLIT) {
#pragma unroll
                    for (int kb = 0; kb < C::KBLK; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // hi_i . lo_j
                            umma_bf16_ts(tmem_s + a * BN, tmem_a + (kb * 4 + ks) * 8,
                                         smem_desc(b_addr + (C::KBLK + kb) * (BN * 128) + ks * 32, 16, 1024), idesc1, 1u);
#pragma unroll
                    for (int kb = 0; kb < C::KBLK; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // lo_i (shared memory, K-major) . hi_j
                            umma_bf16(tmem_s + a * BN, smem_desc(alo_addr + kb * (TC_BM * 128) + ks * 32, 16, 1024),
                                      smem_desc(b_addr + kb * (BN * 128) + ks * 32, 16, 1024), idesc1, 1u);
                }
                umma_commit(s_full + a);
                ring1.next();
                sacc.next();
            };
            issue_mma1();
            Ring<C::NSTAGE> ring2;
            Ring<C::NS> wr;
            for (int t = 0; t < nt; ++t, ring2.next(), wr.next()) {
                if (t + 1 < nt) issue_mma1();                       // keep the tensor pipe busy while the epilogue works
                const uint32_t s = ring2.stage(), w = wr.stage();
                mbar_wait(w_full + w, wr.phase());
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks)
                    // A: W of tile t sits in S slot t % NS: half h at columns h*32 .. h*32+15, 8 columns per k-step.
                    // B: Zj read MN-major: N = d (D/64 groups of 64, LBO = BN*128), K = j (8-row groups, SBO = 1024); one
                    // k-step = 16 rows of j = 2048 B.
                    umma_bf16_ts(tmem_dz, tmem_s + (uint32_t)(t % C::NS) * BN + (ks >> 1) * 32 + (ks & 1) * 8,
                                 smem_desc(b_addr + ks * 2048, BN * 128, 1024), idesc2, (t | ks) != 0);
                if constexpr (SPLIT) {
#pragma unroll
                    for (int ks = 0; ks < BN / 16; ++ks)      // W_lo (columns +16 of the half) . Zj_hi
                        umma_bf16_ts(tmem_dz, tmem_s + (uint32_t)(t % C::NS) * BN + (ks >> 1) * 32 + 16 + (ks & 1) * 8,
                                     smem_desc(b_addr + ks * 2048, BN * 128, 1024), idesc2, 1u);
#pragma unroll
                    for (int ks = 0; ks < BN / 16; ++ks)      // W_hi . Zj_lo (K blocks KBLK.. of the stage)
                        umma_bf16_ts(tmem_dz, tmem_s + (uint32_t)(t % C::NS) * BN + (ks >> 1) * 32 + (ks & 1) * 8,
                                     smem_desc(b_addr + C::KBLK * (BN * 128) + ks * 2048, BN * 128, 1024), idesc2, 1u);
                }
                umma_commit(b_empty + s);
            }
            umma_commit(dz_full);
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int row = q * 32 + lane;
        const int gi = row0 + row;
        const int gic = min(gi, N - 1);
        const int32_t my_lab = labels[gic];
        const float4 xi = xstat[gic];
        const float coef_i = xi.z * WS, invc_i = xi.y * WS;
        const float aux_i = (VARIANT == CY_SUPCON_EXCLUDE) ? xi.w : xi.x;      // exclude: A_i;  self-paced: log-denominator
        {
            // this thread's half row of Zi (D/2 elements = D/4 packed columns) -> tensor memory lanes q*32.., columns h*D/4..
            const uint4* src = reinterpret_cast<const uint4*>(zrows + (size_t)gic * ldz + h * (D / 2));
#pragma unroll
            for (int c = 0; c < D / 128; ++c) {
                uint32_t r[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint4 v = __ldg(src + c * 8 + e);
                    r[4 * e] = v.x; r[4 * e + 1] = v.y; r[4 * e + 2] = v.z; r[4 * e + 3] = v.w;
                }
                tmem_st_32x32(tmem_a + (uint32_t(q * 32) << 16) + h * (D / 4) + c * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        float* wcol = sCol + ew * 128;                // [0,32) labels (as int bits), [32,64) coef_j, [64,96) invc_j, [96,128) aux_j
        const int32_t row_lo = __reduce_min_sync(0xffffffffu, my_lab), row_hi = __reduce_max_sync(0xffffffffu, my_lab);
        int jn = min(t0 * BN + h * 32 + lane, N - 1);
        float nx_lab = __int_as_float(labels[jn]);
        float4 nx = xstat[jn];
        Ring<C::NS> sacc;
        for (int t = 0; t < nt; ++t, sacc.next()) {
            const uint32_t a = sacc.stage();
            const int jbase = (t0 + t) * BN + h * 32;
            __syncwarp();
            wcol[lane] = nx_lab;
            wcol[32 + lane] = nx.z * WS;
            wcol[64 + lane] = nx.y * WS;
            wcol[96 + lane] = (VARIANT == CY_SUPCON_EXCLUDE) ? nx.w : nx.x;
            const bool ragged = (t0 + t + 1) * BN > N;            // uniform per CTA: columns past N are TMA zero fill
            const bool may_have_pos = ragged || (__reduce_max_sync(0xffffffffu, __float_as_int(nx_lab)) >= row_lo &&
                                                 __reduce_min_sync(0xffffffffu, __float_as_int(nx_lab)) <= row_hi);
            __syncwarp();
            if (t + 1 < nt) {                             // next tile's column statistics travel during this tile
                jn = min(jbase + BN + lane, N - 1);
                nx_lab = __int_as_float(labels[jn]);
                nx = xstat[jn];
            }
            mbar_wait(s_full + a, sacc.phase());
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            uint32_t packed[SPLIT ? 32 : 16];            // [0,16): W (hi) pairs; SPLIT: [16,32): the bf16 remainders
            auto put = [&](int idx, float a, float b) {
                const uint32_t hi = pack2<F16>(a, b);
                packed[idx] = hi;
                if constexpr (SPLIT) packed[16 + idx] = pack2<false>(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
            };
            if (!may_have_pos) {                          // no positive pair in this 32 x 32 block: W = E (coef_i + coef_j)
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 cj = *reinterpret_cast<const float4*>(wcol + 32 + e4 * 4);
                    const float w0 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 0]), c1, -c1)) * (coef_i + cj.x);
                    const float w1 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 1]), c1, -c1)) * (coef_i + cj.y);
                    const float w2 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 2]), c1, -c1)) * (coef_i + cj.z);
                    const float w3 = ex2_approx(fmaf(__uint_as_float(r[e4 * 4 + 3]), c1, -c1)) * (coef_i + cj.w);
                    put(e4 * 2, w0, w1);
                    put(e4 * 2 + 1, w2, w3);
                }
            } else {
                const bool diag_tile = __any_sync(0xffffffffu, (gi >= jbase) && (gi < jbase + 32));
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const int4 lj = *reinterpret_cast<const int4*>(wcol + e4 * 4);
                    const float4 cj = *reinterpret_cast<const float4*>(wcol + 32 + e4 * 4);
                    const float4 ij = *reinterpret_cast<const float4*>(wcol + 64 + e4 * 4);
                    const float4 aj = *reinterpret_cast<const float4*>(wcol + 96 + e4 * 4);
                    const int lv[4] = {lj.x, lj.y, lj.z, lj.w};
                    const float cv[4] = {cj.x, cj.y, cj.z, cj.w};
                    const float iv[4] = {ij.x, ij.y, ij.z, ij.w};
                    const float av[4] = {aj.x, aj.y, aj.z, aj.w};
                    float wv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float sv = __uint_as_float(r[e4 * 4 + u]);
                        const float E = ex2_approx(fmaf(sv, c1, -c1));
                        const bool pos = lv[u] == my_lab;
                        float v = E * (coef_i + cv[u]);
                        if (VARIANT == CY_SUPCON) {
                            if (pos) v -= invc_i + iv[u];
                        } else if (VARIANT == CY_SUPCON_EXCLUDE) {
                            // positive pair: G_ij = (1/c_i)(E/(E + A_i + 1e-16) - 1); negative pair: coef_i E (contrastive.py:87-90)
                            if (pos) v = invc_i * (E / (E + aux_i + 1e-16f) - 1.f) + iv[u] * (E / (E + av[u] + 1e-16f) - 1.f);
                        } else {
                            if (pos) {                    // self-paced: the positive term carries the no-grad weights w_ij, w_ji
                                const float L = fmaf(sv, inv_t, -inv_t);
                                v -= sp_weight_tc(VARIANT, L - aux_i, gamma) * invc_i + sp_weight_tc(VARIANT, L - av[u], gamma) * iv[u];
                            }
                        }
                        wv[u] = v;
                    }
                    if (diag_tile || ragged) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = jbase + e4 * 4 + u;
                            if (j == gi || j >= N) wv[u] = 0.f;
                        }
                    }
                    put(e4 * 2, wv[0], wv[1]);
                    put(e4 * 2 + 1, wv[2], wv[3]);
                }
            }
            // in place: this warp's 32 x 32 block of S (columns h*32 ..) becomes 32 x 32 bf16 weights in columns h*32 .. +15
            // (SPLIT: and their bf16 remainders in columns h*32+16 .. +31)
            tc_fence_after();
            if constexpr (SPLIT) tmem_st_32x32(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, packed);
            else tmem_st_32x16(tmem_s + (uint32_t(q * 32) << 16) + a * BN + h * 32, packed);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_full + a);
        }
        // dZ rows of this CTA: TMEM -> registers -> bf16 -> global
        mbar_wait(dz_full, 0);
        tc_fence_after();
        const bool store_row = gi < row_end;
        if (dz32 == nullptr && SPLIT) {
            const float scale = gscale[0] * out_scale;
            float* out = reinterpret_cast<float*>(dz_v) + (size_t)gi * lddz + h * (D / 2);
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * (D / 2) + c * 32, r);
                tmem_ld_wait();
                if (store_row) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(out + c * 32 + e) =
                            make_float4(__uint_as_float(r[e]) * scale, __uint_as_float(r[e + 1]) * scale,
                                        __uint_as_float(r[e + 2]) * scale, __uint_as_float(r[e + 3]) * scale);
                }
            }
        } else if (dz32 == nullptr) {
            const float scale = gscale[0] * out_scale;
            uint16_t* out = dz + (size_t)gi * lddz + h * (D / 2);
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * (D / 2) + c * 32, r);
                tmem_ld_wait();
                if (store_row) {
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            pk[u] = pack2<F16>(__uint_as_float(r[e + 2 * u]) * scale, __uint_as_float(r[e + 2 * u + 1]) * scale);
                        *reinterpret_cast<uint4*>(out + c * 32 + e) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        } else {
            // column-split launch: every split owns an fp32 slab [rows_total][D]; infonce_tc_convert_kernel sums the slabs
            // in split order (no atomics: bitwise reproducible)
            float* out = dz32 + ((size_t)blockIdx.y * rows_total + (size_t)(gi - row_begin)) * D + h * (D / 2);
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dz + (uint32_t(q * 32) << 16) + h * (D / 2) + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 4)      // (rows past row_end land in the slab's own padding rows)
                    *reinterpret_cast<uint4*>(out + c * 32 + e) = make_uint4(r[e], r[e + 1], r[e + 2], r[e + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// fp32 split slabs -> bf16 gradient rows (scaled by gscale / (t N)); slabs summed in split order
template <int D, bool F16, bool OUT32 = false>
__global__ void infonce_tc_convert_kernel(const float* __restrict__ dz32, int nsplit, int64_t rows, int64_t rows_pad, int64_t row_begin,
                                          const float* __restrict__ gscale, float out_scale, uint16_t* __restrict__ dz,
                                          int64_t lddz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 8 consecutive columns
    if (idx >= rows * (D / 8)) return;
    const int64_t r = idx / (D / 8);
    const int c = (int)(idx % (D / 8)) * 8;
    const float scale = gscale[0] * out_scale;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int s = 0; s < nsplit; ++s) {
        const float* p = dz32 + ((size_t)s * rows_pad + r) * D + c;
        const float4 pa = *reinterpret_cast<const float4*>(p), pb = *reinterpret_cast<const float4*>(p + 4);
        a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
        b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
    }
    if constexpr (OUT32) {
        float* o = reinterpret_cast<float*>(dz) + (row_begin + r) * lddz + c;
        *reinterpret_cast<float4*>(o) = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
        *reinterpret_cast<float4*>(o + 4) = make_float4(b.x * scale, b.y * scale, b.z * scale, b.w * scale);
    } else {
        *reinterpret_cast<uint4*>(dz + (row_begin + r) * lddz + c) =
            make_uint4(pack2<F16>(a.x * scale, a.y * scale), pack2<F16>(a.z * scale, a.w * scale),
                       pack2<F16>(b.x * scale, b.y * scale), pack2<F16>(b.z * scale, b.w * scale));
    }
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

// [N, cols] bf16 / fp16, row pitch ldz elements; box = 64 columns (128 B, the swizzle span) x 64 rows; rows past N read as zero.
// cols = d, or 2 d for CY_F32_SPLIT rows ([hi | lo] bf16 halves)
static int make_tmap(CUtensorMap* m, const void* z, int64_t N, int64_t cols, int64_t ldz, int dtype) {
    EncodeTiledFn fn = tensor_map_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CY_ERR_DEVICE; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)N};
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

static inline bool is_split(int dtype) { return dtype == CY_F32_SPLIT; }
static inline int fwd_bn(int dtype) { return is_split(dtype) ? 64 : FWD_BN; }

bool infonce_tc_supported(int dtype, int64_t N, int64_t d, int64_t ldz, const uint8_t* codes, int variant) {
    if (!((dtype == CY_BF16 || dtype == CY_F16 || dtype == CY_F32_SPLIT) && (d == 256 || d == 128) && N >= 256 && (ldz % 8) == 0 &&
          codes == nullptr && N < (int64_t(1) << 30)))
        return false;
    if (is_split(dtype) && ldz < 2 * d) return false;
    if (variant != CY_SUPCON && N > (int64_t)P2_LISTCAP * fwd_bn(dtype)) return false;      // pass-2 tile list capacity
    return true;
}

static int sm_count() { return device_sm_count(); }

// column splits of the backward: minimise waves x tiles-per-CTA (one CTA per SM), at least 16 tiles per CTA.
// max_tps > 0 caps the tiles one CTA accumulates in tensor memory: the fp32 (split) path needs 1e-4 gradients, and the
// tensor core's fp32 accumulator loses ~2^-26 of the running sum per accumulation step on average (measured: 1.2e-4 of the
// gradient max-norm after 12288 steps at N = 65536 with 2 splits) — shorter chains, summed by the conversion kernel with
// round-to-nearest adds, keep it at ~1e-5.
static int bwd_splits(int64_t N, int64_t row_blocks, int64_t max_tps = 0) {
    const int64_t sms = sm_count(), rb = row_blocks, nt = (N + BWD_BN - 1) / BWD_BN;
    int best = 1;
    double best_cost = 1e30;
    const int smax = max_tps > 0 ? 64 : 16;
    for (int s = 1; s <= smax; ++s) {
        if (s > 1 && nt / s < 16) break;
        const int64_t tps = (nt + s - 1) / s;
        if (max_tps > 0 && tps > max_tps && nt / (s + 1) >= 16 && s < smax) continue;      // chain too long: split further
        if ((int64_t)(s - 1) * tps >= nt) continue;                  // the last split would be empty
        const int64_t waves = (rb * s + sms - 1) / sms;
        const double cost = (double)waves * (double)(tps + 6);      // + fixed per-CTA cost (A load, drain) in tile units
        if (cost < best_cost * 0.97) { best_cost = cost; best = s; }
    }
    return best;
}

// column splits of the forward: like the backward, minimise waves x (tiles per CTA + the fixed per-CTA cost — Zi into tensor
// memory, pipeline fill and drain, ~6 tiles), at least 8 column tiles per CTA, at most 16 splits (the workspace holds 16
// slabs).  Measured at N = 65536 (profiles/r2_fwd_splits_sweep.log): 512 row blocks 1.570 ms at 3 splits -> 1.527 at 2; one
// rank's strip of an 8-GPU run (64 row blocks) 0.220 ms at 16 splits -> 0.204 at 9.
static int fwd_splits(int64_t N, int64_t row_blocks, int bn) {
    const int64_t sms = sm_count(), rb = row_blocks, ctiles = (N + bn - 1) / bn;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= 16; ++s) {
        if (s > 1 && ctiles / s < 8) break;
        const int64_t tps = (ctiles + s - 1) / s;
        if ((int64_t)(s - 1) * tps >= ctiles) continue;              // the last split would be empty
        const int64_t waves = (rb * s + sms - 1) / sms;
        const double cost = (double)waves * (double)(tps + 6);
        if (cost < best_cost) { best_cost = cost; best = s; }
    }
    static int forced = -1;      // CY_FWD_SPLITS=n pins the forward's column splits (A/B sweeps)
    if (forced < 0) { const char* e = getenv("CY_FWD_SPLITS"); forced = e ? atoi(e) : 0; }
    if (forced > 0 && forced <= 16 && forced <= ctiles) best = forced;
    return best;
}

constexpr size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

constexpr int64_t SPLIT_MAX_TPS = 128;     // fp32 path: at most 128 column tiles (8192 columns) accumulated per CTA

size_t infonce_tc_workspace_bytes(int64_t N, int64_t d, bool split) {
    if (d != 256 && d != 128) return 0;
    // forward: worst case over row ranges is a single 128-row block -> the most column splits
    const int smax = 16;
    const size_t fwd = (size_t)smax * 2 * 4 * (size_t)N * sizeof(float);
    // pass 2: [tile ranges | 2 slots x 2 values x N]
    const size_t p2 = align256((size_t)((N + 63) / 64) * sizeof(int2)) + (size_t)4 * N * sizeof(float);
    // backward: fp32 slabs, splits x row blocks x 128 x d — worst case over the row-block count
    const int64_t rb_all = (N + TC_BM - 1) / TC_BM;
    size_t bwd = 0;
    for (int64_t rb = 1; rb <= rb_all; ++rb) {
        const int s = bwd_splits(N, rb, split ? SPLIT_MAX_TPS : 0);
        if (s > 1) {
            const size_t b = (size_t)s * rb * TC_BM * d * sizeof(float);
            if (b > bwd) bwd = b;
        }
    }
    size_t m = fwd > bwd ? fwd : bwd;
    if (p2 > m) m = p2;
    return m;
}

template <int D, int PASS, int VARIANT, bool SPLIT>
static int launch_fwd_tc(const CUtensorMap& tmap, const int32_t* labels, int N, int row_begin, int ct_begin, int ct_end, int tps,
                         float inv_t, float gamma, float* part, int slot_base, int fmt, const void* z, int64_t ldz,
                         const int2* tile_range, const float4* xstat, dim3 grid, cudaStream_t st) {
    using S = FwdCfg<D, SPLIT>;
    auto k = infonce_fwd_tc_kernel<D, PASS, VARIANT, SPLIT>;
    static SmemAttrCache attr;
    if (attr.need(S::TOTAL)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
        if (e != cudaSuccess) { set_error("fwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set(S::TOTAL);
    }
    k<<<grid, TC_THREADS, S::TOTAL, st>>>(tmap, labels, N, row_begin, ct_begin, ct_end, tps, inv_t * LOG2E, inv_t, gamma, part, slot_base,
                                          idesc_f16kind_f32(TC_BM, S::BN, 0, 0, fmt), reinterpret_cast<const uint16_t*>(z), ldz,
                                          tile_range, xstat);
    CY_CHECK_LAUNCH("infonce_fwd_tc");
    return CY_OK;
}

static int check_tc_rows(int64_t N, int64_t row_begin, int64_t row_end) {
    CY_CHECK_ARG((row_begin % TC_BM) == 0 && ((row_end % TC_BM) == 0 || row_end == N),
                 "tcgen05 path: row range must start on a multiple of 128 and end on one (or at N)");
    return CY_OK;
}

// pass 1: raw row sums of rows [row_begin, row_end) against all N columns, then the per-row statistics (xstat)
int infonce_fwd_tc(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                   float inv_t, int variant, float* stats, float* xstat, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    int rc = check_tc_rows(N, row_begin, row_end);
    if (rc) return rc;
    CY_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0, "tcgen05 path: z must be 16-byte aligned");
    const bool split = is_split(dtype);
    const int bn = fwd_bn(dtype);
    const int64_t rb = (rows + TC_BM - 1) / TC_BM;
    const int splits = fwd_splits(N, rb, bn);
    const int nslot = splits * 2;
    const size_t need = (size_t)nslot * 4 * (size_t)N * sizeof(float);
    CY_CHECK_ARG(workspace && workspace_bytes >= need, "infonce_fwd_tc: workspace %zu < %zu", workspace_bytes, need);
    CUtensorMap tmap;
    rc = make_tmap(&tmap, z, N, split ? 2 * d : d, ldz, dtype);
    if (rc) return rc;
    const int fmt = dtype == CY_F16 ? 0 : 1;
    const int ctiles = (int)((N + bn - 1) / bn);
    const int tps = (ctiles + splits - 1) / splits;
    dim3 grid((unsigned)rb, (unsigned)splits);
    float* part = reinterpret_cast<float*>(workspace);
#define CY_P1(DV, SP)                                                                                                              \
    if (d == DV && split == SP)                                                                                                    \
        rc = launch_fwd_tc<DV, 1, CY_SUPCON, SP>(tmap, labels, (int)N, (int)row_begin, 0, ctiles, tps, inv_t, 0.f, part, 0, fmt, z, ldz, \
                                                 nullptr, nullptr, grid, st);
    rc = CY_ERR_UNSUPPORTED;
    CY_P1(256, false) CY_P1(128, false) CY_P1(256, true) CY_P1(128, true)
#undef CY_P1
    if (rc) return rc;
    infonce_rowstats_kernel<1><<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(variant, (int)N, (int)row_begin, (int)row_end, inv_t, part,
                                                                              nslot, stats, reinterpret_cast<float4*>(xstat));
    CY_CHECK_LAUNCH("infonce_rowstats<1>");
    return CY_OK;
}

// pass 2 (exclude / self-paced): positive-pair sums that depend on the pass-1 row statistics
int infonce_fwd2_tc(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                    float inv_t, int variant, float gamma, float* stats, float* xstat, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    int rc = check_tc_rows(N, row_begin, row_end);
    if (rc) return rc;
    const bool split = is_split(dtype);
    const int bn = fwd_bn(dtype);
    const int ctiles = (int)((N + bn - 1) / bn);
    const size_t range_bytes = align256((size_t)ctiles * sizeof(int2));
    const size_t need = range_bytes + (size_t)4 * N * sizeof(float);
    CY_CHECK_ARG(workspace && workspace_bytes >= need, "infonce_fwd2_tc: workspace %zu < %zu", workspace_bytes, need);
    CY_CHECK_ARG(ctiles <= P2_LISTCAP, "infonce_fwd2_tc: N too large for the tile list");
    int2* tile_range = reinterpret_cast<int2*>(workspace);
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + range_bytes);
    CUtensorMap tmap;
    rc = make_tmap(&tmap, z, N, split ? 2 * d : d, ldz, dtype);
    if (rc) return rc;
    infonce_tile_range_kernel<<<ctiles, 32, 0, st>>>(labels, (int)N, bn, tile_range);
    CY_CHECK_LAUNCH("infonce_tile_range");
    const int fmt = dtype == CY_F16 ? 0 : 1;
    const int64_t rb = (rows + TC_BM - 1) / TC_BM;
    dim3 grid((unsigned)rb, 1);
    const float4* xs = reinterpret_cast<const float4*>(xstat);
    rc = CY_ERR_UNSUPPORTED;
#define CY_P2(DV, VAR, SP)                                                                                                        \
    if (d == DV && variant == VAR && split == SP)                                                                                 \
        rc = launch_fwd_tc<DV, 2, VAR, SP>(tmap, labels, (int)N, (int)row_begin, 0, ctiles, ctiles, inv_t, gamma, part, 0, fmt, z, ldz, \
                                           tile_range, xs, grid, st);
#define CY_P2V(DV, SP) CY_P2(DV, CY_SUPCON_EXCLUDE, SP) CY_P2(DV, CY_SELFPACED_HARD, SP) CY_P2(DV, CY_SELFPACED_SOFT, SP)
    CY_P2V(256, false) CY_P2V(128, false) CY_P2V(256, true) CY_P2V(128, true)
#undef CY_P2V
#undef CY_P2
    if (rc) {
        if (rc == CY_ERR_UNSUPPORTED) set_error("infonce_fwd2_tc: no instantiation for d=%lld variant=%d", (long long)d, variant);
        return rc;
    }
    infonce_rowstats_kernel<2><<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(variant, (int)N, (int)row_begin, (int)row_end, inv_t, part, 2,
                                                                              stats, reinterpret_cast<float4*>(xstat));
    CY_CHECK_LAUNCH("infonce_rowstats<2>");
    return CY_OK;
}

// row statistics from the SoA sums a SIMT sweep left in `stats` (same kernel, part == nullptr)
int infonce_rowstats(int64_t N, int64_t row_begin, int64_t row_end, float inv_t, int variant, int pass, float* stats, float* xstat,
                     cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    if (pass == 1)
        infonce_rowstats_kernel<1><<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(variant, (int)N, (int)row_begin, (int)row_end, inv_t,
                                                                                  nullptr, 0, stats, reinterpret_cast<float4*>(xstat));
    else
        infonce_rowstats_kernel<2><<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(variant, (int)N, (int)row_begin, (int)row_end, inv_t,
                                                                                  nullptr, 0, stats, reinterpret_cast<float4*>(xstat));
    CY_CHECK_LAUNCH("infonce_rowstats");
    return CY_OK;
}

size_t infonce_loss_workspace_bytes(int64_t N) { return (size_t)((N + LOSS_THREADS - 1) / LOSS_THREADS) * 4 * sizeof(float) + 256; }

int infonce_loss(int64_t N, int variant, const float* xstat, float* out8, int32_t* bad_rows, const int32_t* overflow,
                 void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int nblk = (int)((N + LOSS_THREADS - 1) / LOSS_THREADS);
    CY_CHECK_ARG(workspace && workspace_bytes >= (size_t)nblk * 4 * sizeof(float), "infonce_loss: workspace %zu too small", workspace_bytes);
    float* partials = reinterpret_cast<float*>(workspace);
    infonce_loss_partial_kernel<<<nblk, LOSS_THREADS, 0, st>>>(variant, (int)N, reinterpret_cast<const float4*>(xstat), partials);
    CY_CHECK_LAUNCH("infonce_loss_partial");
    infonce_loss_final_kernel<<<1, 32, 0, st>>>(nblk, (int)N, partials, out8, bad_rows, overflow);
    CY_CHECK_LAUNCH("infonce_loss_final");
    return CY_OK;
}

template <int D, bool F16, int VARIANT, bool SPLIT>
static int launch_bwd_tc(const CUtensorMap& tmap, const int32_t* labels, const float4* xstat, int N, int row_begin, int row_end, int tps,
                         float inv_t, float gamma, const float* gscale, float out_scale, void* dz, int64_t lddz, float* dz32,
                         int rows_total, const void* z, int64_t ldz, dim3 grid, cudaStream_t st) {
    using C = BwdCfg<D, SPLIT>;
    auto k = infonce_bwd_tc_kernel<D, F16, VARIANT, SPLIT>;
    static SmemAttrCache attr;
    if (attr.need(C::TOTAL)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::TOTAL);
        if (e != cudaSuccess) { set_error("bwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set(C::TOTAL);
    }
    const int fmt = F16 ? 0 : 1;
    k<<<grid, TC_THREADS, C::TOTAL, st>>>(tmap, labels, xstat, N, row_begin, row_end, tps, inv_t * LOG2E, inv_t, gamma, gscale, out_scale, dz,
                                          lddz, dz32, rows_total, idesc_f16kind_f32(TC_BM, BWD_BN, 0, 0, fmt),
                                          idesc_f16kind_f32(TC_BM, D, 0, 1, fmt), reinterpret_cast<const uint16_t*>(z), ldz);
    CY_CHECK_LAUNCH("infonce_bwd_tc");
    return CY_OK;
}

int infonce_bwd_tc(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, int64_t row_begin, int64_t row_end,
                   float inv_t, int variant, float gamma, const float* xstat, const float* gscale, void* dz, int64_t lddz, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return CY_OK;
    int rc = check_tc_rows(N, row_begin, row_end);
    if (rc) return rc;
    const bool split = is_split(dtype);
    CY_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (lddz % (split ? 4 : 8)) == 0,
                 "tcgen05 path: z / dz must be 16-byte aligned");
    CUtensorMap tmap;
    rc = make_tmap(&tmap, z, N, split ? 2 * d : d, ldz, dtype);
    if (rc) return rc;
    const bool f16 = dtype == CY_F16;
    const float out_scale = (inv_t / (float)N) * (f16 ? (1.f / 1024.f) : 1.f);
    const int64_t rb = (rows + TC_BM - 1) / TC_BM;
    const int splits = bwd_splits(N, rb, split ? SPLIT_MAX_TPS : 0);
    const int nt = (int)((N + BWD_BN - 1) / BWD_BN);
    const int tps = (nt + splits - 1) / splits;
    float* dz32 = nullptr;
    const int64_t rows_pad = rb * TC_BM;
    if (splits > 1) {
        const size_t need = (size_t)splits * rows_pad * d * sizeof(float);
        CY_CHECK_ARG(workspace && workspace_bytes >= need, "infonce_bwd_tc: workspace %zu < %zu", workspace_bytes, need);
        dz32 = reinterpret_cast<float*>(workspace);
    }
    dim3 grid((unsigned)rb, (unsigned)splits);
    const float4* xs = reinterpret_cast<const float4*>(xstat);
    rc = CY_ERR_UNSUPPORTED;
#define CY_BW(DV, F, VAR, SP)                                                                                                    \
    if (d == DV && f16 == F && variant == VAR && split == SP)                                                                    \
        rc = launch_bwd_tc<DV, F, VAR, SP>(tmap, labels, xs, (int)N, (int)row_begin, (int)row_end, tps, inv_t, gamma, gscale, out_scale, dz, \
                                           lddz, dz32, (int)rows_pad, z, ldz, grid, st);
#define CY_BW4(DV, F, SP) CY_BW(DV, F, CY_SUPCON, SP) CY_BW(DV, F, CY_SUPCON_EXCLUDE, SP) CY_BW(DV, F, CY_SELFPACED_HARD, SP) \
    CY_BW(DV, F, CY_SELFPACED_SOFT, SP)
    CY_BW4(256, false, false) CY_BW4(256, true, false) CY_BW4(128, false, false) CY_BW4(128, true, false)
    CY_BW4(256, false, true) CY_BW4(128, false, true)
#undef CY_BW4
#undef CY_BW
    if (rc) {
        if (rc == CY_ERR_UNSUPPORTED) set_error("infonce_bwd_tc: no instantiation for d=%lld variant=%d", (long long)d, variant);
        return rc;
    }
    if (dz32) {
        const int64_t n8 = rows * (d / 8);
        const unsigned g = (unsigned)((n8 + 255) / 256);
        uint16_t* out = reinterpret_cast<uint16_t*>(dz);
        if (split && d == 256) infonce_tc_convert_kernel<256, false, true><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        else if (split) infonce_tc_convert_kernel<128, false, true><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        else if (d == 256 && f16) infonce_tc_convert_kernel<256, true><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        else if (d == 256) infonce_tc_convert_kernel<256, false><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        else if (f16) infonce_tc_convert_kernel<128, true><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        else infonce_tc_convert_kernel<128, false><<<g, 256, 0, st>>>(dz32, splits, rows, rows_pad, row_begin, gscale, out_scale, out, lddz);
        CY_CHECK_LAUNCH("infonce_tc_convert");
    }
    return CY_OK;
}

}  // namespace cy
