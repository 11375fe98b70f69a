// IIC adjoint (padding = 1, K <= 16, fp32 maps) on the tcgen05 tensor cores — the default adjoint of cy_iic_bwd.
//
//   dL/dy[k2, h, w] = sum_{k1, dy, dx} G[k1, k2, dy, dx] * x[k1, h + dy - 1, w + dx - 1]      (and the mirrored sum for dL/dx)
//
// is a 3x3 convolution with K -> K channels: out[o, p] = sum_k A[p, k] * Wt[o, k], k = (dy, dx, c).  Pixels are the M
// dimension of the MMA (one TMEM lane per pixel), the K output channels its N dimension (N = 16) and the 9 K taps its
// reduction dimension.  The legacy mma.sync adjoint (iic_mma.cu) is bound by the warp schedulers (an HMMA holds the issue
// port for its 8 pipe cycles on sm_100a, so tensor and CUDA-core instructions add up); tcgen05.mma is asynchronous and
// issued by one thread, which leaves the schedulers to the conversion work.  One persistent CTA per SM, three warp roles:
//
//   converters (4 sets x 4 warps, input rows round-robin)  thread = input pixel = TMEM lane.  An input row of a lane quarter
//                         ([K channels][36 floats], image columns w0-4 .. w0+31) is staged with 16-byte cp.async (zero fill
//                         outside the image; three rows in flight per warp, tracked by cp.async groups), the thread reads its K
//                         channel values, splits them into bf16 hi / lo parts (one F2FP + shift / mask / FADD / PRMT per channel
//                         pair), takes the copies shifted by one and two columns from the neighbour lanes by shuffle and writes ONE
//                         A slot [lane = pixel][k = (dx, c)] x {hi, lo} to tensor memory (tcgen05.st).  A warp owns 28 output
//                         pixels + the halo of its shifts, so shuffles never cross a warp (224 = 8 x 28).  The store drain and the
//                         slot-free probe are software-pipelined off the warp's critical path.
//   issuers    (4 warps, output rows round-robin)  output row h takes its three dy taps from the A slots of input rows h-1, h,
//                         h+1: per output row 3 x KS x 3 MMAs (M 128, N 16, K 16; A from TMEM, B = weights from shared memory),
//                         hi*hi + lo*hi + hi*lo, accumulating into one 16-column D slot.  An A slot is written once and read by
//                         three output rows; its "empty" barrier counts their three tcgen05.commit arrivals.  The four barrier
//                         probes of a row are issued back to back so that their latencies overlap.
//   epilogue   (2 sets x 4 warps)  thread = output pixel: tcgen05.ld of the K live columns of a finished D slot, K coalesced
//                         112-byte stores.  No state is carried between rows; waits back off (nanosleep).
//
// Warp order = scheduler priority (highest warp id first): issuers, converters, epilogue.
// Work split: the 2 (sides) x B x ceil(W / 112) column strips of H rows form one linear row space that is cut into equal
// contiguous ranges, one per CTA; a range that starts or ends inside a strip converts two extra halo rows.
//
// Measured on B200 (profiles/probes/probe_umma_issue2, profiles/README.md): N = 16 MMAs execute in 9 clk, one thread issues one
// every ~20 clk, a 32-column tcgen05.st + wait costs ~58 clk per warp, a 16-column tcgen05.ld ~39 clk.  Config 3: 61 us against
// 90 us for the mma.sync adjoint (HBM floor 39 us).  What bounds it now is the SM's instruction issue rate: ~1150 instructions
// per 128-pixel row (conversion 4 x 71, control / addressing 4 x 80, epilogue 4 x 70, barrier polling ~200) at 65 % issue
// utilisation; the tensor pipe is 28 % busy.  -DCY_TC_TIMING builds per-phase cycle counters (never in the product build).
//
// Precision: v = hi + lo (hi = bf16 rounded, lo = bf16 of the exact remainder); products hi*hi + lo*hi + hi*lo are good to
// 2^-16 relative, fp32 accumulation in TMEM: gradients agree with the float64 oracle to ~1e-5 of their max-norm.
// Reference: the autograd adjoint of compute_joint_2D's F.conv2d (contrastyou/losses/discreteMI.py:225-243).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cy {

using namespace tc;

namespace {

constexpr int T_QPX = 28;                        // output pixels per lane quarter (lanes 28..31: halo of the column shifts)
constexpr int T_TWO = 4 * T_QPX;                 // output columns per strip
constexpr int T_NA = 8;                          // A slots (input rows) in tensor memory
constexpr int T_ND = 8;                          // D slots (output rows)
// the role split and the poll back-off times are -D overridable for A/B builds (profiles/probes/iic_variants.sh)
#ifndef CY_TC_NISS
#define CY_TC_NISS 4
#endif
#ifndef CY_TC_NCS
#define CY_TC_NCS 4
#endif
#ifndef CY_TC_NES
#define CY_TC_NES 2
#endif
#ifndef CY_TC_PF
#define CY_TC_PF 3
#endif
#ifndef CY_TC_SLEEP_ISS
#define CY_TC_SLEEP_ISS 64
#endif
#ifndef CY_TC_SLEEP_EPI
#define CY_TC_SLEEP_EPI 256
#endif
#ifndef CY_TC_SLEEP_CONV
#define CY_TC_SLEEP_CONV 0
#endif
constexpr int T_NISS = CY_TC_NISS;               // issuer warps (output rows round-robin)
constexpr int T_NCS = CY_TC_NCS;                 // converter sets (input rows round-robin)
constexpr int T_NES = CY_TC_NES;                 // epilogue sets (output rows round-robin)
// Warp roles.  The scheduler of an SM sub-partition picks the eligible warp with the HIGHEST warp id first (guide: "arbiter
// priority: hi-wid-first"), so the roles that must never be starved sit at the top: the MMA issuers, whose progress frees
// the A slots every converter waits for, then the converters, then the epilogue.
constexpr int T_EPI0 = 0;                        // warps [0, 4 * T_NES): epilogue
constexpr int T_CONV0 = 4 * T_NES;               // then 4 * T_NCS converter warps (warp & 3 = TMEM lane quarter in both roles)
constexpr int T_ISS0 = T_CONV0 + 4 * T_NCS;      // then the issuers
constexpr int T_WALLOC = T_ISS0;                 // the first issuer warp also owns the TMEM allocation
constexpr int T_THREADS = 32 * (T_ISS0 + T_NISS);
constexpr int T_PF = CY_TC_PF;                   // converter prefetch depth (own rows): cp.async groups in flight per warp
constexpr int T_WTILE = 512;                     // one weight tile: [n 16][k 16] bf16, no-swizzle K-major core matrices

constexpr int T_MAXHEADS = 8;                    // sub-head pairs of one launch (cy_iic_bwd_heads)
struct TcGeom {
    int B, K, H, W;
    int S;                                       // sub-head pairs: S independent (x, y, dL/dJ) problems of one shape
    int TW2;                                     // strips per image row
    int rows_total;                              // S * 2 * B * TW2 * H
    long long dj_stride;                         // floats between the heads' dL/dJ tables
    float inv_T;                                 // != 0: the maps are softmax(logits / T) and the outputs are dL/dlogits (fused softmax backward)
};
struct TcPtrs {                                  // per head: the two maps and the two gradient outputs
    const float* x[T_MAXHEADS];
    const float* y[T_MAXHEADS];
    float* dx[T_MAXHEADS];
    float* dy[T_MAXHEADS];
};

// The row-space range [R0, R1) of one CTA is walked as segments (a segment never crosses a strip).  Segment k starts at R0
// (k = 0) or at the k-th strip boundary after it; it produces n_out output rows from n_out + 2 input rows.
struct Seg {
    int unit, hb, n_out;
};
__device__ __forceinline__ bool seg_at(int R0, int R1, int H, int k, Seg& s) {
    const int start = k == 0 ? R0 : (R0 / H + k) * H;
    if (start >= R1) return false;
    s.unit = start / H;
    const int ue = (s.unit + 1) * H, end = ue < R1 ? ue : R1;
    s.hb = start - s.unit * H;
    s.n_out = end - start;
    return true;
}

// unit -> (side, image, strip).  Images run backwards and the two sides alternate: the joint kernel has just streamed x and y
// front to back, so the tail of both tensors is what the 126 MB L2 still holds.
__device__ __forceinline__ void unit_decode(const TcGeom& g, int unit, int& head, int& side, int& b, int& tw) {
    tw = unit % g.TW2;
    side = (unit / g.TW2) & 1;
    const int rest = unit / (2 * g.TW2);
    b = g.B - 1 - rest % g.B;
    head = rest / g.B;
}

// (v0, v1) -> packed bf16 pairs (v0 in the low half), v = hi + lo: hi = bf16 round-to-nearest (one F2FP per pair: 5 per warp
// and row, nothing for the XU pipe at this rate — unlike in the mma.sync kernels, where every fragment element was split
// three times), lo = truncated bf16 of the exact remainder
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(hi << 16);
    const float r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    lo = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// mbarrier wait for a role that is expected to wait (the epilogue is faster than its producers): a failed probe backs off
// for ~100 ns instead of re-issuing at once — in the first profile of this kernel the spin loops were 30 % of all issued
// instructions of an issue-bound SM.  The D ring (8 slots) absorbs the added wake-up latency.
template <int NS>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
        if constexpr (NS > 0) __nanosleep(NS);
    }
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const float* src, uint32_t bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ float lds_f32_own(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

template <int NW>
__device__ __forceinline__ void tmem_st_words(uint32_t taddr, const uint32_t* w) {
    if constexpr (NW == 8) {
        tmem_st_32x8(taddr, *reinterpret_cast<const uint32_t (*)[8]>(w));
    } else if constexpr (NW == 16) {
        tmem_st_32x16(taddr, *reinterpret_cast<const uint32_t (*)[16]>(w));
    } else {
        static_assert(NW == 24, "A slot plane: 8, 16 or 24 words");
        tmem_st_32x16(taddr, *reinterpret_cast<const uint32_t (*)[16]>(w));
        tmem_st_32x8(taddr + 16, *reinterpret_cast<const uint32_t (*)[8]>(w + 16));
    }
}

// -DCY_TC_TIMING (never in the product build): per-phase cycle counters of CTA 0, printed by the host after the launch
#ifdef CY_TC_TIMING
#define TC_T(slot, stmt) do { const long long t0__ = clock64(); stmt; tacc[slot] += clock64() - t0__; } while (0)
#define TC_TDECL long long tacc[6] = {0, 0, 0, 0, 0, 0}
#define TC_TDUMP(base, n) do { if (blockIdx.x == 0 && lane == 0) for (int q_ = 0; q_ < (n); ++q_) tc_dbg[(base) + q_] = tacc[q_]; } while (0)
__device__ long long tc_dbg[64];
#else
#define TC_T(slot, stmt) do { stmt; } while (0)
#define TC_TDECL
#define TC_TDUMP(base, n)
#endif

template <int KH>                                // channel pairs: K == 2 * KH or 2 * KH - 1 (the host picks KH = ceil(K / 2))
__global__ void __launch_bounds__(T_THREADS, 1)
iic_bwd_tc_kernel(const __grid_constant__ TcPtrs ptrs, const __grid_constant__ TcGeom g, const float* __restrict__ djoint,
                  const float* __restrict__ gscale) {
    constexpr int KC = 2 * KH;                   // channels carried per pixel (zero past K)
    constexpr int KS = (3 * KH + 7) / 8;         // K-steps of 16 per input row: k = dxx * KC + c
    constexpr int PW = KS * 8;                   // 32-bit words per plane (hi or lo) of an A slot
    constexpr int ACOLS = 2 * PW;
    constexpr uint32_t A0 = T_ND * 16;           // TMEM columns: [0, 128) D slots, [128, 128 + 8 * ACOLS) A slots
    static_assert(A0 + T_NA * ACOLS <= 512, "TMEM budget");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;                         // weights [head S][side 2][dy 3][ks KS][part 2] tiles of T_WTILE bytes
    constexpr int HEAD_WBYTES = 2 * 3 * KS * 2 * T_WTILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)g.S * HEAD_WBYTES);
    uint64_t* a_full = bars;                     // [T_NA]  the four quarter warps have written the slot
    uint64_t* a_empty = a_full + T_NA;           // [T_NA]  the MMAs of the three output rows that read the slot are complete
    uint64_t* d_full = a_empty + T_NA;           // [T_ND]  the MMAs of the output row are complete
    uint64_t* d_empty = d_full + T_ND;           // [T_ND]  the four epilogue warps have read the slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + T_ND);
    float* ring0 = reinterpret_cast<float*>(tmem_slot + 4);   // converter staging: [warp 4 * T_NCS][stage T_PF][channel KC][lane 32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = g.K;
#ifdef CY_TC_TIMING
    const long long t_entry = clock64();
#endif

    // this CTA's range of the row space
    const int R0 = (int)(((long long)g.rows_total * blockIdx.x) / gridDim.x);
    const int R1 = (int)(((long long)g.rows_total * (blockIdx.x + 1)) / gridDim.x);
    const int plane = g.H * g.W;                 // host checks K * H * W < 2^31

    // ---- one-time setup: weights in UMMA layout (hi / lo), barriers, tensor memory.  Every thread runs it exactly once (it ends
    // in the CTA-wide barrier); the converter warps call it AFTER they have issued their first staged rows, so the cold-HBM
    // latency of those loads overlaps the ~2 us of set-up instead of following it.
    auto setup = [&]() {
        const float scale = gscale[0] * (g.inv_T != 0.f ? g.inv_T : 1.f);      // (1/T of the fused softmax backward rides on the weights)
        for (int i = threadIdx.x; i < g.S * 2 * 3 * KS * 16 * 16; i += T_THREADS) {
            const int k = i % 16, n = (i / 16) % 16, ks = (i / 256) % KS, dyy = (i / (256 * KS)) % 3, side = (i / (256 * KS * 3)) & 1;
            const int head = i / (256 * KS * 3 * 2);
            const float* dj = djoint + (size_t)head * g.dj_stride;
            const int kk = ks * 16 + k, dxx = kk / KC, c = kk % KC, o = n;
            float w = 0.f;
            if (dxx < 3 && c < K && o < K) {
                // side 0 (dL/dy from x): Wt[o = k2][c = k1][dy][dx] = G[k1, k2, dy, dx]
                // side 1 (dL/dx from y): Wt[o = k1][c = k2][dy][dx] = G[k1, k2, 2 - dy, 2 - dx]
                w = side == 0 ? dj[((c * K + o) * 3 + dyy) * 3 + dxx] : dj[((o * K + c) * 3 + (2 - dyy)) * 3 + (2 - dxx)];
                w *= scale;
            }
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            const uint32_t tile = (uint32_t)((((head * 2 + side) * 3 + dyy) * KS + ks) * 2) * T_WTILE;
            const uint32_t off = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(wsm + tile + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(wsm + tile + T_WTILE + off) = lo;
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < T_NA; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 3); }
            for (int i = 0; i < T_ND; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
            fence_barrier_init();
        }
        if (warp == T_WALLOC) {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    };
    const bool is_converter = warp >= T_CONV0 && warp < T_ISS0;
    if (!is_converter) setup();
#ifdef CY_TC_TIMING
    const long long t_start = clock64();
#endif

    if (warp >= T_ISS0 && warp < T_ISS0 + T_NISS) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        const uint32_t tmem = *tmem_slot;
        if (elect_one()) {
            TC_TDECL;
#ifdef CY_TC_TIMING
            const long long tl0 = clock64();
#endif
            constexpr uint32_t idesc = idesc_bf16_f32(128, 16, 0, 0);
            const uint64_t wdesc0 = smem_desc_noswz(smem_u32(wsm), 128, 256);
            uint32_t ar0 = 0, orow = 0;
            Seg sg;
            for (int k = 0; seg_at(R0, R1, g.H, k, sg); ++k) {
                int head, side, b, tw;
                unit_decode(g, sg.unit, head, side, b, tw);
                const int n_out = sg.n_out;
                const uint64_t wside = wdesc0 + (uint64_t)(((head * 2 + side) * 3 * KS * 2 * T_WTILE) >> 4);
                for (int i = 0; i < n_out; ++i, ++orow) {
                    if ((int)(orow % T_NISS) != warp - T_ISS0) continue;
                    const uint32_t a = ar0 + (uint32_t)i;
                    // probe all four barriers first (the ~100-cycle TRYWAIT latencies overlap), then block on the ones that failed
                    uint64_t* const b0 = a_full + (a % T_NA);
                    uint64_t* const b1 = a_full + ((a + 1) % T_NA);
                    uint64_t* const b2 = a_full + ((a + 2) % T_NA);
                    uint64_t* const bd = d_empty + (orow % T_ND);
                    const uint32_t p0 = (a / T_NA) & 1u, p1 = ((a + 1) / T_NA) & 1u, p2 = ((a + 2) / T_NA) & 1u;
                    const uint32_t pd = ((orow / T_ND) & 1u) ^ 1u;
#ifdef CY_TC_TIMING
                    const long long tw0 = clock64();
#endif
                    const bool r0 = mbar_try_wait(b0, p0), r1 = mbar_try_wait(b1, p1), r2 = mbar_try_wait(b2, p2), rd = mbar_try_wait(bd, pd);
                    if (!r0) mbar_wait_backoff<CY_TC_SLEEP_ISS>(b0, p0);
                    if (!r1) mbar_wait_backoff<CY_TC_SLEEP_ISS>(b1, p1);
                    if (!r2) mbar_wait_backoff<CY_TC_SLEEP_ISS>(b2, p2);
                    if (!rd) mbar_wait_backoff<CY_TC_SLEEP_ISS>(bd, pd);
#ifdef CY_TC_TIMING
                    tacc[0] += clock64() - tw0;
                    const long long ti0 = clock64();
#endif
                    tc_fence_after();
                    const uint32_t d = tmem + (orow % T_ND) * 16;
#pragma unroll
                    for (int dyy = 0; dyy < 3; ++dyy) {
                        const uint32_t ah = tmem + A0 + ((a + dyy) % T_NA) * ACOLS, al = ah + PW;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t bh = wside + (uint64_t)(((dyy * KS + ks) * 2 * T_WTILE) >> 4);
                            const uint64_t bl = bh + (uint64_t)(T_WTILE >> 4);
                            if (dyy == 0 && ks == 0) umma_bf16_ts_c<0>(d, ah + ks * 8, bh, idesc);
                            else umma_bf16_ts_c<1>(d, ah + ks * 8, bh, idesc);
                            umma_bf16_ts_c<1>(d, al + ks * 8, bh, idesc);
                            umma_bf16_ts_c<1>(d, ah + ks * 8, bl, idesc);
                        }
                    }
                    umma_commit(d_full + (orow % T_ND));
                    // every A slot expects three arrivals; at the ends of a segment a slot has fewer readers and the first /
                    // last output row stands in for the missing ones
#pragma unroll
                    for (int dyy = 0; dyy < 3; ++dyy) {
                        const int cnt = 1 + (i == 0 ? 2 - dyy : 0) + (i == n_out - 1 ? dyy : 0);
                        for (int c = 0; c < cnt; ++c) umma_commit(a_empty + ((a + dyy) % T_NA));
                    }
#ifdef CY_TC_TIMING
                    tacc[3] += clock64() - ti0;
#endif
                }
                ar0 += (uint32_t)n_out + 2u;
            }
#ifdef CY_TC_TIMING
            tacc[4] = clock64() - tl0;
            if (blockIdx.x == 0) for (int q_ = 0; q_ < 6; ++q_) tc_dbg[(warp - T_ISS0) * 8 + q_] = tacc[q_];
#endif
        }
        __syncwarp();
    } else if (warp >= T_CONV0 && warp < T_ISS0) {
        // ------------------------------------------------------------------------------------------ converters
        const int cset = (warp - T_CONV0) >> 2, quarter = warp & 3;
        // A cursor over this warp's own input rows (global input-row index == cset mod T_NCS).  Everything that needs a
        // division is done once per segment; a step is a pointer increment.
        struct Cursor {
            const float* p;      // &in[b, 0, h, w0q]: first output column of this quarter, channel 0, image row h
            int h, j, n_in, k;   // image row, index inside the segment, input rows of the segment, segment number
            uint32_t ar0;        // global index of the segment's first input row
            uint32_t cmask;      // bit i: this lane's i-th staging chunk lies inside the image row (and below channel K)
            bool done;
        };
        // Staging: one input row of a quarter is [KC channels][CW = 36 floats] = image columns w0q-4 .. w0q+31, copied as
        // 16-byte chunks (9 per channel) by cp.async with zero fill outside the image; W % 4 == 0 makes every chunk lie
        // entirely inside or outside a row.  Lane l copies chunks l, l + 32, l + 64; lane l reads float 3 + l of every channel.
        constexpr int CW = 36, NCHUNK = KC * (CW / 4), NIT = (NCHUNK + 31) / 32, STAGE_BYTES = KC * CW * 4;
        int coff[NIT], ccol[NIT];
        bool clive[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int idx = lane + 32 * it, ch = idx / (CW / 4), jc = idx % (CW / 4);
            clive[it] = idx < NCHUNK && ch < K;
            ccol[it] = 4 * jc - 4;
            coff[it] = ch * plane + ccol[it];
        }
        auto enter = [&](Cursor& c) {                                  // position on the first own row of segment c.k or later
            for (;;) {
                Seg sg;
                c.done = !seg_at(R0, R1, g.H, c.k, sg);
                if (c.done) return;
                c.n_in = sg.n_out + 2;
                c.j = (int)((cset + T_NCS - c.ar0 % T_NCS) % T_NCS);
                if (c.j >= c.n_in) {                                   // a one-row segment (3 input rows) has no row for one of the
                    c.ar0 += (uint32_t)c.n_in;                         // T_NCS = 4 sets: go on to the next segment (writing a row
                    ++c.k;                                             // of its own here would arrive twice on that A slot)
                    continue;
                }
                int head, side, b, tw;
                unit_decode(g, sg.unit, head, side, b, tw);
                const int w0q = tw * T_TWO + quarter * T_QPX;
                c.cmask = 0u;
#pragma unroll
                for (int it = 0; it < NIT; ++it)
                    if (clive[it] && w0q + ccol[it] >= 0 && w0q + ccol[it] < g.W) c.cmask |= 1u << it;
                c.h = sg.hb - 1 + c.j;
                c.p = (side ? ptrs.y : ptrs.x)[head] + (size_t)b * K * plane + w0q + (long long)c.h * g.W;
                return;
            }
        };
        auto step = [&](Cursor& c) {
            c.j += T_NCS;
            if (c.j < c.n_in) {
                c.h += T_NCS;
                c.p += T_NCS * g.W;
            } else {
                c.ar0 += (uint32_t)c.n_in;
                ++c.k;
                enter(c);
            }
        };
        const uint32_t ring = smem_u32(ring0) + (uint32_t)((warp - T_CONV0) * T_PF * STAGE_BYTES);
        const uint32_t ring_wr = ring + (uint32_t)lane * 16u, ring_rd = ring + (uint32_t)(3 + lane) * 4u;
        auto issue_row = [&](const Cursor& c, int stage) {
            const uint32_t m = (!c.done && (unsigned)c.h < (unsigned)g.H) ? c.cmask : 0u;
#pragma unroll
            for (int it = 0; it < NIT; ++it)        // size 0: nothing is read (the address may lie outside the tensor), zeros are written
                if (lane + 32 * it < NCHUNK)
                    cp_async_16(ring_wr + (uint32_t)(stage * STAGE_BYTES + it * 512), c.p + coff[it], (m >> it) & 1u ? 16u : 0u);
            cp_async_commit();
        };
        TC_TDECL;
        Cursor cur, pf;
        cur.k = 0; cur.ar0 = 0; cur.cmask = 0u; cur.p = ptrs.x[0]; cur.h = 0; cur.j = 0; cur.n_in = 0;
        enter(cur);
        pf = cur;
#pragma unroll
        for (int u = 0; u < T_PF; ++u) {
            issue_row(pf, u);
            if (!pf.done) step(pf);
        }
        setup();
        const uint32_t tmem = *tmem_slot;
        // Software pipeline: the tcgen05.st of row n are published (wait::st, fence, arrive) only after the shared-memory reads
        // of row n+1 have been issued, and the a_empty probe of a row is issued before its conversion work, so neither the
        // ~60-cycle store drain nor the ~100-cycle barrier probe sits on the warp's critical path.
        int pending = -1;                                             // A slot written but not yet published
        auto publish = [&]() {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + pending);
        };
        while (!cur.done) {
#pragma unroll
            for (int u = 0; u < T_PF; ++u) {
                if (cur.done) break;
#ifdef CY_TC_TIMING
                const long long tc0 = clock64();
#endif
                const uint32_t ar = cur.ar0 + (uint32_t)cur.j, slot = ar % T_NA, par = ((ar / T_NA) & 1u) ^ 1u;
                const bool free_now = mbar_try_wait(a_empty + slot, par);
                TC_T(0, cp_async_wait<T_PF - 1>());                   // the oldest group (this stage) has landed ...
                __syncwarp();                                         // ... for every lane of the warp
                float v[KC];
#pragma unroll
                for (int c = 0; c < KC; ++c) v[c] = lds_f32_own(ring_rd + (uint32_t)(u * STAGE_BYTES + c * CW * 4));
                if (pending >= 0) publish();
                uint32_t th[PW], tl[PW];
#pragma unroll
                for (int i = 3 * KH; i < PW; ++i) { th[i] = 0u; tl[i] = 0u; }
                // own pixel: KH hi words + KH lo words; the copies shifted by one and two columns come from lanes l+1, l+2 by
                // shuffle (a shared-memory exchange needs fewer instructions on paper, but its 16-byte loads do not land in the
                // consecutive registers tcgen05.st wants: +40 register moves per row and 30 % more LSU wavefronts, measured)
#pragma unroll
                for (int i = 0; i < KH; ++i) {
                    uint32_t hw, lw;
                    split2(v[2 * i], v[2 * i + 1], hw, lw);
                    th[i] = hw;
                    tl[i] = lw;
                    th[KH + i] = __shfl_down_sync(0xffffffffu, hw, 1);
                    tl[KH + i] = __shfl_down_sync(0xffffffffu, lw, 1);
                    th[2 * KH + i] = __shfl_down_sync(0xffffffffu, hw, 2);
                    tl[2 * KH + i] = __shfl_down_sync(0xffffffffu, lw, 2);
                }
                issue_row(pf, u);                                     // refill this stage (every lane is past its reads: the
                if (!pf.done) step(pf);                               // shuffles above are warp-synchronous)
#ifdef CY_TC_TIMING
                tacc[1] += clock64() - tc0;                           // wait for the data + split + shuffles + refill
#endif
                if (!free_now) TC_T(2, mbar_wait_backoff<CY_TC_SLEEP_CONV>(a_empty + slot, par));
#ifdef CY_TC_TIMING
                const long long tc1 = clock64();
#endif
                tc_fence_after();
                const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + A0 + slot * ACOLS;
                tmem_st_words<PW>(ta, th);
                tmem_st_words<PW>(ta + PW, tl);
                pending = (int)slot;
                step(cur);
#ifdef CY_TC_TIMING
                tacc[3] += clock64() - tc1;                           // store + step
                tacc[4] += 1;
#endif
            }
        }
        if (pending >= 0) publish();
        cp_async_wait<0>();
        if (warp == T_CONV0 || warp == T_CONV0 + 5) TC_TDUMP(32 + (warp == T_CONV0 ? 0 : 8), 5);   // w+5: set 1, quarter 1
    } else if (warp < T_CONV0) {
        // ------------------------------------------------------------------------------------------ epilogue
        const int eset = (warp - T_EPI0) >> 2, quarter = warp & 3;
        const uint32_t tmem = *tmem_slot;
        // staging of the fused softmax backward: [warp][3 stages][KC channels][32 floats], behind the converters' ring (only
        // allocated — and only touched — when g.inv_T != 0)
        constexpr int E_STAGE_BYTES = KC * 32 * 4;
        const uint32_t ering = smem_u32(ring0) + (uint32_t)(4 * T_NCS * T_PF * KC * 36 * 4) + (uint32_t)((warp - T_EPI0) * 3 * E_STAGE_BYTES);
        uint32_t orow = 0;
        Seg sg;
        TC_TDECL;
        for (int k = 0; seg_at(R0, R1, g.H, k, sg); ++k) {
            int head, side, b, tw;
            unit_decode(g, sg.unit, head, side, b, tw);
            const int col = tw * T_TWO + quarter * T_QPX + lane;
            const bool col_ok = lane < T_QPX && col < g.W;
            int i = (int)((eset + T_NES - orow % T_NES) % T_NES);      // first own row of the segment
            const size_t off0 = (size_t)b * K * plane + (size_t)sg.hb * g.W + (col_ok ? col : 0);     // row 0 of the segment
            float* const pout = (side ? ptrs.dx : ptrs.dy)[head] + off0;
            // fused softmax backward (cy_iic_bwd_logits_heads): the output map's own probabilities of the row, staged through a
            // small shared-memory ring by cp.async two own rows ahead.  (Plain loads into registers do not work here: they share
            // a scoreboard with the tcgen05.ld of the row, whose wait then exposes the whole L2 / HBM latency every row.)
            const float* const pin = (side ? ptrs.x : ptrs.y)[head] + (size_t)b * K * plane + (size_t)sg.hb * g.W;
            const bool sm_bwd = g.inv_T != 0.f;
            const int col0 = tw * T_TWO + quarter * T_QPX;                 // first column of this quarter
            constexpr int ECH = KC * (T_QPX / 4);                          // 16-byte chunks of one staged row: KC channels x 28 floats
            auto stage_row = [&](int ii, int stg) {                        // (every lane commits a group, also an empty one)
                if (sm_bwd && ii < sg.n_out) {
                    const float* q = pin + (size_t)ii * g.W + col0;
#pragma unroll
                    for (int it = 0; it < (ECH + 31) / 32; ++it) {
                        const int idx = lane + 32 * it, ch = idx / (T_QPX / 4), jc = idx % (T_QPX / 4);
                        if (idx < ECH) {
                            const bool ok = ch < K && col0 + 4 * jc < g.W;
                            cp_async_16(ering + (uint32_t)(stg * E_STAGE_BYTES + (ch * 32 + 4 * jc) * 4), ok ? q + (size_t)ch * plane + 4 * jc : pin,
                                        ok ? 16u : 0u);
                        }
                    }
                }
                cp_async_commit();
            };
            int stg = 0;
            if (sm_bwd) {
                stage_row(i, 0);
                stage_row(i + T_NES, 1);
            }
            for (; i < sg.n_out; i += T_NES) {
                const uint32_t r_ = orow + (uint32_t)i, slot = r_ % T_ND;
                float pv[KC];
                if (sm_bwd) {
                    stage_row(i + 2 * T_NES, stg == 0 ? 2 : stg - 1);
                    cp_async_wait<2>();                                    // this row's group has landed ...
                    __syncwarp();                                          // ... for every lane
#pragma unroll
                    for (int o = 0; o < KC; ++o) pv[o] = lds_f32_own(ering + (uint32_t)(stg * E_STAGE_BYTES + (o * 32 + lane) * 4));
                    stg = stg == 2 ? 0 : stg + 1;
                }
                TC_T(0, mbar_wait_backoff<CY_TC_SLEEP_EPI>(d_full + slot, (r_ / T_ND) & 1u));
#ifdef CY_TC_TIMING
                const long long te0 = clock64();
#endif
                tc_fence_after();
                uint32_t r[16];
                const uint32_t td = tmem + ((uint32_t)(quarter * 32) << 16) + slot * 16;
                if constexpr (KC <= 8) {
                    tmem_ld_cols<KC>(td, r);
                } else if constexpr (KC <= 10) {
                    tmem_ld_32x8(td, *reinterpret_cast<uint32_t (*)[8]>(r));
                    tmem_ld_32x2(td + 8, r + 8);
                } else {
                    tmem_ld_32x16(td, r);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty + slot);
                if (col_ok) {
                    if (sm_bwd) {                // dL/dlogit_o = p_o (dL/dp_o - sum_k p_k dL/dp_k) / T   (1/T is in the weights; channels past K: p = 0)
                        float dot = 0.f;
#pragma unroll
                        for (int o = 0; o < KC; ++o) dot = fmaf(pv[o], __uint_as_float(r[o]), dot);
#pragma unroll
                        for (int o = 0; o < KC; ++o) r[o] = __float_as_uint(pv[o] * (__uint_as_float(r[o]) - dot));
                    }
                    float* q = pout + (size_t)i * g.W;
#pragma unroll
                    for (int o = 0; o < KC; ++o) {
                        if (o < KC - 1 || o < K) *q = __uint_as_float(r[o]);
                        q += plane;
                    }
                }
#ifdef CY_TC_TIMING
                tacc[1] += clock64() - te0;
                tacc[2] += 1;
#endif
            }
            if (sm_bwd) {
                cp_async_wait<0>();                                        // nothing in flight when the next segment reuses the ring
                __syncwarp();
            }
            orow += (uint32_t)sg.n_out;
        }
        if (warp == T_EPI0) TC_TDUMP(48, 3);
    }
    tc_fence_before();
    __syncthreads();
#ifdef CY_TC_TIMING
    if (threadIdx.x == 0) {
        const long long t_end = clock64();
        if (blockIdx.x == 0) { tc_dbg[56] = t_end - t_start; tc_dbg[58] = t_start - t_entry; }
        atomicMax(reinterpret_cast<unsigned long long*>(tc_dbg + 57), (unsigned long long)(t_end - t_start));
        atomicMax(reinterpret_cast<unsigned long long*>(tc_dbg + 59), (unsigned long long)(t_start - t_entry));
    }
#endif
    if (warp == T_WALLOC) tmem_dealloc(*tmem_slot, 512);
}

// weight tiles of every head + barriers + the converters' staging ring
size_t tc_smem_bytes(int KH, int S, bool softmax_bwd) {
    const int KS = (3 * KH + 7) / 8;
    return (size_t)S * 2 * 3 * KS * 2 * T_WTILE + 512 + (size_t)4 * T_NCS * T_PF * (2 * KH) * 36 * 4 +
           (softmax_bwd ? (size_t)4 * T_NES * 3 * (2 * KH) * 32 * 4 : 0);
}

template <int KH>
int launch_bwd_tc(const TcPtrs& ptrs, const TcGeom& g, const float* djoint, const float* gscale, cudaStream_t st) {
    // one CTA per SM: the register file (>= 52 registers x 896 threads) does not hold two, so the 512-column TMEM allocation
    // never waits for a co-resident CTA
    size_t smem = tc_smem_bytes(KH, g.S, g.inv_T != 0.f);
    auto k = iic_bwd_tc_kernel<KH>;
    static SmemAttrCache attr;
    if (attr.need(smem)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("iic_bwd_tc smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr.set(smem);
    }
    const int sms = device_sm_count();
    const int grid = g.rows_total < sms ? g.rows_total : sms;
    k<<<grid, T_THREADS, smem, st>>>(ptrs, g, djoint, gscale);
    CY_CHECK_LAUNCH("iic_bwd_tc");
#ifdef CY_TC_TIMING
    {
        long long h[64];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, tc_dbg, sizeof(h));
        const double rows = (double)g.rows_total / grid;
        fprintf(stderr, "[tc timing] main loop CTA0 %lld cycles, max over CTAs %lld; setup CTA0 %lld, max %lld\n", h[56], h[57], h[58], h[59]);
        { long long z[64] = {0}; cudaMemcpyToSymbol(tc_dbg, z, sizeof(z)); }
        fprintf(stderr, "[tc timing] issuer0 loop total %lld, first-row wait %lld; issuer1 %lld, %lld\n", h[4], h[5], h[12], h[13]);
        fprintf(stderr, "[tc timing, cycles per OWN row, CTA 0 (%.0f rows)] issuer0: waits %.0f issue+commit %.0f | issuer1: %.0f %.0f\n"
                "  converter w%d (%lld own rows): data wait %.0f, top-to-a_empty %.0f, a_empty wait %.0f, store+arrive %.0f | w%d: %.0f %.0f %.0f %.0f\n"
                "  epilogue w%d (%lld own rows): d_full wait %.0f, ld+store %.0f\n",
                rows, h[0] / rows * T_NISS, h[3] / rows * T_NISS, h[8] / rows * T_NISS, h[11] / rows * T_NISS,
                T_CONV0, h[36], (double)h[32] / h[36], (double)h[33] / h[36], (double)h[34] / h[36], (double)h[35] / h[36],
                T_CONV0 + 5, (double)h[40 + 0] / h[44], (double)h[41] / h[44], (double)h[42] / h[44], (double)h[43] / h[44],
                T_EPI0, h[50], (double)h[48] / h[50], (double)h[49] / h[50]);
    }
#endif
    return CY_OK;
}

}  // namespace

// n_heads independent (x, y, dL/dJ) problems of one shape in ONE launch (the sub-head stack of IIDSegmentationLoss callers,
// SURVEY.md 8(f2)); head s reads djoint + s * dj_stride.  Returns CY_ERR_UNSUPPORTED when the shape is not eligible or the heads'
// weight tiles do not fit beside the staging ring (the caller then loops over the heads).
int iic_bwd_tc_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                     const float* djoint, long long dj_stride, const float* gscale, void* const* dxs, void* const* dys,
                     float softmax_inv_T, cudaStream_t st) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (dtype != CY_F32 || pad != 1 || K > 16 || K < 1 || (W % 4) != 0 || n_heads < 1 || n_heads > T_MAXHEADS) return CY_ERR_UNSUPPORTED;
    TcGeom g;
    TcPtrs ptrs;
    for (int s = 0; s < T_MAXHEADS; ++s) {
        const int t = s < n_heads ? s : 0;
        if (!al16(xs[t]) || !al16(ys[t])) return CY_ERR_UNSUPPORTED;
        ptrs.x[s] = reinterpret_cast<const float*>(xs[t]);
        ptrs.y[s] = reinterpret_cast<const float*>(ys[t]);
        ptrs.dx[s] = reinterpret_cast<float*>(dxs[t]);
        ptrs.dy[s] = reinterpret_cast<float*>(dys[t]);
    }
    g.B = B; g.K = K; g.H = H; g.W = W; g.S = n_heads; g.dj_stride = dj_stride; g.inv_T = softmax_inv_T;
    g.TW2 = (W + T_TWO - 1) / T_TWO;
    const long long rows = 2LL * n_heads * B * g.TW2 * H;
    if (rows <= 0 || rows > 0x7fffffffLL / 2 || (long long)K * H * W > 0x7fffffffLL / 8) return CY_ERR_UNSUPPORTED;
    g.rows_total = (int)rows;
    const int KH = (K + 1) / 2;
    if (tc_smem_bytes(KH, n_heads, softmax_inv_T != 0.f) > (size_t)227 * 1024) return CY_ERR_UNSUPPORTED;
    switch (KH) {
        case 1: return launch_bwd_tc<1>(ptrs, g, djoint, gscale, st);
        case 2: return launch_bwd_tc<2>(ptrs, g, djoint, gscale, st);
        case 3: return launch_bwd_tc<3>(ptrs, g, djoint, gscale, st);
        case 4: return launch_bwd_tc<4>(ptrs, g, djoint, gscale, st);
        case 5: return launch_bwd_tc<5>(ptrs, g, djoint, gscale, st);
        case 6: return launch_bwd_tc<6>(ptrs, g, djoint, gscale, st);
        case 7: return launch_bwd_tc<7>(ptrs, g, djoint, gscale, st);
        default: return launch_bwd_tc<8>(ptrs, g, djoint, gscale, st);
    }
}

// returns CY_ERR_UNSUPPORTED when the shape is not eligible (the caller then takes the mma.sync / CUDA-core kernels)
int iic_bwd_tc(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
               const float* gscale, void* dx, void* dy, cudaStream_t st) {
    return iic_bwd_tc_heads(&x, &y, 1, dtype, B, K, H, W, pad, djoint, 0, gscale, &dx, &dy, 0.f, st);
}

}  // namespace cy
