// IIC joint / adjoint on the tensor pipe: warp-level mma.sync.m16n8k16 (bf16 hi/lo split, fp32 accumulate) fed from
// TMA-staged fp32 halo boxes.  Reference: contrastyou/losses/discreteMI.py:225-243 (compute_joint_2D) and its autograd.
//
// Why mma.sync and not tcgen05 here: the contraction is tiny in M and N (K*T = 30 rows/columns at config 3) and huge in
// the reduction dimension (pixels).  profiles/probes/probe_umma_small.cu measured a floor of ~45 cycles per tcgen05.mma
// for M<=128, N<=64 tiles read from shared memory (operand fetch is bound by the 128 B/clk port, which the fp32->bf16
// conversion pass would share), i.e. >= 2.8 clk/pixel/SM, against 3.0 clk/pixel/SM for the register-fed HMMA form below,
// which needs no operand re-layout at all (shifted fragments are plain shared-memory loads).
//
// Precision: every fp32 value v is split as v = hi + lo, hi = bf16(v) rounded, lo = bf16(v - hi) (signed, so the dropped
// lo*lo term and the truncation of lo are zero-mean); products use hi*hi + lo*hi + hi*lo: error <= 2^-16 relative per
// product, fp32 accumulation — well inside the 1e-4 parity bar.
//
// Forward (iic_joint_mma_kernel):  J[k1,k2,dy,dx] = sum x[k1,h+dy-p,w'] * y[k2,h,w'-dx+p]   (w' = w+dx-p)
//   D[(dx,k2) , (k1,dy)] += A[(dx,k2), pixel] * B[pixel, (k1,dy)]        A: shifted rows of y,  B: rows of x
//   K*T = 30 combinations fill 30 of the 32 rows / columns of 2 x 4 MMA tiles (94 %).  The k index of one MMA covers 16
//   consecutive pixels of one image row; k-slot (2q, 2q+1, 2q+8, 2q+9) of lane quad q holds pixels (q, q+4, q+8, q+12),
//   so every fragment element is one conflict-free LDS.32 and a displacement is an address offset.
// Backward (iic_bwd_mma_kernel):   out[o,h,w] = sum_{c,dy,dx} Wt[o,c,dy,dx] * in[c,h+dy-p,w+dx-p]
//   D[o (16), pixel (8)] += Wt_dy[o, (c,dx)] * col[(c,dx), pixel] per box row, accumulated over dy in rolling HMMA
//   accumulators (direct form), or D[(dy,o) (32), pixel (8)] with the dy contributions rolled through the C operands
//   (T form, K <= 10, default); the weights (dL/dJ, scaled; flipped for dL/dx) stay in registers as pre-split A
//   fragments for the whole kernel, half of the warps produce dL/dy from the x box and half dL/dx from the y box.
//
// What bounds these kernels (profiles/README.md): on sm_100a an mma.sync holds the issue port of its scheduler for its 8
// pipe cycles, so per scheduler time = 8 * HMMAs + CUDA-core instructions; neither TMA nor HBM is the limiter.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cy {

using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encode_fn();

namespace {

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// first MMA of a chain: C = 0
__device__ __forceinline__ void mma_bf16_16816_z(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
// plain (non-volatile) shared-memory load from a 32-bit shared address: ptxas may hoist it above the volatile MMAs of
// the previous step (software pipelining), but not above the mbarrier wait (memory clobber)
// compiler-level ordering point between volatile asm statements (no instruction is emitted)
__device__ __forceinline__ void mma_order_fence() { asm volatile("" ::: "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    return *reinterpret_cast<const float*>(__cvta_shared_to_generic(addr));
}

// (v0, v1) -> packed bf16 pairs (v0 in the low half), v = hi + lo.
// hi = v rounded to bf16 (round-half-up on the bit pattern), lo = v - hi (exact in fp32, signed, |lo| <= 2^-9 |v|)
// truncated to bf16.  Integer / FADD / PRMT only: cvt.rn.bf16x2.f32 (F2FP) issues on the XU pipe at 1/16 rate and was the
// measured bottleneck of the first version of these kernels (profiles/README.md).
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const uint32_t h0 = (__float_as_uint(v0) + 0x8000u) & 0xffff0000u;
    const uint32_t h1 = (__float_as_uint(v1) + 0x8000u) & 0xffff0000u;
    hi = __byte_perm(h0, h1, 0x7632);
    const float r0 = v0 - __uint_as_float(h0);
    const float r1 = v1 - __uint_as_float(h1);
    lo = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
}
// Position-pinned variants (volatile asm keeps the source order, which is the issue order of an in-order warp): used where
// CUDA-core work is interleaved by hand between dependent HMMAs.
__device__ __forceinline__ uint32_t lds_b32_pinned(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void split_pair_pinned(uint32_t v0, uint32_t v1, uint32_t& hi, uint32_t& lo) {
    asm volatile(
        "{\n\t.reg .b32 h0, h1, r0, r1;\n\t"
        "add.u32 h0, %2, 0x8000;\n\tadd.u32 h1, %3, 0x8000;\n\t"
        "and.b32 h0, h0, 0xffff0000;\n\tand.b32 h1, h1, 0xffff0000;\n\t"
        "prmt.b32 %0, h0, h1, 0x7632;\n\t"
        "sub.f32 r0, %2, h0;\n\tsub.f32 r1, %3, h1;\n\t"
        "prmt.b32 %1, r0, r1, 0x7632;\n\t}"
        : "=r"(hi), "=r"(lo) : "r"(v0), "r"(v1));
}
// one-time split of the adjoint weights: both parts round-to-nearest
__device__ __forceinline__ void split_pair_rn(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(hi << 16);
    const float r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

struct MmaGeom {
    int B, K, H, W, pad, T;
    int TH, TW;              // tile: TH image rows x TW columns
    int BH;                  // rows of a staged box without halo (>= TH; 9 keeps the bank layout)
    int HH;                  // BH + 2 * pad
    int XW;                  // forward: x box pitch (TW + 4);  backward: pitch of both boxes
    int YW, CO;              // forward: y box pitch (TW + 12), first column = w0 - CO; backward: CO of both boxes
    int tiles_h, tiles_w, n_tiles;
    int x_stage_floats, y_stage_floats;
    int n_combo;             // K * T
    int stages;
    int debug_skip;          // CY_IIC_DEBUG_SKIP=1: stage the boxes but skip the arithmetic (data-movement floor; results invalid)
};

// ring of `n` stages with a runtime depth (the depth is chosen on the host from the shared-memory budget)
struct RingN {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ uint32_t stage() const { return s; }
    __device__ __forceinline__ uint32_t phase() const { return ph; }
    __device__ __forceinline__ void next(uint32_t n) {
        if (++s == n) { s = 0; ph ^= 1u; }
    }
};
constexpr int MAX_STAGES = 4;
constexpr int FW_ROWS = 8;       // image rows per forward tile = row-warps per split (the staged boxes keep 9 / 9 + 2p rows: bank layout)

// ------------------------------------------------------------------------------------------------------ forward
// MT m-tiles (16 rows of (dx,k2) each), NT n-tiles (8 columns of (k1,dy)) split over NSPLIT warps per pixel row, NCH chunks
// of 16 pixels per tile row.  The HMMA accumulator truncates on every add, so a chain is kept to the 3 split terms of
// one 16-pixel step (the first MMA takes C = 0) and the result is added to the running sums with round-to-nearest FADDs:
// the joint then carries no bias that grows with the number of pixels (needed for the padding = 0 loss, whose value is
// a 1e-3 residual of O(1) terms).
// blockIdx.y = sub-head: S independent (x, y) pairs of one shape share a launch, each with gridDim.x CTAs and its own partial
// joints partials[head][cta][nj] (cy_iic_joint_heads; gridDim.y = 1 for the plain call)
constexpr int MMA_MAXHEADS = 8;
struct JointMaps {
    CUtensorMap x[MMA_MAXHEADS], y[MMA_MAXHEADS];
};
template <int MT, int NTW, int NSPLIT, int NCH>
__global__ void __launch_bounds__(FW_ROWS * NSPLIT * 32, NSPLIT == 1 ? 2 : 1)
iic_joint_mma_kernel(const __grid_constant__ JointMaps maps, MmaGeom g, float* __restrict__ partials) {
    const CUtensorMap& tmx = maps.x[blockIdx.y];
    const CUtensorMap& tmy = maps.y[blockIdx.y];
    partials += (size_t)blockIdx.y * gridDim.x * (size_t)(g.K * g.K * g.T * g.T);
    constexpr int NCW = FW_ROWS * NSPLIT;
    constexpr int THREADS = NCW * 32;
    // the partial-joint reduction behind this kernel is a programmatic dependent launch: let its CTAs take the SMs as ours
    // retire (they wait in cudaGridDependencySynchronize until this whole grid has completed)
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float* stage0 = reinterpret_cast<float*>(smem);
    const int stage_floats = g.x_stage_floats + g.y_stage_floats;
    const int K = g.K, T = g.T;
    const int nj = K * K * T * T;
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)g.stages * stage_floats);
    uint64_t* empty = full + MAX_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < g.stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NCW); }
        fence_barrier_init();
        prefetch_tmap(&tmx);
        prefetch_tmap(&tmy);
    }
    __syncthreads();

    const uint32_t stage_bytes = (uint32_t)(K * g.HH * g.XW + K * g.BH * g.YW) * 4u;
    // thread 0 is also the TMA producer.  produce(it, block) requests the it-th tile of this CTA into stage it % stages
    // once every warp has released that stage; with block == false it gives up (returns false) if they have not yet.
    auto produce = [&](int it, bool block) -> bool {
        const int tile = blockIdx.x + it * (int)gridDim.x;
        if (tile >= g.n_tiles) return true;
        const uint32_t s = (uint32_t)it % (uint32_t)g.stages, ph = ((uint32_t)it / (uint32_t)g.stages) & 1u;
        if (block) mbar_wait(empty + s, ph ^ 1u);
        else if (!mbar_try_wait(empty + s, ph ^ 1u)) return false;
        const int b = tile / (g.tiles_h * g.tiles_w);
        const int trem = tile % (g.tiles_h * g.tiles_w);
        const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * g.TW;
        if (g.debug_skip == 3) { mbar_arrive(full + s); return true; }      // arithmetic on stale shared memory, no loads
        mbar_arrive_expect_tx(full + s, stage_bytes);
        float* xs = stage0 + (size_t)s * stage_floats;
        tma_load_3d(xs, &tmx, full + s, w0, h0 - g.pad, b * K);
        tma_load_3d(xs + g.x_stage_floats, &tmy, full + s, w0 - g.CO, h0, b * K);
        return true;
    };
    if (threadIdx.x == 0)
        for (int it = 0; it < g.stages - 1; ++it) produce(it, true);

    float acc[MT][NTW][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NTW; ++b2)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b2][c] = 0.f;
    const int gq = lane >> 2, q = lane & 3;          // fragment row / column group, k-slot quad
    const int split = warp % NSPLIT;
    {
        const int row = warp / NSPLIT;
        // shared-memory byte offsets (relative to the stage) of this lane's fragment rows; combinations past K*T read
        // in-bounds garbage whose accumulator rows / columns are never written out
        const int limit = stage_floats - 16 * NCH - 16;
        uint32_t xoff[NTW], yoff[MT][2];
#pragma unroll
        for (int tn = 0; tn < NTW; ++tn) {
            const int cb = 8 * (split * NTW + tn) + gq;               // (k1, dy) = (cb / T, cb % T)
            const int o = ((cb / T) * g.HH + row + (cb % T)) * g.XW + q;
            xoff[tn] = 4u * (uint32_t)((cb < g.n_combo || o < limit) ? o : limit);
        }
#pragma unroll
        for (int tm = 0; tm < MT; ++tm)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int ca = 16 * tm + 8 * hf + gq;                 // (dx, k2) = (ca / K, ca % K)
                int o = g.x_stage_floats + ((ca % K) * g.BH + row) * g.YW + q + g.CO + g.pad - (ca / K);
                if (ca >= g.n_combo) o = o < limit ? (o < 0 ? 0 : o) : limit;
                yoff[tm][hf] = 4u * (uint32_t)o;
            }
        const uint32_t stage0_addr = smem_u32(stage0);
        const uint32_t stage_stride = (uint32_t)stage_floats * 4u;

        RingN ring;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ring.next(g.stages), ++it) {
            const uint32_t s = ring.stage();
            mbar_wait(full + s, ring.phase());
            const uint32_t sb = stage0_addr + s * stage_stride;
            // refill the stage released one iteration ago as early as the other warps allow (without stalling on them)
            bool refilled = threadIdx.x != 0 || produce(it + g.stages - 1, false);
            // Software pipeline over the NCH * MT sub-steps (16 pixels x one m-tile) of this warp's tile row.  The warp
            // issues in order, so source order is the schedule: raw loads of the NEXT sub-step's operands go out before
            // the 3 * NTW MMAs of the current one, their hi/lo split follows the MMAs (in the shadow of 8 HMMA-pipe
            // cycles each) and only the flush FADDs wait for the MMA results.
            if (g.debug_skip != 1) {
                uint32_t bh[NTW][2], bl[NTW][2], ah[4], al[4];
                float rb[NTW][4], ra[2][4];
                auto load_b = [&](int ch) {
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn) {
                        const uint32_t p = sb + xoff[tn] + ch * 64;
                        rb[tn][0] = lds_f32(p); rb[tn][1] = lds_f32(p + 16); rb[tn][2] = lds_f32(p + 32); rb[tn][3] = lds_f32(p + 48);
                    }
                };
                auto split_b = [&]() {
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn) {
                        split_pair(rb[tn][0], rb[tn][1], bh[tn][0], bl[tn][0]);
                        split_pair(rb[tn][2], rb[tn][3], bh[tn][1], bl[tn][1]);
                    }
                };
                auto load_a = [&](int ch, int tm) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t p = sb + yoff[tm][hf] + ch * 64;
                        ra[hf][0] = lds_f32(p); ra[hf][1] = lds_f32(p + 16); ra[hf][2] = lds_f32(p + 32); ra[hf][3] = lds_f32(p + 48);
                    }
                };
                auto split_a = [&]() {
                    split_pair(ra[0][0], ra[0][1], ah[0], al[0]);
                    split_pair(ra[1][0], ra[1][1], ah[1], al[1]);
                    split_pair(ra[0][2], ra[0][3], ah[2], al[2]);
                    split_pair(ra[1][2], ra[1][3], ah[3], al[3]);
                };
                load_b(0);
                load_a(0, 0);
                split_b();
                split_a();
#pragma unroll
                for (int ss = 0; ss < NCH * MT; ++ss) {
                    const int ch = ss / MT, tm = ss % MT;
                    const bool more = ss + 1 < NCH * MT, new_chunk = more && (ss + 1) % MT == 0;
                    if (tm == 0 && ch > 0 && !refilled) refilled = produce(it + g.stages - 1, false);
                    if (more) load_a((ss + 1) / MT, (ss + 1) % MT);
                    if (new_chunk) load_b(ch + 1);
                    float d[NTW][4];           // NTW independent chains of 3 dependent HMMAs, issued interleaved
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn) mma_bf16_16816_z(d[tn], ah, bh[tn]);
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn) mma_bf16_16816(d[tn], al, bh[tn]);
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn) mma_bf16_16816(d[tn], ah, bl[tn]);
                    mma_order_fence();
                    if (more) split_a();
                    if (new_chunk) split_b();
                    mma_order_fence();
#pragma unroll
                    for (int tn = 0; tn < NTW; ++tn)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[tm][tn][c] += d[tn][c];
                }
            }
            if (!refilled) produce(it + g.stages - 1, true);       // every warp has moved on from that stage by now
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
    // per-CTA reduction over the compute warps through shared memory (the ring is idle now), in a fixed order
    __syncthreads();
    constexpr int NV = MT * NTW * 4;                 // accumulator values per lane
    float* red = stage0;                             // [NCW][NV][32]
#pragma unroll
    for (int tm = 0; tm < MT; ++tm)
#pragma unroll
        for (int tn = 0; tn < NTW; ++tn)
#pragma unroll
            for (int c = 0; c < 4; ++c) red[(warp * NV + (tm * NTW + tn) * 4 + c) * 32 + lane] = acc[tm][tn][c];
    for (int i = threadIdx.x; i < nj; i += THREADS) partials[(size_t)blockIdx.x * nj + i] = 0.f;
    __syncthreads();
    for (int e = threadIdx.x; e < NSPLIT * NV * 32; e += THREADS) {
        const int sp = e / (NV * 32), v = (e / 32) % NV, ln = e % 32;
        float sum = 0.f;
        for (int r = 0; r < FW_ROWS; ++r) sum += red[((r * NSPLIT + sp) * NV + v) * 32 + ln];
        const int c = v & 3, tn = (v >> 2) % NTW, tm = (v >> 2) / NTW;
        const int ca = 16 * tm + (ln >> 2) + (c >> 1) * 8;
        const int cb = 8 * (sp * NTW + tn) + 2 * (ln & 3) + (c & 1);
        if (ca < g.n_combo && cb < g.n_combo) {
            const int dx = ca / K, k2 = ca % K, k1 = cb / T, dy = cb % T;
            partials[(size_t)blockIdx.x * nj + ((k1 * K + k2) * T + dy) * T + dx] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------------------ backward
// padding = 1, K <= 16.  One warp walks an 8-pixel column strip of the box row by row.  For box row r' the B operand
// col[(c,dx), pixel] = in[c, r', pixel + dx - 1] (K*3 k-slots, KS k-steps of 16) is loaded and split ONCE and used by
// the three output rows it feeds: acc[r' - dy] += Wt_dy[o, (c,dx)] * col for dy = 0..2 — three independent HMMA chains
// whose accumulators roll through the strip, so an output row is complete (18 chained HMMAs at K = 10) when its third
// box row has been consumed and goes straight from the accumulator registers to global memory.
// The weights (dL/dJ * gscale; flipped for dL/dx) stay in registers as pre-split A fragments for the whole kernel; half
// of the warps produce dL/dy from the x box and half dL/dx from the y box.
constexpr int BW_WARPS_PER_GROUP = 4;
constexpr int BW_THREADS = 2 * BW_WARPS_PER_GROUP * 32;

#ifndef CY_BW_CTAS
#define CY_BW_CTAS 2
#endif
// TFORM (K <= 10, default there): the weight rows are (dy, o) — 3*K <= 30 of the 32 rows of two m-tiles — so ONE set of
// 6*KS MMAs per box row yields the contributions D[(dy,o), pixel] to all three output rows at once (12 instead of 18 MMAs
// at K = 10).  Row slots 0..2 hold (dy = slot, o = lane group < 8): the three contributions to an output row meet in the
// same thread, as the C operand of the next box row's MMA chain.  Slot 3 holds the (dy, o >= 8) rows; their partial sums
// move between lane groups with one shuffle per box row (see the loop).  Measured 90 us vs 94 us (direct) on config 3.
template <int TWV, int KS, bool TFORM>
__global__ void __launch_bounds__(BW_THREADS, CY_BW_CTAS)
iic_bwd_mma_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, MmaGeom g,
                   const float* __restrict__ djoint, const float* __restrict__ gscale, float* __restrict__ dx_out,
                   float* __restrict__ dy_out) {
    constexpr int T = 3, TH = 9, HH = TH + 2, XW = TWV + 8, NSTRIP = TWV / 8, CO = 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float* stage0 = reinterpret_cast<float*>(smem);
    const int K = g.K;
    const int stage_floats = 2 * g.x_stage_floats;
    // [2][K][K][T][T] scaled weights: only read while the A fragments are built, so the table borrows the last stage
    // of the ring (the prologue requests tiles for stages 0 .. stages-2 only; a CTA-wide barrier ends the borrow)
    float* wtab = stage0 + (size_t)(g.stages - 1) * stage_floats;
    const int nj = K * K * T * T;
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)g.stages * stage_floats);
    uint64_t* empty = full + MAX_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NCW = 2 * BW_WARPS_PER_GROUP;
    if (threadIdx.x == 0) {
        for (int i = 0; i < g.stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NCW); }
        fence_barrier_init();
    }
    const float scale = gscale[0];
    for (int i = threadIdx.x; i < nj; i += BW_THREADS) {
        // group 0 (dL/dy from x): Wt[o=k2][c=k1][dy][dx] = G[k1,k2,dy,dx]
        // group 1 (dL/dx from y): Wt[o=k1][c=k2][dy][dx] = G[k1,k2,T-1-dy,T-1-dx]
        const int dxx = i % T, dyy = (i / T) % T, c = (i / (T * T)) % K, o = i / (T * T * K);
        wtab[i] = djoint[((c * K + o) * T + dyy) * T + dxx] * scale;
        wtab[nj + i] = djoint[((o * K + c) * T + (T - 1 - dyy)) * T + (T - 1 - dxx)] * scale;
    }
    __syncthreads();

    const uint32_t box_bytes = (uint32_t)(K * HH * XW) * 4u;
    // thread 0 is also the TMA producer (see iic_joint_mma_kernel)
    auto produce = [&](int it, bool block) -> bool {
        const int tile = blockIdx.x + it * (int)gridDim.x;
        if (tile >= g.n_tiles) return true;
        const uint32_t s = (uint32_t)it % (uint32_t)g.stages, ph = ((uint32_t)it / (uint32_t)g.stages) & 1u;
        if (block) mbar_wait(empty + s, ph ^ 1u);
        else if (!mbar_try_wait(empty + s, ph ^ 1u)) return false;
        const int b = tile / (g.tiles_h * g.tiles_w);
        const int trem = tile % (g.tiles_h * g.tiles_w);
        const int h0 = (trem / g.tiles_w) * TH, w0 = (trem % g.tiles_w) * TWV;
        if (g.debug_skip == 3) { mbar_arrive(full + s); return true; }      // arithmetic on stale shared memory, no loads
        mbar_arrive_expect_tx(full + s, 2 * box_bytes);
        float* xs = stage0 + (size_t)s * stage_floats;
        tma_load_3d(xs, &tmx, full + s, w0 - CO, h0 - 1, b * K);
        tma_load_3d(xs + g.x_stage_floats, &tmy, full + s, w0 - CO, h0 - 1, b * K);
        return true;
    };
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmx);
        prefetch_tmap(&tmy);
        for (int it = 0; it < g.stages - 1; ++it) produce(it, true);
    }
    {
        const int gq = lane >> 2, q = lane & 3;
        const int group = warp / BW_WARPS_PER_GROUP, wq = warp % BW_WARPS_PER_GROUP;
        const float* wt = wtab + group * nj;
        float* out = group ? dx_out : dy_out;
        // A fragments.  k-slot pairs of k-step ks <-> this lane quad's own elements e = 4*ks + {0,1 | 2,3}, element e
        // <-> K-row R = q + 4*e = c * 3 + dx.  Direct form: one m-tile per dy, row o = gq (+8).  T form: two m-tiles whose
        // row slots (8 rows each) are (dy = 0), (dy = 1), (dy = 2) for o = gq, and slot 3 = (dy3, o3) = (gq / KX, 8 + gq % KX).
        constexpr int NA = TFORM ? 2 : 3;
        const int KX = K > 8 ? K - 8 : 0;
        const int dy3 = (TFORM && KX && gq < 3 * KX) ? gq / KX : 3;      // 3: this lane has no slot-3 row
        const int o3 = 8 + (KX ? gq % KX : 0);
        uint32_t ah[NA][KS][4], al[NA][KS][4];
#pragma unroll
        for (int ia = 0; ia < NA; ++ia)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                float w[2][4];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    int o, dyy;
                    if constexpr (TFORM) {
                        const int slot = 2 * ia + hf;
                        dyy = slot < 3 ? slot : dy3;
                        o = slot < 3 ? gq : o3;
                    } else {
                        dyy = ia;
                        o = gq + 8 * hf;
                    }
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const int R = q + 4 * (4 * ks + e4);
                        const int c = R / 3, dxx = R % 3;
                        w[hf][e4] = (dyy < 3 && o < K && c < K) ? wt[((o * K + c) * T + dyy) * T + dxx] : 0.f;
                    }
                }
                split_pair_rn(w[0][0], w[0][1], ah[ia][ks][0], al[ia][ks][0]);
                split_pair_rn(w[1][0], w[1][1], ah[ia][ks][1], al[ia][ks][1]);
                split_pair_rn(w[0][2], w[0][3], ah[ia][ks][2], al[ia][ks][2]);
                split_pair_rn(w[1][2], w[1][3], ah[ia][ks][3], al[ia][ks][3]);
            }
        // byte offsets (within a box) of this lane's 4*KS K-rows at box row 0, strip 0; rows past K*3 carry zero weights
        // and re-read row 0 (finite data)
        uint32_t roff[4 * KS];
#pragma unroll
        for (int j = 0; j < 4 * KS; ++j) {
            int R = q + 4 * j;
            if (R >= K * 3) R = 0;
            roff[j] = 4u * (uint32_t)((R / 3) * HH * XW + (R % 3) + gq + CO - 1);
        }
        const uint32_t stage0_addr = smem_u32(stage0);
        __syncthreads();                                  // every warp has built its fragments: the last stage is free
        RingN ring;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ring.next(g.stages), ++it) {
            const int b = tile / (g.tiles_h * g.tiles_w);
            const int trem = tile % (g.tiles_h * g.tiles_w);
            const int h0 = (trem / g.tiles_w) * TH, w0 = (trem % g.tiles_w) * TWV;
            const uint32_t s = ring.stage();
            mbar_wait(full + s, ring.phase());
            const uint32_t box = stage0_addr + 4u * (uint32_t)(s * stage_floats + group * g.x_stage_floats);
            bool refilled = threadIdx.x != 0 || produce(it + g.stages - 1, false);
#pragma unroll 1
            for (int strip = wq; strip < (g.debug_skip == 1 ? 0 : NSTRIP); strip += BW_WARPS_PER_GROUP) {
                const uint32_t sbase = box + strip * 32;
                const int w = w0 + 8 * strip + 2 * q;
                float* op = out + (((size_t)b * K + gq) * g.H + h0) * g.W + w;       // (o = gq, row h0); + W per finished row
                const size_t ostride8 = (size_t)8 * g.H * g.W;
                const int rows_ok = g.H - h0;                 // W % 4 == 0 and w even: a pixel pair is inside the row together
                const bool st0 = w < g.W && gq < K && g.debug_skip != 2, st1 = w < g.W && gq + 8 < K && g.debug_skip != 2;
                if constexpr (!TFORM) {
                // Software pipeline (the warp issues in order, so source order is the schedule): the loads of box row
                // rp+1 are issued BEFORE the MMAs of row rp and converted AFTER them, i.e. in the shadow of the 18 x 8
                // HMMA-pipe cycles; only the stores of a finished row wait for the MMA results.
                float acc[3][4];                              // rolling: output row r lives in acc[r % 3]
                uint32_t bh[KS][2], bl[KS][2];
                float v[4 * KS];
#pragma unroll
                for (int j = 0; j < 4 * KS; ++j) v[j] = lds_f32(sbase + roff[j]);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    split_pair(v[4 * ks], v[4 * ks + 1], bh[ks][0], bl[ks][0]);
                    split_pair(v[4 * ks + 2], v[4 * ks + 3], bh[ks][1], bl[ks][1]);
                }
#pragma unroll
                for (int rp = 0; rp < HH; ++rp) {
                    if ((rp == 3 || rp == 7) && !refilled) refilled = produce(it + g.stages - 1, false);
                    if (rp + 1 < HH) {
#pragma unroll
                        for (int j = 0; j < 4 * KS; ++j) v[j] = lds_f32(sbase + roff[j] + (rp + 1) * XW * 4);
                    }
                    // three independent chains (one per dy), issued interleaved term by term
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
#pragma unroll
                            for (int dyy = 0; dyy < 3; ++dyy) {
                                const int r = rp - dyy;               // output row fed by (box row rp, dy)
                                if (r < 0 || r >= TH) continue;
                                const uint32_t (&a)[4] = term == 1 ? al[dyy][ks] : ah[dyy][ks];
                                const uint32_t (&bb)[2] = term == 2 ? bl[ks] : bh[ks];
                                if (dyy == 0 && ks == 0 && term == 0) mma_bf16_16816_z(acc[r % 3], a, bb);    // first touch of row r
                                else mma_bf16_16816(acc[r % 3], a, bb);
                            }
                        }
                    }
                    if (rp + 1 < HH) {
                        uint32_t nh[KS][2], nl[KS][2];
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            split_pair(v[4 * ks], v[4 * ks + 1], nh[ks][0], nl[ks][0]);
                            split_pair(v[4 * ks + 2], v[4 * ks + 3], nh[ks][1], nl[ks][1]);
                        }
                        mma_order_fence();                            // keep the conversion below the MMAs above in the issue order
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            bh[ks][0] = nh[ks][0]; bh[ks][1] = nh[ks][1];
                            bl[ks][0] = nl[ks][0]; bl[ks][1] = nl[ks][1];
                        }
                    }
                    const int rc = rp - 2;                            // output row completed by this box row
                    if (rc >= 0) {
                        if (st0 && rc < rows_ok) *reinterpret_cast<float2*>(op) = make_float2(acc[rc % 3][0], acc[rc % 3][1]);
                        if (st1 && rc < rows_ok) *reinterpret_cast<float2*>(op + ostride8) = make_float2(acc[rc % 3][2], acc[rc % 3][3]);
                        op += g.W;
                    }
                }
                } else {
                // ---- T form.  Per box row rp ONE accumulate chain per m-tile, straight into the rolling partial sums:
                //   m-tile 0: C = {0 (row rp, dy 0), p1 (row rp-1 after dy 0)}        -> D = {p1', p2'}
                //   m-tile 1: C = {p2 (row rp-2 after dy 0,1), s3 (slot-3 pass)}      -> D = {row rp-2 complete, s3'}
                // Slot 3 (channels 8 .. K-1) is a 3-stage systolic pass over the lane groups dy3 = 0, 1, 2: the partial sum
                // of output row r starts in the dy3 = 0 lanes at box row r and moves up KX row groups per box row (one
                // shuffle per register), so that it completes in the dy3 = 2 lanes at box row r+2, like the in-lane rows.
                // The warp issues in order and a chain of 3*KS dependent HMMAs leaves ~12 idle issue cycles per pair, so the
                // CUDA-core work of the row (loads of box row rp+1, stores of the row finished one step ago, the hi/lo
                // split of row rp+1) is pinned BETWEEN the HMMA pairs instead of after them.
                const int lim0 = st0 ? (rows_ok < TH ? rows_ok : TH) : 0;                       // rows this lane stores
                const int lim3 = (KX && w < g.W && dy3 == 2 && g.debug_skip != 2) ? (rows_ok < TH ? rows_ok : TH) : 0;
                float* op3 = out + (((size_t)b * K + o3) * g.H + h0) * g.W + w;
                uint32_t bh[2][KS][2], bl[2][KS][2], v[4 * KS];
#pragma unroll
                for (int j = 0; j < 4 * KS; ++j) v[j] = lds_b32_pinned(sbase + roff[j]);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    split_pair_pinned(v[4 * ks], v[4 * ks + 1], bh[0][ks][0], bl[0][ks][0]);
                    split_pair_pinned(v[4 * ks + 2], v[4 * ks + 3], bh[0][ks][1], bl[0][ks][1]);
                }
                float p1[2] = {0.f, 0.f}, p2[2] = {0.f, 0.f}, s3[2] = {0.f, 0.f}, done[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int rp = 0; rp < HH; ++rp) {
                    const int cur = rp & 1, nxt = cur ^ 1;
                    const bool more = rp + 1 < HH;
                    if ((rp == 3 || rp == 7) && !refilled) refilled = produce(it + g.stages - 1, false);
                    float d[2][4] = {{0.f, 0.f, p1[0], p1[1]}, {p2[0], p2[1], s3[0], s3[1]}};
                    int piece = 0;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                        for (int term = 0; term < 3; ++term, ++piece) {
#pragma unroll
                            for (int tm = 0; tm < 2; ++tm) {
                                const uint32_t (&a)[4] = term == 1 ? al[tm][ks] : ah[tm][ks];
                                const uint32_t (&bb)[2] = term == 2 ? bl[cur][ks] : bh[cur][ks];
                                mma_bf16_16816(d[tm], a, bb);
                            }
                            // CUDA-core piece that follows HMMA pair number `piece`
                            if (piece == 0 && more) {
#pragma unroll
                                for (int j = 0; j < 4 * KS; ++j) v[j] = lds_b32_pinned(sbase + roff[j] + (rp + 1) * XW * 4);
                            }
                            if (piece == 1 && rp >= 3) {              // output row rp - 3 was finished by the previous box row
                                mma_order_fence();
                                if (rp - 3 < lim0) *reinterpret_cast<float2*>(op + (size_t)(rp - 3) * g.W) = make_float2(done[0], done[1]);
                                if (rp - 3 < lim3) *reinterpret_cast<float2*>(op3 + (size_t)(rp - 3) * g.W) = make_float2(done[2], done[3]);
                                mma_order_fence();
                            }
                            if (more) {
                                const int first = 3 * KS - 2 * KS;            // the last 2*KS pairs carry one split each
                                if (piece >= first) {
                                    const int e = piece - first, ks2 = e / 2, hf2 = e % 2;
                                    split_pair_pinned(v[4 * ks2 + 2 * hf2], v[4 * ks2 + 2 * hf2 + 1], bh[nxt][ks2][hf2], bl[nxt][ks2][hf2]);
                                }
                            }
                        }
                    }
                    done[0] = d[1][0]; done[1] = d[1][1]; done[2] = d[1][2]; done[3] = d[1][3];
                    p1[0] = d[0][0]; p1[1] = d[0][1];
                    p2[0] = d[0][2]; p2[1] = d[0][3];
                    if (KX) {
                        const float t0 = __shfl_up_sync(0xffffffffu, d[1][2], 4 * KX), t1 = __shfl_up_sync(0xffffffffu, d[1][3], 4 * KX);
                        s3[0] = dy3 == 0 ? 0.f : t0;
                        s3[1] = dy3 == 0 ? 0.f : t1;
                    }
                }
                if (HH - 3 < lim0) *reinterpret_cast<float2*>(op + (size_t)(HH - 3) * g.W) = make_float2(done[0], done[1]);
                if (HH - 3 < lim3) *reinterpret_cast<float2*>(op3 + (size_t)(HH - 3) * g.W) = make_float2(done[2], done[3]);
                }
            }
            if (!refilled) produce(it + g.stages - 1, true);       // every warp has moved on from that stage by now
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
}

int make_map3d(CUtensorMap* m, const void* base, int B, int K, int H, int W, int box_w, int box_h) {
    EncodeTiledFn fn = tensor_map_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CY_ERR_DEVICE; }
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)K * (cuuint64_t)B};
    cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)K};
    cuuint32_t estr[3] = {1, 1, 1};
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("CY_IIC_L2PROMO"); promo = e ? atoi(e) : 2; }      // 0 none, 1 64 B, 2 128 B, 3 256 B
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d) failed (%d)", (int)r); return CY_ERR_ARG; }
    return CY_OK;
}

// CY_IIC_DEBUG_SKIP (phase timing: skips arithmetic / stores / loads, results INVALID) exists only in `make EXPERIMENTAL=1`
// builds; the product library always computes.
int debug_skip() {
#ifdef CY_EXPERIMENTAL
    static int v = -1;
    if (v < 0) { const char* e = getenv("CY_IIC_DEBUG_SKIP"); v = e ? atoi(e) : 0; }
    return v;
#else
    return 0;
#endif
}

int sm_count_mma() { return device_sm_count(); }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// tile width: 32 or 64 columns, whichever wastes fewer columns on the ragged edge (ties -> 64: fewer barrier rounds)
int pick_tw(int W) {
    const int w32 = (W + 31) / 32 * 32, w64 = (W + 63) / 64 * 64;
    return w64 <= w32 ? 64 : 32;
}

bool fwd_geom(int B, int K, int H, int W, int pad, MmaGeom* g, int* mt, int* ntw, int* nsplit, size_t* smem) {
    const int T = 2 * pad + 1;
    // padding = 0 stays on the fp32 CUDA-core kernels: its loss is a ~1e-3 residual of O(1) terms, so the joint needs
    // full fp32 products (the bf16 hi/lo split is good to ~1e-6 only), and at K*K MACs per pixel it is HBM-bound there
    if (pad < 1 || pad > 3 || (W % 4) != 0 || K > 256) return false;
    const int nc = K * T;
    *mt = (nc + 15) / 16;
    const int nt = (nc + 7) / 8;
    if (*mt > 4) return false;
    if (*mt * nt <= 8) { *nsplit = 1; *ntw = nt; }
    else if (*mt * ((nt + 1) / 2) <= 16) { *nsplit = 2; *ntw = (nt + 1) / 2; }
    else return false;
    g->B = B; g->K = K; g->H = H; g->W = W; g->pad = pad; g->T = T;
    g->TH = FW_ROWS;
    g->BH = 9;                        // HH = 9 + 2p == T (mod 8): fragment rows (k1, dy) land on distinct bank groups
    g->HH = g->BH + 2 * pad;
    g->CO = 4;
    g->n_combo = nc;
    const size_t budget = (*nsplit == 1) ? 113 * 1024 : 225 * 1024;
    // widest tile (fewest barrier rounds) that does not waste columns and still leaves a 2-deep ring
    const int first = pick_tw(W);
    for (int tw = first; tw >= 32; tw -= 32) {
        g->TW = tw;
        g->XW = tw + 4;               // == 4 (mod 32)
        g->YW = tw + 12;              // (YW / 4) odd: the K planes of one row start on distinct bank groups
        g->tiles_h = (H + g->TH - 1) / g->TH;
        g->tiles_w = (W + tw - 1) / tw;
        g->n_tiles = B * g->tiles_h * g->tiles_w;
        g->x_stage_floats = (K * g->HH * g->XW + 31) & ~31;
        g->y_stage_floats = (K * g->BH * g->YW + 31) & ~31;
        for (int st = MAX_STAGES - 1; st >= 2; --st) {
            *smem = (size_t)st * (g->x_stage_floats + g->y_stage_floats) * 4 + 2 * MAX_STAGES * 8 + 128 + 64;
            const size_t red = (size_t)FW_ROWS * *nsplit * *mt * *ntw * 4 * 32 * 4;      // epilogue reduction buffer reuses the ring
            if (*smem <= budget && (size_t)st * (g->x_stage_floats + g->y_stage_floats) * 4 >= red) { g->stages = st; return true; }
        }
    }
    return false;
}

template <int MT, int NTW, int NSPLIT, int NCH>
int launch_joint_mma(const JointMaps& maps, int n_heads, const MmaGeom& g, size_t smem, int grid, float* partials, cudaStream_t st) {
    auto k = iic_joint_mma_kernel<MT, NTW, NSPLIT, NCH>;
    static SmemAttrCache attr;            // raise the opt-in limit only when it grows (remembered per device)
    if (attr.need(smem)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("iic_joint_mma smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr.set(smem);
    }
    k<<<dim3((unsigned)grid, (unsigned)n_heads), FW_ROWS * NSPLIT * 32, smem, st>>>(maps, g, partials);
    CY_CHECK_LAUNCH("iic_joint_mma");
    return CY_OK;
}

}  // namespace

// returns CY_ERR_UNSUPPORTED when the shape is not eligible (the caller then takes the CUDA-core kernels).
// n_heads (x, y) pairs of one shape in one launch: partials is [n_heads][*n_partials][nj].
int iic_joint_mma_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                        float* partials, int* n_partials, cudaStream_t st) {
    MmaGeom g; int mt, ntw, nsplit; size_t smem;
    if (dtype != CY_F32 || n_heads < 1 || n_heads > MMA_MAXHEADS || !fwd_geom(B, K, H, W, pad, &g, &mt, &ntw, &nsplit, &smem))
        return CY_ERR_UNSUPPORTED;
    for (int s = 0; s < n_heads; ++s)
        if (!aligned16(xs[s]) || !aligned16(ys[s])) return CY_ERR_UNSUPPORTED;
    g.debug_skip = debug_skip();
    JointMaps maps;
    for (int s = 0; s < MMA_MAXHEADS; ++s) {
        if (s >= n_heads) { maps.x[s] = maps.x[0]; maps.y[s] = maps.y[0]; continue; }
        int rc = make_map3d(&maps.x[s], xs[s], B, K, H, W, g.XW, g.HH);
        if (rc) return rc;
        rc = make_map3d(&maps.y[s], ys[s], B, K, H, W, g.YW, g.BH);
        if (rc) return rc;
    }
    int cap = sm_count_mma() * (nsplit == 1 ? 2 : 1) / n_heads;       // resident CTAs shared evenly by the heads
    if (cap < 1) cap = 1;
    const int grid = g.n_tiles < cap ? g.n_tiles : cap;
    *n_partials = grid;
#define CY_JM(MTV, NTWV, NSV)                                                                                     \
    if (mt == MTV && ntw == NTWV && nsplit == NSV) {                                                              \
        if (g.TW == 32) return launch_joint_mma<MTV, NTWV, NSV, 2>(maps, n_heads, g, smem, grid, partials, st);  \
        if (g.TW == 64) return launch_joint_mma<MTV, NTWV, NSV, 4>(maps, n_heads, g, smem, grid, partials, st);  \
    }
    CY_JM(1, 1, 1) CY_JM(1, 2, 1) CY_JM(2, 3, 1) CY_JM(2, 4, 1) CY_JM(3, 3, 2) CY_JM(4, 4, 2)
#undef CY_JM
    return CY_ERR_UNSUPPORTED;
}

int iic_joint_mma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, float* partials, int* n_partials,
                  cudaStream_t st) {
    return iic_joint_mma_heads(&x, &y, 1, dtype, B, K, H, W, pad, partials, n_partials, st);
}

int iic_joint_mma_max_partials() { return 2 * sm_count_mma(); }

namespace {

template <int TWV, int KS, bool TFORM>
int launch_bwd_mma(const CUtensorMap& tmx, const CUtensorMap& tmy, const MmaGeom& g, size_t smem, int grid, const float* djoint,
                   const float* gscale, float* dx, float* dy, cudaStream_t st) {
    auto k = iic_bwd_mma_kernel<TWV, KS, TFORM>;
    static SmemAttrCache attr;
    if (attr.need(smem)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("iic_bwd_mma smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr.set(smem);
    }
    k<<<grid, BW_THREADS, smem, st>>>(tmx, tmy, g, djoint, gscale, dx, dy);
    CY_CHECK_LAUNCH("iic_bwd_mma");
    return CY_OK;
}

}  // namespace

int iic_bwd_mma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
                const float* gscale, void* dx, void* dy, cudaStream_t st) {
    if (dtype != CY_F32 || pad != 1 || (W % 4) != 0 || K > 16 || !aligned16(x) || !aligned16(y) || !aligned16(dx) || !aligned16(dy))
        return CY_ERR_UNSUPPORTED;
    const int T = 3;
    MmaGeom g;
    g.B = B; g.K = K; g.H = H; g.W = W; g.pad = pad; g.T = T;
    g.TH = 9;
    g.HH = g.TH + 2;
    g.CO = 4;
    g.n_combo = K * T;
    const int nj = K * K * T * T;
    size_t smem = 0;
    bool ok = false;
    for (int tw = pick_tw(W); tw >= 32 && !ok; tw -= 32) {
        g.TW = tw;
        g.XW = tw + 8;                  // == 8 (mod 32) and HH * XW == 24 (mod 32): the 4 K-rows of a load hit disjoint banks
        g.YW = g.XW;
        g.tiles_h = (H + g.TH - 1) / g.TH;
        g.tiles_w = (W + tw - 1) / tw;
        g.n_tiles = B * g.tiles_h * g.tiles_w;
        g.x_stage_floats = (K * g.HH * g.XW + 31) & ~31;
        g.y_stage_floats = g.x_stage_floats;
        for (int stg = MAX_STAGES - 1; stg >= 2 && !ok; --stg) {
            smem = (size_t)stg * 2 * g.x_stage_floats * 4 + 2 * MAX_STAGES * 8 + 128 + 64;
            if (smem <= (size_t)(227 / CY_BW_CTAS - 1) * 1024 && 2 * g.x_stage_floats >= 2 * nj) { g.stages = stg; ok = true; }
        }
    }
    if (!ok) return CY_ERR_UNSUPPORTED;
    g.debug_skip = debug_skip();
    CUtensorMap tmx, tmy;
    int rc = make_map3d(&tmx, x, B, K, H, W, g.XW, g.HH);
    if (rc) return rc;
    rc = make_map3d(&tmy, y, B, K, H, W, g.XW, g.HH);
    if (rc) return rc;
    const int cap = CY_BW_CTAS * sm_count_mma();
    const int grid = g.n_tiles < cap ? g.n_tiles : cap;
    float* dxf = reinterpret_cast<float*>(dx);
    float* dyf = reinterpret_cast<float*>(dy);
    const int ks = (K * T + 15) / 16;              // k-steps: K <= 5 -> 1, K <= 10 -> 2, K <= 16 -> 3
    static int tform = -1;
    // A/B switch.  Measured at config 3 (B200): direct form 94 us, T form 90 us (12 instead of 18 MMAs per box row, partial
    // sums rolled through the C operands, CUDA-core work pinned between the HMMA pairs).  CY_IIC_TFORM=0 selects the
    // direct form, which also serves 10 < K <= 16.
    if (tform < 0) { const char* e = getenv("CY_IIC_TFORM"); tform = (e && e[0] == '0') ? 0 : 1; }
#define CY_BW(TWV, KSV, TF) return launch_bwd_mma<TWV, KSV, TF>(tmx, tmy, g, smem, grid, djoint, gscale, dxf, dyf, st)
    if (K <= 10 && tform) {
        if (g.TW == 32 && ks == 1) CY_BW(32, 1, true);
        if (g.TW == 32 && ks == 2) CY_BW(32, 2, true);
        if (g.TW == 64 && ks == 1) CY_BW(64, 1, true);
        if (g.TW == 64 && ks == 2) CY_BW(64, 2, true);
    }
    if (g.TW == 32 && ks == 1) CY_BW(32, 1, false);
    if (g.TW == 32 && ks == 2) CY_BW(32, 2, false);
    if (g.TW == 32 && ks == 3) CY_BW(32, 3, false);
    if (g.TW == 64 && ks == 1) CY_BW(64, 1, false);
    if (g.TW == 64 && ks == 2) CY_BW(64, 2, false);
    if (g.TW == 64 && ks == 3) CY_BW(64, 3, false);
#undef CY_BW
    return CY_ERR_UNSUPPORTED;
}

}  // namespace cy
