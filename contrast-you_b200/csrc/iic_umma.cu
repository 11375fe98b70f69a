// IIC adjoint (padding = 1, K <= 16, fp32 inputs) on the tcgen05 tensor cores.
//
//   dL/dx[k1, p] = sum_{k2, dy, dx} dJ[k1, k2, dy, dx] * y[k2, p + (dy, dx) - 1]        (and the mirrored sum for dL/dy)
//
// is a 3x3 convolution with K -> K channels.  The mma.sync version (iic_mma.cu) is bound by the warp schedulers: on
// sm_100a a legacy HMMA holds the issue port for its 8 pipe cycles (profiles/probes/probe_hmma.cu), so tensor and CUDA-core
// work of a warp scheduler add up instead of overlapping.  tcgen05.mma is asynchronous, which turns the adjoint into a
// pipeline of four warp roles per CTA (one persistent CTA per SM):
//
//   producer  (1 thread)  TMA: fp32 box [K planes][TH+2 rows][120 cols] of the SOURCE tensor of a tile -> stage ring
//   converters (8 warps)  one box row at a time (two sets of 4 warps alternate rows): thread = pixel = TMEM lane.  The K
//                         channel values of the pixel are split into bf16 hi / lo parts (channel pairs packed per 32-bit
//                         word, the A-in-TMEM operand format), the copies shifted by one and two columns come from the
//                         neighbour lanes by shuffle, and the six A tiles {hi, lo} x {3 column shifts} are written to tensor
//                         memory with tcgen05.st.  A warp owns 28 output pixels + the 2-pixel halo of its shifts, so the
//                         shuffles never cross a warp (lanes 28..31 of a quarter carry unused rows).
//   issuer    (1 thread)  per box row 9 MMAs, A from TMEM, B = weights [(dy, k1) = 48 rows][K = 16 channels] from shared
//                         memory (hi and lo copies), D[pixel][(dy, k1)] in TMEM: hi*hi + lo*hi + hi*lo per column shift
//   epilogue  (8 warps)   thread = pixel; two sets of 4 warps own half of the channels each.  Reads the (dy, channel)
//                         columns of a finished box row (TMEM reads run at 64 B/clk per SM, so only the live columns are
//                         read) and rolls the three dy contributions of an output row through registers; an output row
//                         completes two box rows after it started.
//
// Shared memory only carries the TMA boxes (read once by the converters) and the 9 KB of weights per side; the operand
// traffic of the MMAs stays in tensor memory.  A tile is one side (dL/dy from x, or dL/dx from y) x one image x TH rows x
// 112 columns.  Reference: the autograd adjoint of IIDSegmentationLoss (contrastyou/losses/discreteMI.py:96-145 in the
// reference); weights as in iic_mma.cu.
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

constexpr int U_TWO = 112;                       // output columns per tile: 4 lane quarters x 28 pixels
constexpr int U_QPX = 28;                        // output pixels per lane quarter (lanes 28..31: halo / unused)
constexpr int U_BOXW = 120;                      // staged columns: image columns w0-4 .. w0+115
constexpr int U_RING = 5;                        // box rows in flight: TMEM slots of 96 columns (48 D + 6 x 8 A)
constexpr int U_SLOT = 96;
constexpr int U_STAGES = 2;
constexpr int U_CSETS = 3;                       // converter sets (4 warps each); set c converts box rows rc % U_CSETS == c
constexpr int U_NISS = 2;                        // MMA issuer threads (one warp each); issuer i takes box rows rc % U_NISS == i
constexpr int U_CONV0 = 1 + U_NISS;              // first converter warp
constexpr int U_EPI0 = U_CONV0 + 4 * U_CSETS;    // first epilogue warp
constexpr int U_ESETS = 4;                       // epilogue sets (4 warps each); set e owns channels [e * HC, e * HC + HC)
constexpr int U_THREADS = 32 * (U_EPI0 + 4 * U_ESETS);
constexpr int U_WPART_BYTES = 48 * 32;           // weights of one column shift and part: [n 48][k 16] bf16, K-major core matrices
constexpr int U_WDX_BYTES = 2 * U_WPART_BYTES;
constexpr int U_WSIDE_BYTES = 3 * U_WDX_BYTES;
constexpr int U_NBAR = 2 * U_STAGES + 4 * U_RING;

struct UmmaGeom {
    int B, K, H, W;
    int TH, HH;                                  // output rows per tile, staged rows (TH + 2)
    int tiles_h, tiles_w, n_tiles;               // n_tiles counts both sides
    int stage_bytes;                             // K * HH * U_BOXW * 4 rounded up to 128
    int debug_skip;
};

__device__ __forceinline__ void tma_load_box(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// (v0, v1) -> packed bf16 pairs (v0 in the low half): hi = round-half-up bf16, lo = truncated bf16 of the exact remainder
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const uint32_t h0 = (__float_as_uint(v0) + 0x8000u) & 0xffff0000u;
    const uint32_t h1 = (__float_as_uint(v1) + 0x8000u) & 0xffff0000u;
    hi = __byte_perm(h0, h1, 0x7632);
    lo = __byte_perm(__float_as_uint(v0 - __uint_as_float(h0)), __float_as_uint(v1 - __uint_as_float(h1)), 0x7632);
}

template <int KC>                                 // channel count rounded up to {4, 6, 8, 10, 12, 16}
__global__ void __launch_bounds__(U_THREADS, 1)
iic_bwd_umma_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, UmmaGeom g,
                    const float* __restrict__ djoint, const float* __restrict__ gscale, float* __restrict__ dx_out,
                    float* __restrict__ dy_out, long long* __restrict__ dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    long long tph[4] = {0, 0, 0, 0};                 // CY_IIC_DEBUG_SKIP=9: cycles per phase of each role (CTA 0)
    const bool timing = dbg != nullptr && blockIdx.x == 0;
#define CY_T(i, stmt) do { if (timing) { const long long t0__ = clock64(); stmt; tph[i] += clock64() - t0__; } else { stmt; } } while (0)
    const int K = g.K, HH = g.HH;
    uint8_t* stage0 = smem;
    uint8_t* wsm = stage0 + (size_t)U_STAGES * g.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + 2 * U_WSIDE_BYTES);
    uint64_t* full = bars;                       // [U_STAGES]  TMA box landed
    uint64_t* empty = full + U_STAGES;           // [U_STAGES]  converters are done with the box
    uint64_t* cfull = empty + U_STAGES;          // [U_RING]    A tiles of the slot written
    uint64_t* cempty = cfull + U_RING;           // [U_RING]    MMAs that read the A tiles have completed
    uint64_t* afull = cempty + U_RING;           // [U_RING]    D of the slot holds a finished box row
    uint64_t* aempty = afull + U_RING;           // [U_RING]    epilogue has read D
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + U_NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: weights in UMMA layout, barriers, TMEM
    {
        const float scale = gscale[0];
        for (int i = threadIdx.x; i < 2 * 3 * 2 * 48 * 16; i += U_THREADS) {
            const int k = i % 16, n = (i / 16) % 48, part = (i / (16 * 48)) % 2, dxx = (i / (16 * 96)) % 3, side = i / (16 * 96 * 3);
            const int dyy = n / 16, o = n % 16, c = k;
            float w = 0.f;
            if (o < K && c < K) {
                // side 0 (dL/dy from x): Wt[o=k2][c=k1][dy][dx] = G[k1,k2,dy,dx];  side 1 (dL/dx from y): G[k1=o,k2=c,2-dy,2-dx]
                w = side == 0 ? djoint[((c * K + o) * 3 + dyy) * 3 + dxx] : djoint[((o * K + c) * 3 + (2 - dyy)) * 3 + (2 - dxx)];
                w *= scale;
            }
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            const uint32_t off = side * U_WSIDE_BYTES + dxx * U_WDX_BYTES + part * U_WPART_BYTES + (n / 8) * 256 + (k / 8) * 128 +
                                 (n % 8) * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(wsm + off) = part ? lo : hi;
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < U_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 4 * U_CSETS); }
            for (int i = 0; i < U_RING; ++i) {
                mbar_init(cfull + i, 4); mbar_init(cempty + i, 1);
                mbar_init(afull + i, 1); mbar_init(aempty + i, 4 * U_ESETS);
            }
            fence_barrier_init();
            prefetch_tmap(&tmx);
            prefetch_tmap(&tmy);
        }
        if (warp == 1) {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = *tmem_slot;
    const int per_image = g.tiles_h * g.tiles_w;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
                const int side = tile & 1, u = tile >> 1;
                const int b = u / per_image, trem = u % per_image;
                const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * U_TWO;
                const uint32_t s = (uint32_t)it % U_STAGES, ph = ((uint32_t)it / U_STAGES) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                if (g.debug_skip == 3) { mbar_arrive(full + s); continue; }
                mbar_arrive_expect_tx(full + s, (uint32_t)(K * HH * U_BOXW * 4));
                tma_load_box(stage0 + (size_t)s * g.stage_bytes, side ? &tmy : &tmx, full + s, w0 - 4, h0 - 1, b * K);
            }
        }
    } else if (warp < U_CONV0) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        // A single thread needs ~30 cycles per tcgen05.mma it issues, so the box rows alternate between U_NISS issuer
        // threads (rows use disjoint TMEM slots; the 9 accumulating MMAs of one row stay in one thread, in order).
        {   // the whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one elected
            // lane issues
            constexpr uint32_t idesc = idesc_bf16_f32(128, 48, 0, 0);
            const uint32_t w_addr = smem_u32(wsm);
            const int iset = warp - 1;
            uint64_t bdesc[2][3][2];                            // [side][dxx][part]
#pragma unroll
            for (int sd = 0; sd < 2; ++sd)
#pragma unroll
                for (int dxx = 0; dxx < 3; ++dxx)
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt)
                        bdesc[sd][dxx][pt] = smem_desc_noswz(w_addr + sd * U_WSIDE_BYTES + dxx * U_WDX_BYTES + pt * U_WPART_BYTES, 128, 256);
            uint32_t rc = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                const int sd = tile & 1;
                uint64_t bh[3], bl[3];
#pragma unroll
                for (int dxx = 0; dxx < 3; ++dxx) {
                    bh[dxx] = sd ? bdesc[1][dxx][0] : bdesc[0][dxx][0];
                    bl[dxx] = sd ? bdesc[1][dxx][1] : bdesc[0][dxx][1];
                }
                for (int rp = 0; rp < HH; ++rp, ++rc) {
                    if ((int)(rc % U_NISS) != iset) continue;
                    const uint32_t slot = rc % U_RING, ph = (rc / U_RING) & 1u;
                    CY_T(0, mbar_wait(aempty + slot, ph ^ 1u));
                    CY_T(1, mbar_wait(cfull + slot, ph));
                    tc_fence_after();
                    const uint32_t d = tmem + slot * U_SLOT, a = d + 48;     // A tile (part, dxx) at a + (part * 3 + dxx) * 8
                    const long long tm0 = timing ? clock64() : 0;
                    if (elect_one()) {
                    if (g.debug_skip != 1) {
                        umma_bf16_ts_c<0>(d, a, bh[0], idesc);
                        umma_bf16_ts_c<1>(d, a + 24, bh[0], idesc);
                        umma_bf16_ts_c<1>(d, a, bl[0], idesc);
                        umma_bf16_ts_c<1>(d, a + 8, bh[1], idesc);
                        umma_bf16_ts_c<1>(d, a + 32, bh[1], idesc);
                        umma_bf16_ts_c<1>(d, a + 8, bl[1], idesc);
                        umma_bf16_ts_c<1>(d, a + 16, bh[2], idesc);
                        umma_bf16_ts_c<1>(d, a + 40, bh[2], idesc);
                        umma_bf16_ts_c<1>(d, a + 16, bl[2], idesc);
                    }
                    umma_commit(cempty + slot);
                    umma_commit(afull + slot);
                    }
                    __syncwarp();
                    if (timing) tph[2] += clock64() - tm0;
                }
            }
            if (timing && iset == 0 && lane == 0) { dbg[0] = tph[0]; dbg[1] = tph[1]; dbg[2] = tph[2]; }
        }
    } else if (warp < U_EPI0) {
        // ------------------------------------------------------------------------------------------ converters
        const int cset = (warp - U_CONV0) >> 2, quarter = warp & 3;
        const int jj = quarter * U_QPX + lane;                 // this lane's pixel <-> image column w0-1+jj, box column jj+3
        constexpr int NP = KC / 2;                             // live channel pairs
        uint32_t rc = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = (uint32_t)it % U_STAGES, sph = ((uint32_t)it / U_STAGES) & 1u;
            mbar_wait(full + s, sph);
            const float* box = reinterpret_cast<const float*>(stage0 + (size_t)s * g.stage_bytes) + jj + 3;
            const int pstride = HH * U_BOXW;
            for (int rp = 0; rp < HH; ++rp, ++rc) {
                if ((int)(rc % U_CSETS) != cset) continue;
                const uint32_t slot = rc % U_RING, ph = (rc / U_RING) & 1u;
                uint32_t t[6][8];                              // A tiles (part * 3 + dxx): 8 words = 16 channels of one pixel
                const long long tc0 = timing ? clock64() : 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t hw = 0u, lw = 0u;
                    if (i < NP && g.debug_skip != 1) {
                        const float* bp = box + rp * U_BOXW + (2 * i) * pstride;
                        const float v0 = bp[0];
                        const float v1 = 2 * i + 1 < K ? bp[pstride] : 0.f;
                        split2(v0, v1, hw, lw);
                    }
                    t[0][i] = hw;
                    t[3][i] = lw;
                    if (i < NP) {
                        t[1][i] = __shfl_down_sync(0xffffffffu, hw, 1);
                        t[2][i] = __shfl_down_sync(0xffffffffu, hw, 2);
                        t[4][i] = __shfl_down_sync(0xffffffffu, lw, 1);
                        t[5][i] = __shfl_down_sync(0xffffffffu, lw, 2);
                    } else {
                        t[1][i] = 0u; t[2][i] = 0u; t[4][i] = 0u; t[5][i] = 0u;
                    }
                }
                if (timing) tph[0] += clock64() - tc0;
                CY_T(1, mbar_wait(cempty + slot, ph ^ 1u));
                const long long tc1 = timing ? clock64() : 0;
                tc_fence_after();
                const uint32_t a = tmem + ((uint32_t)(quarter * 32) << 16) + slot * U_SLOT + 48;
#pragma unroll
                for (int j = 0; j < 6; ++j) tmem_st_32x8(a + j * 8, t[j]);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(cfull + slot);
                if (timing) tph[2] += clock64() - tc1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (timing && warp == U_CONV0 && lane == 0) { dbg[4] = tph[0]; dbg[5] = tph[1]; dbg[6] = tph[2]; }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue
        constexpr int HC = (KC + U_ESETS - 1) / U_ESETS;
        const int c0 = ((warp - U_EPI0) >> 2) * HC;
        const int quarter = warp & 3, px = quarter * U_QPX + lane;
        uint32_t rc = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
            const int side = tile & 1, u = tile >> 1;
            const int b = u / per_image, trem = u % per_image;
            const int h0 = (trem / g.tiles_w) * g.TH, w0 = (trem % g.tiles_w) * U_TWO;
            const size_t plane = (size_t)g.H * g.W;
            float* out = (side ? dx_out : dy_out) + ((size_t)b * K + c0) * plane + (size_t)h0 * g.W + w0 + px;
            const bool col_ok = lane < U_QPX && w0 + px < g.W && g.debug_skip != 2;
            float p1[HC], p2[HC];
#pragma unroll
            for (int o = 0; o < HC; ++o) { p1[o] = 0.f; p2[o] = 0.f; }
            // software pipeline: the TMEM loads of box row rp+1 are in flight while row rp is combined and stored
            constexpr int VN = HC < 8 ? 8 : HC;
            uint32_t v[3][VN], nv[3][VN];
            auto fetch = [&](uint32_t rcx, uint32_t (&dst)[3][VN]) {
                const uint32_t slot = rcx % U_RING, ph = (rcx / U_RING) & 1u;
                CY_T(0, mbar_wait(afull + slot, ph));
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + slot * U_SLOT + c0;
#pragma unroll
                for (int dyy = 0; dyy < 3; ++dyy) tmem_ld_cols<HC>(taddr + dyy * 16, dst[dyy]);
            };
            auto release = [&](uint32_t rcx) {
                const long long te0 = timing ? clock64() : 0;
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(aempty + rcx % U_RING);
                if (timing) tph[1] += clock64() - te0;
            };
            fetch(rc, v);
            release(rc);
            for (int rp = 0; rp < HH; ++rp, ++rc) {
                const bool more = rp + 1 < HH;
                const long long te1 = timing ? clock64() : 0;
                if (more) fetch(rc + 1, nv);
                // box row rp feeds output rows rp (dy 0), rp-1 (dy 1) and rp-2 (dy 2, which completes it)
                const int r = rp - 2;
                float fin[HC];
#pragma unroll
                for (int o = 0; o < HC; ++o) {
                    fin[o] = p2[o] + __uint_as_float(v[2][o]);
                    p2[o] = p1[o] + __uint_as_float(v[1][o]);
                    p1[o] = __uint_as_float(v[0][o]);
                }
                if (r >= 0 && r < g.TH && h0 + r < g.H) {              // warp-uniform
                    if (col_ok) {
                        float* p = out + (size_t)r * g.W;
#pragma unroll
                        for (int o = 0; o < HC; ++o) {
                            if (c0 + o < K) *p = fin[o];
                            p += plane;
                        }
                    }
                }
                if (timing) tph[2] += clock64() - te1;
                if (more) {
                    release(rc + 1);
#pragma unroll
                    for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
                        for (int o = 0; o < HC; ++o) v[dyy][o] = nv[dyy][o];
                }
            }
        }
        if (timing && warp == U_EPI0 && lane == 0) { dbg[8] = tph[0]; dbg[9] = tph[1]; dbg[10] = tph[2]; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

int make_map_box(CUtensorMap* m, const void* base, int B, int K, int H, int W, int box_w, int box_h) {
    EncodeTiledFn fn = tensor_map_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CY_ERR_DEVICE; }
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)K * (cuuint64_t)B};
    cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)K};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(adjoint box) failed (%d)", (int)r); return CY_ERR_ARG; }
    return CY_OK;
}

}  // namespace

// returns CY_ERR_UNSUPPORTED when the shape is not eligible (the caller then takes the mma.sync / CUDA-core kernels)
int iic_bwd_umma(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
                 const float* gscale, void* dx, void* dy, cudaStream_t st) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (dtype != CY_F32 || pad != 1 || (W % 4) != 0 || K > 16 || W < U_BOXW || H < 16 || !al16(x) || !al16(y) || !al16(dx) || !al16(dy))
        return CY_ERR_UNSUPPORTED;
    UmmaGeom g;
    g.B = B; g.K = K; g.H = H; g.W = W;
    const size_t fixed = (size_t)2 * U_WSIDE_BYTES + U_NBAR * 8 + 64 + 1024;
    const size_t budget = 227 * 1024;
    // tallest tile that fits (fewest halo rows), ties broken by the fewest staged rows over the image
    int best_th = 0;
    long best_rows = 0;
    for (int th = 14; th >= 4; --th) {
        const size_t stage = ((size_t)K * (th + 2) * U_BOXW * 4 + 127) & ~(size_t)127;
        if (fixed + U_STAGES * stage > budget) continue;
        const long rows = (long)((H + th - 1) / th) * (th + 2);
        if (!best_th || rows < best_rows) { best_th = th; best_rows = rows; }
    }
    if (!best_th) return CY_ERR_UNSUPPORTED;
    g.TH = best_th;
    g.HH = best_th + 2;
    g.stage_bytes = (int)(((size_t)K * g.HH * U_BOXW * 4 + 127) & ~(size_t)127);
    g.tiles_h = (H + g.TH - 1) / g.TH;
    g.tiles_w = (W + U_TWO - 1) / U_TWO;
    g.n_tiles = 2 * B * g.tiles_h * g.tiles_w;
    {
        static int v = -1;
        if (v < 0) { const char* e = getenv("CY_IIC_DEBUG_SKIP"); v = e ? atoi(e) : 0; }
        g.debug_skip = v;
    }
    const size_t smem = fixed + (size_t)U_STAGES * g.stage_bytes;
    CUtensorMap tmx, tmy;
    int rc = make_map_box(&tmx, x, B, K, H, W, U_BOXW, g.HH);
    if (rc) return rc;
    rc = make_map_box(&tmy, y, B, K, H, W, U_BOXW, g.HH);
    if (rc) return rc;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaSuccess;
#define CY_ATTR(KCV) if (e == cudaSuccess) e = cudaFuncSetAttribute(iic_bwd_umma_kernel<KCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        CY_ATTR(4); CY_ATTR(6); CY_ATTR(8); CY_ATTR(10); CY_ATTR(12); CY_ATTR(16);
#undef CY_ATTR
        if (e != cudaSuccess) { set_error("iic_bwd_umma smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr_smem = smem;
    }
    const int grid = g.n_tiles < sms ? g.n_tiles : sms;
    float* dxf = reinterpret_cast<float*>(dx);
    float* dyf = reinterpret_cast<float*>(dy);
    static long long* dbg = nullptr;
    if (g.debug_skip == 9 && !dbg) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
#define CY_GO(KCV) iic_bwd_umma_kernel<KCV><<<grid, U_THREADS, smem, st>>>(tmx, tmy, g, djoint, gscale, dxf, dyf, dbg)
    if (K <= 4) CY_GO(4);
    else if (K <= 6) CY_GO(6);
    else if (K <= 8) CY_GO(8);
    else if (K <= 10) CY_GO(10);
    else if (K <= 12) CY_GO(12);
    else CY_GO(16);
#undef CY_GO
    CY_CHECK_LAUNCH("iic_bwd_umma");
    if (dbg) {
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const double rows = (double)((g.n_tiles + grid - 1) / grid) * g.HH;
        fprintf(stderr, "[umma phases, cycles per box row of CTA 0] issuer 0 (own rows x2): wait D free %.0f, wait A %.0f, issue %.0f | converter (own rows x%d): "
                "convert %.0f, wait A free %.0f, store %.0f | epilogue: wait D %.0f, ld-wait+release %.0f, fetch+combine+store %.0f\n",
                h[0] / rows * U_NISS, h[1] / rows * U_NISS, h[2] / rows * U_NISS, U_CSETS, h[4] / rows * U_CSETS, h[5] / rows * U_CSETS, h[6] / rows * U_CSETS,
                h[8] / rows, h[9] / rows, h[10] / rows);
    }
    return CY_OK;
}

}  // namespace cy
