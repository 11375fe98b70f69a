// Push all-gather over NVLink peer memory: every rank copies the byte ranges it owns from its local symmetric buffer to the
// same offsets of all other ranks' buffers with plain 16-byte stores through the peer mappings.  Used in place of the NCCL
// all-gathers of the row-sharded InfoNCE step (embeddings + labels: one launch; row statistics: one launch) and of the
// all-reduce of the IIC joint (each rank pushes its partial joint into its slot; cy_iic_epilogue sums the slots).  The
// caller's signal-pad barrier after the launch publishes the stores (include/contrastyou_b200.h, cy_p2p_push) — or, with
// cy_p2p_push_barrier, the kernel itself: its last block signals every peer (release store of the call's epoch into the peer's
// flag array, which lives in the same symmetric buffer) and spins until every peer's flag has reached the epoch.  That is one
// launch per exchange instead of two and no dependence on another library's barrier kernel.
#include "common.cuh"

namespace cy {

struct PushRanges {
    unsigned long long off[4], bytes[4];
    int n;
};

// grid (blocks, world - 1): blockIdx.y picks the destination rank; a block walks the 16-byte chunks of all ranges
__global__ void __launch_bounds__(256) p2p_push_kernel(void* const* __restrict__ peer_bufs, int world, int rank, PushRanges r) {
    const int peer = ((int)blockIdx.y + rank + 1) % world;          // start with the right-hand neighbour: spreads the links
    const uint8_t* src = reinterpret_cast<const uint8_t*>(peer_bufs[rank]);
    uint8_t* dst = reinterpret_cast<uint8_t*>(peer_bufs[peer]);
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthr = (unsigned long long)gridDim.x * blockDim.x;
    for (int k = 0; k < r.n; ++k) {
        const uint4* s = reinterpret_cast<const uint4*>(src + r.off[k]);
        uint4* d = reinterpret_cast<uint4*>(dst + r.off[k]);
        const unsigned long long n16 = r.bytes[k] >> 4;
        for (unsigned long long i = tid; i < n16; i += nthr) d[i] = s[i];
    }
}

// flags: uint32 [world] at byte offset flag_off of EVERY rank's buffer, flags_of_rank_p[r] = last epoch rank r has published to
// p; counter: one device int (zero between calls) that elects the last block of the grid.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) p2p_push_barrier_kernel(void* const* __restrict__ peer_bufs, int world, int rank, PushRanges r,
                                                               unsigned long long flag_off, unsigned int* counter, uint32_t epoch) {
    const int peer = ((int)blockIdx.y + rank + 1) % world;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(peer_bufs[rank]);
    uint8_t* dst = reinterpret_cast<uint8_t*>(peer_bufs[peer]);
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthr = (unsigned long long)gridDim.x * blockDim.x;
    for (int k = 0; k < r.n; ++k) {
        const uint4* s = reinterpret_cast<const uint4*>(src + r.off[k]);
        uint4* d = reinterpret_cast<uint4*>(dst + r.off[k]);
        const unsigned long long n16 = r.bytes[k] >> 4;
        for (unsigned long long i = tid; i < n16; i += nthr) d[i] = s[i];
    }
    // grid-wide completion: every thread's stores are ordered system-wide before its block is counted
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence_system();
    // one lane per peer: publish, then wait for the peer (epochs only grow; the difference test survives the 32-bit wrap)
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const int p = (int)threadIdx.x;
        st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(peer_bufs[p]) + flag_off) + rank, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(peer_bufs[rank]) + flag_off) + p;
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        }
    }
    __syncthreads();
    __threadfence_system();
}

int p2p_push_barrier(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges,
                     unsigned long long flag_off, unsigned int* counter, unsigned int epoch, cudaStream_t st) {
    if (world == 1) return CY_OK;
    CY_CHECK_ARG(world <= 256 && counter != nullptr && (flag_off & 15) == 0, "cy_p2p_push_barrier: world <= 256, counter and a 16-byte aligned flag offset are required");
    PushRanges r;
    r.n = n_ranges;
    unsigned long long total = 0;
    for (int k = 0; k < n_ranges; ++k) {
        r.off[k] = ranges[2 * k];
        r.bytes[k] = ranges[2 * k + 1];
        CY_CHECK_ARG((r.off[k] & 15) == 0 && (r.bytes[k] & 15) == 0, "cy_p2p_push_barrier: ranges must be multiples of 16 bytes");
        total += r.bytes[k];
    }
    unsigned long long want = (total + 4095) / 4096;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 64 ? 64 : want));
    dim3 grid(blocks, (unsigned)(world - 1));
    p2p_push_barrier_kernel<<<grid, 256, 0, st>>>(peer_bufs, world, rank, r, flag_off, counter, epoch);
    CY_CHECK_LAUNCH("p2p_push_barrier");
    return CY_OK;
}

int p2p_push(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges, cudaStream_t st) {
    if (world == 1) return CY_OK;
    PushRanges r;
    r.n = n_ranges;
    unsigned long long total = 0;
    for (int k = 0; k < n_ranges; ++k) {
        r.off[k] = ranges[2 * k];
        r.bytes[k] = ranges[2 * k + 1];
        CY_CHECK_ARG((r.off[k] & 15) == 0 && (r.bytes[k] & 15) == 0, "cy_p2p_push: ranges must be multiples of 16 bytes");
        total += r.bytes[k];
    }
    // enough CTAs per destination to keep its link busy, not more than the copy has 4 KB pieces
    unsigned long long want = (total + 4095) / 4096;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 64 ? 64 : want));
    dim3 grid(blocks, (unsigned)(world - 1));
    p2p_push_kernel<<<grid, 256, 0, st>>>(peer_bufs, world, rank, r);
    CY_CHECK_LAUNCH("p2p_push");
    return CY_OK;
}

}  // namespace cy
