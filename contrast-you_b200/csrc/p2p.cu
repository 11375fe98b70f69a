// Push all-gather over NVLink peer memory: every rank copies the byte ranges it owns from its local symmetric buffer to the
// same offsets of all other ranks' buffers with plain 16-byte stores through the peer mappings.  Used in place of the NCCL
// all-gathers of the row-sharded InfoNCE step (embeddings + labels: one launch; row statistics: one launch) and of the
// all-reduce of the IIC joint (each rank pushes its partial joint into its slot; cy_iic_epilogue sums the slots).  The
// caller's signal-pad barrier after the launch publishes the stores (include/contrastyou_b200.h, cy_p2p_push).
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
