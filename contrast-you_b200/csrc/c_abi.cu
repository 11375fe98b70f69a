// extern "C" surface of libcontrastyou_b200.so (see include/contrastyou_b200.h for the contract of every call).
#include <stdarg.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace cy {

// NVTX range around every compute entry point (header-only NVTX v3: a no-op unless a profiler is attached), so that an
// nsys / ncu timeline shows the loss path by C-ABI call and the kernels nest under it.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define CY_NVTX(name) cy::NvtxRange nvtx_range__(name)

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev;
}

int device_sm_count() {
    static int sms[CY_MAX_DEVICES] = {};
    const int dev = current_device() % CY_MAX_DEVICES;
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}

// infonce_simt.cu
int infonce_fwd_simt(const void*, int, int64_t, int64_t, int64_t, const int32_t*, const uint8_t*, int64_t, int64_t, float, int,
                     int, float, float*, cudaStream_t);
int infonce_bwd_simt(const void*, int, int64_t, int64_t, int64_t, const int32_t*, const uint8_t*, int64_t, int64_t, float, int,
                     float, const float*, const float*, void*, int64_t, cudaStream_t);
int infonce_masks(int64_t, const int32_t*, const uint8_t*, float*, float*, cudaStream_t);
int labels_canonicalize(const void*, int, int64_t, int32_t*, int32_t*, cudaStream_t);
int infonce_pack(const void*, const void*, int, int64_t, int64_t, int64_t, int64_t, const int64_t*, void*, int*, float*, const int64_t*, int64_t,
                 cudaStream_t);
int infonce_pack_split(const void*, const void*, int64_t, int64_t, int64_t, int64_t, const int64_t*, void*, int*, cudaStream_t);
int infonce_unpack(const void*, int, int64_t, int64_t, int64_t, const int64_t*, void*, void*, const void*, const float*, const float*,
                   const int64_t*, int64_t, cudaStream_t);
// infonce_tc.cu
bool infonce_tc_supported(int dtype, int64_t N, int64_t d, int64_t ldz, const uint8_t* codes, int variant);
size_t infonce_tc_workspace_bytes(int64_t N, int64_t d, bool split);
size_t infonce_loss_workspace_bytes(int64_t N);
int infonce_fwd_tc(const void*, int, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, float, int, float*, float*, void*, size_t,
                   cudaStream_t);
int infonce_fwd2_tc(const void*, int, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, float, int, float, float*, float*, void*,
                    size_t, cudaStream_t);
int infonce_rowstats(int64_t, int64_t, int64_t, float, int, int, float*, float*, cudaStream_t);
int infonce_loss(int64_t, int, const float*, float*, int32_t*, const int32_t*, void*, size_t, cudaStream_t);
int infonce_bwd_tc(const void*, int, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, float, int, float, const float*,
                   const float*, void*, int64_t, void*, size_t, cudaStream_t);
// iic.cu
size_t iic_workspace_bytes(int, int, int, int, int);
int iic_joint(const void*, const void*, int, int, int, int, int, int, double*, void*, size_t, cudaStream_t);
size_t iic_epilogue_workspace_bytes(int, int);
int iic_epilogue(const double*, int, int, int, int, float, float, double, float*, float*, float*, float*, void*, size_t, cudaStream_t);
// imsat.cu
size_t imsat_workspace_bytes(int);
int imsat_fwd(const void*, int, int64_t, int, int64_t, float, float*, float*, void*, size_t, cudaStream_t);
int imsat_bwd(const void*, int, int64_t, int, int64_t, float, const float*, const float*, void*, cudaStream_t);
// p2p.cu
int p2p_push(void* const*, int, int, const unsigned long long*, int, cudaStream_t);
int p2p_push_barrier(void* const*, int, int, const unsigned long long*, int, unsigned long long, unsigned int*, unsigned int, cudaStream_t);
int iic_bwd(const void*, const void*, int, int, int, int, int, int, const float*, const float*, void*, void*, cudaStream_t);
int softmax_t_fwd(const void* const*, void* const*, int, int, int, int, int, int, float, cudaStream_t);
int iic_joint_heads(const void* const*, const void* const*, int, int, int, int, int, int, int, double*, long long, void*, size_t,
                    cudaStream_t);
int iic_epilogue_heads(const double*, long long, long long, int, int, int, int, int, float, float, double, float*, float*, float*, long long,
                       void*, size_t, cudaStream_t);
int iic_bwd_heads(const void* const*, const void* const*, int, int, int, int, int, int, int, const float*, long long, const float*,
                  void* const*, void* const*, float, cudaStream_t);

static int check_infonce_args(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                              const uint8_t* codes, int64_t row_begin, int64_t row_end, int variant) {
    CY_CHECK_ARG(z != nullptr, "z is null");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16 || dtype == CY_F32_SPLIT, "unknown dtype %d", dtype);
    CY_CHECK_ARG(N >= 2 && (N % 2) == 0, "N=%lld must be even (two stacked views)", (long long)N);
    CY_CHECK_ARG(d >= 1 && ldz >= d, "d=%lld ldz=%lld", (long long)d, (long long)ldz);
    CY_CHECK_ARG(labels != nullptr || codes != nullptr, "need labels or codes");
    CY_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "rows [%lld,%lld) outside [0,%lld)",
                 (long long)row_begin, (long long)row_end, (long long)N);
    CY_CHECK_ARG(variant >= CY_SUPCON && variant <= CY_SELFPACED_SOFT, "unknown variant %d", variant);
    return CY_OK;
}

// 0 = simt, 1 = tcgen05; negative = error
static int resolve_path(int path, int dtype, int64_t N, int64_t d, int64_t ldz, const uint8_t* codes, int variant,
                        int64_t row_begin, int64_t row_end) {
    // the tensor kernels work on 128-row blocks: a row range that does not start on a multiple of 128 and end on one (or at
    // N) — row-sharded ranks with an odd local batch — takes the CUDA-core kernels under CY_PATH_AUTO instead of failing;
    // forward, pass 2 and backward see the same arguments and therefore take the same path
    const bool rows_ok = (row_begin % 128) == 0 && ((row_end % 128) == 0 || row_end == N);
    const bool tc_ok = rows_ok && infonce_tc_supported(dtype, N, d, ldz, codes, variant);
    if (path == CY_PATH_TCGEN05) {
        if (!tc_ok) {
            set_error("tcgen05 path needs bf16 / fp16, d in {128, 256}, N >= 256, a 128-aligned row range and label masks "
                      "(got dtype=%d N=%lld d=%lld variant=%d rows=[%lld,%lld) codes=%d)",
                      dtype, (long long)N, (long long)d, variant, (long long)row_begin, (long long)row_end, codes != nullptr);
            return CY_ERR_UNSUPPORTED;
        }
        return 1;
    }
    if (path == CY_PATH_AUTO && tc_ok && N >= 1024) return 1;
    if (dtype == CY_F32_SPLIT) {
        set_error("CY_F32_SPLIT rows are a tensor-path format (d in {128, 256}, N >= 256, 128-aligned row range, label masks)");
        return CY_ERR_UNSUPPORTED;
    }
    if (d > 256) {
        set_error("SIMT path supports d <= 256 (got %lld)", (long long)d);
        return CY_ERR_UNSUPPORTED;
    }
    return 0;
}

}  // namespace cy

using namespace cy;

extern "C" {

int cy_abi_version(void) { return CY_ABI_VERSION; }

const char* cy_last_error(void) { return g_err; }

int cy_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

unsigned long long cy_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

size_t cy_infonce_workspace_bytes(int64_t N, int64_t d, int dtype, int variant, int path) {
    (void)variant;
    const size_t loss = infonce_loss_workspace_bytes(N);
    if (path == CY_PATH_SIMT || dtype == CY_F32) return loss;
    const size_t tc = infonce_tc_workspace_bytes(N, d, dtype == CY_F32_SPLIT);
    return (tc > loss ? tc : loss) + 16;
}

int cy_infonce_fwd(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, const uint8_t* codes,
                   int64_t row_begin, int64_t row_end, float inv_t, int variant, int path, float* stats, float* xstat,
                   void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_infonce_fwd");
    int rc = check_infonce_args(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, variant);
    if (rc) return rc;
    CY_CHECK_ARG(stats != nullptr && xstat != nullptr, "stats / xstat is null");
    const int p = resolve_path(path, dtype, N, d, ldz, codes, variant, row_begin, row_end);
    if (p < 0) return p;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p == 1)
        return infonce_fwd_tc(z, dtype, N, d, ldz, labels, row_begin, row_end, inv_t, variant, stats, xstat, workspace, workspace_bytes, st);
    rc = infonce_fwd_simt(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, inv_t, variant, 1, 0.f, stats, st);
    if (rc) return rc;
    return infonce_rowstats(N, row_begin, row_end, inv_t, variant, 1, stats, xstat, st);
}

int cy_infonce_fwd_pass2(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                         const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant, float gamma,
                         int path, float* stats, float* xstat, void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_infonce_fwd_pass2");
    int rc = check_infonce_args(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, variant);
    if (rc) return rc;
    CY_CHECK_ARG(variant != CY_SUPCON, "CY_SUPCON has no second pass");
    CY_CHECK_ARG(stats != nullptr && xstat != nullptr, "stats / xstat is null");
    const int p = resolve_path(path, dtype, N, d, ldz, codes, variant, row_begin, row_end);
    if (p < 0) return p;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p == 1)
        return infonce_fwd2_tc(z, dtype, N, d, ldz, labels, row_begin, row_end, inv_t, variant, gamma, stats, xstat, workspace,
                               workspace_bytes, st);
    rc = infonce_fwd_simt(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, inv_t, variant, 2, gamma, stats, st);
    if (rc) return rc;
    return infonce_rowstats(N, row_begin, row_end, inv_t, variant, 2, stats, xstat, st);
}

int cy_infonce_loss(int64_t N, int variant, const float* xstat, float* out8, int32_t* bad_rows, const int32_t* overflow,
                    void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_infonce_loss");
    CY_CHECK_ARG(xstat && out8 && N >= 2, "null pointer");
    CY_CHECK_ARG(variant >= CY_SUPCON && variant <= CY_SELFPACED_SOFT, "unknown variant %d", variant);
    return infonce_loss(N, variant, xstat, out8, bad_rows, overflow, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_bwd(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels, const uint8_t* codes,
                   int64_t row_begin, int64_t row_end, float inv_t, int variant, float gamma, int path, const float* xstat,
                   const float* gscale, void* dz, int64_t lddz, void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_infonce_bwd");
    int rc = check_infonce_args(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, variant);
    if (rc) return rc;
    CY_CHECK_ARG(xstat && gscale && dz && lddz >= d, "null pointer or lddz < d");
    const int p = resolve_path(path, dtype, N, d, ldz, codes, variant, row_begin, row_end);
    if (p < 0) return p;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p == 1)
        return infonce_bwd_tc(z, dtype, N, d, ldz, labels, row_begin, row_end, inv_t, variant, gamma, xstat, gscale, dz, lddz, workspace,
                              workspace_bytes, st);
    return infonce_bwd_simt(z, dtype, N, d, ldz, labels, codes, row_begin, row_end, inv_t, variant, gamma, xstat, gscale, dz,
                            lddz, st);
}

int cy_infonce_masks(int64_t N, const int32_t* labels, const uint8_t* codes, float* pos_mask, float* neg_mask, void* stream) {
    CY_NVTX("cy_infonce_masks");
    CY_CHECK_ARG(N >= 2 && (N % 2) == 0 && (labels || codes), "bad arguments");
    return infonce_masks(N, labels, codes, pos_mask, neg_mask, reinterpret_cast<cudaStream_t>(stream));
}

int cy_labels_canonicalize(const void* src, int src_kind, int64_t n, int32_t* dst, int32_t* overflow, void* stream) {
    CY_NVTX("cy_labels_canonicalize");
    CY_CHECK_ARG(src && dst && n >= 1 && src_kind >= 0 && src_kind <= 2, "bad arguments");
    CY_CHECK_ARG(src_kind != 2 || overflow != nullptr, "int64 labels need the overflow counter");
    return labels_canonicalize(src, src_kind, n, dst, overflow, reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_pack(const void* f1, const void* f2, int dtype, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                    const int64_t* order, void* z, int32_t* bad_rows, float* inv_norm, void* stream) {
    CY_NVTX("cy_infonce_pack");
    CY_CHECK_ARG(f1 && f2 && z && n >= 1 && d >= 1 && ld1 >= d && ld2 >= d, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    return infonce_pack(f1, f2, dtype, n, d, ld1, ld2, order, z, bad_rows, inv_norm, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_pack_split(const void* f1, const void* f2, int64_t n, int64_t d, int64_t ld1, int64_t ld2, const int64_t* order,
                          void* zs, int32_t* bad_rows, void* stream) {
    CY_NVTX("cy_infonce_pack_split");
    CY_CHECK_ARG(f1 && f2 && zs && n >= 1 && d >= 1 && ld1 >= d && ld2 >= d, "bad arguments");
    return infonce_pack_split(f1, f2, n, d, ld1, ld2, order, zs, bad_rows, reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_pack_gather(const void* map1, const void* map2, int dtype, int64_t n, int64_t d, const int64_t* pix_off,
                           int64_t chan_stride, const int64_t* order, void* z, int32_t* bad_rows, void* stream) {
    CY_NVTX("cy_infonce_pack_gather");
    CY_CHECK_ARG(map1 && map2 && z && pix_off && n >= 1 && d >= 1 && chan_stride >= 1, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    return infonce_pack(map1, map2, dtype, n, d, 0, 0, order, z, bad_rows, nullptr, pix_off, chan_stride,
                        reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_unpack(const void* dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* order, void* g1,
                      void* g2, const void* z, const float* inv_norm, const float* gscale, void* stream) {
    CY_NVTX("cy_infonce_unpack");
    CY_CHECK_ARG(dz && g1 && g2 && n >= 1 && d >= 1 && lddz >= d, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    CY_CHECK_ARG((z == nullptr) == (inv_norm == nullptr), "z and inv_norm go together");
    return infonce_unpack(dz, dtype, n, d, lddz, order, g1, g2, z, inv_norm, gscale, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

int cy_infonce_unpack_scatter(const void* dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* order, void* gmap1,
                              void* gmap2, const int64_t* pix_off, int64_t chan_stride, const float* gscale, void* stream) {
    CY_NVTX("cy_infonce_unpack_scatter");
    CY_CHECK_ARG(dz && gmap1 && gmap2 && pix_off && n >= 1 && d >= 1 && lddz >= d && chan_stride >= 1, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    return infonce_unpack(dz, dtype, n, d, lddz, order, gmap1, gmap2, nullptr, nullptr, gscale, pix_off, chan_stride,
                          reinterpret_cast<cudaStream_t>(stream));
}

size_t cy_iic_workspace_bytes(int B, int K, int H, int W, int pad) {
    if (B < 1 || K < 1 || H < 1 || W < 1 || pad < 0) return 0;
    return iic_workspace_bytes(B, K, H, W, pad);
}

static int check_iic(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad) {
    CY_CHECK_ARG(x && y, "null input");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    CY_CHECK_ARG(B >= 1 && K >= 1 && H >= 1 && W >= 1, "bad shape [%d,%d,%d,%d]", B, K, H, W);
    CY_CHECK_ARG(pad >= 0, "negative padding %d", pad);
    return CY_OK;
}

int cy_iic_joint(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, double* joint, void* workspace,
                 size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_iic_joint");
    int rc = check_iic(x, y, dtype, B, K, H, W, pad);
    if (rc) return rc;
    CY_CHECK_ARG(joint != nullptr, "joint is null");
    return iic_joint(x, y, dtype, B, K, H, W, pad, joint, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

size_t cy_iic_epilogue_workspace_bytes(int K, int pad) {
    if (K < 1 || pad < 0) return 0;
    return iic_epilogue_workspace_bytes(K, pad);
}

int cy_iic_epilogue(const double* joint, int n_slots, int K, int pad, int symmetric, float lamda, float eps, double n_pixels,
                    float* loss, float* p00, float* p_ij, float* djoint, void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_iic_epilogue");
    CY_CHECK_ARG(joint && loss && p00 && K >= 1 && pad >= 0 && n_slots >= 1, "bad arguments");
    return iic_epilogue(joint, n_slots, K, pad, symmetric, lamda, eps, n_pixels, loss, p00, p_ij, djoint, workspace, workspace_bytes,
                        reinterpret_cast<cudaStream_t>(stream));
}

size_t cy_imsat_workspace_bytes(int K) { return K >= 1 ? imsat_workspace_bytes(K) : 0; }

int cy_imsat_fwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, float* out2, float* q, void* workspace,
                 size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_imsat_fwd");
    CY_CHECK_ARG(pred && out2 && q && N >= 1 && S >= 1, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    return imsat_fwd(pred, dtype, N, K, S, eps, out2, q, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int cy_imsat_bwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, const float* q, const float* g2, void* grad,
                 void* stream) {
    CY_NVTX("cy_imsat_bwd");
    CY_CHECK_ARG(pred && q && g2 && grad && N >= 1 && S >= 1, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    return imsat_bwd(pred, dtype, N, K, S, eps, q, g2, grad, reinterpret_cast<cudaStream_t>(stream));
}

int cy_p2p_push(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges, void* stream) {
    CY_NVTX("cy_p2p_push");
    CY_CHECK_ARG(peer_bufs && ranges && world >= 1 && rank >= 0 && rank < world && n_ranges >= 1 && n_ranges <= 4, "bad arguments");
    return p2p_push(peer_bufs, world, rank, ranges, n_ranges, reinterpret_cast<cudaStream_t>(stream));
}

int cy_p2p_push_barrier(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges,
                        unsigned long long flag_off, unsigned int* counter, unsigned int epoch, void* stream) {
    CY_NVTX("cy_p2p_push_barrier");
    CY_CHECK_ARG(peer_bufs && ranges && world >= 1 && rank >= 0 && rank < world && n_ranges >= 1 && n_ranges <= 4, "bad arguments");
    return p2p_push_barrier(peer_bufs, world, rank, ranges, n_ranges, flag_off, counter, epoch, reinterpret_cast<cudaStream_t>(stream));
}

int cy_iic_bwd(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
               const float* gscale, void* dx, void* dy, void* stream) {
    CY_NVTX("cy_iic_bwd");
    int rc = check_iic(x, y, dtype, B, K, H, W, pad);
    if (rc) return rc;
    CY_CHECK_ARG(djoint && gscale && dx && dy, "null pointer");
    return iic_bwd(x, y, dtype, B, K, H, W, pad, djoint, gscale, dx, dy, reinterpret_cast<cudaStream_t>(stream));
}

int cy_iic_joint_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                       double* joint, long long joint_stride, void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_iic_joint_heads");
    CY_CHECK_ARG(xs && ys && n_heads >= 1 && joint && joint_stride >= 0, "bad arguments");
    for (int s = 0; s < n_heads; ++s) {
        const int rc = check_iic(xs[s], ys[s], dtype, B, K, H, W, pad);
        if (rc) return rc;
    }
    return iic_joint_heads(xs, ys, n_heads, dtype, B, K, H, W, pad, joint, joint_stride, workspace, workspace_bytes,
                           reinterpret_cast<cudaStream_t>(stream));
}

int cy_iic_epilogue_heads(const double* joint, long long joint_stride, long long slot_stride, int n_heads, int n_slots, int K, int pad,
                          int symmetric, float lamda, float eps, double n_pixels, float* loss, float* p00, float* djoint,
                          long long out_stride, void* workspace, size_t workspace_bytes, void* stream) {
    CY_NVTX("cy_iic_epilogue_heads");
    CY_CHECK_ARG(joint && loss && p00 && n_heads >= 1 && n_slots >= 1 && K >= 1 && pad >= 0 && n_pixels > 0 && slot_stride >= 0,
                 "bad arguments");
    return iic_epilogue_heads(joint, joint_stride, slot_stride, n_heads, n_slots, K, pad, symmetric, lamda, eps, n_pixels, loss, p00, djoint,
                              out_stride, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int cy_iic_bwd_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                     const float* djoint, long long djoint_stride, const float* gscale, void* const* dxs, void* const* dys,
                     void* stream) {
    CY_NVTX("cy_iic_bwd_heads");
    CY_CHECK_ARG(xs && ys && dxs && dys && n_heads >= 1 && djoint && gscale, "bad arguments");
    for (int s = 0; s < n_heads; ++s) {
        const int rc = check_iic(xs[s], ys[s], dtype, B, K, H, W, pad);
        if (rc) return rc;
        CY_CHECK_ARG(dxs[s] && dys[s], "null pointer");
    }
    return iic_bwd_heads(xs, ys, n_heads, dtype, B, K, H, W, pad, djoint, djoint_stride, gscale, dxs, dys, 0.f,
                         reinterpret_cast<cudaStream_t>(stream));
}

int cy_iic_bwd_logits_heads(const void* const* pxs, const void* const* pys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                            const float* djoint, long long djoint_stride, const float* gscale, float T, void* const* dlxs,
                            void* const* dlys, void* stream) {
    CY_NVTX("cy_iic_bwd_logits_heads");
    CY_CHECK_ARG(pxs && pys && dlxs && dlys && n_heads >= 1 && djoint && gscale, "bad arguments");
    CY_CHECK_ARG(T > 0.f && isfinite(T), "temperature %g", (double)T);
    for (int s = 0; s < n_heads; ++s) {
        const int rc = check_iic(pxs[s], pys[s], dtype, B, K, H, W, pad);
        if (rc) return rc;
        CY_CHECK_ARG(dlxs[s] && dlys[s] && dlxs[s] != pxs[s] && dlys[s] != pys[s], "null or aliased gradient pointer");
    }
    return iic_bwd_heads(pxs, pys, n_heads, dtype, B, K, H, W, pad, djoint, djoint_stride, gscale, dlxs, dlys, 1.0f / T,
                         reinterpret_cast<cudaStream_t>(stream));
}

int cy_softmax_t_fwd(const void* const* logits, void* const* probs, int n_maps, int dtype, int B, int K, int H, int W, float T,
                     void* stream) {
    CY_NVTX("cy_softmax_t_fwd");
    CY_CHECK_ARG(logits && probs && n_maps >= 1, "bad arguments");
    CY_CHECK_ARG(dtype == CY_F32 || dtype == CY_BF16 || dtype == CY_F16, "unknown dtype %d", dtype);
    CY_CHECK_ARG(B >= 1 && K >= 1 && H >= 1 && W >= 1, "bad shape [%d,%d,%d,%d]", B, K, H, W);
    CY_CHECK_ARG(T > 0.f && isfinite(T), "temperature %g", (double)T);
    for (int i = 0; i < n_maps; ++i) CY_CHECK_ARG(logits[i] && probs[i], "null map %d", i);
    return softmax_t_fwd(logits, probs, n_maps, dtype, B, K, H, W, T, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
