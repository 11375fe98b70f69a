// placeholder until the tcgen05 kernels land (next commit): reports "unsupported" so AUTO resolves to SIMT
#include "common.cuh"
namespace cy {
bool infonce_tc_supported(int, int64_t, int64_t, int64_t, const uint8_t*, int) { return false; }
size_t infonce_tc_workspace_bytes(int64_t, int64_t) { return 0; }
int infonce_fwd_tc(const void*, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, float, float*, void*, size_t,
                   cudaStream_t) { set_error("tcgen05 path not built"); return CY_ERR_UNSUPPORTED; }
int infonce_bwd_tc(const void*, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, float, const float*, const float*,
                   void*, int64_t, void*, size_t, cudaStream_t) { set_error("tcgen05 path not built"); return CY_ERR_UNSUPPORTED; }
}
