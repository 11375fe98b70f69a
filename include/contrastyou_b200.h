/*
 * contrastyou_b200.h — C ABI of libcontrastyou_b200.so: the sm_100a kernels behind the drop-in InfoNCE / IIC
 * loss modules of Contrast-You.
 *
 * The reference is pure Python/PyTorch: it has no FFI of its own.  Each entry point below replaces the chain of
 * ATen calls that a reference function lowers to (cited as file:line under /root/reference); the binding a
 * maintainer adds on the reference side is a ctypes stub inside a torch.autograd.Function.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (torch allocator) unless
 *     marked [host]; no allocation, no host synchronisation and no host<->device copy happens inside these calls;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 on success, CY_ERR_* (negative) on invalid arguments, or the positive cudaError_t of a failed
 *     launch; cy_last_error() returns a thread-local description of the last failure;
 *   - dtype codes describe the element type of embeddings / probability maps; accumulation is always fp32.
 */
#ifndef CONTRASTYOU_B200_H
#define CONTRASTYOU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CY_ABI_VERSION 7

/* element types of embeddings / probability maps */
#define CY_F32 0
#define CY_BF16 1
#define CY_F16 2
/* fp32 embeddings split for the tensor kernels: z is [N, 2 d] bf16, row = [hi (d) | lo (d)] with fp32 value = hi + lo
 * (cy_infonce_pack_split writes it); products hi.hi + hi.lo + lo.hi are good to 2^-16, the gradient dz is fp32 [N, d].
 * Accepted by cy_infonce_fwd / _fwd_pass2 / _bwd on the tcgen05 path only (d in {128, 256}). */
#define CY_F32_SPLIT 3

/* InfoNCE variants (cy_infonce_*: `variant`) */
#define CY_SUPCON 0         /* SupConLoss1, exclude_other_pos=False      contrastive.py:92            */
#define CY_SUPCON_EXCLUDE 1 /* SupConLoss1, exclude_other_pos=True       contrastive.py:87-90         */
#define CY_SELFPACED_HARD 2 /* SelfPacedSupConLoss, weight_update="hard" contrastive.py:197-204       */
#define CY_SELFPACED_SOFT 3 /* SelfPacedSupConLoss, weight_update="soft"                              */

/* kernel family selection (cy_infonce_*: `path`) */
#define CY_PATH_AUTO 0
#define CY_PATH_SIMT 1    /* fp32 CUDA-core tiles: every variant, every mask source, any dtype, d <= 256               */
#define CY_PATH_TCGEN05 2 /* TMA -> smem -> tcgen05.mma -> TMEM: every variant, label masks, bf16 / fp16, d in {128,256}, */
                          /* N >= 256 (ragged N allowed), row range starting on a multiple of 128 and ending on one or at N */

/* error codes */
#define CY_OK 0
#define CY_ERR_ARG (-1)         /* inconsistent sizes / null pointers                                  */
#define CY_ERR_UNSUPPORTED (-2) /* shape or variant outside what the selected path implements          */
#define CY_ERR_DEVICE (-3)      /* not an sm_100 device                                                */

int cy_abi_version(void);
const char* cy_last_error(void);
/* number of SMs of the current device (grid sizing of persistent kernels); <0 on error */
int cy_device_sm_count(void);
/* number of kernels this library has launched in the calling process so far (every launch site counts itself);
 * monotonically increasing, used by bench.py to report `gpu_launches` as a measured difference */
unsigned long long cy_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * InfoNCE / SupCon family.   Replaces exp_sim_temperature + SupConLoss1._forward / SelfPacedSupConLoss._forward
 * (contrastyou/losses/contrastive.py:14-20, :51-100, :136-204) and their autograd backward.
 *
 *   z        [N, d] embeddings, both views stacked (rows [0,n) = proj_feat1, [n,2n) = proj_feat2, N = 2n), row
 *            stride ldz elements, L2-normalised rows (asserted by the caller like contrastive.py:58)
 *   labels   [N] int32, tiled over the views (labels[i+n] == labels[i]); P_ij = [labels_i == labels_j][i != j],
 *            Neg_ij = [labels_i != labels_j]  (contrastive.py:38-44, :62-71).  SimCLR (:45-48) = arange(n) tiled.
 *   codes    optional [n, n] uint8 for the explicit mask= path (:33-36): 1 = positive, 0 = negative, anything else
 *            = neither; looked up as codes[(i % n) * n + (j % n)], diagonal i == j always cleared.  When non-NULL
 *            it overrides labels.
 *   rows     the call covers rows [row_begin, row_end) of the N x N problem against all N columns (single GPU:
 *            0..N; row-sharded multi-GPU: the rank's block).
 *
 * Forward = cy_infonce_fwd (+ cy_infonce_fwd_pass2 for the variants that need the row sums first), then cy_infonce_loss.
 * Two caller-owned arrays carry the row statistics:
 *   stats  [CY_NSTAT][N] float — raw per-row sums of the sweeps (scratch of the forward; owned rows only);
 *   xstat  [N][4] float (16-byte aligned) — per row, everything the BACKWARD needs when it meets the row as a row or as a
 *          column, plus the row's loss term.  cy_infonce_fwd / _pass2 fill the owned rows; in the row-sharded multi-GPU form
 *          ONE all-gather of the owned rows of xstat is the whole exchange (the loss is then reduced from the gathered array
 *          by every rank, identically: no scalar collective).  Slots (CY_XS_*):
 *            variant            [0]                 [1]     [2]                          [3]
 *            SUPCON             log-denominator     1/c_i   1/den_i                      loss term of row i
 *            SELFPACED_*        log-denominator     1/c_i   sw_i/(c_i den_i)             loss term of row i
 *            SUPCON_EXCLUDE     loss term of row i  1/c_i   u_i/(c_i (ratio_i+1e-4))     A_i = neg_sum_i/(ratio_i+1e-4)
 * ---------------------------------------------------------------------------------------------------------------- */
#define CY_NSTAT 8
#define CY_STAT_LOGDEN 0 /* log(sum_j M_ij E_ij + 1e-16), E = exp(S - 1/t), M = P + Neg                      */
#define CY_STAT_INVC 1   /* 1 / c_i,  c_i = sum_j P_ij                                                       */
#define CY_STAT_COEF 2   /* coefficient of E_ij in dL/dS_ij for row i (1/den_i, sw_i/(c_i den_i), ...)       */
#define CY_STAT_AUX 3    /* pass 1: sum_j Neg_ij E_ij; exclude -> A_i = neg_sum_i/(ratio_i+1e-4)             */
#define CY_STAT_POSL 4   /* sum_j P_ij (S_ij - 1/t)   (pass 1)  /  sum_j P_ij w_ij logp_ij (pass 2)          */
#define CY_STAT_NEGC 5   /* sum_j Neg_ij                                                                     */
#define CY_STAT_POSE 6   /* sum_j P_ij E_ij, then c_i                                                        */
#define CY_STAT_SW 7     /* pass 2: sum_j P_ij w_ij (self-paced)  /  sum_j u_ij (exclude)                    */

/* bytes of scratch the forward / loss / backward need for this problem (split-column partial sums, fp32 gradient slabs
 * of a column-split backward, block partials of the loss reduction): one buffer of this size serves every call */
size_t cy_infonce_workspace_bytes(int64_t N, int64_t d, int dtype, int variant, int path);

/* pass 1: raw row sums of rows [row_begin, row_end) against all N columns, then their row statistics (xstat complete for
 * CY_SUPCON) */
int cy_infonce_fwd(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                   const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant, int path,
                   float* stats, float* xstat, void* workspace, size_t workspace_bytes, void* stream);

/* second forward sweep for CY_SUPCON_EXCLUDE / CY_SELFPACED_*: positive-pair sums that depend on the pass-1 statistics of
 * the owned rows; completes their xstat rows.  gamma = self-paced age parameter (contrastive.py:206-212).  On the tensor
 * path a row block visits only the column tiles whose label range intersects its own: O(N) work when rows are sorted by
 * label. */
int cy_infonce_fwd_pass2(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                         const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant,
                         float gamma, int path, float* stats, float* xstat, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Loss reduction over ALL N rows of xstat (the caller has all-gathered them in the sharded case), fixed summation order.
 * out8 [8] float: [0] = loss = sum_i term_i / N, [1] = sum_ij P_ij w_ij, [2] = sum_ij P_ij (self-paced downgrade ratio =
 * [1]/[2], contrastive.py:179-181; 0 for the other variants), [3] = number of non-finite row terms (NaN check, :98-99),
 * [4] = *bad_rows (un-normalised rows counted by cy_infonce_pack, :58), [5] = *overflow (cy_labels_canonicalize), [6..7] = 0:
 * everything the reference's per-step assertions need, in one 32-byte device->host read.  bad_rows / overflow may be NULL;
 * *bad_rows is reset to zero once it has been read (a persistent counter serves every step without a memset). */
int cy_infonce_loss(int64_t N, int variant, const float* xstat, float* out8, int32_t* bad_rows, const int32_t* overflow,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Backward: dz[row_begin:row_end, :] = gscale[0] * (1/t) * sum_j (G_ij + G_ji) z_j  with G = dLoss/dS built on the
 * fly from the xstat rows of i and j.  gscale is a DEVICE scalar: the upstream gradient
 * of the loss (carries hook weight and GradScaler scale; for correct_grad also 1/downgrade_ratio).  dz has the dtype
 * of z and row stride lddz; only the owned rows are written.  Bitwise reproducible (no atomics). */
int cy_infonce_bwd(const void* z, int dtype, int64_t N, int64_t d, int64_t ldz, const int32_t* labels,
                   const uint8_t* codes, int64_t row_begin, int64_t row_end, float inv_t, int variant, float gamma,
                   int path, const float* xstat, const float* gscale, void* dz, int64_t lddz, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Materialise the [N, N] fp32 positive / negative masks with the SAME device predicate the loss kernels evaluate
 * in-register (contrastive.py:62-71: repeat(2,2) and cleared diagonal).  Backs the lazily evaluated module attributes
 * pos_mask / neg_mask (read by semi_seg/hooks/infonce.py:235-242 on the first batch of an epoch) and the bit-exact
 * mask parity tests.  Either output may be NULL. */
int cy_infonce_masks(int64_t N, const int32_t* labels, const uint8_t* codes, float* pos_mask, float* neg_mask,
                     void* stream);

/* Canonicalise a label vector to int32 so that integer equality == the reference's comparison.
 * src_kind: 0 = float32 values (python lists go through torch.Tensor(list), contrastive.py:40: -0.0 == +0.0, NaN
 * never equal), 1 = int32 (copied), 2 = int64 (torch's default integer dtype: narrowed; *overflow is incremented for
 * every value outside the int32 range so that the caller can raise instead of comparing truncated labels).
 * Writes dst[0:n] and dst[n:2n] (tiling over the two views).  overflow may be NULL for kinds 0 and 1. */
int cy_labels_canonicalize(const void* src, int src_kind, int64_t n, int32_t* dst, int32_t* overflow, void* stream);

/* Feeder of cy_infonce_fwd: stacks the two views into z [2n, d] (contiguous), z[i] = source row order[i] of
 * cat(f1, f2) (contrastive.py:15; `order` = a permutation of 0..2n-1 as int64, or NULL for the identity — the modules
 * pass the label sort so that equal labels are adjacent), and runs the reference's `is_normalized` assertion
 * (contrastive.py:9-11, :58) on the device: *bad_rows is incremented once per source row whose L2 norm, rounded to the
 * element type like torch's `norm`, is not within 1e-8 + 1e-5 of 1.  ld1 / ld2: row pitches of f1 / f2 in elements.
 * bad_rows may be NULL (python -O: the reference's assert is stripped too).
 * inv_norm != NULL fuses the projector's L2-normalise tail (contrastyou/projectors/nn.py:47-54, heads.py:20,116-117)
 * into the same pass: z[i] = f / max(||f||_2, 1e-12) and inv_norm[i] [2n] keeps the reciprocal norm for the adjoint;
 * the is_normalized check is skipped in that mode. */
int cy_infonce_pack(const void* f1, const void* f2, int dtype, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                    const int64_t* order, void* z, int32_t* bad_rows, float* inv_norm, void* stream);

/* fp32 feeder of the tensor kernels: as cy_infonce_pack for float32 views (cat, permutation, is_normalized counted on the fp32
 * values), but every row is written as [hi | lo] bf16 halves (CY_F32_SPLIT): zs [2n, 2 d] bf16. */
int cy_infonce_pack_split(const void* f1, const void* f2, int64_t n, int64_t d, int64_t ld1, int64_t ld2, const int64_t* order,
                          void* zs, int32_t* bad_rows, void* stream);

/* Adjoint of cy_infonce_pack: scatters dz [2n, d] (row pitch lddz) back to the two views, g(order[i]) = dz[i];
 * g1, g2 are contiguous [n, d].  With (z, inv_norm) from a normalising pack it also applies the Jacobian of the
 * normalisation, g = (dz - z (z . dz)) * inv_norm; pass NULL, NULL otherwise.  gscale (device scalar, may be NULL)
 * multiplies the result: the modules run cy_infonce_bwd with unit upstream gradient right behind the forward sweep (no
 * host round trip between the two) and apply the actual upstream gradient here. */
int cy_infonce_unpack(const void* dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* order, void* g1,
                      void* g2, const void* z, const float* inv_norm, const float* gscale, void* stream);

/* Dense-feature variant of the feeder: the gather of `region_extractor`
 * (semi_seg/hooks/infonce.py:31-46, call sites :262-263) fused into the pack.  map1 / map2: [B, C, h, w] feature maps of the
 * two views (the dense projector's output); row i of a view = the C-vector of one sampled pixel: element c at
 * map[pix_off[i] + c * chan_stride] (pix_off [n] int64 = b*C*h*w + y*w + x, chan_stride = h*w; the same pixel list serves
 * both views, like the shared seed in the hook).  Everything else as cy_infonce_pack (d = C).  The adjoint scatters dz into
 * the caller's ZEROED [B, C, h, w] gradient maps (the sampled pixels of an image are distinct: plain stores). */
int cy_infonce_pack_gather(const void* map1, const void* map2, int dtype, int64_t n, int64_t d, const int64_t* pix_off,
                           int64_t chan_stride, const int64_t* order, void* z, int32_t* bad_rows, void* stream);
int cy_infonce_unpack_scatter(const void* dz, int dtype, int64_t n, int64_t d, int64_t lddz, const int64_t* order,
                              void* gmap1, void* gmap2, const int64_t* pix_off, int64_t chan_stride, const float* gscale,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * IIC discrete-MI segmentation loss.  Replaces compute_joint_2D / compute_joint_2D_with_padding_zeros
 * (contrastyou/losses/discreteMI.py:225-261), IIDSegmentationLoss.forward (:139-165) and their autograd backward.
 *
 *   x, y     [B, K, H, W] contiguous probability maps (x_out, x_tf_out)
 *   joint    raw joint [K, K, T, T] in DOUBLE, T = 2*pad+1:  J[k1,k2,dy,dx] = sum_{b,h,w} x[b,k1,h+dy-pad,w+dx-pad] *
 *            y[b,k2,h,w]  — what the F.conv2d call at discreteMI.py:229-232 returns.  It stays in double between the two
 *            calls because the epilogue's global min-shift (:233) turns a relative error e of J into ~e * J / (J - min J)
 *            of the normalised joint — a factor ~sqrt(pixels), 300-1000x at config 3 — and a float32 rounding of J
 *            (6e-8) alone then costs 1e-4 of gradient accuracy.  Multi-GPU: every rank accumulates its images, the
 *            caller all-reduces `joint` (900 doubles), then every rank runs the epilogue.
 * ---------------------------------------------------------------------------------------------------------------- */
size_t cy_iic_workspace_bytes(int B, int K, int H, int W, int pad);

int cy_iic_joint(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, double* joint,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Epilogue on the (global) raw joint: min-shift + 1e-8, per-displacement and global normalisation, optional
 * symmetrisation (pad > 0, :233-243) or division by n_pixels (pad == 0, :246-261); marginals; loss (:154-165).
 * Outputs: loss[1]; p00 [K,K] = p_i_j[0][0] (:152, get_joint_matrix); p_ij [T,T,K,K] (may be NULL);
 * djoint [K,K,T,T] = dLoss/dJoint (may be NULL).  n_pixels = global B*H*W.  One CTA, fp64; its few arrays live in
 * shared memory unless K*K*T*T is large, in which case cy_iic_epilogue_workspace_bytes() is non-zero and the caller
 * passes that much device scratch.  n_slots >= 1: `joint` is [n_slots][K,K,T,T] and the slots are summed first, in slot
 * order — the peer-memory form of the multi-GPU all-reduce: every rank pushes its partial joint into slot `rank` of every
 * peer (cy_p2p_push), and each rank then forms the identical global joint inside this kernel. */
size_t cy_iic_epilogue_workspace_bytes(int K, int pad);
int cy_iic_epilogue(const double* joint, int n_slots, int K, int pad, int symmetric, float lamda, float eps,
                    double n_pixels, float* loss, float* p00, float* p_ij, float* djoint, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Backward: dx, dy [B,K,H,W] (dtype of x) = gscale[0] * adjoint of cy_iic_joint applied to djoint. */
int cy_iic_bwd(const void* x, const void* y, int dtype, int B, int K, int H, int W, int pad, const float* djoint,
               const float* gscale, void* dx, void* dy, void* stream);

/* Sub-head stack.  The hooks evaluate the criterion once per sub-head pair of a cluster head and
 * average (semi_seg/hooks/discretemi.py:111, contrastyou/arch/unet.py cluster heads): n_heads independent problems of ONE
 * shape.  These three calls take the heads together: xs / ys / dxs / dys are HOST arrays of n_heads device pointers
 * ([B,K,H,W] each); head s uses joint + s*joint_stride (doubles, [n_slots][K,K,T,T] inside), loss / p00 / djoint + s*out_stride
 * (floats) and djoint + s*djoint_stride.  One joint launch + one reduction launch, one epilogue launch (a CTA per head) and one
 * adjoint launch when the tensor-core kernels take the shape (heads in chunks of 8); otherwise they run head by head through
 * the single-head entry points — results are identical to n_heads single-head calls either way.
 * Workspace of cy_iic_joint_heads: n_heads * cy_iic_workspace_bytes().  cy_iic_epilogue_heads has no p_ij output; slot sl of head s
 * (multi-GPU: rank sl's partial joint) lies at joint + s*joint_stride + sl*slot_stride, slot_stride = 0 meaning K*K*T*T (slots back
 * to back) — the peer-memory exchange lays the stack out as [world][n_heads][K,K,T,T], i.e. joint_stride = K*K*T*T and slot_stride =
 * n_heads*K*K*T*T. */
int cy_iic_joint_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                       double* joint, long long joint_stride, void* workspace, size_t workspace_bytes, void* stream);
int cy_iic_epilogue_heads(const double* joint, long long joint_stride, long long slot_stride, int n_heads, int n_slots, int K, int pad,
                          int symmetric, float lamda, float eps, double n_pixels, float* loss, float* p00, float* djoint,
                          long long out_stride, void* workspace, size_t workspace_bytes, void* stream);
int cy_iic_bwd_heads(const void* const* xs, const void* const* ys, int n_heads, int dtype, int B, int K, int H, int W, int pad,
                     const float* djoint, long long djoint_stride, const float* gscale, void* const* dxs, void* const* dys,
                     void* stream);

/* Fused SoftmaxWithT backward (contrastyou/projectors/nn.py:36-44 feeds IIDSegmentationLoss through
 * DenseClusterHead, heads.py:151-172).  pxs / pys are the PROBABILITY maps p = softmax(logits / T) over the K planes that the
 * forward consumed; dlxs / dlys receive dLoss/dlogits = p * (dL/dp - sum_k p_k dL/dp_k) / T, formed in the adjoint's epilogue
 * where a thread holds all K channels of a pixel — the dL/dp maps are never written and the separate softmax backward pass
 * (read dL/dp, read p, write dL/dlogits: 3 maps of traffic per side) disappears.  Shapes the tensor-core adjoint does not take
 * run cy_iic_bwd followed by an in-place softmax-backward kernel.  n_heads = 1 is the plain single-pair call.
 * (The forward half of the fusion is deliberately absent: the joint kernel is issue-bound, not HBM-bound — DESIGN.md §9.) */
/* SoftmaxWithT forward (nn.py:36-44): probs[i][b, :, h, w] = softmax(logits[i][b, :, h, w] / T) over the K planes, for n_maps
 * maps of one shape [B,K,H,W] (both views of every sub-head) in ONE streaming launch: each map read once, written once.
 * probs[i] may equal logits[i] (in place, like the reference's `input /= T`).  Host arrays of device pointers. */
int cy_softmax_t_fwd(const void* const* logits, void* const* probs, int n_maps, int dtype, int B, int K, int H, int W, float T,
                     void* stream);

int cy_iic_bwd_logits_heads(const void* const* pxs, const void* const* pys, int n_heads, int dtype, int B, int K, int H, int W,
                            int pad, const float* djoint, long long djoint_stride, const float* gscale, float T,
                            void* const* dlxs, void* const* dlys, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * IMSAT entropies.  Replaces imsat_loss / the marginal + conditional entropy pair of
 * contrastyou/losses/discreteMI.py:275-297 (used by IMSATLoss :20-52 and IMSATDynamicWeight :55-87) and their autograd.
 *   pred   [N, K, S] contiguous simplex over K (S = product of trailing dims; 1 for [N, K] classification outputs), K <= 64
 *   out2   {marginal = H(mean_{n,s} pred), conditional = mean_{n,s} H(pred[n,:,s])},  H(p) = -sum_k p log(p + eps)
 *   q      [K] class means, kept for the backward
 * One streaming pass each way; per-block partial sums reduced in a fixed order.  g2 = device {dL/dmarginal, dL/dconditional}.
 * ---------------------------------------------------------------------------------------------------------------- */
size_t cy_imsat_workspace_bytes(int K);
int cy_imsat_fwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, float* out2, float* q, void* workspace,
                 size_t workspace_bytes, void* stream);
int cy_imsat_bwd(const void* pred, int dtype, int64_t N, int K, int64_t S, float eps, const float* q, const float* g2,
                 void* grad, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Exchange over NVLink peer memory (row-sharded InfoNCE, batch-sharded IIC).  The reference has no distributed path; the
 * north_star's all-gather of embeddings / statistics and all-reduce of the joint are done here by the ranks' own stores
 * into each other's buffers instead of NCCL collectives.
 *
 *   peer_bufs  DEVICE array [world] of base pointers: the same symmetric allocation as mapped on every rank
 *              (torch.distributed._symmetric_memory: buffer_ptrs_dev); peer_bufs[rank] is the local buffer
 *   ranges     HOST array of n_ranges (<= 4) pairs (byte offset, byte count), both multiples of 16
 * Copies every range from the local buffer to the same offsets of all OTHER ranks' buffers (push all-gather: each rank
 * owns the ranges it pushes).  Completion on the peers is established by the caller's signal-pad barrier after the call. */
int cy_p2p_push(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges, void* stream);

/* cy_p2p_push + the barrier in ONE launch: after its copies the kernel's last block publishes `epoch` to every peer (release
 * store into the peer's flag array) and waits until every peer has published an epoch >= `epoch` to this rank.
 *   flag_off  byte offset (multiple of 16) inside the symmetric buffer of a uint32 [world] flag array; zero it on every rank
 *             (and barrier once) before the first call
 *   counter   one LOCAL device uint32, zero before the first call (the kernel leaves it zero)
 *   epoch     1, 2, 3, ... — the same sequence on every rank (one value per call)
 * On return (in stream order) every rank's ranges are visible in this rank's buffer. */
int cy_p2p_push_barrier(void* const* peer_bufs, int world, int rank, const unsigned long long* ranges, int n_ranges,
                        unsigned long long flag_off, unsigned int* counter, unsigned int epoch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRASTYOU_B200_H */
