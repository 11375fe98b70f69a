"""Multi-GPU sharding of the two losses: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the exchange.

The reference has no distributed path at all (it never calls ``init_process_group``); under a DDP launch
each rank would contrast its local batch only.  What is built here is the row-sharded global-batch form of
BASELINE.json's config 4:

InfoNCE — rank r owns the rows of its local samples.
  1. all-gather the embeddings (bf16: N*d*2 bytes, 32 MiB at N=65536) and the raw labels;
  2. every rank runs the forward sweep for its row block against all N columns -> per-row statistics, partial loss;
  3. all-gather three N-float row-statistic vectors, all-reduce the 4-float scalar block;
  4. backward: every rank computes the COMPLETE gradient of its own rows with one more strip sweep
     (dZ_i = (1/t) sum_k (G_ik + G_ki) z_k needs only the statistics of rows i and k, SURVEY.md §8e design (ii)),
     so no gradient collective is needed; the north_star's reduce-scatter of dZ (design (i)) would move 64 MiB to
     produce the identical numbers.
  Row order of the gathered problem is rank-major ([r0 view1; r0 view2; r1 view1; ...]): for label-derived masks the
  order of rows is immaterial (P_ij depends on the labels only, the diagonal on i == j).

IIC — images are independent summands of the raw joint: every rank accumulates its images, the [K,K,T,T] joint is
  all-reduced (900 floats) BEFORE the non-linear epilogue, which every rank then evaluates redundantly; the
  backward is purely local.

The returned loss is the global-batch loss on every rank and each rank's backward yields d(global loss)/d(its own
samples).  Under DDP (which averages parameter gradients over ranks) pass ``grad_scale=world_size`` to recover exactly
the single-process gradient of the global-batch loss.
"""
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L

__all__ = ["row_range", "rank_major_labels", "gather_rank_major", "make_stats_exchange", "exchange_strip_stats", "make_joint_reduce",
           "ShardedSupConLoss", "shard_iic_loss"]


def _ws(group):
    return dist.get_world_size(group), dist.get_rank(group)


def row_range(n_local: int, group=None):
    """rows of the gathered 2*n_local*G problem owned by this rank (rank-major order)."""
    _, rank = _ws(group)
    return rank * 2 * n_local, (rank + 1) * 2 * n_local


class _GatherRows(torch.autograd.Function):
    """all-gather [m, d] row blocks into [G*m, d]; backward hands this rank's row block of the gradient back (every
    rank already holds the complete gradient of its own rows, so nothing is reduced)."""

    @staticmethod
    def forward(ctx, local, group):
        world, rank = _ws(group)
        ctx.rows = (rank * local.shape[0], (rank + 1) * local.shape[0])
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        rb, re = ctx.rows
        return grad[rb:re], None


def gather_rank_major(local: Tensor, group=None) -> Tensor:
    return _GatherRows.apply(local, group)


def rank_major_labels(raw_all: Tensor, world: int, canonicalize) -> Tensor:
    """raw labels of all ranks [G * n_local] -> canonical int32 labels of the rank-major stacked problem
    [G * 2 * n_local].  ``canonicalize(raw, n)`` returns the view-major tiling [raw; raw] as int32 (cy_labels_canonicalize)."""
    n_all = raw_all.shape[0]
    n_local = n_all // world
    view_major = canonicalize(raw_all, n_all)                      # [2, G, n_local]
    return view_major.view(2, world, n_local).permute(1, 0, 2).reshape(-1).contiguous()


def make_stats_exchange(n_local: int, group=None, stat_rows=(L.CY_STAT_LOGDEN, L.CY_STAT_INVC, L.CY_STAT_COEF, L.CY_STAT_AUX)):
    """callback for losses.contrastive.info_nce(gather_stats=...): all-gather the row statistics the backward needs
    for foreign columns, all-reduce the scalar block (partial loss, self-paced sums, NaN count)."""
    rb, re = row_range(n_local, group)

    def exchange(stats: Tensor, out4: Tensor):
        for s in stat_rows:
            dist.all_gather_into_tensor(stats[s], stats[s, rb:re].clone(), group=group)
        dist.all_reduce(out4, group=group)

    return exchange


def make_joint_reduce(group=None):
    """callback for IIDSegmentationLoss: sum the raw joint over ranks; the pixel count scales with the world size
    (equal per-rank batches, as under a DistributedSampler)."""
    world, _ = _ws(group)

    def reduce(joint: Tensor, n_pixels: float) -> float:
        dist.all_reduce(joint, group=group)
        return n_pixels * world

    return reduce


class _ShardedInfoNCE(torch.autograd.Function):
    """z_loc [2*n_loc, d] (this rank's stacked views, rows already in their final order) -> (loss, out4).

    forward : all-gather z (rank-major), forward strip of the owned rows, ONE all-gather that carries the four row-statistic
              vectors of the strip plus the four partial scalars of every rank.
    backward: one backward strip -> the complete gradient of the owned rows, written straight into a [2*n_loc, d]
              buffer (no N x d zero fill, no gradient collective)."""

    @staticmethod
    def forward(ctx, z_loc, labels_all, inv_t, path, group):
        lib = L.lib()
        world, rank = _ws(group)
        rows_loc, d = z_loc.shape
        N = world * rows_loc
        rb, re = rank * rows_loc, (rank + 1) * rows_loc
        dev = z_loc.device
        dt = L.dtype_code(z_loc)
        st = L.stream_ptr(dev)
        z_all = torch.empty(N, d, dtype=z_loc.dtype, device=dev)
        dist.all_gather_into_tensor(z_all, z_loc.contiguous(), group=group)
        stats = torch.empty(L.CY_NSTAT, N, dtype=torch.float32, device=dev)
        out4 = torch.zeros(4, dtype=torch.float32, device=dev)
        ws_bytes = lib.cy_infonce_workspace_bytes(N, d, dt, L.CY_SUPCON, path)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        L.check(lib.cy_infonce_fwd(z_all.data_ptr(), dt, N, d, d, labels_all.data_ptr(), None, rb, re, inv_t, L.CY_SUPCON, path,
                                   stats.data_ptr(), ws.data_ptr(), ws_bytes, st), "cy_infonce_fwd")
        L.check(lib.cy_infonce_finalize(N, rb, re, inv_t, L.CY_SUPCON, 1, stats.data_ptr(), out4.data_ptr(), st),
                "cy_infonce_finalize")
        out4 = exchange_strip_stats(stats, out4, rb, re, group)
        ctx.save_for_backward(z_all, labels_all, stats, ws)
        ctx.cfg = (inv_t, path, rb, re)
        ctx.mark_non_differentiable(out4)
        return out4[0].clone(), out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_out4):
        lib = L.lib()
        z_all, labels_all, stats, ws = ctx.saved_tensors
        inv_t, path, rb, re = ctx.cfg
        N, d = z_all.shape
        gscale = grad_loss
        if gscale.dtype != torch.float32 or gscale.numel() != 1 or not gscale.is_contiguous():
            gscale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        dz_loc = torch.empty(re - rb, d, dtype=z_all.dtype, device=z_all.device)
        # the kernels index dz by GLOBAL row: hand them the address row rb would have in a full [N, d] gradient
        dz_base = dz_loc.data_ptr() - rb * d * dz_loc.element_size()
        L.check(lib.cy_infonce_bwd(z_all.data_ptr(), L.dtype_code(z_all), N, d, d, labels_all.data_ptr(), None, rb, re, inv_t,
                                   L.CY_SUPCON, 0.0, path, stats.data_ptr(), gscale.data_ptr(), dz_base, d, ws.data_ptr(),
                                   ws.numel(), L.stream_ptr(z_all.device)), "cy_infonce_bwd")
        return dz_loc, None, None, None, None


_STAT_ROWS = (L.CY_STAT_LOGDEN, L.CY_STAT_INVC, L.CY_STAT_COEF, L.CY_STAT_AUX)


def exchange_strip_stats(stats: Tensor, out4: Tensor, rb: int, re: int, group=None) -> Tensor:
    """ONE all-gather for what ``make_stats_exchange`` does with five collectives: every rank sends
    [4 statistic rows of its strip | its 4 partial scalars]; on return ``stats`` holds the four rows for all N columns
    (rank-major strips) and the summed scalar block is returned."""
    world, _ = _ws(group)
    rows_loc = re - rb
    N = stats.shape[1]
    assert _STAT_ROWS == (0, 1, 2, 3)      # the four rows are the leading rows of `stats`: plain slices, no index tensors
    send = torch.cat([stats[:4, rb:re].reshape(-1), out4])          # (a python index list would cost a pageable H2D + sync)
    recv = torch.empty(world * send.numel(), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, send.numel())
    stats[:4].view(4, world, rows_loc).copy_(recv[:, :4 * rows_loc].view(world, 4, rows_loc).permute(1, 0, 2))
    return recv[:, 4 * rows_loc:].sum(dim=0)


class ShardedSupConLoss(torch.nn.Module):
    """Global-batch SupConLoss1 over all ranks of ``group`` (label / SimCLR masks).  Same forward signature as the
    single-process module; every rank passes its local views and labels.

    Per step and rank: canonical labels of all ranks (one tiny all-gather), a LOCAL sort of the owned rows by label
    (the loss is invariant under row permutations inside a rank's block; the other ranks' blocks arrive sorted), one
    pack kernel (cat + permutation + is_normalized), then ``_ShardedInfoNCE`` (two more collectives)."""

    def __init__(self, temperature=0.07, *, group=None, grad_scale: float = 1.0, path: str = "auto",
                 deferred_checks: bool = False):
        super().__init__()
        self._t = temperature
        self._group = group
        self._grad_scale = float(grad_scale)
        # deferred_checks: as in SupConLoss1 — device-side counters instead of the per-step host read, which makes the step
        # (collectives included) capturable in a CUDA graph when the labels are a tensor
        self._deferred_checks = deferred_checks
        self._flags = None
        self._path = {"auto": L.CY_PATH_AUTO, "simt": L.CY_PATH_SIMT, "tcgen05": L.CY_PATH_TCGEN05}[path]

    def forward(self, proj_feat1: Tensor, proj_feat2: Tensor, target=None, mask: Optional[Tensor] = None, **kwargs):
        from .losses.contrastive import _canonical_labels, _PackViews
        if mask is not None:
            raise NotImplementedError("the sharded loss derives masks from labels (explicit [n,n] masks are per-process)")
        L.require_cuda(proj_feat1, proj_feat2)
        assert proj_feat1.shape == proj_feat2.shape, (proj_feat1.shape, proj_feat2.shape)
        assert proj_feat1.dim() == 2 and proj_feat1.dtype == proj_feat2.dtype
        world, rank = _ws(self._group)
        n_local, d = proj_feat1.shape
        rows_loc = 2 * n_local
        N = world * rows_loc
        device = proj_feat1.device
        if target is None:      # SimCLR: globally unique ids
            raw = torch.arange(rank * n_local, (rank + 1) * n_local, dtype=torch.int32, device=device)
        elif isinstance(target, list):
            raw = torch.tensor(target, dtype=torch.float32, device=device)      # contrastive.py:39-40
        else:
            raw = target.to(device)
        raw_all = torch.empty(world * n_local, dtype=raw.dtype, device=device)
        dist.all_gather_into_tensor(raw_all, raw.contiguous(), group=self._group)
        labels_all = rank_major_labels(raw_all, world, lambda r, n: _canonical_labels(r, n, device))      # [N]
        order = None
        tc = (proj_feat1.dtype in (torch.bfloat16, torch.float16) and d == 256 and rows_loc % 128 == 0 and N >= 256
              and (self._path == L.CY_PATH_TCGEN05 or (self._path == L.CY_PATH_AUTO and N >= 1024)))
        if tc:      # sort every rank's block by label; the permutation is needed for the owned block only
            sorted_blocks, perm = labels_all.view(world, rows_loc).sort(dim=1)
            order = perm[rank]
            labels_all = sorted_blocks.reshape(-1)
        f1 = proj_feat1 if proj_feat1.stride(1) == 1 else proj_feat1.contiguous()
        f2 = proj_feat2 if proj_feat2.stride(1) == 1 else proj_feat2.contiguous()
        z_loc, bad = _PackViews.apply(f1, f2, order, __debug__, False)
        loss, _ = _ShardedInfoNCE.apply(z_loc, labels_all, float(1.0 / self._t), self._path, self._group)
        if self._deferred_checks:
            nan = torch.isnan(loss.detach()).to(torch.int32).reshape(1)
            cur = torch.cat((bad.to(torch.int32).reshape(1) if bad.numel() else torch.zeros_like(nan), nan))
            if self._flags is None or self._flags.device != cur.device:
                self._flags = cur.clone()
            else:
                self._flags.add_(cur)       # in place: the counters keep their address across CUDA-graph replays
            return loss * self._grad_scale if self._grad_scale != 1.0 else loss
        if __debug__:
            nbad, val = torch.stack((bad[0].to(torch.float32), loss.detach())).tolist()
            assert nbad == 0, f"features need to be normalized first"
        else:
            val = loss.item()
        if val != val:
            raise RuntimeError(loss)
        return loss * self._grad_scale if self._grad_scale != 1.0 else loss

    def raise_if_flagged(self):
        """deferred_checks mode: one host read of this rank's accumulated (un-normalised rows, NaN losses) counters"""
        if self._flags is None:
            return
        nbad, nan = self._flags.tolist()
        self._flags.zero_()
        assert nbad == 0, f"features need to be normalized first"
        if nan:
            raise RuntimeError(f"loss was NaN in {nan} forward call(s)")


def shard_iic_loss(criterion, group=None):
    """Make an ``IIDSegmentationLoss`` batch-sharded: its raw joint is all-reduced over ``group`` before the epilogue."""
    criterion._reduce_joint = make_joint_reduce(group)
    return criterion
