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

__all__ = ["row_range", "rank_major_labels", "gather_rank_major", "make_stats_exchange", "make_joint_reduce",
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


class ShardedSupConLoss(torch.nn.Module):
    """Global-batch SupConLoss1 over all ranks of ``group`` (label / SimCLR masks).  Same forward signature as the
    single-process module; every rank passes its local views and labels."""

    def __init__(self, temperature=0.07, *, group=None, grad_scale: float = 1.0, path: str = "auto"):
        super().__init__()
        self._t = temperature
        self._group = group
        self._grad_scale = float(grad_scale)
        self._path = {"auto": L.CY_PATH_AUTO, "simt": L.CY_PATH_SIMT, "tcgen05": L.CY_PATH_TCGEN05}[path]

    def forward(self, proj_feat1: Tensor, proj_feat2: Tensor, target=None, mask: Optional[Tensor] = None, **kwargs):
        from .losses.contrastive import info_nce, _canonical_labels, is_normalized, sort_rows_by_label, tensor_core_eligible
        if mask is not None:
            raise NotImplementedError("the sharded loss derives masks from labels (explicit [n,n] masks are per-process)")
        assert is_normalized(proj_feat1) and is_normalized(proj_feat2), f"features need to be normalized first"
        assert proj_feat1.shape == proj_feat2.shape, (proj_feat1.shape, proj_feat2.shape)
        world, rank = _ws(self._group)
        n_local = proj_feat1.shape[0]
        device = proj_feat1.device
        if target is None:      # SimCLR: globally unique ids
            raw = torch.arange(rank * n_local, (rank + 1) * n_local, dtype=torch.int32, device=device)
        elif isinstance(target, list):
            raw = torch.tensor(target, dtype=torch.float32, device=device)      # contrastive.py:39-40
        else:
            raw = target.to(device)
        raw_all = torch.empty(world * n_local, dtype=raw.dtype, device=device)
        dist.all_gather_into_tensor(raw_all, raw.contiguous(), group=self._group)
        labels = rank_major_labels(raw_all, world, lambda r, n: _canonical_labels(r, n, device))
        z_all = gather_rank_major(torch.cat([proj_feat1, proj_feat2], dim=0), self._group)
        if tensor_core_eligible(z_all, labels, None, L.CY_SUPCON, self._path):
            # sort every rank's row block by label (ownership of rows is unchanged): see sort_rows_by_label
            for r in range(world):
                z_all, labels = sort_rows_by_label(z_all, labels, r * 2 * n_local, (r + 1) * 2 * n_local)
        loss, _ = info_nce(z_all, labels, None, self._t, L.CY_SUPCON, path=self._path, rows=row_range(n_local, self._group),
                           gather_stats=make_stats_exchange(n_local, self._group))
        if torch.isnan(loss):
            raise RuntimeError(loss)
        return loss * self._grad_scale if self._grad_scale != 1.0 else loss


def shard_iic_loss(criterion, group=None):
    """Make an ``IIDSegmentationLoss`` batch-sharded: its raw joint is all-reduced over ``group`` before the epilogue."""
    criterion._reduce_joint = make_joint_reduce(group)
    return criterion
