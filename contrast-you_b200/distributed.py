"""Multi-GPU sharding of the two losses: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the exchange.

The reference has no distributed path at all (it never calls ``init_process_group``); under a DDP launch
each rank would contrast its local batch only.  What is built here is the row-sharded global-batch form of
BASELINE.json's config 4:

InfoNCE — rank r owns the rows of its local samples (rank-major order of the gathered problem:
  [r0 view1; r0 view2; r1 view1; ...]; for label-derived masks the order of rows is immaterial).  Per step:
  1. local: canonical labels, local sort by label, one pack kernel (cat + permutation + is_normalized) that writes the
     rank's rows straight into its slice of the [N, d] buffer;
  2. all-gather (in place) of the embeddings (bf16: N*d*2 bytes, 32 MiB at N=65536) and of the sorted labels;
  3. forward sweep(s) of the owned row block against all N columns -> the owned rows of ``xstat`` [N, 4];
  4. all-gather (in place) of ``xstat`` (16 B per row: 1 MiB at N=65536) — the ONLY exchange of statistics; every rank
     then reduces the loss from the gathered array in the same fixed order (identical bits, no scalar collective);
  5. backward: every rank computes the COMPLETE gradient of its own rows with one more strip sweep
     (dZ_i = (1/t) sum_k (G_ik + G_ki) z_k needs only the statistics of rows i and k, SURVEY.md §8e design (ii)),
     so no gradient collective is needed.  ``variant="reduce_scatter"`` runs the north_star's design (i) instead — each
     rank produces its strip's contribution to ALL N rows and the [N, d] gradient is reduce-scattered — for comparison;
     it moves N*d*4 bytes per rank to produce the same numbers.

IIC — images are independent summands of the raw joint: every rank accumulates its images, the [K,K,T,T] joint (double) is
  all-reduced (900 values) BEFORE the non-linear epilogue, which every rank then evaluates redundantly; the backward
  is purely local.

The returned loss is the global-batch loss on every rank and each rank's backward yields d(global loss)/d(its own
samples).  Under DDP (which averages parameter gradients over ranks) pass ``grad_scale=world_size`` to recover exactly
the single-process gradient of the global-batch loss.
"""
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L

__all__ = ["row_range", "gather_rank_major", "gather_rows_", "local_view_major", "make_joint_reduce", "PeerExchange",
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


def gather_rows_(full: Tensor, group=None) -> Tensor:
    """IN-PLACE all-gather of a rank-major array: ``full`` [G * m, ...] holds this rank's block at rows [rank*m, (rank+1)*m)
    on entry and every rank's block on return (the collective reads the own block where it already lies)."""
    world, rank = _ws(group)
    m = full.shape[0] // world
    dist.all_gather_into_tensor(full, full[rank * m:(rank + 1) * m], group=group)
    return full


def local_view_major(raw_local: Tensor, canonicalize) -> Tensor:
    """raw labels of this rank [n_local] -> canonical int32 labels of its row block [2 * n_local] (view-major tiling,
    ``canonicalize(raw, n)`` = cy_labels_canonicalize).  Canonicalisation is element-wise, so doing it per rank gives the
    same integers as doing it on the gathered vector."""
    return canonicalize(raw_local, raw_local.shape[0])


class PeerExchange:
    """Symmetric (peer-mapped) scratch of one process group: the ranks write into each other's buffers over NVLink with their
    own kernels (cy_p2p_push) and meet at signal-pad barriers — no NCCL collective on the data path.

    Built on ``torch.distributed._symmetric_memory`` (cuMem allocation + peer mapping + signal pads); the rendezvous is a
    collective, so every rank must create / grow the exchange at the same call.  ``acquire(nbytes)`` returns the local view of
    one of TWO alternating halves: rank A may already be writing step k+1's blocks into rank B's memory while B still reads
    step k's — a half is reused only two steps later, after a barrier that every rank reaches behind its own readers."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = _ws(group)
        self.buf = None
        self.hdl = None
        self.half_bytes = 0
        self.step = 0

    @staticmethod
    def available() -> bool:
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
            return True
        except Exception:  # noqa
            return False

    FLAG_BYTES = 4096      # uint32 [world] epoch flags of cy_p2p_push_barrier, behind the two data halves

    def _grow(self, nbytes: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.half_bytes = (nbytes + (1 << 20) - 1) >> 20 << 20
        self.buf = symm_mem.empty(2 * self.half_bytes + self.FLAG_BYTES, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.step = 0
        # the exchange's own barrier state: epoch flags inside the symmetric buffer (zeroed on every rank before anyone can
        # publish into them), the last-block counter of the push kernel, the epoch of the next call, and what each half's
        # label block currently holds (ShardedSupConLoss skips re-sending unchanged labels)
        self.buf[2 * self.half_bytes:].zero_()
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.counter_ptr = self.counter.data_ptr()
        self.dev_index = self.buf.device.index
        self.stream_of = torch._C._cuda_getCurrentRawStream
        self.epoch = 0
        self.labels_token = [None, None]
        self.hdl.barrier(channel=0)

    def acquire(self, nbytes: int, device):
        """-> (local uint8 view of this step's half, its byte offset inside the symmetric buffer)"""
        if self.buf is None or nbytes > self.half_bytes:
            self._grow(nbytes, device)
        self.half = self.step & 1
        off = self.half * self.half_bytes
        self.step += 1
        return self.buf[off:off + self.half_bytes], off

    def push(self, ranges):
        """copy the (offset, bytes) ranges of the local buffer into every peer's buffer (offsets relative to the whole
        symmetric buffer) and meet the peers — ONE launch (cy_p2p_push_barrier: the kernel's last block publishes this call's
        epoch to every peer and waits for theirs): on return (in stream order) every rank's ranges are visible here.  An empty
        range list is a plain barrier."""
        import ctypes
        flat = []
        for off, nb in ranges:
            assert off % 16 == 0 and nb % 16 == 0, (off, nb)
            flat += [off, nb]
        if not flat:
            flat = [0, 0]
        self.push_ranges((ctypes.c_ulonglong * len(flat))(*flat), len(flat) // 2)

    def push_ranges(self, arr, n_ranges: int):
        """``push`` with the (offset, bytes) pairs already in a ctypes array (callers that push the same ranges every step)"""
        self.epoch += 1
        L.check(L.lib().cy_p2p_push_barrier(self.ptrs_dev, self.world, self.rank, arr, n_ranges, 2 * self.half_bytes,
                                            self.counter_ptr, self.epoch & 0xffffffff, self.stream_of(self.dev_index)),
                "cy_p2p_push_barrier")


def make_joint_reduce(group=None, exchange: str = "auto"):
    """callback for IIDSegmentationLoss: make the raw joint global; the pixel count scales with the world size (equal per-rank
    batches, as under a DistributedSampler).  Protocol: ``reduce(compute_into, shape, n_pixels, device) -> (joint, n_slots,
    n_pixels)`` where ``compute_into(out)`` runs cy_iic_joint into ``out``.

    exchange "p2p" (default when symmetric memory is available): every rank computes its partial joint straight into ITS
    slot of a peer-mapped [world, K,K,T,T] array, pushes the slot to all peers (cy_p2p_push + signal-pad barrier) and
    cy_iic_epilogue sums the slots in rank order — identical bits on every rank, no NCCL call.  "nccl": all-reduce in place."""
    world, rank = _ws(group)
    state = {"px": None, "failed": exchange == "nccl", "plans": {}}

    def reduce(compute_into, shape, n_pixels: float, device):
        # (a captured graph would replay ONE half of the alternating buffer every step: captures take the collective form)
        if not state["failed"] and world > 1 and not torch.cuda.is_current_stream_capturing():
            try:
                if state["px"] is None:
                    if not PeerExchange.available():
                        raise RuntimeError("torch.distributed._symmetric_memory is not available")
                    state["px"] = PeerExchange(group)
                px = state["px"]
                shape = tuple(shape)
                plan = state["plans"].get(shape)
                if plan is None or plan[2] is not px.buf:
                    # per shape, once: the tensor views of both halves and the push ranges (the step is host-bound at the bench
                    # size — ~10 view / slice calls and a ctypes array per step were a fifth of its enqueue time)
                    import ctypes
                    nj = 1
                    for v in shape:
                        nj *= v
                    slot_bytes = (nj * 8 + 15) // 16 * 16
                    need = world * slot_bytes
                    if px.buf is None or need > px.half_bytes:
                        px._grow(need, device)
                        state["plans"].clear()
                    halves = []
                    for half in (0, 1):
                        base = half * px.half_bytes
                        slots = px.buf[base:base + need].view(torch.float64).view(world, slot_bytes // 8)
                        mine = slots[rank, :nj].view(shape)
                        whole = slots.view(world, *shape) if slot_bytes == nj * 8 else None
                        arr = (ctypes.c_ulonglong * 2)(base + rank * slot_bytes, slot_bytes)
                        halves.append((slots, mine, whole, arr))
                    plan = (halves, nj, px.buf)
                    state["plans"][shape] = plan
                px.half = px.step & 1
                px.step += 1
                slots, mine, whole, arr = plan[0][px.half]
                compute_into(mine)
                px.push_ranges(arr, 1)
                if whole is not None:
                    return whole, world, n_pixels * world
                return slots[:, :plan[1]].contiguous().view(world, *shape), world, n_pixels * world
            except Exception as e:  # noqa  (rendezvous refused on this box: fall back for good, every rank alike)
                if exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); using NCCL all-reduce for the IIC joint")
                state["failed"] = True
        joint = torch.empty(shape, dtype=torch.float64, device=device)
        compute_into(joint)
        if world > 1:
            dist.all_reduce(joint, group=group)
        return joint, 1, n_pixels * world

    return reduce


class _ShardedInfoNCE(torch.autograd.Function):
    """(f1, f2) [n_loc, d] local views -> (loss, out8).  One autograd node and one uninterrupted kernel / collective sequence
    per step: pack, the three in-place all-gathers, forward sweep(s), loss reduction and — when a gradient will be asked
    for — the strip backward with unit upstream gradient right behind it (losses/contrastive.py _FusedInfoNCE explains
    why); ``backward`` only scatters the stored rows back to the two views, times the upstream gradient."""

    @staticmethod
    def forward(ctx, f1, f2, labels_loc, order, inv_t, variant, gamma, path, group, check, design, overflow, status, px, scratch,
                lab_cached):
        from .losses.contrastive import _unit_scale
        lib = L.lib()
        world, rank = _ws(group)
        n_loc, d = f1.shape
        rows_loc = 2 * n_loc
        N = world * rows_loc
        rb, re = rank * rows_loc, (rank + 1) * rows_loc
        dev = f1.device
        dt = L.dtype_code(f1)
        need_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        with L.guard(f1):
            st = L.stream_ptr(dev)
            esz = f1.element_size()
            if px is not None:
                # peer-memory exchange: the three gathered arrays live in this step's half of the symmetric buffer
                zb, lb, xb = N * d * esz, N * 4, N * 16
                view, base = px.acquire(zb + lb + xb, dev)
                z_all = view[:zb].view(f1.dtype).view(N, d)
                labels_all = view[zb:zb + lb].view(torch.int32)
                xstat = view[zb + lb:zb + lb + xb].view(torch.float32).view(N, 4)
            else:
                z_all = torch.empty(N, d, dtype=f1.dtype, device=dev)
                labels_all = torch.empty(N, dtype=torch.int32, device=dev)
                xstat = torch.empty(N, 4, dtype=torch.float32, device=dev)
            # per-shape scratch of the module (row statistics, kernel workspace) and its persistent un-normalised-row counter
            # (cy_infonce_loss resets it once read): no allocation / memset launches per step
            key = (N, d, dt, variant, path, str(dev))
            if scratch.get("key") != key:
                ws_bytes = lib.cy_infonce_workspace_bytes(N, d, dt, variant, path)
                scratch.clear()
                scratch.update(key=key, stats=torch.empty(L.CY_NSTAT, N, dtype=torch.float32, device=dev), ws_bytes=ws_bytes,
                               ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
                               bad=torch.zeros(1, dtype=torch.int32, device=dev))
            stats, ws, ws_bytes = scratch["stats"], scratch["ws"], scratch["ws_bytes"]
            out8 = torch.empty(8, dtype=torch.float32, device=dev)
            bad = scratch["bad"] if check else None
            L.check(lib.cy_infonce_pack(f1.data_ptr(), f2.data_ptr(), dt, n_loc, d, f1.stride(0), f2.stride(0), L.ptr(order),
                                        z_all.data_ptr() + rb * d * esz, L.ptr(bad), None, st), "cy_infonce_pack")
            # the owned label block: a block that this half of the exchange buffer already holds (same cached label tensor,
            # unchanged since) is neither copied nor sent again — every rank decides that for ITS block only
            held = px.labels_token[px.half] if px is not None else None      # (tensor, version, rb): the tensor is kept alive
            send_labels = not (held is not None and held[0] is labels_loc and held[1] == labels_loc._version and held[2] == rb)
            if send_labels:
                labels_all[rb:re].copy_(labels_loc)
                if px is not None:
                    px.labels_token[px.half] = (labels_loc, labels_loc._version, rb) if lab_cached else None
            if px is not None:      # ONE launch pushes this rank's embedding rows (and labels) to every peer and meets the peers
                px.push([(base + rb * d * esz, rows_loc * d * esz)] + ([(base + zb + rb * 4, rows_loc * 4)] if send_labels else []))
            else:
                gather_rows_(z_all, group)
                gather_rows_(labels_all, group)
            zp, lp = z_all.data_ptr(), labels_all.data_ptr()
            L.check(lib.cy_infonce_fwd(zp, dt, N, d, d, lp, None, rb, re, inv_t, variant, path, stats.data_ptr(), xstat.data_ptr(),
                                       ws.data_ptr(), ws_bytes, st), "cy_infonce_fwd")
            if variant != L.CY_SUPCON:
                L.check(lib.cy_infonce_fwd_pass2(zp, dt, N, d, d, lp, None, rb, re, inv_t, variant, gamma, path, stats.data_ptr(),
                                                 xstat.data_ptr(), ws.data_ptr(), ws_bytes, st), "cy_infonce_fwd_pass2")
            if px is not None:
                px.push([(base + zb + lb + rb * 16, rows_loc * 16)])
            else:
                gather_rows_(xstat, group)
            L.check(lib.cy_infonce_loss(N, variant, xstat.data_ptr(), out8.data_ptr(), L.ptr(bad), L.ptr(overflow), ws.data_ptr(),
                                        ws_bytes, st), "cy_infonce_loss")
            loss = out8[0].clone()
            if status is not None:
                status.post(out8)
            dz_loc = None
            if need_grad:
                one = _unit_scale(dev)
                if design == "reduce_scatter":
                    dz_loc = _strip_contribution_reduce_scatter(z_all, labels_all, xstat, one, inv_t, variant, gamma, rb, re, group)
                else:
                    dz_loc = torch.empty(re - rb, d, dtype=z_all.dtype, device=dev)
                    # the kernels index dz by GLOBAL row: hand them the address row rb would have in a full [N, d] gradient
                    dz_base = dz_loc.data_ptr() - rb * d * dz_loc.element_size()
                    L.check(lib.cy_infonce_bwd(zp, dt, N, d, d, lp, None, rb, re, inv_t, variant, gamma, path, xstat.data_ptr(),
                                               one.data_ptr(), dz_base, d, ws.data_ptr(), ws_bytes, st), "cy_infonce_bwd")
        if need_grad:
            ctx.save_for_backward(dz_loc, *([order] if order is not None else []))
        ctx.n_loc = n_loc
        ctx.mark_non_differentiable(out8)
        return loss, out8

    @staticmethod
    def backward(ctx, grad_loss, _g8):
        lib = L.lib()
        dz_loc, *rest = ctx.saved_tensors
        order = rest[0] if rest else None
        n_loc, d = ctx.n_loc, dz_loc.shape[1]
        dev = dz_loc.device
        gscale = grad_loss
        if gscale.dtype != torch.float32 or gscale.numel() != 1 or not gscale.is_contiguous():
            gscale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        with L.guard(dz_loc):
            g1 = torch.empty(n_loc, d, dtype=dz_loc.dtype, device=dev)
            g2 = torch.empty(n_loc, d, dtype=dz_loc.dtype, device=dev)
            L.check(lib.cy_infonce_unpack(dz_loc.data_ptr(), L.dtype_code(dz_loc), n_loc, d, d, L.ptr(order), g1.data_ptr(),
                                          g2.data_ptr(), None, None, gscale.data_ptr(), L.stream_ptr(dev)), "cy_infonce_unpack")
        return g1, g2, None, None, None, None, None, None, None, None, None, None, None, None, None, None


def _strip_contribution_reduce_scatter(z_all, labels_all, xstat, gscale, inv_t, variant, gamma, rb, re, group):
    """north_star design (i), kept for the side-by-side timing: this rank evaluates dL/dS for ITS strip of
    rows only (G_ij for i owned, every j) and scatters both factors of dZ = (1/t)(G Z + G^T Z) — the owned rows get G Z,
    every column j gets its G^T Z share — into a full [N, d] fp32 gradient, which is then reduce-scattered over the
    ranks.  Evaluated with stock torch ops on the strip (tile by tile, fp32): it is the contract-named data flow, not the
    optimised path; numbers are identical to design (ii) up to summation order."""
    world, rank = _ws(group)
    N, d = z_all.shape
    rows = re - rb
    lab_i = labels_all[rb:re]
    xs_i = xstat[rb:re]
    zf = z_all.float()
    zi = zf[rb:re]
    dz = torch.zeros(N, d, dtype=torch.float32, device=z_all.device)
    if variant != L.CY_SUPCON:
        raise NotImplementedError("the reduce-scatter comparison design covers SupConLoss1's default variant")
    step = 4096
    for j0 in range(0, N, step):
        j1 = min(N, j0 + step)
        s = zi @ zf[j0:j1].t()
        e = torch.exp((s - 1.0) * inv_t)
        pos = (lab_i[:, None] == labels_all[None, j0:j1])
        g = e * xs_i[:, 2:3] - pos.to(torch.float32) * xs_i[:, 1:2]            # G_ij = coef_i E_ij - P_ij / c_i
        if rb < j1 and re > j0:                                                 # clear the diagonal of this block
            lo, hi = max(rb, j0), min(re, j1)
            idx = torch.arange(lo, hi, device=z_all.device)
            g[idx - rb, idx - j0] = 0.0
        dz[rb:re] += g @ zf[j0:j1]
        dz[j0:j1] += g.t() @ zi
    dz *= gscale * (inv_t / N)
    out = torch.empty(rows, d, dtype=torch.float32, device=z_all.device)
    dist.reduce_scatter_tensor(out, dz, group=group)
    return out.to(z_all.dtype)


class ShardedSupConLoss(torch.nn.Module):
    """Global-batch SupConLoss1 over all ranks of ``group`` (label / SimCLR masks, every variant of the family through
    ``exclude_other_pos`` / ``self_paced``).  Same forward signature as the single-process module; every rank passes its
    local views and labels.

    ``backward_design``: "local" (default; complete gradient of the owned rows from the gathered row statistics, no gradient
    collective — SURVEY.md §8e (ii)) or "reduce_scatter" (north_star's design (i), for comparison)."""

    def __init__(self, temperature=0.07, exclude_other_pos=False, *, group=None, grad_scale: float = 1.0, path: str = "auto",
                 deferred_checks: bool = False, backward_design: str = "local", exchange: str = "auto"):
        super().__init__()
        self._t = temperature
        self._group = group
        self._grad_scale = float(grad_scale)
        self._variant = L.CY_SUPCON_EXCLUDE if exclude_other_pos else L.CY_SUPCON
        self._gamma = 1e6
        # deferred_checks: as in SupConLoss1 — device-side counters instead of the per-step host read, which makes the step
        # (collectives included) capturable in a CUDA graph when the labels are a tensor
        self._deferred_checks = deferred_checks
        self._flags = None
        assert backward_design in ("local", "reduce_scatter")
        self._design = backward_design
        self._path = {"auto": L.CY_PATH_AUTO, "simt": L.CY_PATH_SIMT, "tcgen05": L.CY_PATH_TCGEN05}[path]
        self._cache = None
        self._status = None
        # exchange: "p2p" = the ranks' own stores over NVLink peer memory + signal-pad barriers (PeerExchange), "nccl" = three
        # in-place all-gathers, "auto" = p2p when torch's symmetric memory rendezvous works on this box, else nccl
        assert exchange in ("auto", "p2p", "nccl")
        self._exchange = exchange
        self._px = None
        self._scratch = {}

    def _local_labels(self, target, n_local, rank, device, sort):
        """canonical labels of the owned row block (+ the local sort), cached while the same label tensor comes back"""
        from .losses.contrastive import _canonical_labels, _TensorLabelCache
        overflow = None
        if target is None:      # SimCLR: globally unique ids
            raw = torch.arange(rank * n_local, (rank + 1) * n_local, dtype=torch.int32, device=device)
        elif isinstance(target, list):
            raw = torch.tensor(target, dtype=torch.float32, device=device)      # contrastive.py:39-40
        else:
            raw = target
        cacheable = isinstance(target, Tensor) and not torch.cuda.is_current_stream_capturing()      # see _ContrastBase._prepare
        if cacheable:
            if self._cache is None:
                self._cache = _TensorLabelCache()
            hit = self._cache.get(target, n_local, device, sort)
            if hit is not None:
                return hit[0], hit[1], None, True
        if isinstance(raw, Tensor) and raw.dtype == torch.int64:
            overflow = torch.zeros(1, dtype=torch.int32, device=device)
        labels = _canonical_labels(raw, n_local, device, overflow)
        order = None
        if sort:
            order = torch.argsort(labels)
            labels = labels.index_select(0, order)
        if cacheable and overflow is None:
            self._cache.put(target, n_local, device, sort, (labels, order))
            return labels, order, overflow, True
        return labels, order, overflow, False

    def forward(self, proj_feat1: Tensor, proj_feat2: Tensor, target=None, mask: Optional[Tensor] = None, **kwargs):
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
        # sort every rank's block by label when the tensor kernels will run (every rank decides the same way)
        tc = (proj_feat1.dtype in (torch.bfloat16, torch.float16) and d in (128, 256) and rows_loc % 128 == 0 and N >= 256
              and (self._variant == L.CY_SUPCON or N <= 4096 * 128)
              and (self._path == L.CY_PATH_TCGEN05 or (self._path == L.CY_PATH_AUTO and N >= 1024)))
        labels_loc, order, overflow, lab_cached = self._local_labels(target, n_local, rank, device, tc)
        f1 = proj_feat1 if proj_feat1.stride(1) == 1 else proj_feat1.contiguous()
        f2 = proj_feat2 if proj_feat2.stride(1) == 1 else proj_feat2.contiguous()
        status = None
        if not self._deferred_checks:
            from .losses.contrastive import _HostStatus
            if self._status is None:
                self._status = _HostStatus()
            status = self._status
        px = self._peer_exchange(rows_loc % 4 == 0 and (rows_loc * d * proj_feat1.element_size()) % 16 == 0)
        loss, out8 = _ShardedInfoNCE.apply(f1, f2, labels_loc, order, float(1.0 / self._t), self._variant, float(self._gamma),
                                           self._path, self._group, __debug__, self._design, overflow, status, px, self._scratch,
                                           lab_cached)
        if status is None:
            cur = out8[3:6].detach().clone()          # [non-finite terms, un-normalised rows, label overflows]
            cur[0] = cur[0] + torch.isnan(loss.detach()).to(cur.dtype)
            if self._flags is None or self._flags.device != cur.device:
                self._flags = cur
            else:
                self._flags.add_(cur)       # in place: the counters keep their address across CUDA-graph replays
            return loss * self._grad_scale if self._grad_scale != 1.0 else loss
        host = status.wait()                  # one 32-byte read, posted BEFORE the backward sweep was launched
        val, nbad, over = host[0], host[4], host[5]
        assert nbad == 0, f"features need to be normalized first"
        if over:
            raise ValueError("int64 labels outside the int32 range are not supported (pass int32 labels or a python list)")
        if val != val:
            raise RuntimeError(loss)
        return loss * self._grad_scale if self._grad_scale != 1.0 else loss

    def _peer_exchange(self, aligned: bool):
        # (a captured graph would replay ONE half of the alternating buffer every step: captures take the collective form)
        if self._exchange == "nccl" or not aligned or _ws(self._group)[0] == 1 or torch.cuda.is_current_stream_capturing():
            return None
        if self._px is None:
            try:
                if not PeerExchange.available():
                    raise RuntimeError("torch.distributed._symmetric_memory is not available")
                px = PeerExchange(self._group)
                px._grow(1 << 20, torch.device("cuda", torch.cuda.current_device()))      # rendezvous now: fail here, not mid-step
                self._px = px
            except Exception as e:  # noqa  (same outcome on every rank of one box)
                if self._exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); ShardedSupConLoss uses NCCL all-gathers")
                self._exchange = "nccl"
                return None
        return self._px

    def raise_if_flagged(self):
        """deferred_checks mode: one host read of this rank's accumulated (un-normalised rows, NaN losses) counters"""
        if self._flags is None:
            return
        nan, nbad, over = self._flags.tolist()
        self._flags.zero_()
        assert nbad == 0, f"features need to be normalized first"
        if over:
            raise ValueError("int64 labels outside the int32 range are not supported (pass int32 labels or a python list)")
        if nan:
            raise RuntimeError(f"loss was NaN in {nan} forward call(s)")


def shard_iic_loss(criterion, group=None, exchange: str = "auto"):
    """Make an ``IIDSegmentationLoss`` batch-sharded: its raw joint is summed over ``group`` before the epilogue
    (``exchange``: "auto" | "p2p" | "nccl", see make_joint_reduce)."""
    criterion._reduce_joint = make_joint_reduce(group, exchange)
    return criterion
