"""Drop-in IIC discrete mutual-information losses backed by libcontrastyou_b200.so.

Mirrors ``contrastyou/losses/discreteMI.py``: ``IIDSegmentationLoss`` (:127-170), ``IIDLoss`` (:90-124) and the
joint builders ``compute_joint_2D`` (:225-243), ``compute_joint_2D_with_padding_zeros`` (:246-261), ``compute_joint``
(:201-222), with unchanged signatures / attributes / exceptions (call sites: semi_seg/hooks/discretemi.py:47-76,
:99-113, ccblock.py:315-339, midl.py:44-54, cc.py:39,:144-146).

The reference builds the joint with ``F.conv2d`` using a B x H x W "filter"; here one streaming kernel accumulates
the K x K x T x T joint straight from the two probability maps (cy_iic_joint), a single-CTA kernel does the 900-float
normalisation / entropy epilogue together with its analytic derivative (cy_iic_epilogue) and one kernel applies the
adjoint to produce both input gradients (cy_iic_bwd).
"""
import math
import ctypes
import sys

import torch
from torch import Tensor, nn

from .. import _lib as L

__all__ = ["IIDSegmentationLoss", "IIDLoss", "softmax_with_t", "compute_joint", "compute_joint_2D", "compute_joint_2D_with_padding_zeros",
           "raw_joint", "simplex"]


def simplex(t: Tensor, axis=1) -> bool:
    """contrastyou/utils/general.py:68-77."""
    _sum = t.sum(axis).type(torch.float32)
    return torch.allclose(_sum, torch.ones_like(_sum, dtype=torch.float32), rtol=1e-4, atol=1e-4)


def _check_pair(x: Tensor, y: Tensor):
    L.require_cuda(x, y)
    if x.dim() != 4 or x.shape != y.shape:
        raise ValueError(f"expected two [B, K, H, W] maps of equal shape, got {tuple(x.shape)} and {tuple(y.shape)}")
    if y.dtype != x.dtype:
        y = y.to(x.dtype)
    return x.contiguous(), y.contiguous()


_WORKSPACES = {}


def _workspace(nbytes: int, device, stream: int):
    """Scratch for the per-CTA partial joints, cached per (device, stream): launches on one stream are ordered, so the
    buffer can be reused by the next call without a fresh allocation (the loss is called once per hook per batch).
    Not while a CUDA graph is being captured: that buffer belongs to the graph's private pool and its capture stream may
    be recycled for eager work later, so captured calls get their own allocation."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
    key = (device.index, stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


_SIZES = {}


def _sizes(lib, B, K, H, W, padding):
    """(partial-joint workspace bytes, epilogue workspace bytes) of a shape — asked once per shape, not once per step"""
    key = (B, K, H, W, padding)
    v = _SIZES.get(key)
    if v is None:
        v = _SIZES[key] = (lib.cy_iic_workspace_bytes(B, K, H, W, padding), lib.cy_iic_epilogue_workspace_bytes(K, padding))
    return v


def _joint_forward(x, y, padding, joint=None):
    lib = L.lib()
    B, K, H, W = x.shape
    T = 2 * padding + 1
    with L.guard(x):
        if joint is None:      # double: see include/contrastyou_b200.h (the epilogue's min-shift amplifies a float32 rounding of J)
            joint = torch.empty(K, K, T, T, dtype=torch.float64, device=x.device)
        st = L.stream_ptr(x.device)
        ws_bytes = _sizes(lib, B, K, H, W, padding)[0]
        ws = _workspace(ws_bytes, x.device, st)
        L.check(lib.cy_iic_joint(x.data_ptr(), y.data_ptr(), L.dtype_code(x), B, K, H, W, padding, joint.data_ptr(),
                                 ws.data_ptr(), ws_bytes, st), "cy_iic_joint")
    return joint


def _joint_backward(x, y, padding, djoint, gscale):
    lib = L.lib()
    B, K, H, W = x.shape
    with L.guard(x):
        dx, dy = torch.empty_like(x), torch.empty_like(y)
        L.check(lib.cy_iic_bwd(x.data_ptr(), y.data_ptr(), L.dtype_code(x), B, K, H, W, padding, djoint.data_ptr(),
                               gscale.data_ptr(), dx.data_ptr(), dy.data_ptr(), L.stream_ptr(x.device)), "cy_iic_bwd")
    return dx, dy


class _RawJoint(torch.autograd.Function):
    """Differentiable raw joint J[k1,k2,dy,dx] = sum_{b,h,w} x[b,k1,h+dy-p,w+dx-p] y[b,k2,h,w] — what the reference's
    ``F.conv2d(x^T, weight=y^T, padding=p)`` returns (discreteMI.py:227-232)."""

    @staticmethod
    def forward(ctx, x, y, padding):
        ctx.save_for_backward(x, y)
        ctx.padding = padding
        return _joint_forward(x, y, padding).to(torch.float32)      # what F.conv2d hands back in the reference

    @staticmethod
    def backward(ctx, gJ):
        x, y = ctx.saved_tensors
        one = torch.ones(1, dtype=torch.float32, device=x.device)
        dx, dy = _joint_backward(x, y, ctx.padding, gJ.to(torch.float32).contiguous(), one)
        return dx, dy, None


def raw_joint(x_out: Tensor, x_tf_out: Tensor, padding: int = 0) -> Tensor:
    x, y = _check_pair(x_out, x_tf_out)
    return _RawJoint.apply(x, y, int(padding))


def compute_joint_2D(x_out: Tensor, x_tf_out: Tensor, *, symmetric: bool = True, padding: int = 0):
    """discreteMI.py:225-243 — [T, T, K, K] joint with the reference's min-shift / normalisation quirks."""
    p_i_j = raw_joint(x_out, x_tf_out, padding)
    p_i_j = p_i_j - p_i_j.min().detach() + 1e-8
    p_i_j = p_i_j.permute(2, 3, 0, 1)
    p_i_j = p_i_j / p_i_j.sum(dim=[2, 3], keepdim=True)
    if symmetric:
        p_i_j = (p_i_j + p_i_j.permute(0, 1, 3, 2)) / 2.0
    p_i_j = p_i_j / p_i_j.sum()
    return p_i_j.contiguous()


def compute_joint_2D_with_padding_zeros(x_out: Tensor, x_tf_out: Tensor, *, symmetric: bool = True):
    """discreteMI.py:246-261 — [1, 1, K, K], divided by the pixel count, no shift / renormalisation."""
    k = x_out.shape[1]
    n = x_out.shape[0] * x_out.shape[2] * x_out.shape[3]
    p_i_j = raw_joint(x_out, x_tf_out, 0).view(k, k) / n
    if symmetric:
        p_i_j = (p_i_j + p_i_j.t()) / 2.0
    return p_i_j.view(1, 1, k, k).contiguous()


def compute_joint(x_out: Tensor, x_tf_out: Tensor, symmetric=True) -> Tensor:
    """discreteMI.py:201-222 — [K, K] joint of two [bn, K] simplices."""
    assert simplex(x_out), f"x_out not normalized."
    assert simplex(x_tf_out), f"x_tf_out not normalized."
    bn, k = x_out.shape
    assert x_tf_out.size()[0] == bn and x_tf_out.size()[1] == k
    p_i_j = raw_joint(x_out.reshape(bn, k, 1, 1), x_tf_out.reshape(bn, k, 1, 1), 0).view(k, k)
    if symmetric:
        p_i_j = (p_i_j + p_i_j.t()) / 2.0
    p_i_j = p_i_j / p_i_j.sum()
    return p_i_j.contiguous()


def softmax_with_t(logits, T: float = 1.0):
    """``[softmax(l / T, dim=1) for l in logits]`` for [B, K, H, W] maps of one shape and dtype — SoftmaxWithT (reference
    projectors/nn.py:36-44) for all maps in ONE streaming launch (cy_softmax_t_fwd).  No autograd: the backward is fused into
    the IIC adjoint (``IIDSegmentationLoss.forward_logits``)."""
    logits = [m.contiguous() for m in logits]
    m0 = logits[0]
    L.require_cuda(*logits)
    for m in logits:
        assert m.shape == m0.shape and m.dtype == m0.dtype and m.dim() == 4, "maps must share shape and dtype"
    B, K, H, W = m0.shape
    out = [torch.empty_like(m) for m in logits]
    n = len(logits)
    with L.guard(m0):
        L.check(L.lib().cy_softmax_t_fwd((ctypes.c_void_p * n)(*[m.data_ptr() for m in logits]),
                                         (ctypes.c_void_p * n)(*[m.data_ptr() for m in out]), n, L.dtype_code(m0), B, K, H, W,
                                         float(T), L.stream_ptr(m0.device)), "cy_softmax_t_fwd")
    return tuple(out)


class _IIDSegFunction(torch.autograd.Function):
    """(x, y) -> (loss, p_i_j[0][0]) with the fused epilogue; ``reduce_joint`` all-reduces the raw joint when the batch
    is sharded over ranks (the epilogue is non-linear in J, so it must see the global joint)."""

    @staticmethod
    def forward(ctx, x, y, padding, symmetric, lamda, eps, reduce_joint):
        lib = L.lib()
        B, K, H, W = x.shape
        T = 2 * padding + 1
        nj = K * K * T * T
        # the returned loss must not share storage (and hence a version counter) with the saved dL/dJ: callers scale losses in
        # place (`loss *= w`), which would otherwise invalidate the backward — so it is its own allocation (no clone kernel)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        buf = torch.empty(K * K + nj, dtype=torch.float32, device=x.device)
        p00 = buf[:K * K].view(K, K)
        djoint = buf[K * K:].view(K, K, T, T)
        n_pixels = float(B * H * W)
        n_slots = 1
        if reduce_joint is None:
            joint = _joint_forward(x, y, padding)                                    # [K, K, T, T] float64
        else:
            # batch-sharded: the callback supplies where the partial joint is written and makes the other ranks' partials
            # visible — an NCCL all-reduce in place (1 slot) or peer-memory stores into per-rank slots that the epilogue sums
            joint, n_slots, n_pixels = reduce_joint(lambda out: _joint_forward(x, y, padding, out), (K, K, T, T), n_pixels, x.device)
        with L.guard(x):
            ws_bytes = _sizes(lib, B, K, H, W, padding)[1]
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
            L.check(lib.cy_iic_epilogue(joint.data_ptr(), n_slots, K, padding, int(bool(symmetric)), float(lamda), float(eps),
                                        n_pixels, loss.data_ptr(), p00.data_ptr(), None, djoint.data_ptr(), L.ptr(ws), ws_bytes,
                                        L.stream_ptr(x.device)), "cy_iic_epilogue")
        ctx.save_for_backward(x, y, djoint)
        ctx.padding = padding
        ctx.mark_non_differentiable(p00)
        return loss, p00

    @staticmethod
    def backward(ctx, grad_loss, _grad_p00):
        x, y, djoint = ctx.saved_tensors
        gscale = grad_loss
        if gscale.dtype != torch.float32 or gscale.numel() != 1 or not gscale.is_contiguous():
            gscale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        dx, dy = _joint_backward(x, y, ctx.padding, djoint, gscale)
        return dx, dy, None, None, None, None, None


class _IIDSegMultiFunction(torch.autograd.Function):
    """S sub-head pairs in one autograd node: (x_0, y_0, ..., x_{S-1}, y_{S-1}) -> (mean loss, p00 of head 0).

    The hooks evaluate ``sum(criterion(x1, x2) for x1, x2 in zip(prob1, prob2)) / len(prob1)`` over the sub-heads of a
    cluster head (semi_seg/hooks/discretemi.py:111, ccblock.py:208-218): S python-level criterion calls, S autograd nodes
    and ~8 small allocations each.  Here the S joints, the S epilogues and the S adjoints are ONE launch each (plus the joint's
    reduction): cy_iic_joint_heads / cy_iic_epilogue_heads / cy_iic_bwd_heads, with one output buffer; the adjoint gets the
    head's share 1/S of the upstream gradient through ``gscale``.

    ``temperature`` not None: the inputs are the cluster heads' LOGITS (``DenseClusterHead(..., skip_softmax=True)``); the
    probabilities softmax(logits / T) are formed here and only they are kept for the backward, whose adjoint launch applies the
    softmax backward in its epilogue (cy_iic_bwd_logits_heads) — dL/dp is never written to memory."""

    @staticmethod
    def forward(ctx, padding, symmetric, lamda, eps, temperature, reduce_joint, *maps):
        lib = L.lib()
        if temperature is not None:
            maps = softmax_with_t(maps, temperature)
        S = len(maps) // 2
        x0 = maps[0]
        B, K, H, W = x0.shape
        T = 2 * padding + 1
        nj = K * K * T * T
        per = 1 + K * K + nj
        if x0.device.index != torch.cuda.current_device():
            torch.cuda.set_device(x0.device)      # (rare) model on a non-current device: make it current for the launches
        buf = torch.empty(S, per, dtype=torch.float32, device=x0.device)
        st = L.stream_ptr(x0.device)
        ws_bytes = lib.cy_iic_workspace_bytes(B, K, H, W, padding)
        ews_bytes = lib.cy_iic_epilogue_workspace_bytes(K, padding)
        ews = torch.empty(ews_bytes, dtype=torch.uint8, device=x0.device) if ews_bytes else None
        dt = L.dtype_code(x0)
        base = buf.data_ptr()
        # one joint launch (+ its reduction), one epilogue launch (a CTA per head); cy_iic_*_heads fall back to head-by-head
        # launches inside the library for shapes the tensor-core kernels do not take
        xs = (ctypes.c_void_p * S)(*[maps[2 * s].data_ptr() for s in range(S)])
        ys = (ctypes.c_void_p * S)(*[maps[2 * s + 1].data_ptr() for s in range(S)])
        ws = _workspace(S * ws_bytes, x0.device, st)

        def joints_into(out):                     # out: [S, K, K, T, T] float64, contiguous
            L.check(lib.cy_iic_joint_heads(xs, ys, S, dt, B, K, H, W, padding, out.data_ptr(), nj, ws.data_ptr(), S * ws_bytes, st),
                    "cy_iic_joint_heads")
        n_pixels, n_slots = float(B * H * W), 1
        if reduce_joint is None:
            joints = torch.empty(S, nj, dtype=torch.float64, device=x0.device)      # raw joints stay in double (contrastyou_b200.h)
            joints_into(joints)
        else:
            # batch-sharded: the stack of S partial joints is exchanged as ONE array — per-rank slots [world][S][K,K,T,T] in peer
            # memory (summed by the epilogue in rank order) or one in-place NCCL all-reduce
            joints, n_slots, n_pixels = reduce_joint(joints_into, (S, K, K, T, T), n_pixels, x0.device)
        L.check(lib.cy_iic_epilogue_heads(joints.data_ptr(), nj, S * nj, S, n_slots, K, padding, int(bool(symmetric)), float(lamda),
                                          float(eps), n_pixels, base, base + 4, base + 4 * (1 + K * K), per, L.ptr(ews), ews_bytes, st),
                "cy_iic_epilogue_heads")
        ctx.save_for_backward(buf, *maps)
        ctx.cfg = (padding, S, K, T, per, temperature)
        p00 = buf[0, 1:1 + K * K].view(K, K)
        ctx.mark_non_differentiable(p00)
        return buf[:, 0].mean(), p00

    @staticmethod
    def backward(ctx, grad_loss, _grad_p00):
        lib = L.lib()
        buf, *maps = ctx.saved_tensors
        padding, S, K, T, per, temperature = ctx.cfg
        B, _, H, W = maps[0].shape
        if buf.device.index != torch.cuda.current_device():
            torch.cuda.set_device(buf.device)
        gscale = (grad_loss.detach().to(torch.float32) / S).reshape(1).contiguous()
        st = L.stream_ptr(buf.device)
        dt = L.dtype_code(maps[0])
        grads = [torch.empty_like(m) for m in maps]
        xs = (ctypes.c_void_p * S)(*[maps[2 * s].data_ptr() for s in range(S)])
        ys = (ctypes.c_void_p * S)(*[maps[2 * s + 1].data_ptr() for s in range(S)])
        dxs = (ctypes.c_void_p * S)(*[grads[2 * s].data_ptr() for s in range(S)])
        dys = (ctypes.c_void_p * S)(*[grads[2 * s + 1].data_ptr() for s in range(S)])
        dj = buf.data_ptr() + (1 + K * K) * 4
        if temperature is None:
            L.check(lib.cy_iic_bwd_heads(xs, ys, S, dt, B, K, H, W, padding, dj, per, gscale.data_ptr(), dxs, dys, st), "cy_iic_bwd_heads")
        else:
            L.check(lib.cy_iic_bwd_logits_heads(xs, ys, S, dt, B, K, H, W, padding, dj, per, gscale.data_ptr(), float(temperature),
                                                dxs, dys, st), "cy_iic_bwd_logits_heads")
        return (None, None, None, None, None, None, *grads)


class IIDSegmentationLoss(nn.Module):
    """discreteMI.py:127-170."""

    def __init__(self, lamda=1.0, padding=0, eps: float = 1e-5, symmetric: bool = False) -> None:
        super(IIDSegmentationLoss, self).__init__()
        self.lamda = lamda
        self.padding = padding
        self._eps = eps
        self.symmetric = symmetric
        self._reduce_joint = None      # set by contrast_you_b200.distributed.shard_iic_loss

    def forward(self, x_out: Tensor, x_tf_out: Tensor, mask: Tensor = None) -> Tensor:
        if mask is not None:
            x_out *= mask              # in place, like the reference (:142-144)
            x_tf_out *= mask
        if self.padding < 0:
            raise ValueError(self.padding)
        x, y = _check_pair(x_out, x_tf_out)
        loss, p00 = _IIDSegFunction.apply(x, y, int(self.padding), bool(self.symmetric), float(self.lamda), float(self._eps),
                                          self._reduce_joint)
        self.__dict__["_p_i_j"] = p00      # (plain attribute; Module.__setattr__ costs ~5 us of type checks per step)
        return loss

    def forward_logits(self, x_logits: Tensor, x_tf_logits: Tensor, T: float = 1.0) -> Tensor:
        """``self(softmax(x_logits / T, 1), softmax(x_tf_logits / T, 1))`` with SoftmaxWithT's backward fused into the adjoint
        kernel (extension).  Feed it ``DenseClusterHead(...)(features, skip_softmax=True)``."""
        return self.forward_heads([x_logits], [x_tf_logits], logits_T=T)

    def forward_heads(self, x_outs, x_tf_outs, logits_T: float = None) -> Tensor:
        """mean of ``self(x, y)`` over the sub-head pairs, evaluated as ONE autograd node (extension).
        Equivalent to ``sum(self(x, y) for x, y in zip(x_outs, x_tf_outs)) / len(x_outs)``; ``_p_i_j`` is head 0's.
        ``logits_T`` not None: the inputs are logits and every pair is softmax(. / logits_T) first (see ``forward_logits``)."""
        if self.padding < 0:
            raise ValueError(self.padding)
        if logits_T is not None and not float(logits_T) > 0:
            raise ValueError(f"temperature {logits_T}")
        assert len(x_outs) == len(x_tf_outs) and len(x_outs) >= 1
        maps = []
        for x, y in zip(x_outs, x_tf_outs):
            x, y = _check_pair(x, y)
            assert x.shape == x_outs[0].shape and x.dtype == x_outs[0].dtype, "sub-heads must share shape and dtype"
            maps += [x, y]
        # (a criterion sharded with shard_iic_loss exchanges the whole stack of partial joints between the two kernels)
        loss, p00 = _IIDSegMultiFunction.apply(int(self.padding), bool(self.symmetric), float(self.lamda), float(self._eps),
                                               None if logits_T is None else float(logits_T), self._reduce_joint, *maps)
        self.__dict__["_p_i_j"] = p00      # (plain attribute; Module.__setattr__ costs ~5 us of type checks per step)
        return loss

    def get_joint_matrix(self):
        if not hasattr(self, "_p_i_j"):
            raise RuntimeError()
        return self._p_i_j.detach().cpu().numpy()


class IIDLoss(nn.Module):
    """discreteMI.py:90-124 — returns (loss, loss_no_lamb, p_i_j); the 1e-10 guards are hard-coded in the reference."""

    def __init__(self, lamb: float = 1.0, eps: float = sys.float_info.epsilon):
        super().__init__()
        self.lamb = float(lamb)
        self.eps = float(eps)

    def forward(self, x_out: Tensor, x_tf_out: Tensor):
        if x_out.dim() != 2:
            raise AssertionError(x_out.shape)
        joint = compute_joint(x_out, x_tf_out)                   # asserts both simplices (discreteMI.py:108-110, :210-211)
        guard = 1e-10                                            # hard-coded in the reference (:118-123), not self.eps
        log_rows = torch.log(joint.sum(dim=1, keepdim=True) + guard)      # [K, 1]
        log_cols = torch.log(joint.sum(dim=0, keepdim=True) + guard)      # [1, K]
        neg_entropy = (joint * torch.log(joint + guard)).sum()
        marginals = (joint * (log_rows + log_cols)).sum()
        # closed form of  -sum J (log J - lamb log p_j - lamb log p_i)  for lamb and for lamb = 1
        return self.lamb * marginals - neg_entropy, marginals - neg_entropy, joint
