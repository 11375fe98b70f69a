"""Losses of the reference that reuse the IIC joint with a different epilogue.

Each class mirrors its reference counterpart (same ctor / forward signatures, attributes, exceptions); the only heavy
step — the raw joint J[k1,k2,dy,dx] = sum_{b,h,w} x[b,k1,h+dy-p,w+dx-p] y[b,k2,h,w] — goes through ``raw_joint``
(cy_iic_joint / cy_iic_bwd, the same sm_100a kernels as ``IIDSegmentationLoss``); the K x K (x T x T) epilogues are a
handful of tiny torch ops on the joint, differentiated by autograd.

* ``RedundancyCriterion``  contrastyou/losses/redundancy_reduction.py:12-54  (hook: semi_seg/hooks/ccblock.py:342-376)
* ``PUISegLoss``           contrastyou/losses/pica_loss.py:43-80
* ``IMSATLoss`` / ``IMSATDynamicWeight`` / ``imsat_loss`` / ``imsat_with_entropy``
                           contrastyou/losses/discreteMI.py:20-87, :275-297   (hooks: ccblock.py:428-472, midl.py:83-90)
"""
import math
import sys

import torch
from torch import Tensor, nn

from .discreteMI import compute_joint_2D_with_padding_zeros, raw_joint, simplex

__all__ = ["RedundancyCriterion", "PUISegLoss", "IMSATLoss", "IMSATDynamicWeight", "imsat_loss", "imsat_with_entropy",
           "Entropy"]


class Entropy(nn.Module):
    """contrastyou/losses/kl.py:31-62 — -sum p log(p + eps) over dim 1."""

    def __init__(self, reduction="mean", eps=1e-16):
        super().__init__()
        assert reduction in ("mean", "sum", "none")
        self._eps = eps
        self._reduction = reduction

    def forward(self, input_: Tensor) -> Tensor:
        if input_.dim() < 2:
            raise AssertionError(f"Entropy expects [N, K, ...], got {tuple(input_.shape)}")
        assert simplex(input_), f"Entropy input should be a simplex"
        per_sample = -(input_ * torch.log(input_ + self._eps)).sum(dim=1)      # [N, ...]: class axis reduced
        reduce = {"mean": per_sample.mean, "sum": per_sample.sum, "none": lambda: per_sample}
        return reduce[self._reduction]()


entropy_criterion = Entropy(reduction="none", eps=1e-8)      # semi_seg/hooks/midl.py:13


class RedundancyCriterion(nn.Module):
    """redundancy_reduction.py:12-54.  With J the padding-0 joint (K x K, total mass 1), r = J.sum(1), c = J.sum(0):

        loss = -sum(target * log(J + eps)) + lamda * sum(J * (log(c + eps)[None, :] + log(r + eps)[:, None]))
        target = alpha * I / K + (1 - alpha) * J          (the target is NOT detached in the reference: it carries gradient)
    """

    def __init__(self, *, eps: float = 1e-5, symmetric: bool = True, lamda: float = 1, alpha: float) -> None:
        super().__init__()
        self._eps = eps
        self.symmetric = symmetric
        self.lamda = lamda
        self.alpha = alpha

    def forward(self, x_out: Tensor, x_tf_out: Tensor):
        k = x_out.shape[1]
        joint = compute_joint_2D_with_padding_zeros(x_out, x_tf_out, symmetric=self.symmetric).view(k, k)
        self._p_i_j = joint
        eye = torch.eye(k, device=joint.device, dtype=joint.dtype)
        target = eye * (self.alpha / k) + joint * (1 - self.alpha)
        log_rows = torch.log(joint.sum(dim=1, keepdim=True) + self._eps)       # [k, 1]
        log_cols = torch.log(joint.sum(dim=0, keepdim=True) + self._eps)       # [1, k]
        marginal_term = self.lamda * (joint * (log_cols + log_rows)).sum()
        match_term = -(target * torch.log(joint + self._eps)).sum()
        return match_term + marginal_term

    def kl_criterion(self, dist: Tensor, prior: Tensor):
        return -(prior * (dist + self._eps).log() + (1 - prior) * (1 - dist + self._eps).log()).mean()

    def get_joint_matrix(self):
        if not hasattr(self, "_p_i_j"):
            raise RuntimeError()
        return self._p_i_j.detach().cpu().numpy()

    def set_ratio(self, alpha: float):
        """0: entropy minimisation ... 1: Barlow-twins-like identity target (redundancy_reduction.py:46-54)"""
        assert 0 <= alpha <= 1, alpha
        self.alpha = alpha


class PUISegLoss(nn.Module):
    """pica_loss.py:43-80.  J = the IIC conv-joint with padding p, shifted by its global minimum (+1e-16), normalised per
    displacement, symmetrised over (k1, k2) and averaged over the T x T displacements;
    loss = mean(-I * log(J + 1e-16)) + lamda * (log n + sum(q log q)) with q = the class-mean of x at every pixel
    (the reference reduces dim 0 of the PERMUTED [K, N, H, W] map, :70 — i.e. over the classes; kept literally)."""

    def __init__(self, lamda=2.0, padding=3):
        super().__init__()
        self.lamda = lamda
        self.padding = padding

    def forward(self, x_out, x_tf_out):
        assert x_out.shape == x_tf_out.shape, ('Inputs are required to have same shape')
        raw = raw_joint(x_out, x_tf_out, self.padding)                      # [K, K, T, T] == F.conv2d(x^T, weight=y^T)
        shifted = raw - raw.min().detach() + 1e-16
        per_disp = shifted / shifted.sum(dim=(0, 1), keepdim=True)
        sym = 0.5 * (per_disp + per_disp.transpose(0, 1))
        joint = sym.mean(dim=(2, 3))                                        # [K, K]
        q = x_out.mean(dim=1).reshape(-1)
        balance = math.log(q.numel()) + (q * q.log()).sum()
        return self.kl(joint) + self.lamda * balance

    def kl(self, joint_p):
        k = joint_p.size(0)
        return -(torch.log(joint_p.diagonal() + 1e-16)).sum() / (k * k)


def _class_major(prediction: Tensor) -> Tensor:
    """[N, K, ...] -> [K, N * ...] (classification and segmentation inputs alike)"""
    return prediction.moveaxis(0, 1).reshape(prediction.shape[1], -1)


class _IMSATEntropies(torch.autograd.Function):
    """prediction [N, K, ...] -> (marginal entropy, conditional entropy) in one streaming kernel each way (cy_imsat_fwd /
    cy_imsat_bwd) instead of the eager graph's ~10 passes over the map."""

    @staticmethod
    def forward(ctx, prediction, eps):
        from .. import _lib as L
        lib = L.lib()
        p = prediction.contiguous()
        N, K = p.shape[0], p.shape[1]
        S = p.numel() // (N * K)
        with L.guard(p):
            out2 = torch.empty(2, dtype=torch.float32, device=p.device)
            q = torch.empty(K, dtype=torch.float32, device=p.device)
            ws_bytes = lib.cy_imsat_workspace_bytes(K)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
            L.check(lib.cy_imsat_fwd(p.data_ptr(), L.dtype_code(p), N, K, S, float(eps), out2.data_ptr(), q.data_ptr(), ws.data_ptr(),
                                     ws_bytes, L.stream_ptr(p.device)), "cy_imsat_fwd")
        ctx.save_for_backward(p, q)
        ctx.eps = float(eps)
        return out2[0].clone(), out2[1].clone()

    @staticmethod
    def backward(ctx, g_marginal, g_conditional):
        from .. import _lib as L
        lib = L.lib()
        p, q = ctx.saved_tensors
        N, K = p.shape[0], p.shape[1]
        S = p.numel() // (N * K)
        with L.guard(p):
            g2 = torch.stack((g_marginal.detach().to(torch.float32).reshape(()), g_conditional.detach().to(torch.float32).reshape(())))
            grad = torch.empty_like(p)
            L.check(lib.cy_imsat_bwd(p.data_ptr(), L.dtype_code(p), N, K, S, ctx.eps, q.data_ptr(), g2.data_ptr(), grad.data_ptr(),
                                     L.stream_ptr(p.device)), "cy_imsat_bwd")
        return grad, None


def imsat_with_entropy(prediction: Tensor):
    """discreteMI.py:288-297 -> (entropy of the mean prediction, mean entropy of the predictions), eps 1e-8.
    CUDA float32 / half maps with K <= 64 take the streaming kernels; anything else (CPU tensors, float64 — the host-side
    mirror the CPU fixtures pin) evaluates the same two formulas with stock torch ops."""
    if prediction.is_cuda and prediction.dtype in (torch.float32, torch.bfloat16, torch.float16) and prediction.shape[1] <= 64:
        return _IMSATEntropies.apply(prediction, 1e-8)
    per_class = _class_major(prediction)
    marginal = entropy_criterion(per_class.mean(1, keepdim=True).t()).mean()
    conditional = entropy_criterion(per_class.t()).mean()
    return marginal, conditional


def imsat_loss(prediction: Tensor, lamda: float = 1.0):
    """discreteMI.py:275-285: -(lamda * H(mean prediction) - mean H(prediction))."""
    marginal, conditional = imsat_with_entropy(prediction)
    return conditional - lamda * marginal


class IMSATLoss(nn.Module):
    """discreteMI.py:20-52: the IMSAT objective of one map, or the average over the two views when both are given."""

    def __init__(self, lamda: float = 1.0, eps: float = sys.float_info.epsilon):
        super().__init__()
        self.eps = float(eps)
        self.lamda = float(lamda)

    def forward(self, x_out: Tensor, x_tf_out: Tensor = None):
        single = x_tf_out is None
        other = x_out if single else x_tf_out
        assert len(x_out.shape) == 2, x_out.shape
        assert simplex(x_out), f"x_out not normalized."
        assert simplex(other), f"x_tf_out not normalized."
        self.x_out, self.x_tf_out = x_out, other
        if single:
            return imsat_loss(x_out, lamda=self.lamda)
        return 0.5 * (imsat_loss(x_out, lamda=self.lamda) + imsat_loss(other, lamda=self.lamda))

    def get_joint_matrix(self):
        bn, k = self.x_out.shape
        return compute_joint_2D_with_padding_zeros(self.x_out.reshape(bn, k, 1, 1), self.x_tf_out.reshape(bn, k, 1, 1),
                                                   symmetric=False).squeeze().detach().cpu().numpy()


class IMSATDynamicWeight(IMSATLoss):
    """discreteMI.py:55-87: -w * H(mean) + mean H, where the buffer w drifts by 0.01 * (log K - H(mean)) per call
    (``dynamic_weight`` is part of the reference's state_dict)."""

    def __init__(self, lamda: float = 1.0, use_dynamic: bool = True, eps: float = sys.float_info.epsilon):
        super().__init__(lamda, eps)
        self.register_buffer("dynamic_weight", torch.tensor(lamda))
        self.use_dynamic_weight = use_dynamic

    def forward(self, x_out: Tensor, **kwargs):
        self.dynamic_weight = self.dynamic_weight.to(device=x_out.device, dtype=x_out.dtype)
        assert len(x_out.shape) == 2, x_out.shape
        assert simplex(x_out), f"x_out not normalized."
        self.x_out = self.x_tf_out = x_out
        marginal, conditional = imsat_with_entropy(x_out)
        value = conditional - self.dynamic_weight * marginal
        if self.use_dynamic_weight:
            with torch.no_grad():
                self.dynamic_weight = self.dynamic_weight + 0.01 * (math.log(x_out.shape[1]) - marginal.detach())
        return value
