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
        assert input_.shape.__len__() >= 2
        b, _, *s = input_.shape
        assert simplex(input_), f"Entropy input should be a simplex"
        e = input_ * (input_ + self._eps).log()
        e = -1.0 * e.sum(1)
        assert e.shape == torch.Size([b, *s])
        if self._reduction == "mean":
            return e.mean()
        elif self._reduction == "sum":
            return e.sum()
        return e


entropy_criterion = Entropy(reduction="none", eps=1e-8)      # semi_seg/hooks/midl.py:13


class RedundancyCriterion(nn.Module):
    """redundancy_reduction.py:12-54: cross-entropy of the padding-0 joint against alpha * I/k + (1 - alpha) * joint,
    plus the marginal-entropy constraint."""

    def __init__(self, *, eps: float = 1e-5, symmetric: bool = True, lamda: float = 1, alpha: float) -> None:
        super().__init__()
        self._eps = eps
        self.symmetric = symmetric
        self.lamda = lamda
        self.alpha = alpha

    def forward(self, x_out: Tensor, x_tf_out: Tensor):
        k = x_out.shape[1]
        p_i_j = compute_joint_2D_with_padding_zeros(x_out, x_tf_out, symmetric=self.symmetric)
        p_i_j = p_i_j.view(k, k)
        self._p_i_j = p_i_j
        target = ((self.onehot_label(k=k, device=p_i_j.device) / k) * self.alpha + p_i_j * (1 - self.alpha))
        p_i = p_i_j.sum(dim=1).view(k, 1).expand(k, k)
        p_j = p_i_j.sum(dim=0).view(1, k).expand(k, k)
        constrained = (-p_i_j * (- self.lamda * torch.log(p_j + self._eps) - self.lamda * torch.log(p_i + self._eps))).sum()
        pseudo_loss = -(target * (p_i_j + self._eps).log()).sum()
        return pseudo_loss + constrained

    @staticmethod
    def onehot_label(k, device):
        return torch.eye(k, device=device, dtype=torch.bool)

    def kl_criterion(self, dist: Tensor, prior: Tensor):
        return -(prior * (dist + self._eps).log() + (1 - prior) * (1 - dist + self._eps).log()).mean()

    def get_joint_matrix(self):
        if not hasattr(self, "_p_i_j"):
            raise RuntimeError()
        return self._p_i_j.detach().cpu().numpy()

    def set_ratio(self, alpha: float):
        assert 0 <= alpha <= 1, alpha
        self.alpha = alpha


class PUISegLoss(nn.Module):
    """pica_loss.py:43-80: the same conv-joint as IIC (padding p), min-shifted by 1e-16, normalised per displacement,
    symmetrised, averaged over displacements; cross-entropy of its diagonal + the (literal) balance term."""

    def __init__(self, lamda=2.0, padding=3):
        super().__init__()
        self.lamda = lamda
        self.padding = padding

    def forward(self, x_out, x_tf_out):
        assert x_out.shape == x_tf_out.shape, ('Inputs are required to have same shape')
        p_i_j = raw_joint(x_out, x_tf_out, self.padding)                    # [k, k, T, T] == F.conv2d(x^T, weight=y^T)
        p_i_j = p_i_j - p_i_j.min().detach() + 1e-16
        p_i_j = p_i_j.permute(2, 3, 0, 1)
        p_i_j = p_i_j / p_i_j.sum(dim=3, keepdim=True).sum(dim=2, keepdim=True)
        p_i_j = (p_i_j + p_i_j.permute(0, 1, 3, 2)) / 2.0
        p_i_j = p_i_j.mean(dim=[0, 1])
        loss_ce = self.kl(p_i_j)
        # the reference takes the mean over dim 0 of the PERMUTED map (k, n, h, w), i.e. over the classes (:70)
        p = x_out.mean(1).view(-1)
        loss_ne = math.log(p.size(0)) + (p * p.log()).sum()
        return loss_ce + self.lamda * loss_ne

    def kl(self, joint_p):
        diagnal = torch.eye(joint_p.size(0), device=joint_p.device, dtype=torch.float)
        return (-diagnal * torch.log(joint_p + 1e-16)).mean()


def imsat_loss(prediction: Tensor, lamda: float = 1.0):
    """discreteMI.py:275-285."""
    pred = prediction.moveaxis(0, 1).reshape(prediction.shape[1], -1)
    margin = pred.mean(1, keepdims=True)
    mi = -entropy_criterion(pred.t()).mean() + entropy_criterion(margin.t()).mean() * lamda
    return -mi


def imsat_with_entropy(prediction: Tensor):
    """discreteMI.py:288-297."""
    pred = prediction.moveaxis(0, 1).reshape(prediction.shape[1], -1)
    margin = pred.mean(1, keepdims=True)
    return entropy_criterion(margin.t()).mean(), entropy_criterion(pred.t()).mean()


class IMSATLoss(nn.Module):
    """discreteMI.py:20-52."""

    def __init__(self, lamda: float = 1.0, eps: float = sys.float_info.epsilon):
        super().__init__()
        self.eps = float(eps)
        self.lamda = float(lamda)

    def forward(self, x_out: Tensor, x_tf_out: Tensor = None):
        idenity_input = False
        if x_tf_out is None:
            idenity_input = True
            x_tf_out = x_out
        assert len(x_out.shape) == 2, x_out.shape
        assert simplex(x_out), f"x_out not normalized."
        assert simplex(x_tf_out), f"x_tf_out not normalized."
        self.x_out = x_out
        self.x_tf_out = x_tf_out
        if not idenity_input:
            return 0.5 * (imsat_loss(x_out, lamda=self.lamda) + imsat_loss(x_tf_out, lamda=self.lamda))
        return imsat_loss(x_out, lamda=self.lamda)

    def get_joint_matrix(self):
        bn, k = self.x_out.shape
        return compute_joint_2D_with_padding_zeros(self.x_out.reshape(bn, k, 1, 1), self.x_tf_out.reshape(bn, k, 1, 1),
                                                   symmetric=False).squeeze().detach().cpu().numpy()


class IMSATDynamicWeight(IMSATLoss):
    """discreteMI.py:55-87 (the ``dynamic_weight`` buffer is part of the reference's state_dict)."""

    def __init__(self, lamda: float = 1.0, use_dynamic: bool = True, eps: float = sys.float_info.epsilon):
        super().__init__(lamda, eps)
        self.register_buffer("dynamic_weight", torch.tensor(lamda))
        self.use_dynamic_weight = use_dynamic

    def forward(self, x_out: Tensor, **kwargs):
        device, dtype = x_out.device, x_out.dtype
        self.dynamic_weight = self.dynamic_weight.to(device).to(dtype)
        K = x_out.shape[1]
        x_tf_out = x_out
        assert len(x_out.shape) == 2, x_out.shape
        assert simplex(x_out), f"x_out not normalized."
        assert simplex(x_tf_out), f"x_tf_out not normalized."
        self.x_out = x_out
        self.x_tf_out = x_tf_out
        marg, cond = imsat_with_entropy(x_out)
        mi = self.dynamic_weight * marg * -1.0 + cond
        if self.use_dynamic_weight:
            with torch.no_grad():
                increment = (math.log(K) - marg.detach()) * 0.01
                self.dynamic_weight = self.dynamic_weight + increment
        return mi
