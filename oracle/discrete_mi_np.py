"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU oracle (numpy) for the IIC discrete mutual-information losses of the reference,
``contrastyou/losses/discreteMI.py`` (``IIDSegmentationLoss``, ``IIDLoss`` and the joint builders).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it.  Pinned against ``tests/golden/iic_*.npz`` / ``iid_*.npz`` produced by the reference itself
(tests/golden/make_golden.py).  Parity status: PINNED.

The reference builds the joint with ``F.conv2d`` using one softmax map as the "image" and the other as a
B x H x W "filter" (discreteMI.py:225-232); the restatement below is the explicit shifted contraction that
call evaluates, and the gradients are the analytic ones autograd yields for the reference graph.
"""
import numpy as np


def raw_joint_2d(x, y, padding):
    """discreteMI.py:227-232 — J[k1,k2,dy,dx] = sum_{b,h,w} x[b,k1,h+dy-p,w+dx-p] * y[b,k2,h,w], x zero-padded."""
    B, K, H, W = x.shape
    p = int(padding)
    T = 2 * p + 1
    xp = np.zeros((B, K, H + 2 * p, W + 2 * p), dtype=x.dtype)
    xp[:, :, p:p + H, p:p + W] = x
    J = np.empty((K, K, T, T), dtype=x.dtype)
    for dy in range(T):
        for dx in range(T):
            J[:, :, dy, dx] = np.einsum("bihw,bjhw->ij", xp[:, :, dy:dy + H, dx:dx + W], y, optimize=True)
    return J


def joint_epilogue(J, *, padding, symmetric, n_pixels=None):
    """From the raw joint [K,K,T,T] to the normalised p_i_j [T,T,K,K].

    padding > 0 (compute_joint_2D, :233-243): subtract the detached global min, add 1e-8, permute, normalise
    every displacement to mass 1, optionally symmetrise over (k1,k2), normalise the whole tensor to mass 1.
    padding == 0 (compute_joint_2D_with_padding_zeros, :246-261): J / n_pixels, optional symmetrise; no shift and
    no renormalisation."""
    if padding > 0:
        A = J - J.min() + 1e-8
        Q = np.transpose(A, (2, 3, 0, 1))
        s = Q.sum(axis=(2, 3), keepdims=True)
        Bm = Q / s
        C = (Bm + np.transpose(Bm, (0, 1, 3, 2))) / 2.0 if symmetric else Bm
        total = C.sum()
        return C / total, dict(s=s, B=Bm, total=total)
    Q = np.transpose(J, (2, 3, 0, 1)) / n_pixels
    C = (Q + np.transpose(Q, (0, 1, 3, 2))) / 2.0 if symmetric else Q
    return C, dict()


def mi_loss_and_grad_wrt_pij(P, lamda, eps):
    """discreteMI.py:154-165 — loss = sum -P (log(P+eps) - lam log(p_i+eps) - lam log(p_j+eps)) / T^2 where
    p_i = P.sum(dim=2) (over k1) and p_j = P.sum(dim=3) (over k2) of the [T,T,K,K] tensor.  Also dLoss/dP."""
    T = P.shape[0]
    a = P.sum(axis=2, keepdims=True)     # "p_i_mat": function of k2
    b = P.sum(axis=3, keepdims=True)     # "p_j_mat": function of k1
    terms = -P * (np.log(P + eps) - lamda * np.log(a + eps) - lamda * np.log(b + eps))
    loss = terms.sum() / (T * T)
    g = -(np.log(P + eps) + P / (P + eps)
          - lamda * (np.log(a + eps) + a / (a + eps))
          - lamda * (np.log(b + eps) + b / (b + eps))) / (T * T)
    return loss, g


def grad_wrt_raw_joint(gP, P, aux, *, padding, symmetric, n_pixels=None):
    """Back-propagate dLoss/dP [T,T,K,K] through joint_epilogue to dLoss/dJ [K,K,T,T] (min is detached)."""
    if padding > 0:
        gC = (gP - (gP * P).sum()) / aux["total"]
        gB = (gC + np.transpose(gC, (0, 1, 3, 2))) / 2.0 if symmetric else gC
        gA = (gB - (gB * aux["B"]).sum(axis=(2, 3), keepdims=True)) / aux["s"]
        return np.transpose(gA, (2, 3, 0, 1))
    gQ = (gP + np.transpose(gP, (0, 1, 3, 2))) / 2.0 if symmetric else gP
    return np.transpose(gQ, (2, 3, 0, 1)) / n_pixels


def input_grads(x, y, gJ, padding):
    """Adjoint of raw_joint_2d: dL/dx[b,k1,h',w'] = sum_{k2,dy,dx} gJ[k1,k2,dy,dx] y[b,k2,h'-dy+p,w'-dx+p];
    dL/dy[b,k2,h,w] = sum_{k1,dy,dx} gJ[k1,k2,dy,dx] x[b,k1,h+dy-p,w+dx-p]."""
    B, K, H, W = x.shape
    p = int(padding)
    T = 2 * p + 1
    xp = np.zeros((B, K, H + 2 * p, W + 2 * p), dtype=x.dtype)
    xp[:, :, p:p + H, p:p + W] = x
    gxp = np.zeros_like(xp)
    gy = np.zeros_like(y)
    for dy in range(T):
        for dx in range(T):
            g = gJ[:, :, dy, dx]                                           # [k1, k2]
            gy += np.einsum("ij,bihw->bjhw", g, xp[:, :, dy:dy + H, dx:dx + W], optimize=True)
            gxp[:, :, dy:dy + H, dx:dx + W] += np.einsum("ij,bjhw->bihw", g, y, optimize=True)
    return gxp[:, :, p:p + H, p:p + W], gy


def iid_segmentation_loss(x, y, *, lamda=1.0, padding=0, eps=1e-5, symmetric=False, mask=None, dtype=np.float64):
    """IIDSegmentationLoss.forward (discreteMI.py:139-165) + gradients w.r.t. the (pre-mask) inputs.

    returns dict(loss, grad_x, grad_y, joint [K,K] = p_i_j[0][0] (:152), p_i_j [T,T,K,K], raw_joint, grad_raw_joint)"""
    if padding < 0:
        raise ValueError(padding)                                          # :150-151
    x = np.asarray(x, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    if mask is not None:                                                   # :142-144 (in place in the reference)
        mask = np.asarray(mask, dtype=dtype)
        x = x * mask
        y = y * mask
    B, K, H, W = x.shape
    npx = B * H * W
    J = raw_joint_2d(x, y, padding)
    P, aux = joint_epilogue(J, padding=padding, symmetric=symmetric, n_pixels=npx)
    loss, gP = mi_loss_and_grad_wrt_pij(P, lamda, eps)
    gJ = grad_wrt_raw_joint(gP, P, aux, padding=padding, symmetric=symmetric, n_pixels=npx)
    gx, gy = input_grads(x, y, gJ, padding)
    if mask is not None:
        gx = gx * mask
        gy = gy * mask
    return dict(loss=loss, grad_x=gx, grad_y=gy, joint=P[0][0], p_i_j=P, raw_joint=J, grad_raw_joint=gJ)


def softmax_with_t(logits, T=1.0, dtype=np.float64):
    """SoftmaxWithT.forward (contrastyou/projectors/nn.py:36-44): ``input /= T`` then softmax over dim 1."""
    z = np.asarray(logits, dtype=dtype) / T
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def iid_segmentation_loss_from_logits(logits_x, logits_y, T=1.0, **kw):
    """The discrete-MI hook's criterion call on a cluster head's sub-heads (semi_seg/hooks/discretemi.py:106-111:
    ``sum(criterion(x1, x2) for x1, x2 in zip(prob1, prob2)) / len(prob1)``) with the heads' SoftmaxWithT tail
    (projectors/nn.py:36-44; heads.py:64-70) made explicit: logits_x / logits_y are sequences of S maps [B,K,H,W].

    returns dict(loss, grad_x [S,...], grad_y [S,...]) — gradients w.r.t. the LOGITS: dL/dz = p * (g - sum_k p_k g_k) / T."""
    S = len(logits_x)
    loss, gxs, gys = 0.0, [], []
    for lx, ly in zip(logits_x, logits_y):
        px, py = softmax_with_t(lx, T), softmax_with_t(ly, T)
        r = iid_segmentation_loss(px, py, **kw)
        loss += r["loss"] / S
        for p, g, out in ((px, r["grad_x"], gxs), (py, r["grad_y"], gys)):
            g = g / S
            out.append(p * (g - (p * g).sum(axis=1, keepdims=True)) / T)
    return dict(loss=loss, grad_x=np.stack(gxs), grad_y=np.stack(gys))


def compute_joint(x, y, symmetric=True):
    """discreteMI.py:201-222 — J = sum_b x_b y_b^T, optional symmetrise, normalise to mass 1."""
    J = np.einsum("bi,bj->ij", x, y)
    if symmetric:
        J = (J + J.T) / 2.0
    return J / J.sum()


def iid_loss(x, y, lamb=1.0, dtype=np.float64):
    """IIDLoss.forward (discreteMI.py:101-124): returns dict(loss, loss_no_lamb, p_i_j, grad_x, grad_y) with the
    gradients of ``loss`` (the only output the hooks use, semi_seg/hooks/discretemi.py:56-62)."""
    x = np.asarray(x, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    raw = np.einsum("bi,bj->ij", x, y)
    sym = (raw + raw.T) / 2.0
    tot = sym.sum()
    P = sym / tot
    pi = P.sum(axis=1, keepdims=True)     # :114 rows
    pj = P.sum(axis=0, keepdims=True)     # :115 cols
    e = 1e-10                              # hard-coded in the reference, self.eps is unused (:118-123)

    def f(lam):
        return (-P * (np.log(P + e) - lam * np.log(pj + e) - lam * np.log(pi + e))).sum()

    loss, loss_no_lamb = f(lamb), f(1.0)
    gP = -(np.log(P + e) + P / (P + e)
           - lamb * (np.log(pj + e) + pj / (pj + e))
           - lamb * (np.log(pi + e) + pi / (pi + e)))
    gS = (gP - (gP * P).sum()) / tot
    gR = (gS + gS.T) / 2.0
    gx = y @ gR.T
    gy = x @ gR
    return dict(loss=loss, loss_no_lamb=loss_no_lamb, p_i_j=P, grad_x=gx, grad_y=gy)
