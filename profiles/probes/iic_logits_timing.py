"""cfg3 (32 x 10 x 224 x 224 fp32, padding 1): the adjoint with the softmax backward fused into its epilogue
(cy_iic_bwd_logits_heads) vs the plain adjoint followed by torch's softmax backward on both maps; and the whole criterion step
from logits both ways (torch softmax forward in both)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from contrast_you_b200 import _lib as L  # noqa: E402
from contrast_you_b200.losses.discreteMI import IIDSegmentationLoss  # noqa: E402


def timed(fn, flush, reps=30, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    lib = L.lib()
    B, K, H, W, pad = 32, 10, 224, 224, 1
    torch.manual_seed(0)
    lx, ly = 2 * torch.randn(B, K, H, W, device=dev), 2 * torch.randn(B, K, H, W, device=dev)
    px, py = lx.softmax(1), ly.softmax(1)
    dj = torch.randn(K, K, 3, 3, device=dev) * 1e-6
    one = torch.ones(1, device=dev)
    dx, dy = torch.empty_like(px), torch.empty_like(py)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = L.stream_ptr()
    arr = lambda *ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

    def plain():
        L.check(lib.cy_iic_bwd(px.data_ptr(), py.data_ptr(), 0, B, K, H, W, pad, dj.data_ptr(), one.data_ptr(), dx.data_ptr(),
                               dy.data_ptr(), st), "bwd")

    def fused():
        L.check(lib.cy_iic_bwd_logits_heads(arr(px), arr(py), 1, 0, B, K, H, W, pad, dj.data_ptr(), 0, one.data_ptr(), 1.0, arr(dx),
                                            arr(dy), st), "bwd_logits")

    def plain_then_torch():
        plain()
        torch._softmax_backward_data(dx, px, 1, torch.float32)
        torch._softmax_backward_data(dy, py, 1, torch.float32)

    print(f"adjoint alone                      {timed(plain, flush):7.1f} us")
    print(f"adjoint + fused softmax backward   {timed(fused, flush):7.1f} us")
    print(f"adjoint, then torch softmax bwd x2 {timed(plain_then_torch, flush):7.1f} us")
    crit = IIDSegmentationLoss(padding=pad)

    def step_ref():
        a, b = lx.detach().requires_grad_(), ly.detach().requires_grad_()
        crit(torch.softmax(a, 1), torch.softmax(b, 1)).backward()

    def step_fused():
        a, b = lx.detach().requires_grad_(), ly.detach().requires_grad_()
        crit.forward_logits(a, b).backward()

    print(f"logits -> loss -> dlogits, torch softmax both ways  {timed(step_ref, flush):7.1f} us")
    print(f"logits -> loss -> dlogits, forward_logits           {timed(step_fused, flush):7.1f} us")


if __name__ == "__main__":
    main()
