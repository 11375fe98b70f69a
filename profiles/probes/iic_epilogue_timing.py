"""In-stream duration of cy_iic_epilogue at config 3 (K = 10, padding 1): events around the call alone (warm), and the step's
chain joint -> epilogue -> adjoint with an event after every call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from contrast_you_b200 import _lib as L  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib = L.lib()
    B, K, H, W, pad = 32, 10, 224, 224, 1
    torch.manual_seed(0)
    x = (2 * torch.randn(B, K, H, W, device=dev)).softmax(1)
    y = (2 * torch.randn(B, K, H, W, device=dev)).softmax(1)
    nj = K * K * 9
    joint = torch.empty(nj, dtype=torch.float64, device=dev)
    wsb = lib.cy_iic_workspace_bytes(B, K, H, W, pad)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty(1 + K * K + nj, device=dev)
    one = torch.ones(1, device=dev)
    dx, dy = torch.empty_like(x), torch.empty_like(y)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = L.stream_ptr()
    o = out.data_ptr()

    def j():
        L.check(lib.cy_iic_joint(x.data_ptr(), y.data_ptr(), 0, B, K, H, W, pad, joint.data_ptr(), ws.data_ptr(), wsb, st), "joint")

    def e():
        L.check(lib.cy_iic_epilogue(joint.data_ptr(), 1, K, pad, 0, 1.0, 1e-5, float(B * H * W), o, o + 4, None, o + 4 * (1 + K * K), None, 0,
                                    st), "epilogue")

    def b():
        L.check(lib.cy_iic_bwd(x.data_ptr(), y.data_ptr(), 0, B, K, H, W, pad, o + 4 * (1 + K * K), one.data_ptr(), dx.data_ptr(),
                               dy.data_ptr(), st), "bwd")
    j(); e(); b()
    torch.cuda.synchronize()
    reps = 40
    acc = [0.0, 0.0, 0.0, 0.0]
    for _ in range(reps):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda._sleep(400000)                      # let the host queue the whole chain first
        ev[0].record(); j(); ev[1].record(); e(); ev[2].record(); b(); ev[3].record()
        torch.cuda.synchronize()
        for i in range(3):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        acc[3] += ev[0].elapsed_time(ev[3])
    print("chain, queued behind a gate: joint+reduce %.1f us | epilogue %.1f us | adjoint %.1f us | total %.1f us"
          % tuple(a / reps * 1e3 for a in acc))
    tot = 0.0
    for _ in range(reps):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(100000)
        a.record(); e(); c.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(c)
    print("epilogue alone (warm): %.1f us" % (tot / reps * 1e3))


if __name__ == "__main__":
    main()
