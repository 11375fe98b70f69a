"""Kernel-level IIC micro-benchmark (config 3 by default): times cy_iic_joint / cy_iic_bwd through the C ABI with CUDA
events and optionally dumps the results, so that two kernel variants (CY_IIC_MMA=0/1, CY_IIC_TMA=0/1) can be compared
across processes.  Usage: python profiles/probes/iic_micro.py [--B 32 --K 10 --H 224 --W 224 --pad 1] [--reps 20] [--out f.npz]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from contrast_you_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for k, v in (("B", 32), ("K", 10), ("H", 224), ("W", 224), ("pad", 1), ("reps", 20)):
        ap.add_argument("--" + k, type=int, default=v)
    ap.add_argument("--out", default=None)
    ap.add_argument("--flush", default="write", choices=["write", "read", "rotate", "none"],
                    help="between timed launches: write a 256 MiB buffer (dirty L2 lines are written back DURING the timed kernel), "
                         "read it (clean lines), rotate over 3 input/output sets (inputs larger than L2, no flush), or nothing")
    ap.add_argument("--ref", default=None, help="npz written by an earlier run: report max relative differences")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = L.lib()
    st = L.stream_ptr()
    g = torch.Generator(device="cpu").manual_seed(1)
    x = (2 * torch.randn(a.B, a.K, a.H, a.W, generator=g)).softmax(1).to(dev)
    y = (2 * torch.randn(a.B, a.K, a.H, a.W, generator=g)).softmax(1).to(dev)
    T = 2 * a.pad + 1
    joint = torch.empty(a.K, a.K, T, T, device=dev, dtype=torch.float64)
    wsb = lib.cy_iic_workspace_bytes(a.B, a.K, a.H, a.W, a.pad)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    dj = torch.randn(a.K, a.K, T, T, generator=g).to(dev)
    one = torch.ones(1, device=dev)
    dx, dy = torch.empty_like(x), torch.empty_like(y)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nset = 3 if a.flush == "rotate" else 1
    xs = [x] + [x.clone() for _ in range(nset - 1)]
    ys = [y] + [y.clone() for _ in range(nset - 1)]
    dxs = [dx] + [torch.empty_like(x) for _ in range(nset - 1)]
    dys = [dy] + [torch.empty_like(y) for _ in range(nset - 1)]
    cur = [0]

    def joint_call():
        x, y = xs[cur[0]], ys[cur[0]]
        L.check(lib.cy_iic_joint(x.data_ptr(), y.data_ptr(), 0, a.B, a.K, a.H, a.W, a.pad, joint.data_ptr(), ws.data_ptr(), wsb, st), "joint")

    def bwd_call():
        x, y, dx, dy = xs[cur[0]], ys[cur[0]], dxs[cur[0]], dys[cur[0]]
        L.check(lib.cy_iic_bwd(x.data_ptr(), y.data_ptr(), 0, a.B, a.K, a.H, a.W, a.pad, dj.data_ptr(), one.data_ptr(),
                               dx.data_ptr(), dy.data_ptr(), st), "bwd")

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.reps):
            if a.flush == "write":
                flush.zero_()
            elif a.flush == "read":
                flush.sum()
            cur[0] = (cur[0] + 1) % nset
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))

    jm, jmin = timeit(joint_call)
    bm, bmin = timeit(bwd_call)
    by = 2 * a.B * a.K * a.H * a.W * 4
    cur[0] = 0
    print(f"flush={a.flush} env MMA={os.environ.get('CY_IIC_MMA', '1')} TMA={os.environ.get('CY_IIC_TMA', '1')}  "
          f"joint {jm * 1e3:.1f} us (min {jmin * 1e3:.1f}; {by / jm / 1e6:.0f} GB/s)  "
          f"bwd {bm * 1e3:.1f} us (min {bmin * 1e3:.1f}; {2 * by / bm / 1e6:.0f} GB/s)  "
          f"fwd+bwd {3 * by / (jm + bm) / 1e6:.0f} GB/s")
    res = {"joint": joint.cpu().numpy(), "dx": dx[:2].cpu().numpy(), "dy": dy[:2].cpu().numpy(),
           "dx_last": dx[-1, :, -12:, :].cpu().numpy(), "dy_last": dy[-1, :, -12:, :].cpu().numpy()}
    if a.out:
        np.savez(a.out, **res)
    if a.ref:
        r = np.load(a.ref)
        for k in res:
            d = np.abs(res[k].astype(np.float64) - r[k]).max() / np.abs(r[k]).max()
            rel = np.abs(res[k].astype(np.float64) / r[k] - 1).max() if k == "joint" else float("nan")
            print(f"  {k}: max-norm relative diff {d:.3e}  elementwise rel {rel:.3e}")


if __name__ == "__main__":
    main()
