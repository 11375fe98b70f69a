"""Sweep of the InfoNCE forward's column-split count (CY_FWD_SPLITS) for the full problem (1 GPU) and for one rank's strip of
an 8-GPU run: cy_infonce_fwd through the C ABI, CUDA events, L2 flushed.  python profiles/probes/fwd_splits_sweep.py"""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from contrast_you_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0"); N, d = 65536, 256
g = torch.Generator().manual_seed(0)
z = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
lab = torch.sort(torch.randint(0, 4096, (N // 2,), generator=g).to(torch.int32).repeat(2))[0].to(dev)
wsb = lib.cy_infonce_workspace_bytes(N, d, 1, 0, 2); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
stats = torch.empty(4, N, device=dev); xstat = torch.empty(N, 4, device=dev); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = L.stream_ptr(dev)
out = []
for rb, re in ((0, N), (0, N // 8), (3 * N // 8, 4 * N // 8)):
    def f():
        L.check(lib.cy_infonce_fwd(z.data_ptr(), 1, N, d, d, lab.data_ptr(), None, rb, re, 1 / 0.07, 0, 2, stats.data_ptr(), xstat.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
    for _ in range(3): f()
    ts = []
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); out.append(ts[len(ts) // 2])
print("CY_FWD_SPLITS=%%s  full %%.3f ms  strip0 %%.4f ms  strip3 %%.4f ms" %% (__import__("os").environ.get("CY_FWD_SPLITS", "default"), *out))
''' % ROOT
for s in ["", "1", "2", "3", "4", "5", "6", "7", "8", "9", "10", "12", "14", "16"]:
    env = dict(os.environ)
    if s: env["CY_FWD_SPLITS"] = s
    else: env.pop("CY_FWD_SPLITS", None)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print((r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1], flush=True)
