import sys, time, torch
sys.path.insert(0, '/root/repo')
from contrast_you_b200.losses import SupConLoss1, IIDSegmentationLoss
dev = torch.device('cuda:0')
for n, dt in ((18, torch.float32), (90, torch.float32), (64, torch.float32), (512, torch.bfloat16)):
    f1 = torch.nn.functional.normalize(torch.randn(n, 256, device=dev), dim=1).to(dt)
    f2 = torch.nn.functional.normalize(torch.randn(n, 256, device=dev), dim=1).to(dt)
    lab = torch.randint(0, 3, (n,)).tolist()
    crit = SupConLoss1()
    def step():
        a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
        l = crit(a, b, target=lab); l.backward(); return l
    for _ in range(10): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): step()
    torch.cuda.synchronize(); print(f"SupConLoss1 n={n} {dt}: {(time.perf_counter()-t0)/200*1e6:.0f} us/step (fwd+bwd, incl. host checks)")
x = torch.randn(10, 10, 224, 224, device=dev).softmax(1); y = torch.randn(10, 10, 224, 224, device=dev).softmax(1)
crit = IIDSegmentationLoss(padding=1)
def step():
    a, b = x.detach().requires_grad_(), y.detach().requires_grad_()
    l = crit(a, b); l.backward(); return l
for _ in range(10): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize(); print(f"IIDSegmentationLoss 10x10x224x224: {(time.perf_counter()-t0)/200*1e6:.0f} us/step")

# the same small InfoNCE step, strict vs deferred checks vs CUDA graph replay
n = 18
f1 = torch.nn.functional.normalize(torch.randn(n, 256, device=dev), dim=1)
f2 = torch.nn.functional.normalize(torch.randn(n, 256, device=dev), dim=1)
labt = torch.randint(0, 3, (n,), device=dev, dtype=torch.int32)
class W(torch.nn.Module):
    def __init__(s, **kw):
        super().__init__(); s.crit = SupConLoss1(**kw)
    def forward(s, x, y):
        return s.crit(x, y, target=labt)
def bench(fn, name):
    def step():
        a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
        l = fn(a, b); l.backward(); return l
    for _ in range(10): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): step()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter()-t0)/300*1e6:.0f} us/step")
bench(W(), "n=18 strict (one host read per forward)")
bench(W(deferred_checks=True), "n=18 deferred checks")
gm = torch.cuda.make_graphed_callables(W(deferred_checks=True), (f1.clone().requires_grad_(), f2.clone().requires_grad_()))
bench(gm, "n=18 deferred checks + CUDA graph")
