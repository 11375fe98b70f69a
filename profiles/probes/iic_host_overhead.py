import sys, time, torch
sys.path.insert(0, '/root/repo')
from contrast_you_b200.losses import IIDSegmentationLoss
dev = torch.device('cuda:0')
for B in (1, 32):
    x = (2 * torch.randn(B, 10, 224, 224, device=dev)).softmax(1)
    y = (2 * torch.randn(B, 10, 224, 224, device=dev)).softmax(1)
    iic = IIDSegmentationLoss(padding=1)
    def step():
        xa = x.detach().requires_grad_(); ya = y.detach().requires_grad_()
        loss = iic(xa, ya); loss.backward(); return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host issue {1e6*(t1-t0)/n:.0f} us/step, gpu events {1e3*e0.elapsed_time(e1)/n:.0f} us/step, wall {1e6*(t2-t0)/n:.0f}")

if len(sys.argv) > 1 and sys.argv[1] == "profile":
    import cProfile, pstats
    x = (2 * torch.randn(1, 10, 224, 224, device=dev)).softmax(1)
    y = (2 * torch.randn(1, 10, 224, 224, device=dev)).softmax(1)
    iic = IIDSegmentationLoss(padding=1)
    def step():
        xa = x.detach().requires_grad_(); ya = y.detach().requires_grad_()
        loss = iic(xa, ya); loss.backward(); return loss
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300): step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
