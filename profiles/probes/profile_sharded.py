"""torch.profiler table of one sharded InfoNCE step (rank 0), to see what the per-step constant is made of.
torchrun --nproc-per-node 2 profiles/probes/profile_sharded.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from contrast_you_b200 import distributed as cyd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
N, d = 65536, 256
n_loc = N // 2 // world
g = torch.Generator().manual_seed(rank)
f1 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
f2 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
lab = torch.randint(0, 4096, (n_loc,), generator=g).to(torch.int32).to(dev)
crit = cyd.ShardedSupConLoss()
def step():
    a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
    loss = crit(a, b, target=lab); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
import time
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
if rank == 0: print(f"world {world}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step wall")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))
dist.destroy_process_group()
