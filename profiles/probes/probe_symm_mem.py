"""Probe: does torch's symmetric memory (cuMem + peer mapping) rendezvous on this box, and what do a signal-pad barrier and a
peer-store all-gather cost next to NCCL?  Launch under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/probes/probe_symm_mem.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    group = dist.group.WORLD
    nbytes = 64 << 20
    t0 = time.perf_counter()
    buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(buf, group)
    torch.cuda.synchronize()
    print(f"[rank {rank}] rendezvous ok in {time.perf_counter() - t0:.2f} s: world {hdl.world_size} rank {hdl.rank} "
          f"buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} signal_pad {hdl.signal_pad_size} B multicast {hdl.has_multicast_support}",
          flush=True)

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3      # us

    # correctness: every rank writes its block into every peer through the mapped pointers
    m = (32 << 20) // world
    own = buf[rank * m:(rank + 1) * m]
    own.fill_(rank + 1)
    hdl.barrier(channel=0)
    for p in range(world):
        if p != rank:
            peer = hdl.get_buffer(p, (nbytes,), torch.uint8)
            peer[rank * m:(rank + 1) * m].copy_(own)
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    got = [int(buf[r * m].item()) for r in range(world)]
    assert got == [r + 1 for r in range(world)], got
    print(f"[rank {rank}] peer-store all-gather correct: {got}", flush=True)

    t_bar = timeit(lambda: hdl.barrier(channel=0))

    def push():
        for p in range(world):
            if p != rank:
                hdl.get_buffer(p, (nbytes,), torch.uint8)[rank * m:(rank + 1) * m].copy_(own, non_blocking=True)
        hdl.barrier(channel=0)
    t_push = timeit(push)
    full = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    t_nccl = timeit(lambda: dist.all_gather_into_tensor(full, full[rank * m:(rank + 1) * m]))
    small = torch.empty(65536 * 16, dtype=torch.uint8, device=dev)
    ms = small.numel() // world
    t_nccl_small = timeit(lambda: dist.all_gather_into_tensor(small, small[rank * ms:(rank + 1) * ms]))
    j = torch.zeros(900, dtype=torch.float64, device=dev)
    t_ar = timeit(lambda: dist.all_reduce(j))
    if rank == 0:
        print(f"world {world}: symm barrier {t_bar:.1f} us | peer-store all-gather 32 MiB (+barrier) {t_push:.1f} us | "
              f"NCCL all-gather 32 MiB {t_nccl:.1f} us | NCCL all-gather 1 MiB {t_nccl_small:.1f} us | NCCL all-reduce 900 f64 {t_ar:.1f} us",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa
        import traceback
        traceback.print_exc()
        sys.exit(1)
