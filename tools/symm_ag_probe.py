"""probe: pull-based all-gather over torch symmetric memory (copy engines, P2P over NVLink) vs NCCL, for the row blocks
the partitioned encoder exchanges.  torchrun --nproc-per-node N tools/symm_ag_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as sm  # noqa: E402

N, F = 2927963, 50
blk = (N + world - 1) // world
x = torch.full((blk, F), float(rank + 1), device=dev) + torch.arange(F, device=dev)
gname = dist.group.WORLD.group_name


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def nccl():
    full = torch.empty(world * blk, F, device=dev)
    dist.all_gather_into_tensor(full, x)
    return full


def symm():
    out = torch.ops.symm_mem._low_contention_all_gather(x, gname)
    return torch.ops._c10d_functional.wait_tensor(out)


ms_n, ref = timeit(nccl)
try:
    try:
        sm.enable_symm_mem_for_group(gname)
    except Exception as ex:       # newer versions enable it implicitly
        if rank == 0:
            print("enable_symm_mem_for_group:", repr(ex)[:120])
    ms_s, got = timeit(symm)
    ok = bool(torch.equal(ref, got))
    if rank == 0:
        recv = (world - 1) * blk * F * 4
        print(f"world {world}: nccl all_gather {ms_n:.3f} ms ({recv / ms_n / 1e6:.0f} GB/s recv), "
              f"symmetric-memory pull {ms_s:.3f} ms ({recv / ms_s / 1e6:.0f} GB/s recv), equal: {ok}", flush=True)
except Exception as ex:
    if rank == 0:
        print("symmetric memory unavailable:", repr(ex)[:400], flush=True)
dist.destroy_process_group()
