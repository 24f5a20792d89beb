#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/d2h_probe.py : device-to-host bandwidth of every rank, alone and
with all ranks copying at once, with the pinned buffer allocated before / after binding the process
to the GPU's NUMA node (engine.bind_host_to_gpu_node)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from phyloligo_b200 import engine
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
def measure(pinned, reps=4, together=True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pinned.copy_(src, non_blocking=True); torch.cuda.synchronize()
    if world > 1 and together: dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): pinned.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
def report(tag, val):
    t = torch.tensor([val], device="cuda", dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        vals = [float(v.item()) for v in allv]
    else:
        vals = [val]
    if rank == 0:
        print("%-44s per-rank GB/s: %s  sum %.1f" % (tag, " ".join("%.1f" % v for v in vals), sum(vals)), flush=True)
floating = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
floating.fill_(1)
report("all ranks at once, unbound pinned buffer", measure(floating))
node = engine.bind_host_to_gpu_node(local)
bound = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
bound.fill_(1)
report("all ranks at once, buffer on the GPU's node", measure(bound))
report("numa node of each rank", -1.0 if node is None else float(node))
for r in range(world):  # one rank at a time
    if world > 1: dist.barrier()
    v = measure(bound, together=False) if r == rank else 0.0
    if world > 1: dist.barrier()
    report("rank %d alone (bound buffer)" % r, v)
if world > 1:
    dist.destroy_process_group()
