"""bandwidthTest-style probe of the device -> pinned-host path, one process per GPU under torchrun: every rank copies a 1 GiB
device buffer into its own pinned host buffer (cudaMemcpyAsync, CUDA events), first alone (the others idle), then all at
once.  With --bind each rank first pins itself and its host allocations to its GPU's NUMA node (bench.py:numa_bind).
    python -m torch.distributed.run --nproc-per-node 8 tools/d2h_bw.py [--bind]
Prints one JSON line on rank 0: per-rank solo GB/s, per-rank concurrent GB/s, aggregate."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import numa_bind

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
info = numa_bind(local) if "--bind" in sys.argv else {"gpu": local, "bound": False}
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.copy_(d); torch.cuda.synchronize()

def timed(reps=4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

solo = 0.0
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        solo = timed()
if world > 1:
    dist.barrier()
conc = timed(8)
vals = torch.tensor([solo, conc], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
infos = [None] * world
if world > 1:
    dist.all_gather_object(infos, info)
else:
    infos = [info]
if rank == 0:
    print(json.dumps({"bind": "--bind" in sys.argv, "world": world, "solo_gbs": [round(float(v[0]), 2) for v in allv],
                      "concurrent_gbs": [round(float(v[1]), 2) for v in allv], "aggregate_concurrent_gbs": round(sum(float(v[1]) for v in allv), 2),
                      "numa": infos}))
if world > 1:
    dist.destroy_process_group()
