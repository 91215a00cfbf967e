"""Host-link ceiling for the e2e figure (dev tool; VERDICT r01 task 4): every rank copies a pinned 1 GiB buffer to its GPU
10 times, all ranks at once.  Launch like bench.py (`python -m torch.distributed.run --nproc-per-node N tools/time_h2d.py`);
rank 0 prints one JSON line with the per-rank GB/s and what that means for a step that moves 1 GiB per rank."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import numa_pin

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
pin = numa_pin(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28                                     # 1 GiB of fp32
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device=dev)
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
b.record()
b.synchronize()
gbs = 10 * n * 4 / (a.elapsed_time(b) / 1e3) / 1e9
t = torch.tensor([gbs], device=dev, dtype=torch.float64)
out = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(out, t)
else:
    out = [t]
if rank == 0:
    per = [round(float(x), 2) for x in out]
    print(json.dumps({"ranks": world, "h2d_gbs_per_rank": per, "h2d_gbs_total": round(sum(per), 1),
                      "ms_per_gib_slowest_rank": round(1.073741824 / min(per) * 1e3, 2), "host_affinity_rank0": pin}), flush=True)
if world > 1:
    dist.destroy_process_group()
