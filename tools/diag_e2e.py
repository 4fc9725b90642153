import os, time, torch, numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
rank=int(os.environ.get("RANK",0)); world=int(os.environ.get("WORLD_SIZE",1)); local=int(os.environ.get("LOCAL_RANK",0))
torch.cuda.set_device(local)
if world>1: dist.init_process_group("nccl", device_id=torch.device("cuda",local))
a=torch.empty(42_000_000//8, dtype=torch.int64).pin_memory()
d=torch.empty_like(a, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); 
    if world>1: dist.barrier()
    t0=time.perf_counter()
    for _ in range(10): d.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()-t0
    print(f"rank {rank} rep {rep}: H2D {10*a.numel()*8/t/1e9:.1f} GB/s", flush=True)
# chunked with side streams
side=[torch.cuda.Stream() for _ in range(2)]
for rep in range(3):
    torch.cuda.synchronize()
    if world>1: dist.barrier()
    t0=time.perf_counter()
    for _ in range(10):
        for c in range(4):
            with torch.cuda.stream(side[c%2]):
                n=a.numel()//4
                x=a[c*n:(c+1)*n].to("cuda", non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()-t0
    print(f"rank {rank} rep {rep}: chunked to() {10*a.numel()*8/t/1e9:.1f} GB/s", flush=True)
print(rank, "affinity", len(os.sched_getaffinity(0)), flush=True)
if world>1: dist.destroy_process_group()
