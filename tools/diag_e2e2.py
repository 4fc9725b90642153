import os, time, torch, numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import bench
from muygpys_b200 import ops
from muygpys_b200.examples.from_indices import regress_from_indices
from muygpys_b200.gp import MuyGPS
from muygpys_b200.gp.deformation import Isotropy, l2
from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter
from muygpys_b200.gp.kernels import Matern
from muygpys_b200.gp.noise import HomoscedasticNoise
rank=int(os.environ.get("RANK",0)); world=int(os.environ.get("WORLD_SIZE",1)); local=int(os.environ.get("LOCAL_RANK",0))
torch.cuda.set_device(local); dev=torch.device("cuda",local)
use_nccl = os.environ.get("USE_NCCL","1")=="1"
if world>1 and use_nccl: dist.init_process_group("nccl", device_id=dev)
x_h,y_h,_=bench.make_data(2, n=200000)
q_h=np.random.default_rng(rank).uniform(size=(100000,2))
x,y,q=(torch.as_tensor(a).to(dev) for a in (x_h,y_h,q_h))
nn,_=ops.knn(x,q,50)
model = MuyGPS(kernel=Matern(smoothness=Parameter(1.5), deformation=Isotropy(l2, Parameter(0.1))), noise=HomoscedasticNoise(1e-3), scale=AnalyticScale())
q_pin=torch.as_tensor(q_h).pin_memory(); nn_pin=nn.cpu().pin_memory(); idx_pin=torch.arange(100000).pin_memory()
mp_=torch.empty(100000,dtype=torch.float64).pin_memory(); vp=torch.empty(100000,dtype=torch.float64).pin_memory()
def step():
    m,v=regress_from_indices(model, idx_pin, nn_pin, q_pin, x, y); mp_.copy_(m,non_blocking=True); vp.copy_(v,non_blocking=True)
for tag in ("plain","sampler"):
    s=None
    if tag=="sampler":
        s=bench.ClockSampler(local); s.start()
    for _ in range(5): step()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(20):
        t0=time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    print(f"rank {rank} {tag}: e2e wall ms median {np.median(ts):.2f} max {max(ts):.2f} OMP={os.environ.get('OMP_NUM_THREADS')}", flush=True)
    if s: s.stop()
if world>1 and use_nccl: dist.destroy_process_group()
