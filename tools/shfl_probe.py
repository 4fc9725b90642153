#!/usr/bin/env python
"""Dev tool: warp-shuffle issue rate and latency on this GPU (mgp_fp64_probe modes 5, 6)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from muygpys_b200 import ops
sms = torch.cuda.get_device_properties(0).multi_processor_count
clk = 1.965e9
def t(mode, blocks, threads, iters):
    ops.fp64_probe(mode, blocks, threads, iters); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fp64_probe(mode, blocks, threads, iters); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best
out = {}
iters = 20000
for warps_per_sm in (4, 8, 16, 32):
    secs = t(5, sms, warps_per_sm * 32, iters)
    shfl = 8 * iters * warps_per_sm  # warp-level SHFL per SM
    out[f"shfl_per_clk_per_sm_w{warps_per_sm}"] = shfl / (secs * clk)
out["shfl_dep_cycles"] = t(6, sms, 32, iters) / iters * clk
print(json.dumps(out))
