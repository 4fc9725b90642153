#!/usr/bin/env python
"""Pinned H2D rate with and without GPU-local CPU affinity -- dev tool."""
import json
import os
import subprocess
import sys

import torch


def h2d(label, out):
    buf = torch.empty(40_000_000 // 8, dtype=torch.int64).pin_memory()
    buf.fill_(1)
    dst = torch.empty_like(buf, device="cuda")
    dst.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        dst.copy_(buf, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out[label] = 10 * 40e6 / (a.elapsed_time(b) * 1e-3) / 1e9


out = {"cpus": os.cpu_count(), "affinity_before": len(os.sched_getaffinity(0))}
torch.cuda.init()
h2d("default_gbs", out)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + j for i, w in enumerate(words) for j in range(64) if (w >> j) & 1]
    out["gpu_local_cpus"] = f"{cpus[0]}..{cpus[-1]} ({len(cpus)})" if cpus else "none"
    allowed = sorted(set(cpus) & os.sched_getaffinity(0))
    if allowed:
        os.sched_setaffinity(0, allowed)
        h2d("gpu_local_affinity_gbs", out)
except Exception as e:  # noqa: BLE001
    out["nvml_error"] = repr(e)
print(json.dumps(out))
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:])
print(subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"], capture_output=True, text=True).stdout)
