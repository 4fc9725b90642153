#!/usr/bin/env python
"""Measure the FP64 DFMA / DMMA issue rates of the device (roofline denominator).

MEASURED_PEAKS.json has no FP64 entry, so this is the measured FP64 peak the
roofline of the fused neighbourhood kernel is quoted against.  Prints one JSON
object; used by bench.py (imported) and runnable stand-alone under gpurun.
"""

import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from muygpys_b200 import ops


def _time(mode, blocks, threads, iters, reps=5):
    ops.fp64_probe(mode, blocks, threads, iters)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.fp64_probe(mode, blocks, threads, iters)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def measure(iters=20000):
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {"sms": sms}
    blocks, threads = sms * 4, 256
    nthreads = blocks * threads
    nwarps = nthreads // 32
    t = _time(0, blocks, threads, iters)
    out["dfma_tflops"] = 2 * 8 * iters * nthreads / t / 1e12
    t = _time(1, blocks, threads, iters)
    out["dmma_tflops"] = 2 * 256 * 4 * iters * nwarps / t / 1e12
    t = _time(2, blocks, threads, iters)
    fl = (2 * 8 * iters * nthreads / 2) + (2 * 256 * 4 * iters * nwarps / 2)
    out["mixed_tflops"] = fl / t / 1e12
    # latency: one warp per SMSP-ish, dependent chains
    t = _time(3, sms, 32, iters)
    clk = torch.cuda.clock_rate() * 1e6 if hasattr(torch.cuda, "clock_rate") else None
    out["dfma_dep_ns"] = t / iters * 1e9
    t = _time(4, sms, 32, iters)
    out["dmma_dep_ns"] = t / iters * 1e9
    if clk:
        out["sm_clock_mhz_now"] = clk / 1e6
    out["fp64_peak_tflops"] = max(out["dfma_tflops"], out["dmma_tflops"])
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
