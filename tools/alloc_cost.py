#!/usr/bin/env python
"""Dev tool (GPU box): what does one cudaMalloc + cudaFree of a sweep-sized buffer cost, and how much does it vary?
(p6d_sweep_run allocates its chunk buffers per call; see NOTES.md "Smaller follow-ups".)"""
import json, statistics, sys, time
import torch
from cuda import cudart

torch.cuda.init(); torch.zeros(1, device="cuda")
res = {}
for mb in (90, 350, 700):
    t_malloc, t_free, t_touch = [], [], []
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
        t0 = time.perf_counter()
        err, p = cudart.cudaMalloc(mb << 20)
        t1 = time.perf_counter()
        assert err == cudart.cudaError_t.cudaSuccess
        cudart.cudaMemset(p, 0, mb << 20); cudart.cudaDeviceSynchronize()
        t2 = time.perf_counter()
        cudart.cudaFree(p)
        t3 = time.perf_counter()
        t_malloc.append((t1 - t0) * 1e3); t_free.append((t3 - t2) * 1e3); t_touch.append((t2 - t1) * 1e3)
    res[f"{mb}MB"] = {"malloc_ms_median": round(statistics.median(t_malloc), 3), "malloc_ms_max": round(max(t_malloc), 3),
                      "free_ms_median": round(statistics.median(t_free), 3), "free_ms_max": round(max(t_free), 3),
                      # first touch of the fresh range (memset + sync): the first iteration is the interesting one
                      "first_touch_ms_first": round(t_touch[0], 3), "first_touch_ms_median": round(statistics.median(t_touch), 3)}
print(json.dumps(res))
