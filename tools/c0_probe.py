#!/usr/bin/env python
"""where the time of configs[0] (300 Mbases, k=19 d=31, table bits 24) goes: per-category device times
(modgpuModsetTimes) against the wall clock of the call"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modimizer_b200 as mg
from modimizer_b200 import synth

dev = torch.device("cuda:0")
sp = synth.read_spec(12345, 10_000_000, 7, 10_000)
n, L = 30_000, 10_000
buf = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
synth.reads_device(sp, 0, n, False, buf.data_ptr())
torch.cuda.synchronize()
for rep in range(4):
    ms = mg.Modset(24, 19, 31, 17)
    ms.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = ms.add_device(buf.data_ptr(), offs.data_ptr(), n, n * L)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0)
    print("rep %d wall %.3f ms hashes %d times %s" % (rep, wall, tot, ms.times()))
    ms.profile(False)
    t0 = time.perf_counter()
    ms.clear(); tot = ms.add_device(buf.data_ptr(), offs.data_ptr(), n, n * L)
    torch.cuda.synchronize()
    print("   second add (no profiling, buffers allocated) wall %.3f ms" % (1e3 * (time.perf_counter() - t0)))
    ms.close()
