"""probe: modset build from a 30x readset (config-0 shape, scaled up): most k-mers repeat"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modimizer_b200 as mg
from modimizer_b200 import synth
dev = torch.device("cuda:0")
k, d, bits = int(sys.argv[1]) if len(sys.argv) > 1 else 19, int(sys.argv[2]) if len(sys.argv) > 2 else 31, 26
sp = synth.read_spec(12345, 40_000_000, 7, 10_000)
nreads = 120_000
n = nreads * 10_000
buf = torch.empty(n + 64, dtype=torch.uint8, device=dev)
synth.reads_device(sp, 0, nreads, False, buf.data_ptr())
offs = torch.arange(nreads + 1, dtype=torch.int64, device=dev) * 10_000
torch.cuda.synchronize()
ms = mg.Modset(bits, k, d, 17)
ms.profile(True)
best = 1e9
for it in range(5):
    ms.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot = ms.add_device(buf.data_ptr(), offs.data_ptr(), nreads, n)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t = ms.times()
print("reads k=%d d=%d: %.1f Mbases, %d hashes, %d distinct, best %.3f ms = %.0f Gbases/s; per-call ms: select %.3f insert %.3f" %
      (k, d, n / 1e6, tot, ms.max, best, n / best / 1e6, t["select"][0] / 5, t["insert"][0] / 5))
