#!/usr/bin/env python
"""one build pass of the bench genome for profiling the count kernel under ncu:
   python tools/prof_count.py [k d bits] [flags]   (3.1 Gbases resident, 2 warm-up + 2 measured adds)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import modimizer_b200 as mg
from modimizer_b200 import synth

k, d, bits = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (31, 64, 28)
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
gb = float(os.environ.get("GBASES", "3.1"))
dev = torch.device("cuda:0")
nb = int(gb * 1e9); nb -= nb % 32
offs = (np.arange(25, dtype=np.float64) * (nb / 24)).astype(np.uint64); offs[-1] = nb
d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
d_bases = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
synth.genome_device(12345, 0, nb, 1, d_bases.data_ptr())
torch.cuda.synchronize()
ms = mg.Modset(bits, k, d, 17)
ms.set_flags(flags)
for it in range(4):
    ms.clear()
    t = ms.add_device(d_bases.data_ptr(), d_offs.data_ptr(), 24, nb)
    e = ms.max
print("k=%d d=%d flags=%d hashes %d entries %d" % (k, d, flags, t, e))
ms.profile(True)
for it in range(3):
    ms.clear(); ms.add_device(d_bases.data_ptr(), d_offs.data_ptr(), 24, nb); ms.max
tm = ms.times()
print({n: round(v[0] / 3, 4) for n, v in tm.items()})
