#!/usr/bin/env python
"""one-screen summary of a bench.py JSON line (for gpurun logs)"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("N=%d value %.0f Gb/s  %.3f ms/step  select %.3f  insert %.3f  frac %.3f  e2e %s  packed %s  launches %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"], k["select"]["ms_per_step"], k["insert"]["ms_per_step"], d["roofline"]["frac"],
    "%.1f" % d["e2e"]["value"] if d.get("e2e") else None,
    "%.1f" % d["e2e_packed"]["value"] if d.get("e2e_packed") else None, d["gpu_launches"]))
for c in d.get("configs") or []:
    print("  k=%d d=%d bases %.2e: %.0f Gb/s %.3f ms  select %.3f ms frac %.3f  insert %.3f  e2e %.1f" % (
        c["k"], c["d"], c["bases"], c["value"], c["ms_per_step"], c["roofline"]["ms_per_step"], c["roofline"]["frac"],
        c["insert_ms_per_step"], c["e2e"]["value"]))
if d.get("parity"):
    print("  parity", d["parity"]["ok"], d["parity"]["against"])
