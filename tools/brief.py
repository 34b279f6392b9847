import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("value %.0f ms/step %.3f select %.3f insert %.3f e2e %s launches %s" % (d["value"], d["ms_per_step"], k["select"]["ms_per_step"], k["insert"]["ms_per_step"], (d.get("e2e") or {}).get("value"), d["gpu_launches"]))
