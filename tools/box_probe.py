"""probe: is this box one of the 'slow insert' ones?  copy / fill bandwidth, region build time, GPU identity"""
import sys, os, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
x = torch.empty(1 << 30, dtype=torch.uint8, device=dev); y = torch.empty_like(x)
big = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
print("copy 1 GiB: %.0f GB/s (r+w)" % (2 * 1.0737 / t(lambda: y.copy_(x)) * 1e3))
print("fill 2 GiB: %.0f GB/s" % (2.147 / t(lambda: big.zero_()) * 1e3))
print("read-reduce 1 GiB: %.0f GB/s" % (1.0737 / t(lambda: x.view(torch.int64).sum()) * 1e3))
print(subprocess.run(["nvidia-smi", "--query-gpu=name,uuid,pci.bus_id,clocks.sm,clocks.mem,ecc.mode.current,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
print(subprocess.run("nvidia-smi -q | grep -i -E 'remapp|retired|pending|row' | head -12", shell=True, capture_output=True, text=True).stdout)
