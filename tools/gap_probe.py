"""probe: does the region build depend on what ran right before it?  select -> [gap variants] -> build, G = 1"""
import ctypes as C, sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modimizer_b200 as mg
from modimizer_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda:0")
nb = 3100000000
d_bases = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
synth.genome_device(12345, 0, nb, 1, d_bases.data_ptr(), st)
offs = torch.tensor([0, nb], dtype=torch.int64, device=dev)
G = 1
ms = mg.Modset(28, 31, 64, 17); ms.set_stream(st)
R = int(lib.modgpuModsetRegions(ms._p)); expected = nb // 64 + 1
mean = expected / float(G * R); cap = (int(1.1 * mean + 4.0 * math.sqrt(mean) + 8) + 1) & ~1; oc = max(65536, expected // 4)
sb = torch.empty(G * R * cap, dtype=torch.int64, device=dev); sc = torch.zeros(G * R, dtype=torch.int32, device=dev)
so = torch.empty(G * oc, dtype=torch.int64, device=dev); soc = torch.zeros(G, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)
bptr = (C.c_void_p * G)(sb.data_ptr()); optr = (C.c_void_p * G)(so.data_ptr())
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
big = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
def run(mode):
    out = []
    for it in range(5):
        _lib.check(lib.modgpuModsetClear(ms._p))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        _lib.check(lib.modgpuModsetSelectBucketsDevice(ms._p, C.c_void_p(d_bases.data_ptr()), C.c_void_p(offs.data_ptr()), 1, nb, 0, G,
                   C.c_void_p(sb.data_ptr()), cap, C.c_void_p(sc.data_ptr()), C.c_void_p(so.data_ptr()), oc, C.c_void_p(soc.data_ptr()), C.c_void_p(cnt.data_ptr())))
        e[1].record()
        if mode == "sleep": torch.cuda.synchronize(); time.sleep(0.02)
        if mode == "flush": flush.zero_()
        if mode == "sync": torch.cuda.synchronize()
        e[2].record()
        _lib.check(lib.modgpuModsetBuildFromPeers(ms._p, bptr, C.c_void_p(sc.data_ptr()), cap, G, optr, oc, C.c_void_p(soc.data_ptr())))
        e[3].record(); torch.cuda.synchronize()
        out.append((e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])))
    print(mode.ljust(8), "select %.3f build %s" % (min(o[0] for o in out), " ".join("%.3f" % o[1] for o in out)))
for mode in ("b2b", "sync", "sleep", "flush", "b2b"):
    run(mode)
def t(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)
print("fill 2 GiB alone: %.3f ms x3:" % t(lambda: big.zero_()), "%.3f %.3f" % (t(lambda: big.zero_()), t(lambda: big.zero_())))
