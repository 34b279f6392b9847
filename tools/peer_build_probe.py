"""probe: time modgpuModsetBuildFromPeers with LOCAL sources on one GPU (separates kernel cost from NVLink cost)"""
import ctypes as C, sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modimizer_b200 as mg
from modimizer_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda:0")
nb = 3100000000 - 3100000000 % 32
d_bases = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
synth.genome_device(12345, 0, nb, 1, d_bases.data_ptr(), st)
offs = torch.tensor([0, nb], dtype=torch.int64, device=dev)
for G in (1, 2, 4, 8):
    ms = mg.Modset(28, 31, 64, 17)
    ms.set_stream(st)
    R = int(lib.modgpuModsetRegions(ms._p))
    expected = nb // 64 + 1
    mean = expected / float(G * R)
    cap = (int(1.1 * mean + 4.0 * math.sqrt(mean) + 8) + 1) & ~1
    oc = max(65536, expected // 4)
    sb = torch.empty(G * R * cap, dtype=torch.int64, device=dev)
    sc = torch.zeros(G * R, dtype=torch.int32, device=dev)
    so = torch.empty(G * oc, dtype=torch.int64, device=dev)
    soc = torch.zeros(G, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    # pretend every "owner" is this GPU: nSrc = G sources, each the buckets this rank made for owner o
    bptr = (C.c_void_p * G)(*[sb.data_ptr() + o * R * cap * 8 for o in range(G)])
    optr = (C.c_void_p * G)(*[so.data_ptr() + o * oc * 8 for o in range(G)])
    ts, tb = [], []
    for it in range(4):
        _lib.check(lib.modgpuModsetClear(ms._p))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.check(lib.modgpuModsetSelectBucketsDevice(ms._p, C.c_void_p(d_bases.data_ptr()), C.c_void_p(offs.data_ptr()), 1, nb, 0, G,
                                                       C.c_void_p(sb.data_ptr()), cap, C.c_void_p(sc.data_ptr()), C.c_void_p(so.data_ptr()), oc,
                                                       C.c_void_p(soc.data_ptr()), C.c_void_p(cnt.data_ptr())))
        e[1].record()
        _lib.check(lib.modgpuModsetBuildFromPeers(ms._p, bptr, C.c_void_p(sc.data_ptr()), cap, G, optr, oc, C.c_void_p(soc.data_ptr())))
        e[2].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    print("G=%d cap=%d select %.3f ms build-from-(local)-peers %.3f ms entries %d selected %d" % (G, cap, min(ts), min(tb), ms.max, int(cnt.item())))
    ms.close()
