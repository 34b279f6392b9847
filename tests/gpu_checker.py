"""Adapter that gives the CUDA path (through the C ABI / modimizer_b200) the same
surface as tests/harness.Checker, so that golden cases and parity helpers can
run unchanged on the oracle, the reference and the GPU."""
import ctypes as C

import numpy as np


class GpuChecker:
    name = "gpu"

    def __init__(self):
        import torch
        import modimizer_b200 as mg
        from modimizer_b200 import _lib
        mg.require_device()
        self.torch, self.mg, self._lib, self.lib = torch, mg, _lib, _lib.load()

    def hasher(self, k, w, seed):
        s = self.mg.Seqhash(k, w, seed)
        return dict(mask=int(s.mask), shift=int(s.shift1), factor1=int(s.factor1), factor2=int(s._h.factor2))

    def mod_scan(self, k, w, seed, codes, flags=1 | 2):
        torch, lib, _lib = self.torch, self.lib, self._lib
        codes = np.ascontiguousarray(codes, np.uint8)
        n = len(codes)
        h = _lib.Hasher()
        _lib.check(lib.modgpuHasherInit(C.byref(h), k, w, seed))
        dev = torch.device("cuda:0")
        d_bases = torch.from_numpy(codes.copy() if n else np.zeros(1, np.uint8)).to(dev)
        d_offs = torch.tensor([0, n], dtype=torch.int64, device=dev)
        words = lib.modgpuPackedWords(n)
        d_packed = torch.zeros(words, dtype=torch.int64, device=dev)
        d_ends = torch.zeros(words, dtype=torch.int32, device=dev)
        cap = max(n, 1)
        d_k = torch.zeros(cap, dtype=torch.int64, device=dev)
        d_p = torch.zeros(cap, dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        d_ws = torch.zeros(lib.modgpuHashSelectWorkspace(n) // 8 + 8, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.modgpuPack2bit(d_bases.data_ptr(), n, 0, d_packed.data_ptr(), st), "pack")
        _lib.check(lib.modgpuMarkEnds(d_offs.data_ptr(), 1, n, d_ends.data_ptr(), st), "ends")
        _lib.check(lib.modgpuHashSelect(C.byref(h), d_packed.data_ptr(), d_ends.data_ptr(), n, d_k.data_ptr(), d_p.data_ptr(),
                                        cap, d_cnt.data_ptr(), d_ws.data_ptr(), flags, st), "select")
        torch.cuda.synchronize()
        cnt = int(d_cnt.item())
        km = d_k[:cnt].cpu().numpy().view(np.uint64)
        return (km & np.uint64((1 << 62) - 1)), d_p[:cnt].cpu().numpy().astype(np.int32), (km >> np.uint64(63)).astype(np.uint8)

    # ---- modset
    def modset_new(self, bits, k, w, seed):
        return self.mg.Modset(bits, k, w, seed, exact_order=True)

    def modset_add(self, ms, codes, offs):
        return ms.add(np.ascontiguousarray(codes, np.uint8), np.ascontiguousarray(offs, np.uint64), is_ascii=0)

    def _modset_max(self, ms):
        return ms.max

    def modset_export(self, ms):
        return ms.export()

    def modset_sorted(self, ms):
        return ms.sorted_dump()

    def modset_hist(self, ms):
        return ms.histogram()

    def modset_summary(self, ms):
        return ms.summary()

    def _modset_setcopy(self, ms, a, b, c):
        ms.set_copy(a, b, c)

    def _modset_setcopyM(self, ms, c):
        ms.set_copy_m(c)

    def _modset_free(self, ms):
        ms.close()

    # ---- modmap
    def ref_build(self, bits, k, w, seed, codes, offs):
        R = self.mg.Reference(bits, k, w, seed, np.ascontiguousarray(codes, np.uint8), np.ascontiguousarray(offs, np.uint64), is_ascii=0)
        return R, R.counts

    def ref_export(self, R):
        return R.export()

    def _ref_modset(self, R):
        return R.ms

    def ref_query(self, R, codes, offs):
        return R.query(np.ascontiguousarray(codes, np.uint8), np.ascontiguousarray(offs, np.uint64), is_ascii=0)

    def _ref_free(self, R):
        R.close()
