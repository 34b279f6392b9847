import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import modimizer_b200 as mg
import hostemul as he
sp = he.read_spec(778, 2_000_000, 13, 150, sub_ppm=3000, frag_len=400, pair_mode=1, dup_mode=1)
nreads = 2 * 266_000
reads = he.reads(sp, 0, nreads)
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(150)
d_reads = torch.from_numpy(reads).cuda(); d_offs = torch.from_numpy(offs.view(np.int64)).cuda()
res = {}
for flags in (0, 65536):
    ms = mg.Modset(22, 19, 31, 17); ms.set_flags(flags)
    tot = ms.add_device(d_reads.data_ptr(), d_offs.data_ptr(), nreads, len(reads))
    v, d, i = ms.sorted_dump(); res[flags] = (tot, v, d); ms.close()
    print(flags, tot, len(v), int(d.astype(np.int64).sum()))
(t0, v0, d0), (t1, v1, d1) = res[0], res[65536]
if len(v0) == len(v1) and (v0 == v1).all():
    bad = np.nonzero(d0 != d1)[0]
    print("count diffs", [(hex(int(v0[j])), int(d0[j]), int(d1[j])) for j in bad[:10]])
    km = int(v0[bad[0]]) if len(bad) else None
else:
    s0, s1 = set(v0.tolist()), set(v1.tolist())
    print("only gen2", [hex(x) for x in list(s0 - s1)[:5]], "only gen1", [hex(x) for x in list(s1 - s0)[:5]])
    km = list(s0 - s1)[0] if s0 - s1 else None
if km is not None:
    # where does this k-mer (or its reverse complement) occur as a 19-mer in the concatenated batch?
    k = 19
    codes = reads.astype(np.uint64)
    def pack(x): 
        return x
    target = [(km >> (2 * (k - 1 - j))) & 3 for j in range(k)]
    rc = [3 - t for t in target[::-1]]
    arr = reads
    for name, t in (("fwd", target), ("rc", rc)):
        m = np.ones(len(arr) - k + 1, bool)
        for j, b in enumerate(t):
            m &= arr[j:len(arr) - k + 1 + j] == b
        pos = np.nonzero(m)[0]
        print(name, "positions", pos[:20], "mod150", (pos % 150)[:20], "tile", (pos // 2048)[:20], "in-tile", (pos % 2048)[:20])
