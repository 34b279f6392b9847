"""GPU parity tests: the CUDA path through the C ABI (libmodgpu.so) against the
oracle (oracle/liboracle.so) on the same seeded inputs.  Bit-exact everywhere:
the path is integer/byte work, there is no tolerance.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

import harness as H
import hostemul as he

pytestmark = pytest.mark.gpu

KAT_SEQ = ("ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGTACGATCGTAGCTAGCTAGCTGATCGATGCATGCATCGATCGTAGCTAGCTAGCTAGCATCGATGCATGCAAATTTGGGCCCATATCGCGATATCGC")


@pytest.fixture(scope="module")
def mg():
    import modimizer_b200 as m
    m.require_device()
    return m


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


# --------------------------------------------------------------- helpers --
def gpu_select(mg, torch, k, d, seed, data, offs, flags=1, is_ascii=0):
    """device-level K1 + K2 through the C ABI; returns (kmers, gpos) as numpy"""
    import ctypes as C
    from modimizer_b200 import _lib
    lib = _lib.load()
    h = _lib.Hasher()
    _lib.check(lib.modgpuHasherInit(C.byref(h), k, d, seed))
    data = np.ascontiguousarray(data, np.uint8)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = int(offs[-1])
    dev = torch.device("cuda:0")
    d_bases = torch.from_numpy(data if n else np.zeros(1, np.uint8)).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    words = lib.modgpuPackedWords(n)
    d_packed = torch.zeros(words, dtype=torch.int64, device=dev)
    d_ends = torch.zeros(words, dtype=torch.int32, device=dev)
    cap = max(n, 1)
    d_k = torch.zeros(cap, dtype=torch.int64, device=dev)
    d_p = torch.zeros(cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    d_ws = torch.zeros(lib.modgpuHashSelectWorkspace(n) // 8 + 8, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.modgpuPack2bit(d_bases.data_ptr(), n, is_ascii, d_packed.data_ptr(), st), "pack")
    _lib.check(lib.modgpuMarkEnds(d_offs.data_ptr(), len(offs) - 1, n, d_ends.data_ptr(), st), "ends")
    _lib.check(lib.modgpuHashSelect(C.byref(h), d_packed.data_ptr(), d_ends.data_ptr(), n, d_k.data_ptr(), d_p.data_ptr(),
                                    cap, d_cnt.data_ptr(), d_ws.data_ptr(), flags, st), "select")
    torch.cuda.synchronize()
    cnt = int(d_cnt.item())
    km = d_k[:cnt].cpu().numpy().view(np.uint64)
    gp = d_p[:cnt].cpu().numpy().view(np.uint32)
    return km, gp, d_packed.cpu().numpy().view(np.uint64)


def oracle_select(orc, k, d, seed, data, offs):
    ks, ps, fs = [], [], []
    for r in range(len(offs) - 1):
        a, b = int(offs[r]), int(offs[r + 1])
        kk, pp, ff = orc.mod_scan(k, d, seed, data[a:b])
        ks.append(kk); ps.append(pp.astype(np.int64) + a); fs.append(ff)
    if not ks:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64), np.zeros(0, np.uint8)
    return np.concatenate(ks), np.concatenate(ps), np.concatenate(fs)


def random_batch(rng, nseq, maxlen, mode="random"):
    lens = rng.integers(0, maxlen, nseq)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(offs[-1])
    if mode == "polyA":
        data = np.zeros(max(n, 1), np.uint8)
    elif mode == "AT":
        data = np.tile(np.array([0, 3], np.uint8), n // 2 + 1)[:max(n, 1)].copy()
    else:
        data = rng.integers(0, 4, max(n, 1)).astype(np.uint8)
    return data[:n] if n else np.zeros(0, np.uint8), offs


# ------------------------------------------------------------------ K1 K2 --
def test_kat_survey_vectors(mg, torch_cuda, orc):
    """the known-answer vectors of SURVEY section 4 (generated from the reference)"""
    codes = H.codes_from_ascii(KAT_SEQ)
    offs = np.array([0, len(codes)], np.uint64)
    km, gp, _ = gpu_select(mg, torch_cuda, 19, 31, 17, codes, offs, flags=1 | 2)
    assert [int(x) for x in gp] == [3, 41, 49, 84]
    assert [hex(int(x) & ((1 << 62) - 1)) for x in km] == ["0x3e4e58c9c9", "0x2c9c9c9e36", "0x24e4d8d272", "0x1393639c9c"]
    assert [int(x) >> 63 for x in km] == [1, 1, 0, 0]                      # isF
    # len < k -> nothing
    km, gp, _ = gpu_select(mg, torch_cuda, 19, 31, 17, codes[:18], np.array([0, 18], np.uint64))
    assert len(km) == 0
    # ASCII input gives the same answer as codes
    asc = np.frombuffer(KAT_SEQ.encode(), np.uint8)
    km2, gp2, _ = gpu_select(mg, torch_cuda, 19, 31, 17, asc, offs, flags=1, is_ascii=1)
    assert [int(x) for x in gp2] == [3, 41, 49, 84]


def test_pack_layout(mg, torch_cuda):
    rng = np.random.default_rng(5)
    for n in (1, 31, 32, 33, 1000, 8192, 8193, 100003):
        codes = rng.integers(0, 4, n).astype(np.uint8)
        offs = np.array([0, n], np.uint64)
        _, _, packed = gpu_select(mg, torch_cuda, 5, 1, 17, codes, offs)
        nw = (n + 31) // 32
        exp = np.zeros(nw, np.uint64)
        he.lib().hm_pack(codes, n, 0, exp, nw)
        assert np.array_equal(packed[:nw], exp), n
        assert not packed[nw:].any()
        # ASCII, mixed case and N
        asc = np.frombuffer(b"ACGTacgtNn", np.uint8)[rng.integers(0, 10, n)]
        _, _, packed2 = gpu_select(mg, torch_cuda, 5, 1, 17, asc, offs, is_ascii=1)
        exp2 = np.zeros(nw, np.uint64)
        he.lib().hm_pack(H.codes_from_ascii(asc.tobytes()), n, 0, exp2, nw)
        assert np.array_equal(packed2[:nw], exp2), n


def test_pack_misaligned_pointer(mg, torch_cuda):
    """a batch that starts anywhere inside a resident buffer (any pointer alignment)"""
    import ctypes as C
    from modimizer_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(6)
    n = 70001
    dev = torch_cuda.device("cuda:0")
    host = rng.integers(0, 4, n + 64).astype(np.uint8)
    d = torch_cuda.from_numpy(host).to(dev)
    words = lib.modgpuPackedWords(n)
    out = torch_cuda.zeros(words, dtype=torch_cuda.int64, device=dev)
    for off in list(range(0, 18)) + [31, 33, 47]:
        for ascii_ in (0, 1):
            src = host[off:off + n]
            if ascii_:
                asc = np.frombuffer(b"ACGT", np.uint8)[src]
                d2 = torch_cuda.from_numpy(np.concatenate([np.zeros(off, np.uint8), asc, np.zeros(64, np.uint8)])).to(dev)
                ptr = d2.data_ptr() + off
            else:
                ptr = d.data_ptr() + off
            _lib.check(lib.modgpuPack2bit(ptr, n, ascii_, out.data_ptr(), torch_cuda.cuda.current_stream().cuda_stream))
            torch_cuda.cuda.synchronize()
            nw = (n + 31) // 32
            exp = np.zeros(nw, np.uint64)
            he.lib().hm_pack(np.ascontiguousarray(src), n, 0, exp, nw)
            assert np.array_equal(out.cpu().numpy().view(np.uint64)[:nw], exp), (off, ascii_)


@pytest.mark.parametrize("flags", [1, 0, 1 | 4, 4, 1 | 8, 8 | 4])
def test_select_random_params(mg, torch_cuda, orc, flags):
    """property test: GPU selected list == the serial iterator's list, in order
    (ORDERED) or as a multiset (count mode), for random k, d, seed, ragged batches"""
    rng = np.random.default_rng(100 + flags)
    ds = [1, 2, 3, 7, 8, 16, 31, 32, 48, 62, 64, 128, 1000, 4096]
    for trial in range(40):
        k = int(rng.integers(1, 32))
        d = int(rng.choice(ds))
        seed = int(rng.integers(0, 50))
        mode = ["random", "random", "random", "polyA", "AT"][trial % 5]
        data, offs = random_batch(rng, int(rng.integers(1, 40)), 40 if trial % 4 == 0 else 3000, mode)
        ek, ep, ef = oracle_select(orc, k, d, seed, data, offs)
        km, gp, _ = gpu_select(mg, torch_cuda, k, d, seed, data, offs, flags=flags | 2)
        assert len(km) == len(ek), (k, d, seed, trial)
        kk = km & np.uint64((1 << 62) - 1)
        ff = (km >> np.uint64(63)).astype(np.uint8)
        if flags & 1:
            assert np.array_equal(kk, ek) and np.array_equal(gp.astype(np.int64), ep) and np.array_equal(ff, ef), (k, d, seed, trial)
        else:
            o1 = np.lexsort((kk, gp)); o2 = np.lexsort((ek, ep))
            assert np.array_equal(kk[o1], ek[o2]) and np.array_equal(gp[o1].astype(np.int64), ep[o2]), (k, d, seed, trial)


def test_select_large_multi_tile(mg, torch_cuda, orc):
    """many tiles, reads that start at arbitrary (non word-aligned) offsets, both
    headline parameter sets, TMA and plain loads, prefilter and generic"""
    rng = np.random.default_rng(77)
    lens = np.concatenate([rng.integers(100, 20000, 300), [0, 1, 18, 19, 30, 31, 32, 33, 0, 0, 150, 150]])
    rng.shuffle(lens)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rng.integers(0, 4, int(offs[-1])).astype(np.uint8)
    for (k, d) in ((19, 31), (31, 64), (16, 64), (17, 32), (31, 31)):
        ek, ep, ef = oracle_select(orc, k, d, 17, data, offs)
        for flags in (1, 1 | 4, 1 | 8):
            km, gp, _ = gpu_select(mg, torch_cuda, k, d, 17, data, offs, flags=flags)
            assert np.array_equal(km, ek) and np.array_equal(gp.astype(np.int64), ep), (k, d, flags)
        km, gp, _ = gpu_select(mg, torch_cuda, k, d, 17, data, offs, flags=0)
        assert np.array_equal(np.sort(km), np.sort(ek)), (k, d)


def test_select_table_prefilter(mg, torch_cuda, orc):
    """count mode with the shared-memory candidate table (k >= 30, 64-2k+tz <= 8) == oracle == arithmetic
    prefilter (flag 64) == no prefilter (flag 8), TMA and plain loads, ragged multi-tile batches"""
    rng = np.random.default_rng(78)
    lens = np.concatenate([rng.integers(20, 30000, 200), [0, 1, 29, 30, 31, 32, 33, 61, 62, 63, 64, 65, 0, 95, 96, 97]])
    rng.shuffle(lens)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rng.integers(0, 4, int(offs[-1])).astype(np.uint8)
    data[1000:1500] = 0
    data[70000:71000:2] = 3
    for (k, d, seed) in ((31, 64, 17), (31, 64, 0), (31, 32, 3), (31, 8, 5), (31, 192, 17), (30, 16, 17), (30, 8, 1), (31, 128, 9)):
        ek, ep, ef = oracle_select(orc, k, d, seed, data, offs)
        o2 = np.lexsort((ek, ep))
        for flags in (0, 4, 64, 8):
            km, gp, _ = gpu_select(mg, torch_cuda, k, d, seed, data, offs, flags=flags | 2)
            kk = km & np.uint64((1 << 62) - 1)
            ff = (km >> np.uint64(63)).astype(np.uint8)
            o1 = np.lexsort((kk, gp))
            assert len(km) == len(ek), (k, d, flags)
            assert np.array_equal(kk[o1], ek[o2]) and np.array_equal(gp[o1].astype(np.int64), ep[o2]), (k, d, flags)
            assert np.array_equal(ff[o1], ef[o2]), (k, d, flags)


# ------------------------------------------------------------------ modset --
def _check_modset(mg, orc, bits, k, d, seed, data, offs, exact):
    ms = mg.Modset(bits, k, d, seed, exact_order=exact)
    oms = orc.modset_new(bits, k, d, seed)
    try:
        tot = ms.add(data, offs, is_ascii=0)
        otot = orc.modset_add(oms, data, offs)
        assert tot == otot
        assert ms.max == orc._modset_max(oms)
        gv, gd, gi = ms.sorted_dump()
        ov, od, oi = orc.modset_sorted(oms)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od) and np.array_equal(gi, oi)
        if exact:                                            # identical index numbering
            v, dd, ii = ms.export()
            ov2, od2, oi2 = orc.modset_export(oms)
            assert np.array_equal(v, ov2) and np.array_equal(dd, od2)
        assert np.array_equal(ms.histogram(), orc.modset_hist(oms))
        assert ms.summary() == orc.modset_summary(oms)
        return ms, oms
    except Exception:
        ms.close(); orc._modset_free(oms)
        raise


@pytest.mark.parametrize("exact", [False, True])
def test_modset_build_count(mg, orc, exact):
    """config-1 shape, scaled: reads sampled from a genome, both strands, ~30x"""
    glen, rlen, nreads = 200000, 5000, 1200
    sp = he.read_spec(12345, glen, 99, rlen)
    data = he.reads(sp, 0, nreads)
    offs = (np.arange(nreads + 1, dtype=np.uint64) * np.uint64(rlen))
    ms, oms = _check_modset(mg, orc, 22, 19, 31, 17, data, offs, exact)
    try:
        # classification -s and -sM (modutils.c:205-219) + summary tallies
        c = ms.set_copy(10, 45, 75)
        orc._modset_setcopy(oms, 10, 45, 75)
        gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(oms)
        assert np.array_equal(gi, oi)
        assert list(c) == [int((oi == j).sum()) for j in range(4)]
        assert ms.summary() == orc.modset_summary(oms)
        ms.set_copy_m(40)
        orc._modset_setcopyM(oms, 40)
        assert np.array_equal(ms.sorted_dump()[2], orc.modset_sorted(oms)[2])
        assert ms.summary() == orc.modset_summary(oms)
        # a second batch accumulates into the same table
        data2 = he.reads(sp, nreads, 300)
        offs2 = (np.arange(301, dtype=np.uint64) * np.uint64(rlen))
        assert ms.add(data2, offs2, is_ascii=0) == orc.modset_add(oms, data2, offs2)
        gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(oms)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od)
        # lookups (modsetIndexFind(..,false))
        probe = np.concatenate([ov[:1000], ov[:1000] ^ np.uint64(1)])
        gidx, gcp = ms.find(probe)
        oidx = np.array([orc._modset_find(oms, int(x)) for x in probe])
        assert np.array_equal(gidx != 0, oidx != 0)
        if exact:
            assert np.array_equal(gidx, oidx)
    finally:
        ms.close(); orc._modset_free(oms)


@pytest.mark.parametrize("mode", ["bulk", "bulk_nofusepack", "bulk_nofuse", "direct", "partitioned"])
def test_insert_paths_agree(mg, orc, mode):
    """the three insert strategies (shared-memory region build, direct HBM probes,
    region-partitioned direct) give the oracle's modset: fresh table, repeated
    adds into a populated table, heavy skew (bucket overflow), histogram"""
    flags = {"bulk": 255 << 8, "bulk_nofusepack": (255 << 8) | 32, "bulk_nofuse": (255 << 8) | 16, "direct": 1 << 8,
             "partitioned": 7 << 8}[mode]
    sp = he.read_spec(12345, 300000, 42, 3000, 2000)
    nreads = 2500
    data = he.reads(sp, 0, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(3000)
    skew = np.zeros(200000, np.uint8)                     # poly-A: one k-mer, 200k times
    skew[100000:100050] = 1
    ms = mg.Modset(22, 19, 4, 17)
    ms.set_flags(flags)
    oms = orc.modset_new(22, 19, 4, 17)
    try:
        for (d_, o_) in ((data, offs), (data[:3000 * 900], offs[:901]), (skew, np.array([0, 200000], np.uint64)), (data, offs)):
            assert ms.add(d_, o_, is_ascii=0) == orc.modset_add(oms, d_, o_)
            assert ms.max == orc._modset_max(oms)
        # raw ASCII text (mixed case, N) with a ragged end, through the same path
        part = data[:3000 * 77 + 13].copy()
        asc = np.frombuffer(b"ACGT", np.uint8)[part]
        asc[::7] |= 0x20                                      # lower case
        asc[5000:5100] = ord("N"); part[5000:5100] = 0
        po = np.array([0, 3000 * 40 + 1, 3000 * 40 + 1, len(part)], np.uint64)
        assert ms.add(asc, po, is_ascii=1) == orc.modset_add(oms, part, po)
        gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(oms)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od)
        assert np.array_equal(ms.histogram(), orc.modset_hist(oms))
        idx, _ = ms.find(np.concatenate([ov[::50], ov[::50] ^ np.uint64(3)]))
        assert (idx[:len(ov[::50])] != 0).all()
    finally:
        ms.close(); orc._modset_free(oms)


@pytest.mark.parametrize("k,d", [(31, 64), (30, 16), (31, 8)])
def test_fused_pass_table_scan_ascii_and_ragged(mg, torch_cuda, orc, k, d):
    """the bench path end to end - raw bytes -> fused K1 + table-driven K2 + bucket scatter -> region build - on
    ASCII text (mixed case, N) and on codes, ragged batches whose last warp tile is partial, device pointers that
    are and are not 16-byte aligned (the unaligned one takes the separate pack kernel + packed-stream loader)"""
    rng = np.random.default_rng(k * 100 + d)
    lens = np.concatenate([rng.integers(1, 9000, 150), [0, k - 1, k, k + 1, 2047, 2048, 2049, 2080, 4096 + 31]])
    rng.shuffle(lens)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    codes = rng.integers(0, 4, int(offs[-1])).astype(np.uint8)
    asc = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    asc[::5] |= 0x20
    asc[7000:7300] = ord("n"); codes[7000:7300] = 0
    oms = orc.modset_new(22, k, d, 17)
    otot = orc.modset_add(oms, codes, offs)
    ov, od, _ = orc.modset_sorted(oms)
    dev = torch_cuda.device("cuda:0")
    d_offs = torch_cuda.from_numpy(offs.view(np.int64)).to(dev)
    try:
        for (buf, is_ascii) in ((codes, 0), (asc, 1)):
            for shift in (0, 16, 5):                      # device pointer alignment
                ms = mg.Modset(22, k, d, 17)
                ms.set_flags(255 << 8)                    # force the bulk (fused) insert path at this small size
                d_b = torch_cuda.from_numpy(np.concatenate([np.zeros(shift, np.uint8), buf, np.zeros(64, np.uint8)])).to(dev)
                assert ms.add_device(d_b.data_ptr() + shift, d_offs.data_ptr(), len(lens), len(buf), is_ascii) == otot
                gv, gd, _ = ms.sorted_dump()
                assert np.array_equal(gv, ov) and np.array_equal(gd, od), (is_ascii, shift)
                ms.close()
    finally:
        orc._modset_free(oms)


def test_modset_edge_cases(mg, orc):
    rng = np.random.default_rng(3)
    # empty batch, empty reads, len<k, len==k, palindromes, d = 1 (every k-mer), k = 1
    for (k, d, mode) in ((19, 31, "random"), (5, 1, "random"), (1, 1, "random"), (4, 3, "AT"), (31, 64, "polyA"), (12, 8, "AT")):
        data, offs = random_batch(rng, 30, 200, mode)
        ms, oms = _check_modset(mg, orc, 20, k, d, 17, data, offs, True)
        ms.close(); orc._modset_free(oms)
    ms = mg.Modset(20, 19, 31, 17)
    assert ms.add(np.zeros(0, np.uint8), np.array([0], np.uint64)) == 0
    assert ms.add(np.zeros(0, np.uint8), np.array([0, 0, 0], np.uint64)) == 0
    assert ms.max == 0
    ms.close()


@pytest.mark.parametrize("accum", [2, 3, 16])
def test_deferred_build_accumulates_chunks(mg, orc, accum):
    """modgpuModsetSetAccumulate: the k-mers of several chunks wait in the region buckets and the regions are built once.
    Same set as the oracle whatever the grouping: reader-triggered flush, explicit flush, a skewed chunk that overflows
    the overflow list in the middle (rolled back, what was waiting is built, the chunk repeats through the list path),
    clear with k-mers waiting, back to the immediate mode"""
    sp = he.read_spec(777, 300000, 42, 3000, 2000)
    nreads = 2500
    data = he.reads(sp, 0, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(3000)
    skew = np.zeros(400000, np.uint8)                     # poly-A: one k-mer 400k times, 4x what the chunk announces
    skew[100000:100050] = 1
    ms = mg.Modset(22, 19, 4, 17)
    ms.set_flags(255 << 8)                                # always the bucket path
    ms.set_accumulate(accum)
    oms = orc.modset_new(22, 19, 4, 17)
    try:
        parts = [(data[:3000 * 700], offs[:701]), (data[3000 * 700:3000 * 1500], offs[700:1501] - offs[700]),
                 (data, offs), (skew, np.array([0, 400000], np.uint64)), (data[:3000 * 900], offs[:901]),
                 (data[3000 * 1500:], offs[1500:] - offs[1500]), (data, offs)]
        for i, (d_, o_) in enumerate(parts):
            assert ms.add(d_, o_, is_ascii=0) == orc.modset_add(oms, d_, o_), i
            if i == 4:
                assert ms.max == orc._modset_max(oms)     # a reader in the middle: applies what is waiting
        ms.flush()
        gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(oms)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od)
        assert np.array_equal(ms.histogram(), orc.modset_hist(oms))
        # k-mers waiting when the set is cleared are dropped with it
        ms.add(data, offs, is_ascii=0)
        ms.clear()
        assert ms.max == 0
        # a reader without flush(), then back to the immediate mode with k-mers waiting
        d_, o_ = parts[1]
        o2 = orc.modset_new(22, 19, 4, 17)
        assert ms.add(d_, o_, is_ascii=0) == orc.modset_add(o2, d_, o_)
        assert np.array_equal(ms.histogram(), orc.modset_hist(o2))
        assert ms.add(d_, o_, is_ascii=0) == orc.modset_add(o2, d_, o_)
        ms.set_accumulate(1)
        assert ms.add(data, offs, is_ascii=0) == orc.modset_add(o2, data, offs)
        gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(o2)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od)
        orc._modset_free(o2)
    finally:
        ms.close(); orc._modset_free(oms)


def test_deferred_build_reports_a_full_table_at_the_flush(mg):
    """deferred mode moves the reference's die() (modset.c:58) from the add to the flush / first reader"""
    rng = np.random.default_rng(9)
    n = 2000000
    data = rng.integers(0, 4, n).astype(np.uint8)
    ms = mg.Modset(20, 19, 4, 17)          # capacity 2^18 - 2 entries, ~500k distinct modimizers
    ms.set_flags(255 << 8)
    ms.set_accumulate(4)
    ms.add(data, np.array([0, n], np.uint64), is_ascii=0)
    with pytest.raises(mg.ModgpuError):
        ms.flush()
    ms.close()


def test_depth_saturation(mg, orc):
    """depth is U16 saturating at 65535 (modutils.c:26): poly-A read, d = 1"""
    n = 70000 + 18
    data = np.zeros(n, np.uint8)
    offs = np.array([0, n], np.uint64)
    ms = mg.Modset(20, 19, 1, 17)
    oms = orc.modset_new(20, 19, 1, 17)
    assert ms.add(data, offs, is_ascii=0) == orc.modset_add(oms, data, offs) == 70000
    gv, gd, gi = ms.sorted_dump(); ov, od, oi = orc.modset_sorted(oms)
    assert np.array_equal(gv, ov) and np.array_equal(gd, od) and int(gd[0]) == 65535
    assert np.array_equal(ms.histogram(), orc.modset_hist(oms))
    ms.close(); orc._modset_free(oms)


def test_table_overflow_is_an_error(mg):
    """the reference dies when max >= size (modset.c:58); we raise"""
    rng = np.random.default_rng(9)
    n = 400000
    data = rng.integers(0, 4, n).astype(np.uint8)
    ms = mg.Modset(20, 19, 1, 17)          # capacity 2^18 - 2 entries
    with pytest.raises(mg.ModgpuError):
        ms.add(data, np.array([0, n], np.uint64), is_ascii=0)
    ms.close()


# ------------------------------------------------------------------ modmap --
def test_reference_build_and_query(mg, orc):
    """config-2/5 shape, scaled: genome with duplicated segments (copy2 / multi
    classes), reads with errors looked up; every Reference array identical"""
    seqlens = [300000, 150000, 0, 70000, 19, 18]
    total = sum(seqlens)
    genome = he.genome(4242, 0, total, 1)
    # plant small-scale duplicates so that copy2 / multi classes are populated
    genome[150000:160000] = genome[10000:20000]
    genome[310000:312000] = genome[10000:12000]
    genome[400000:405000] = genome[100000:105000]
    offs = np.concatenate([[0], np.cumsum(seqlens)]).astype(np.uint64)
    for (k, d) in ((31, 64), (19, 31)):
        R = mg.Reference(24, k, d, 17, genome, offs, is_ascii=0)
        oref, ocounts = orc.ref_build(24, k, d, 17, genome, offs)
        try:
            assert list(R.counts) == list(ocounts), (k, d)
            assert ocounts[2] > 0 and ocounts[3] > 0
            g = R.export(); o = orc.ref_export(oref)
            for key in ("index", "offset", "id", "depth", "loc", "rev"):
                assert np.array_equal(g[key], o[key]), (k, d, key)
            gv, gd, gi = R.ms.export()
            ov, od, oi = orc.modset_export(orc._ref_modset(oref))
            assert np.array_equal(gv, ov) and np.array_equal(gd, od) and np.array_equal(gi, oi)
            # reads: error-free, 1% and 10% substitution/indel
            for (sub, ins, dele, ont) in ((0, 0, 0, False), (10000, 0, 0, False), (30000, 30000, 40000, True)):
                sp = he.read_spec(4242, 300000, 5, 2000, sub, ins, dele, dup_mode=1)
                reads = he.reads(sp, 0, 200, ont)
                roffs = np.arange(201, dtype=np.uint64) * np.uint64(2000)
                gq = R.query(reads, roffs, is_ascii=0)
                oq = orc.ref_query(oref, reads, roffs)
                for key in ("seedOff", "index", "pos", "hitId", "hitOffset", "counters"):
                    assert np.array_equal(gq[key], oq[key]), (k, d, key, sub)
                assert oq["counters"][:, 1].sum() > 0
        finally:
            R.close(); orc._ref_free(oref)


def test_import_export_roundtrip(mg, orc):
    """sync-to-host / upload: export -> import into a fresh table -> identical set and lookups"""
    rng = np.random.default_rng(21)
    data = rng.integers(0, 4, 300000).astype(np.uint8)
    offs = np.array([0, 100000, 300000], np.uint64)
    ms = mg.Modset(22, 19, 31, 17, exact_order=True)
    ms.add(data, offs, is_ascii=0)
    ms.set_copy(1, 2, 3)
    v, d, i = ms.export()
    ms2 = mg.Modset(22, 19, 31, 17)
    ms2.import_entries(v, d, i)
    v2, d2, i2 = ms2.export()
    assert np.array_equal(v, v2) and np.array_equal(d, d2) and np.array_equal(i, i2)
    idx, cp = ms2.find(v[::7])
    assert np.array_equal(idx, np.arange(1, len(v) + 1, dtype=np.uint32)[::7]) and np.array_equal(cp, i[::7] & 3)
    ms.close(); ms2.close()


def test_device_synth_matches_host(mg, torch_cuda):
    """the device generators produce the bytes the host generator (oracle side) does"""
    from modimizer_b200 import synth
    dev = torch_cuda.device("cuda:0")
    buf = torch_cuda.zeros(200000, dtype=torch_cuda.uint8, device=dev)
    synth.genome_device(12345, 64, 200000, 1, buf.data_ptr())
    torch_cuda.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), he.genome(12345, 64, 200000, 1))
    synth.genome_device(12345, 7, 1999, 0, buf.data_ptr() + 1)
    torch_cuda.cuda.synchronize()
    assert np.array_equal(buf[1:2000].cpu().numpy(), he.genome(12345, 7, 1999, 0))
    for (spec, ont) in ((synth.read_spec(12345, 100000, 7, 150, 5000, frag_len=400, pair_mode=1), False),
                        (synth.read_spec(12345, 100000, 7, 1001, 1000), False),
                        (synth.read_spec(12345, 100000, 7, 500, 30000, 30000, 40000), True)):
        n = 64
        synth.reads_device(spec, 3, n, ont, buf.data_ptr())
        torch_cuda.cuda.synchronize()
        hs = he.read_spec(spec.genomeSeed, spec.genomeLen, spec.readSeed, spec.readLen, spec.subPPM, spec.insPPM,
                          spec.delPPM, spec.fragLen, spec.pairMode, spec.dupMode)
        assert np.array_equal(buf[:n * spec.readLen].cpu().numpy(), he.reads(hs, 3, n, ont))


def test_large_batch_properties(mg, torch_cuda):
    """size-independent properties at a multi-chunk size (no oracle): adding the
    same batch twice doubles every depth and keeps the set; the histogram
    counts every entry once; total hashes = sum of depths (below saturation)."""
    from modimizer_b200 import synth
    n = (1 << 28) + 12345                    # > one host chunk
    dev = torch_cuda.device("cuda:0")
    buf = torch_cuda.zeros(n, dtype=torch_cuda.uint8, device=dev)
    synth.genome_device(777, 0, n, 0, buf.data_ptr())
    nseq = 50
    cuts = np.linspace(0, n, nseq + 1).astype(np.uint64)
    d_offs = torch_cuda.from_numpy(cuts.view(np.int64)).to(dev)
    ms = mg.Modset(26, 31, 64, 17)
    t1 = ms.add_device(buf.data_ptr(), d_offs.data_ptr(), nseq, n)
    v1, d1, _ = ms.sorted_dump()
    assert int(d1.astype(np.int64).sum()) == t1
    assert abs(t1 - n / 64) < 0.02 * n / 64
    host = buf.cpu().numpy()
    t2 = ms.add(host, cuts, is_ascii=0)      # same data through the host path, chunked + pipelined
    assert t2 == t1
    v2, d2, _ = ms.sorted_dump()
    assert np.array_equal(v1, v2) and np.array_equal(d2, 2 * d1)
    h = ms.histogram()
    assert int(h.sum()) == ms.max == len(v1)
    ms.close()


# ------------------------------------------------------------- set operations --
def _build_pair(mg, orc, bits, k, d, data, offs, exact=True):
    ms = mg.Modset(bits, k, d, 17, exact_order=exact)
    oms = orc.modset_new(bits, k, d, 17)
    assert ms.add(data, offs, is_ascii=0) == orc.modset_add(oms, data, offs)
    return ms, oms


def test_prune_and_merge(mg, orc):
    """modsetDepthPrune and modsetMerge (modset.c:64-77, 106-128): identical arrays, index order included"""
    sp = he.read_spec(12345, 150000, 11, 2500, 3000)
    d1 = he.reads(sp, 0, 1500); o1 = np.arange(1501, dtype=np.uint64) * np.uint64(2500)
    d2 = he.reads(sp, 1000, 1200); o2 = np.arange(1201, dtype=np.uint64) * np.uint64(2500)
    for (k, d) in ((19, 31), (31, 64)):
        a, oa = _build_pair(mg, orc, 22, k, d, d1, o1)
        b, ob = _build_pair(mg, orc, 22, k, d, d2, o2)
        a.set_copy(3, 20, 40); orc._modset_setcopy(oa, 3, 20, 40)
        b.set_copy(2, 15, 30); orc._modset_setcopy(ob, 2, 15, 30)
        # prune b, then merge it into a
        b.prune(3, 25); orc._modset_prune(ob, 3, 25)
        for x, y in zip(b.export(), orc.modset_export(ob)):
            assert np.array_equal(x, y), (k, d, "prune")
        assert b.summary() == orc.modset_summary(ob)
        b.prune(5, 0); orc._modset_prune(ob, 5, 0)                    # max = 0: no upper bound
        for x, y in zip(b.export(), orc.modset_export(ob)):
            assert np.array_equal(x, y), (k, d, "prune0")
        assert a.merge(b) and orc._modset_merge(oa, ob) == 1
        for x, y in zip(a.export(), orc.modset_export(oa)):
            assert np.array_equal(x, y), (k, d, "merge")
        assert a.summary() == orc.modset_summary(oa)
        assert np.array_equal(a.histogram(), orc.modset_hist(oa))
        # incompatible hashers are refused (modset.c:111)
        c = mg.Modset(22, k, d + 1, 17)
        assert not a.merge(c)
        for m in (a, b, c):
            m.close()
        orc._modset_free(oa); orc._modset_free(ob)


def test_add_packed_is_add(mg, orc):
    """modgpuModsetAddPacked: the batch in the reference's own 2-bit packing (sqioSeqPack, seqio.c:557-570) gives the
    modset of the same batch as bytes - ragged lengths around the byte and k boundaries, empty sequences, and a batch
    of 150-base reads where nearly every 16-base group of the expansion touches a sequence boundary"""
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 17, 18, 19, 20, 30, 31, 32, 33, 63, 64, 65, 150, 151, 1000, 0, 4097] + [int(x) for x in rng.integers(0, 400, 300)]
    for case in range(2):
        if case == 1:
            lens = [150] * 20000 + [149, 151, 2]
        offs = np.zeros(len(lens) + 1, np.uint64); offs[1:] = np.cumsum(lens)
        codes = rng.integers(0, 4, int(offs[-1])).astype(np.uint8)
        packed, boffs = mg.seqio_pack(codes, offs)
        for (k, d) in ((19, 31), (31, 64), (1, 1), (4, 3)):
            a = mg.Modset(22, k, d, 17); b = mg.Modset(22, k, d, 17); o = orc.modset_new(22, k, d, 17)
            ta = a.add(codes, offs, is_ascii=0)
            tb = b.add_packed(packed, boffs, offs)
            to = orc.modset_add(o, codes, offs)
            assert ta == tb == to, (case, k, d)
            for x, y in zip(a.sorted_dump(), b.sorted_dump()):
                assert np.array_equal(x, y), (case, k, d)
            ov, od, oi = orc.modset_sorted(o)
            gv, gd, gi = b.sorted_dump()
            assert np.array_equal(gv, ov) and np.array_equal(gd, od), (case, k, d)
            a.close(); b.close(); orc._modset_free(o)
    # packed bytes shorter than (len+3)/4 are refused
    bad = mg.Modset(22, 19, 31, 17)
    with pytest.raises(mg.ModgpuError):
        bad.add_packed(packed, np.zeros(len(boffs), np.uint64), offs)
    bad.close()


def test_info_flags_survive_device_ops(mg, tmp_path):
    """Modset.info is a whole byte (modset.h:49-52: MS_MINOR 4, MS_REPEAT 8, MS_INTERNAL 0x10, MS_RDNA 0x20 beside the
    two copy bits): import / export, prune (modset.c:73), merge (modset.c:124-125: touched entries are masked to the
    copy bits, the others keep their flags) and the .mod round trip keep it; with the stock modutils on the box the
    same through its -rt / -p / -m / -wt commands"""
    import subprocess
    rng = np.random.default_rng(99)
    sp = he.read_spec(777, 120000, 5, 2000, 2000)
    d1 = he.reads(sp, 0, 900); o1 = np.arange(901, dtype=np.uint64) * np.uint64(2000)
    d2 = he.reads(sp, 600, 700); o2 = np.arange(701, dtype=np.uint64) * np.uint64(2000)
    flags = np.array([0, 4, 8, 0x10, 0x20, 0x3C, 0x14, 0], np.uint8)

    def flagged(data, offs):
        t = mg.Modset(20, 19, 31, 17)
        t.add(data, offs, is_ascii=0)
        t.set_copy(3, 20, 40)
        v, d, i = t.export()
        t.close()
        i = (i | flags[rng.integers(0, len(flags), len(i))]).astype(np.uint8)
        m = mg.Modset(20, 19, 31, 17)
        m.import_entries(v, d, i)
        return m, v, d, i

    a, va, da, ia = flagged(d1, o1)
    b, vb, db, ib = flagged(d2, o2)
    assert (ia >= 4).sum() > 100 and (ib >= 4).sum() > 100
    for x, y in zip(a.export(), (va, da, ia)):
        assert np.array_equal(x, y), "import/export"
    # .mod round trip (modsetWrite / modsetRead, modset.c:79-104)
    path = str(tmp_path / "flags.mod")
    a.write_mod(path, gzip=True)
    back = mg.Modset.read_mod(path)
    for x, y in zip(back.export(), (va, da, ia)):
        assert np.array_equal(x, y), "mod round trip"
    back.close()
    # prune: survivors in their old order with their whole info byte
    keep = (db >= 3) & (db < 25)
    b.prune(3, 25)
    for x, y in zip(b.export(), (vb[keep], db[keep], ib[keep])):
        assert np.array_equal(x, y), "prune"
    vb, db, ib = vb[keep], db[keep], ib[keep]
    # merge b into a: the reference's loop restated on the arrays
    pos = {int(k): j for j, k in enumerate(va)}
    ev, ed, ei = list(va), [int(x) for x in da], [int(x) for x in ia]
    for k, d, i in zip(vb, db, ib):
        j = pos.get(int(k))
        if j is None:
            j = len(ev); pos[int(k)] = j
            ev.append(k); ed.append(0); ei.append(0)
        ed[j] = min(ed[j] + int(d), 65535)
        ei[j] = (ei[j] & 3) | min((ei[j] & 3) + (int(i) & 3), 3)
    assert a.merge(b)
    gv, gd, gi = a.export()
    assert np.array_equal(gv, np.array(ev, np.uint64)) and np.array_equal(gd, np.array(ed, np.uint16)), "merge"
    assert np.array_equal(gi, np.array(ei, np.uint8)), "merge info"
    assert (gi >= 4).sum() > 50                                        # untouched entries kept their flags
    a.close(); b.close()
    # the stock tool on a flagged set: -rt reads any info value, -p / -wt carry it
    modutils = H.ref_cli("modutils")
    if modutils:
        txt = str(tmp_path / "flags.txt")
        with open(txt, "w") as f:
            f.write("modset bits 20 size %d k 19 w 31 seed 17\n" % (len(va) + 1))
            for j in range(len(va)):
                f.write("%d\t%s\t%d\t%d\n" % (j + 1, H.kmer_string(va[j], 19), da[j], ia[j]))
        smod, sout = str(tmp_path / "stock.mod"), str(tmp_path / "stock_pruned.txt")
        r = subprocess.run([modutils, "-rt", txt, "-w", smod, "-p", "4", "30", "-wt", sout], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        g = mg.Modset.read_mod(smod)
        for x, y in zip(g.export(), (va, da, ia)):
            assert np.array_equal(x, y), "stock .mod with flags"
        g.prune(4, 30)
        gout = str(tmp_path / "gpu_pruned.txt")
        g.write_text(gout)
        assert open(gout).read() == open(sout).read()
        g.close()


def test_mod_file_roundtrip_with_stock_modutils(mg, orc, tmp_path):
    """modsetWrite / modsetRead (modset.c:79-104): a GPU-built .mod is loaded by the UNMODIFIED reference tool
    (oracle/_ref/modutils -r), and a .mod written by the reference tool is loaded by us"""
    import subprocess
    sp = he.read_spec(4242, 100000, 3, 2000, 1000)
    data = he.reads(sp, 0, 900); offs = np.arange(901, dtype=np.uint64) * np.uint64(2000)
    ms, oms = _build_pair(mg, orc, 20, 19, 31, data, offs)
    ms.set_copy(3, 20, 40); orc._modset_setcopy(oms, 3, 20, 40)
    ov, od, oi = orc.modset_export(oms)
    for gz in (False, True):
        path = str(tmp_path / ("g%d.mod" % gz))
        ms.write_mod(path, gzip=gz)
        back = mg.Modset.read_mod(path)                                # our own reader
        for x, y in zip(back.export(), (ov, od, oi)):
            assert np.array_equal(x, y)
        assert back.summary() == orc.modset_summary(oms)
        assert (back.k, back.w, back.seed, back.factor1) == (19, 31, 17, ms.factor1)
        back.close()
    modutils = H.ref_cli("modutils")
    if modutils:
        wt, his = str(tmp_path / "dump.txt"), str(tmp_path / "dump.his")
        r = subprocess.run([modutils, "-r", str(tmp_path / "g1.mod"), "-wt", wt, "-H", his], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exp = ["modset bits 20 size %d k 19 w 31 seed 17" % (len(ov) + 1)]
        exp += ["%d\t%s\t%d\t%d" % (j + 1, H.kmer_string(ov[j], 19), od[j], oi[j]) for j in range(len(ov))]
        assert open(wt).read().splitlines() == exp
        assert orc.modset_summary(oms) in r.stdout
        # the other direction: the reference tool writes (gzip'd through fzopen), we read; and it can
        # keep working on a set we wrote: lookups through ITS index[] (-P refpaint uses modsetIndexFind)
        fa = str(tmp_path / "r.fa")
        H.write_fasta(fa, [data[int(offs[j]):int(offs[j + 1])] for j in range(200)])
        ref_mod = str(tmp_path / "ref.mod")
        subprocess.run([modutils, "-c", "20", "19", "31", "17", "-a", fa, "-w", ref_mod], check=True, capture_output=True)
        theirs = mg.Modset.read_mod(ref_mod)
        o2 = orc.modset_new(20, 19, 31, 17)
        orc.modset_add(o2, data[:int(offs[200])], offs[:201])
        for x, y in zip(theirs.export(), orc.modset_export(o2)):
            assert np.array_equal(x, y)
        theirs.close(); orc._modset_free(o2)
        paint = subprocess.run([modutils, "-r", str(tmp_path / "g0.mod"), "-P", fa], capture_output=True, text=True)
        assert paint.returncode == 0 and paint.stdout.count("\n  ") > 1000          # hits found through the rebuilt index[]
    ms.close(); orc._modset_free(oms)


def test_readset_loop(mg, orc):
    """modasm readsetFileRead hot loop (modasm.c:151-191) against the oracle"""
    g = he.genome(777, 0, 200000, 0)
    offs = np.array([0, 120000, 120000, 200000], np.uint64)
    sp = he.read_spec(777, 200000, 9, 3000, 20000)
    rd = he.reads(sp, 0, 300); roffs = np.arange(301, dtype=np.uint64) * np.uint64(3000)
    for (k, d) in ((19, 31), (31, 64), (15, 4)):
        ms, oms = _build_pair(mg, orc, 22, k, d, g, offs)
        for rep in range(2):                                            # twice: depth is reset and re-counted each time
            gr = ms.readset(rd, roffs, is_ascii=0)
            orr = orc.readset(oms, rd, roffs)
            for key in orr:
                assert np.array_equal(gr[key], orr[key]), (k, d, key)
            assert len(orr["hit"]) > 100
            for x, y in zip(ms.export(), orc.modset_export(oms)):
                assert np.array_equal(x, y), (k, d, "depth recount")
        ms.close(); orc._modset_free(oms)


# ------------------------------------------------------------------ golden --
def _golden_cases():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    return G


@pytest.mark.parametrize("name", sorted(_golden_cases().cases()))
def test_gpu_matches_golden(mg, name):
    """the CUDA path against the vectors generated from the unmodified reference
    (tests/golden/golden_v1.json): identical lists, identical index numbering,
    identical histogram / summary text / Reference arrays / query results"""
    import json, os
    from gpu_checker import GpuChecker
    G = _golden_cases()
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))
    got = json.loads(json.dumps(G.evaluate(GpuChecker(), name, G.cases()[name])))
    assert got == golden[name]


# --------------------------------------------------------------- multi-GPU --
def test_peer_buckets_single_gpu(mg, torch_cuda, orc):
    """fully fused multi-GPU building blocks on one GPU: hash_select into per-(owner, region) buckets with tiny
    bucket capacity (forces the overflow segments), then every owner's share is built into its own table:
    the union of the G tables == the plain modset"""
    import ctypes as C
    from modimizer_b200 import _lib
    lib = _lib.load()
    sp = he.read_spec(12345, 300000, 42, 3000, 2000)
    nreads = 2000
    data = he.reads(sp, 0, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(3000)
    dev = torch_cuda.device("cuda:0")
    d_b = torch_cuda.from_numpy(data).to(dev); d_o = torch_cuda.from_numpy(offs.view(np.int64)).to(dev)
    for (k, d, G, cap) in ((31, 64, 2, 64), (19, 31, 3, 4), (19, 31, 1, 200)):
        src = mg.Modset(22, k, d, 17)
        src.set_stream(torch_cuda.cuda.current_stream().cuda_stream)
        R = lib.modgpuModsetRegions(src._p)
        oms = orc.modset_new(22, k, d, 17)
        tot = orc.modset_add(oms, data, offs)
        ovf_cap = tot + 16
        sb = torch_cuda.zeros(G * R * cap, dtype=torch_cuda.int64, device=dev)
        sc = torch_cuda.zeros(G * R, dtype=torch_cuda.int32, device=dev)
        so = torch_cuda.zeros(G * ovf_cap, dtype=torch_cuda.int64, device=dev)
        soc = torch_cuda.zeros(G, dtype=torch_cuda.int32, device=dev)
        cnt = torch_cuda.zeros(1, dtype=torch_cuda.int64, device=dev)
        _lib.check(lib.modgpuModsetSelectBucketsDevice(src._p, d_b.data_ptr(), d_o.data_ptr(), nreads, len(data), 0, G, sb.data_ptr(), cap,
                                                       sc.data_ptr(), so.data_ptr(), ovf_cap, soc.data_ptr(), cnt.data_ptr()), "selectBuckets")
        torch_cuda.cuda.synchronize()
        assert int(cnt.item()) == tot
        allv, alld = [], []
        for g in range(G):                                   # owner g builds from "its" slice of the send buffer
            t = mg.Modset(22, k, d, 17)
            t.set_stream(torch_cuda.cuda.current_stream().cuda_stream)
            _lib.check(lib.modgpuModsetBuildFromBuckets(t._p, sb.data_ptr() + g * R * cap * 8, sc.data_ptr() + g * R * 4, cap, 1,
                                                        so.data_ptr() + g * ovf_cap * 8, ovf_cap, soc.data_ptr() + g * 4), "build")
            v, dd, _ = t.sorted_dump()
            assert all(lib.modgpuOwnerOf(int(x), G) == g for x in v[:200])
            allv.append(v); alld.append(dd)
            t.close()
        v = np.concatenate(allv); dd = np.concatenate(alld)
        o = np.argsort(v)
        ov, od, _ = orc.modset_sorted(oms)
        assert np.array_equal(v[o], ov) and np.array_equal(dd[o], od), (k, d, G)
        src.close(); orc._modset_free(oms)


def test_owner_segments_single_gpu(mg, torch_cuda, orc):
    """the fused multi-GPU building blocks on one GPU: hash_select into per-owner segments, then a bulk
    insert of those segments (as if received from peers) == the plain modset"""
    import ctypes as C
    from modimizer_b200 import _lib
    lib = _lib.load()
    sp = he.read_spec(12345, 300000, 42, 3000, 2000)
    nreads = 2000
    data = he.reads(sp, 0, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(3000)
    dev = torch_cuda.device("cuda:0")
    d_b = torch_cuda.from_numpy(data).to(dev); d_o = torch_cuda.from_numpy(offs.view(np.int64)).to(dev)
    for (k, d, G) in ((31, 64, 3), (19, 31, 8), (19, 4, 2)):
        ms = mg.Modset(22, k, d, 17)
        ms.set_stream(torch_cuda.cuda.current_stream().cuda_stream)
        oms = orc.modset_new(22, k, d, 17)
        tot = orc.modset_add(oms, data, offs)
        cap = int(tot / G * 1.3) + 1000
        seg = torch_cuda.zeros(G * cap, dtype=torch_cuda.int64, device=dev)
        cnt = torch_cuda.zeros(G, dtype=torch_cuda.int32, device=dev)
        _lib.check(lib.modgpuModsetSelectOwnersDevice(ms._p, d_b.data_ptr(), d_o.data_ptr(), nreads, len(data), 0, G,
                                                      seg.data_ptr(), cap, cnt.data_ptr()), "selectOwners")
        torch_cuda.cuda.synchronize()
        c = cnt.cpu().numpy()
        assert int(c.sum()) == tot and (c <= cap).all()
        segs = seg.cpu().numpy().view(np.uint64).reshape(G, cap)
        for g in range(G):
            assert all(lib.modgpuOwnerOf(int(x), G) == g for x in segs[g, :min(int(c[g]), 300)])
        _lib.check(lib.modgpuModsetInsertSegments(ms._p, seg.data_ptr(), G, cap, cnt.data_ptr(), tot), "insertSegments")
        gv, gd, _ = ms.sorted_dump(); ov, od, _ = orc.modset_sorted(oms)
        assert np.array_equal(gv, ov) and np.array_equal(gd, od), (k, d, G)
        ms.close(); orc._modset_free(oms)



def _sharded_worker(rank, world, port, tmpdir):
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from modimizer_b200 import _lib
    from modimizer_b200.dist import ShardedModset
    _lib.check(_lib.load().modgpuSetDevice(rank))
    import hostemul as he
    sp = he.read_spec(12345, 200000, 3, 5000)
    per = 600 // world
    data = he.reads(sp, rank * per, per)
    offs = np.arange(per + 1, dtype=np.uint64) * np.uint64(5000)
    sm = ShardedModset(22, 19, 31, 17)
    sm.add(data, offs)                                        # host path (fused exchange)
    n1 = sm.synchronize()
    d_b = torch.from_numpy(data).cuda(); d_o = torch.from_numpy(offs.view(np.int64)).cuda()
    sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))        # device path, same data again
    n2 = sm.synchronize()
    assert n1 == n2 and n1 > 0
    p2p = sm.fused_mode == "p2p"                              # peer-memory exchange really in use (CUDA IPC worked)
    sm.fused_mode = "peer"                                    # the same buckets through an NCCL all-to-all
    sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))
    assert sm.synchronize() == n1
    sm.fused_mode = "segments"                                # per-owner segments + scatter at the receiver
    sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))
    assert sm.synchronize() == n1
    sm.fused = False                                          # the list-based exchange, once more
    n3 = sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))
    assert n3 == n1
    # deferred peer build: three batches share one exchange + build, a fourth waits until synchronize()
    sm.synchronize()                                          # (the list exchange also counts towards synchronize())
    sm.fused, sm.fused_mode = True, ("p2p" if p2p else "peer")
    sm.set_accumulate(3)
    for _ in range(4):
        sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))
    assert sm.synchronize() == 4 * n1
    # a skewed group is transactional: rank 0 adds poly-A (every window selected, one k-mer: bucket and overflow
    # segment run over), rank 1 its normal chunk - synchronize() raises on BOTH ranks and NOTHING was applied on either;
    # after set_robust() the same group goes through
    skew_ok = True
    if p2p:
        from modimizer_b200._lib import ModgpuError
        sm.set_accumulate(1)
        skew = np.zeros(len(data), np.uint8)
        d_s = torch.from_numpy(skew).cuda(); d_so = torch.tensor([0, len(skew)], dtype=torch.int64).cuda()
        before = (sm.local.max, sm.histogram().copy())
        def group():
            if rank == 0:
                sm.add_device(d_s.data_ptr(), d_so.data_ptr(), 1, len(skew))
            else:
                sm.add_device(d_b.data_ptr(), d_o.data_ptr(), per, len(data))
        group()
        try:
            sm.synchronize()
            skew_ok = False                                   # must not pass
        except ModgpuError:
            pass
        skew_ok = skew_ok and sm.local.max == before[0] and np.array_equal(sm.histogram(), before[1])
        sm.set_robust(True)
        group()
        sm.synchronize()
        flag = torch.tensor([1 if skew_ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        skew_ok = bool(flag.item())
    v, d, i = sm.gather_sorted_dump()
    h = sm.histogram()
    gmax = sm.global_max()
    if rank == 0:
        np.savez(os.path.join(tmpdir, "sharded.npz"), v=v, d=d, h=h, gmax=gmax, sel=n1, p2p=p2p, skew_ok=skew_ok)
    sm.close()
    dist.destroy_process_group()


def test_sharded_modset_two_gpus(mg, torch_cuda, orc, tmp_path):
    """hash-sharded table over 2 GPUs - peer-memory exchange (region build reading the other rank's buckets over
    NVLink), NCCL bucket all-to-all, per-owner segments, list exchange: union == single-GPU == oracle"""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_sharded_worker, args=(world, 29800 + os.getpid() % 1000, str(tmp_path)), nprocs=world, join=True)
    r = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    sp = he.read_spec(12345, 200000, 3, 5000)
    data = he.reads(sp, 0, 600)
    offs = np.arange(601, dtype=np.uint64) * np.uint64(5000)
    oms = orc.modset_new(22, 19, 31, 17)
    assert bool(r["p2p"]), "the peer-memory exchange fell back to NCCL"
    for _ in range(9):                                        # every rank added its chunk nine times
        orc.modset_add(oms, data, offs)
    assert bool(r["skew_ok"]), "a skewed group was not skipped on every rank with the tables untouched"
    orc.modset_add(oms, np.zeros(300 * 5000, np.uint8), np.array([0, 300 * 5000], np.uint64))      # rank 0's poly-A batch
    orc.modset_add(oms, data[300 * 5000:], offs[:301])                                             # rank 1's chunk of that group
    ov, od, _ = orc.modset_sorted(oms)
    assert np.array_equal(r["v"], ov) and np.array_equal(r["d"], od)
    assert np.array_equal(r["h"], orc.modset_hist(oms)) and int(r["gmax"]) == len(ov)
    orc._modset_free(oms)
