"""The kernels' per-base arithmetic (modimizer_b200/csrc/mg_common.cuh, compiled
for the host by tests/host/) against the oracle: packing layout, window
extraction, canonical choice, the division-free divisibility test, the
power-of-two prefilter, read-boundary masks.  No GPU needed."""
import numpy as np
import pytest

import harness as H
import hostemul as he


def oracle_select(orc, k, d, seed, data, offs):
    ks, ps, fs = [], [], []
    for r in range(len(offs) - 1):
        a, b = int(offs[r]), int(offs[r + 1])
        kk, pp, ff = orc.mod_scan(k, d, seed, data[a:b])
        ks.append(kk); ps.append(pp.astype(np.int64) + a); fs.append(ff)
    return np.concatenate(ks), np.concatenate(ps), np.concatenate(fs)


def test_divisibility_without_division():
    rng = np.random.default_rng(0)
    lib = he.lib()
    for d in (1, 2, 3, 5, 7, 31, 48, 62, 64, 96, 1000, 4096, 65537, 2**31 - 1):
        vals = np.concatenate([rng.integers(0, 2**62, 2000, dtype=np.uint64), np.arange(0, 200, dtype=np.uint64) * np.uint64(d),
                               rng.integers(0, 2**40, 500, dtype=np.uint64) * np.uint64(d) % np.uint64(2**62)])
        for v in vals:
            assert lib.hm_divisible(d, int(v)) == (1 if int(v) % d == 0 else 0), (d, int(v))


def test_khasher_constants():
    out = np.zeros(8, np.uint64)
    he.lib().hm_khasher(31, 64, 0x49308bb9003cb3ad, out)
    assert int(out[1]) == 2 and int(out[2]) == 6 and int(out[5]) == 1          # shift, tz, prefilter on
    he.lib().hm_khasher(19, 31, 0x49308bb9003cb3ad, out)
    assert int(out[2]) == 0 and int(out[5]) == 0 and (int(out[3]) * 31) % 2**64 == 1
    he.lib().hm_khasher(12, 64, 0x49308bb9003cb3ad, out)
    assert int(out[5]) == 0                                                    # 64-2k+tz > 32: generic path


def test_pack_layout_and_ascii():
    rng = np.random.default_rng(1)
    for n in (1, 31, 32, 33, 64, 1000):
        codes = rng.integers(0, 4, n).astype(np.uint8)
        nw = (n + 31) // 32
        w = np.zeros(nw, np.uint64)
        he.lib().hm_pack(codes, n, 0, w, nw)
        for j in range(n):
            assert (int(w[j // 32]) >> (62 - 2 * (j % 32))) & 3 == codes[j]
        asc = np.frombuffer(b"ACGTacgtNn", np.uint8)[rng.integers(0, 10, n)]
        w2 = np.zeros(nw, np.uint64)
        he.lib().hm_pack(asc, n, 1, w2, nw)
        exp = H.codes_from_ascii(asc.tobytes())
        for j in range(n):
            assert (int(w2[j // 32]) >> (62 - 2 * (j % 32))) & 3 == exp[j]


@pytest.mark.parametrize("prefilter", [-1, 0])
def test_select_matches_oracle(orc, prefilter):
    rng = np.random.default_rng(2 + prefilter)
    ds = [1, 2, 3, 7, 8, 16, 31, 32, 48, 62, 64, 128, 1000, 4096]
    for trial in range(250):
        k = int(rng.integers(1, 32)); d = int(rng.choice(ds)); seed = int(rng.integers(0, 100))
        f1 = orc.hasher(k, d, seed)["factor1"]
        lens = rng.integers(0, 40 if trial % 5 == 0 else 300, int(rng.integers(1, 8)))
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        n = int(offs[-1])
        codes = rng.integers(0, 4, max(n, 1)).astype(np.uint8)
        if trial % 7 == 0:
            codes[:] = 0
        if trial % 11 == 0:
            codes = np.tile(np.array([0, 3], np.uint8), n // 2 + 1)[:max(n, 1)].copy()
        ek, ep, ef = oracle_select(orc, k, d, seed, codes, offs)
        km, gp, isf = he.select(k, d, f1, codes[:n], offs, prefilter=prefilter)
        assert np.array_equal(km, ek) and np.array_equal(gp.astype(np.int64), ep) and np.array_equal(isf, ef), (k, d, seed, trial)


@pytest.mark.parametrize("k", [16, 17, 19, 24, 31])
def test_full_scan_32bit_evaluation(orc, k):
    """the kernel's 32-bit evaluation of the generic-d full scan (mg_selected32, k >= 16) against the oracle and, inside
    the host build, against the 64-bit evaluation at every window: odd d, 2*odd, 4*odd, poly-A, AT repeats"""
    rng = np.random.default_rng(100 + k)
    for d in (31, 1, 3, 5, 62, 124, 2, 4, 999, 65537):
        f1 = orc.hasher(k, d, 17)["factor1"]
        lens = rng.integers(0, 4000, 6)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        n = int(offs[-1])
        codes = rng.integers(0, 4, max(n, 1)).astype(np.uint8)
        codes[int(offs[1]):int(offs[2])] = 0
        codes[int(offs[2]):int(offs[3])] = np.tile(np.array([0, 3], np.uint8), n)[:int(offs[3] - offs[2])]
        ek, ep, ef = oracle_select(orc, k, d, 17, codes, offs)
        km, gp, isf = he.select(k, d, f1, codes[:n], offs, prefilter=0)
        assert np.array_equal(km, ek) and np.array_equal(gp.astype(np.int64), ep) and np.array_equal(isf, ef), (k, d)


@pytest.mark.parametrize("k,d", [(31, 64), (31, 32), (31, 8), (31, 192), (30, 16), (30, 8), (31, 128), (19, 32)])
def test_table_prefilter_equals_arithmetic_prefilter(k, d):
    """mg_lut_scan (one shared-memory lookup per 4 positions) marks exactly the windows the
    multiplicative low-word prefilter marks, wherever the kernel would use the table"""
    rng = np.random.default_rng(k * 1000 + d)
    data = rng.integers(0, 4, 200000, dtype=np.uint8)
    data[5000:5200] = 0                      # poly-A and a dinucleotide repeat
    data[9000:9400:2] = 0
    data[9001:9400:2] = 3
    for f1 in (0x49308bb9003cb3ad, 0x6b8b4567327b23c7, 0xffffffffffffffff, 1):
        bad, ncand = he.lut_check(k, d, f1, data)
        applies = (64 - 2 * k) + (d & -d).bit_length() - 1 <= 8 and k >= 30 and (d & -d) >= 8
        if not applies:
            assert bad == -1
            continue
        assert bad == 0
        assert ncand > 0


def test_synth_generators_are_stable():
    """the synthetic inputs are part of the golden contract: pin a few bytes"""
    g = he.genome(12345, 0, 64, 1)
    assert g.tolist()[:20] == [3, 2, 3, 1, 2, 1, 2, 0, 0, 0, 2, 0, 3, 3, 3, 0, 0, 2, 0, 2]
    # duplicated segments exist in dupMode 1
    a = he.genome(4242, 0, 1 << 22, 1)
    segs = a.reshape(-1, 1 << 16)
    import hashlib
    hs = [hashlib.md5(s.tobytes()).hexdigest() for s in segs]
    assert len(set(hs)) < len(hs)
    # pair mode: mate 1 is the reverse complement of the fragment's far end
    sp = he.read_spec(12345, 100000, 7, 150, 0, frag_len=400, pair_mode=1)
    r = he.reads(sp, 0, 2).reshape(2, 150)
    gen = he.genome(12345, 0, 100000, 0)
    fwd = "".join(map(str, r[0])); rc = "".join(map(str, (3 - r[1])[::-1]))
    s = "".join(map(str, gen))
    i, j = s.find(fwd), s.find(rc)
    if i < 0:                                  # fragment on the reverse strand: swap roles
        fwd = "".join(map(str, (3 - r[0])[::-1])); rc = "".join(map(str, r[1]))
        i, j = s.find(fwd), s.find(rc)
    assert i >= 0 and j >= 0 and abs(abs(i - j) - 250) == 0
