"""ctypes bindings of the two CPU checkers (oracle/harness_api.h).

TEST INFRASTRUCTURE: `orc` is oracle/liboracle.so (our restatement), `ref` is
oracle/_ref/libmodref.so (the unmodified reference objects behind the same API).
Neither is ever imported by the product package.
"""
import ctypes as C
import gzip
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")


def build_oracle():
    """(re)build liboracle.so, and oracle/_ref when /root/reference is present"""
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL)


class Checker:
    """One of the two implementations of harness_api.h"""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.name = prefix.rstrip("_")
        f = self._f
        f("hasher", None, [C.c_int, C.c_int, C.c_int, u64p])
        f("mod_scan", C.c_int64, [C.c_int, C.c_int, C.c_int, u8p, C.c_int, u64p, i32p, u8p, C.c_int64])
        f("modset_new", C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int])
        f("modset_free", None, [C.c_void_p])
        f("modset_prefault", None, [C.c_void_p])
        f("modset_add", C.c_uint64, [C.c_void_p, u8p, u64p, C.c_int64])
        f("modset_max", C.c_uint32, [C.c_void_p])
        f("modset_export", None, [C.c_void_p, u64p, u16p, u8p])
        f("modset_find", C.c_uint32, [C.c_void_p, C.c_uint64])
        f("modset_setcopy", None, [C.c_void_p, C.c_int, C.c_int, C.c_int])
        f("modset_setcopyM", None, [C.c_void_p, C.c_int])
        f("modset_hist", None, [C.c_void_p, u32p])
        f("modset_summary", C.c_int, [C.c_void_p, C.c_char_p, C.c_int])
        f("modset_prune", None, [C.c_void_p, C.c_int, C.c_int])
        f("modset_merge", C.c_int, [C.c_void_p, C.c_void_p])
        f("readset", C.c_int64, [C.c_void_p, u8p, u64p, C.c_int64, u64p, u32p, u16p, i32p, C.c_int64])
        f("ref_build", C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, u8p, u64p, C.c_int64, u32p])
        f("ref_free", None, [C.c_void_p])
        f("ref_modset", C.c_void_p, [C.c_void_p])
        f("ref_max", C.c_uint32, [C.c_void_p])
        f("ref_export", None, [C.c_void_p, u32p, u32p, u32p, u32p, u32p, u32p])
        f("ref_query", C.c_int64, [C.c_void_p, u8p, u64p, C.c_int64, u64p, u32p, u32p, u32p, u32p, i32p, C.c_int64])
        if prefix == "ref_":
            f("ref_build_fasta", C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, u32p])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, "_" + name, fn)

    # ---- convenience wrappers -------------------------------------------
    def hasher(self, k, w, seed):
        out = np.zeros(4, np.uint64)
        self._hasher(k, w, seed, out)
        return dict(mask=int(out[0]), shift=int(out[1]), factor1=int(out[2]), factor2=int(out[3]))

    def mod_scan(self, k, w, seed, codes):
        codes = np.ascontiguousarray(codes, np.uint8)
        cap = max(1, len(codes))
        kmer = np.zeros(cap, np.uint64)
        pos = np.zeros(cap, np.int32)
        isf = np.zeros(cap, np.uint8)
        n = self._mod_scan(k, w, seed, codes if len(codes) else np.zeros(1, np.uint8), len(codes), kmer, pos, isf, cap)
        return kmer[:n].copy(), pos[:n].copy(), isf[:n].copy()

    def modset_new(self, bits, k, w, seed):
        return self._modset_new(bits, k, w, seed)

    def modset_add(self, ms, codes, offs):
        codes = np.ascontiguousarray(codes, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        if len(codes) == 0:
            codes = np.zeros(1, np.uint8)
        return int(self._modset_add(ms, codes, offs, len(offs) - 1))

    def modset_export(self, ms):
        n = self._modset_max(ms)
        v = np.zeros(max(n, 1), np.uint64)
        d = np.zeros(max(n, 1), np.uint16)
        i = np.zeros(max(n, 1), np.uint8)
        self._modset_export(ms, v, d, i)
        return v[:n], d[:n], i[:n]

    def modset_sorted(self, ms):
        """(kmer, depth, info&3) sorted by k-mer: the parity key of SURVEY 8(c)"""
        v, d, i = self.modset_export(ms)
        o = np.argsort(v, kind="stable")
        return v[o], d[o], (i[o] & 3)

    def modset_hist(self, ms):
        b = np.zeros(65536, np.uint32)
        self._modset_hist(ms, b)
        return b

    def modset_summary(self, ms):
        buf = C.create_string_buffer(4096)
        n = self._modset_summary(ms, buf, 4096)
        return buf.raw[:n].decode()

    def readset(self, ms, codes, offs):
        codes = np.ascontiguousarray(codes, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        nseq = len(offs) - 1
        cap = max(1, len(codes))
        ho = np.zeros(nseq + 1, np.uint64)
        hit = np.zeros(cap, np.uint32)
        dx = np.zeros(cap, np.uint16)
        miss = np.zeros(max(nseq, 1), np.int32)
        n = self._readset(ms, codes, offs, nseq, ho, hit, dx, miss, cap)
        return dict(hitOff=ho, hit=hit[:n], dx=dx[:n], nMiss=miss[:nseq])

    def ref_build(self, bits, k, w, seed, codes, offs):
        codes = np.ascontiguousarray(codes, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        counts = np.zeros(4, np.uint32)
        r = self._ref_build(bits, k, w, seed, codes, offs, len(offs) - 1, counts)
        return r, counts

    def ref_export(self, r):
        n = self._ref_max(r)
        m = self._modset_max(self._ref_modset(r)) + 1
        a = [np.zeros(max(n, 1), np.uint32) for _ in range(3)]
        depth = np.zeros(m, np.uint32)
        rev = np.zeros(max(n, 1), np.uint32)
        loc = np.zeros(m, np.uint32)
        self._ref_export(r, a[0], a[1], a[2], depth, rev, loc)
        return dict(index=a[0][:n], offset=a[1][:n], id=a[2][:n], depth=depth, rev=rev[:n], loc=loc)

    def ref_query(self, r, codes, offs):
        codes = np.ascontiguousarray(codes, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        nseq = len(offs) - 1
        cap = max(1, len(codes))
        so = np.zeros(nseq + 1, np.uint64)
        si = np.zeros(cap, np.uint32)
        sp = np.zeros(cap, np.uint32)
        hid = np.zeros(2 * cap, np.uint32)
        hoff = np.zeros(2 * cap, np.uint32)
        ctr = np.zeros(4 * max(nseq, 1), np.int32)
        n = self._ref_query(r, codes, offs, nseq, so, si, sp, hid, hoff, ctr, cap)
        return dict(seedOff=so, index=si[:n], pos=sp[:n], hitId=hid[:2 * n].reshape(-1, 2),
                    hitOffset=hoff[:2 * n].reshape(-1, 2), counters=ctr[:4 * nseq].reshape(-1, 4))


_cache = {}


def oracle():
    if "orc" not in _cache:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _cache["orc"] = Checker(path, "orc_")
    return _cache["orc"]


def reference():
    """the unmodified reference behind the harness API, or None when oracle/_ref was never built"""
    if "ref" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libmodref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/seqhash.c"):
            build_oracle()
        _cache["ref"] = Checker(path, "ref_") if os.path.exists(path) else None
    return _cache["ref"]


def ref_cli(name):
    p = os.path.join(ORACLE_DIR, "_ref", name)
    return p if os.path.exists(p) else None


# ---- sequence helpers ----------------------------------------------------
ASCII = np.frombuffer(b"ACGT", np.uint8)


def codes_from_ascii(s):
    """reference dna2indexConv (seqio.c:643-652) with the N->0 patch (modutils.c:39)"""
    t = np.zeros(256, np.uint8)
    for ch, v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("N", 0)):
        t[ord(ch)] = v
        t[ord(ch.lower())] = v
    return t[np.frombuffer(s.encode() if isinstance(s, str) else s, np.uint8)]


def kmer_string(v, k):
    """seqString (seqhash.c:198-206): lower-case, most significant base first"""
    return "".join("acgt"[(int(v) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def write_fasta(path, seqs, names=None, width=0):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s\n" % (names[i] if names else "s%d" % i))
            txt = ASCII[np.asarray(s, np.uint8)].tobytes().decode()
            if width:
                for j in range(0, len(txt), width):
                    f.write(txt[j:j + width] + "\n")
            else:
                f.write(txt + "\n")


def modmap_case(d, seed=33):
    """g.fa / r.fa for the modmap driver tests: three reference sequences with a two-copy, a three-copy and an
    inverted segment; reads of both strands with 3 % substitutions, plus a long two-copy read, len < k, empty
    and an unrelated read"""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, 400000).astype(np.uint8)
    genome[300000:330000] = genome[50000:80000]
    genome[120000:126000] = genome[50000:56000]
    genome[350000:358000] = (3 - genome[10000:18000])[::-1]
    chrom = [genome[:100000], genome[100000:250000], genome[250000:]]
    reads = []
    for _ in range(400):
        s = int(rng.integers(0, len(genome) - 6000)); L = int(rng.integers(300, 6000))
        r = genome[s:s + L].copy()
        if rng.integers(0, 2):
            r = (3 - r)[::-1]
        e = rng.random(len(r)) < 0.03
        r[e] = (r[e] + rng.integers(1, 4, int(e.sum()))) & 3
        reads.append(r)
    reads += [genome[299000:331000].copy(), genome[49000:57000].copy(), (3 - genome[349000:359000])[::-1].copy(),
              genome[:18], genome[7:7], rng.integers(0, 4, 3000).astype(np.uint8)]
    write_fasta(os.path.join(d, "g.fa"), chrom, names=["chrA", "chrB", "chrC"], width=60)
    write_fasta(os.path.join(d, "r.fa"), reads, width=0)


def stable_lines(text):
    """the lines of a tool's output that do not carry rusage numbers (utils.c:176-204)"""
    return [l for l in text.splitlines() if not l.startswith("user\t") and not l.startswith("total resources used")]


def ref_file_masked(raw):
    """a .ref file (modmap.c:136-156) with the heap pointers that arrayWrite and dictWrite dump zeroed: locates the
    trailing Array (32-byte struct, base pointer at 8) and DICT (dim, max, table[1<<dim], names[max+1] pointers, strings)
    from the end of the U32 arrays by scanning for the Array magic"""
    b = bytearray(raw)
    assert b[:8] == b"RFMSHv1\0"
    n = int(np.frombuffer(raw[8:12], np.uint32)[0])
    # index/offset/id/rev have n entries, depth/loc have m: the Array struct follows; its dim/size/max are at +16/+20/+24
    at = None
    for m in range(1, n + 2):
        o = 16 + 16 * n + 8 * m
        if o + 32 <= len(b) and np.frombuffer(raw[o + 20:o + 24], np.int32)[0] == 4 and 0 < np.frombuffer(raw[o + 16:o + 20], np.int32)[0] < (1 << 24) \
                and np.frombuffer(raw[o + 24:o + 28], np.int32)[0] <= np.frombuffer(raw[o + 16:o + 20], np.int32)[0]:
            dim = int(np.frombuffer(raw[o + 16:o + 20], np.int32)[0])
            do = o + 32 + 4 * dim
            if do + 8 <= len(b):
                ddim, dmax = [int(x) for x in np.frombuffer(raw[do:do + 8], np.int32)]
                if 0 < ddim < 30 and 0 < dmax < (1 << ddim) and do + 8 + 4 * (1 << ddim) + 8 * (dmax + 1) <= len(b):
                    at = (o, do, ddim, dmax)
                    break
    assert at, "no Array / DICT tail found in the .ref file"
    o, do, ddim, dmax = at
    b[o + 8:o + 16] = bytes(8)
    p = do + 8 + 4 * (1 << ddim)
    b[p:p + 8 * (dmax + 1)] = bytes(8 * (dmax + 1))
    return bytes(b)


def modmap_driver_vs_stock(stock, driver, d, check_mod):
    """runs the stock modmap and a driver with the same command lines in directory d and compares everything"""
    def run(tool, args):
        r = subprocess.run([tool] + args, cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, (tool, args, r.stdout[-500:], r.stderr[-500:])
        return r
    def rd(name, mode="r"):
        with open(os.path.join(d, name), mode) as f:
            return f.read()
    for (k, w) in ((19, 31), (31, 64), (15, 8)):
        par = ["-K", str(k), "-W", str(w), "-B", "20"]
        a = run(stock, ["-o", "a.out"] + par + ["-f", "g.fa", "-w", "a", "-v", "-q", "r.fa"])
        b = run(driver, ["-o", "b.out"] + par + ["-f", "g.fa", "-w", "b", "-v", "-q", "r.fa"])
        oa, ob = stable_lines(rd("a.out")), stable_lines(rd("b.out"))
        assert oa == ob, (k, w)
        assert any(l.startswith("M\t") for l in oa) and any(l.startswith("Q\t") for l in oa), (k, w)
        assert stable_lines(a.stdout) == stable_lines(b.stdout), (k, w)             # the -v seed lines
        assert any(l.startswith("  ") and "\t" in l for l in a.stdout.splitlines()), (k, w)
        # everything on one stream: seed lines and M lines interleave exactly as in the stock tool
        assert stable_lines(run(stock, par + ["-f", "g.fa", "-v", "-q", "r.fa"]).stdout) == \
               stable_lines(run(driver, par + ["-f", "g.fa", "-v", "-q", "r.fa"]).stdout), (k, w)
        # the index files (gzip streams, utils.c:108-139): same content - first-occurrence numbering, the reference's
        # probe order, packed arrays.  arrayWrite / dictWrite dump heap pointers (array.c:215, dict.c:95): masked.
        ra, rb = ref_file_masked(gzip.decompress(rd("a.ref", "rb"))), ref_file_masked(gzip.decompress(rd("b.ref", "rb")))
        assert ra == rb, (k, w, "ref")
        if check_mod:
            assert gzip.decompress(rd("a.mod", "rb")) == gzip.decompress(rd("b.mod", "rb")), (k, w, "mod")
        if check_mod:                          # and the stock modmap maps from the files written by the driver
            run(stock, ["-o", "c.out", "-r", "b", "-q", "r.fa"])
            pick = lambda ls: [l for l in ls if l[:2] in ("Q\t", "M\t")]
            assert pick(stable_lines(rd("c.out"))) == pick(oa), (k, w)
