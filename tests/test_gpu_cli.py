"""The C host side of the drop-in boundary on a GPU box:

  * modimizer_b200/modutils_shim, modmap_shim - the reference's modutils.c / modmap.c, UNMODIFIED, linked against
    libmodshim.so (seqhash.h + modset.h under their own names, served by the GPU) instead of seqhash.o / modset.o;
  * modimizer_b200/modutils_dropin, modmap_dropin - the same sources with the caller loops of INTEGRATION.md sections
    1-2 swapped in at build time (csrc/shim/Makefile: sed + modutils_hot.c / modmap_hot.c): the batched path;
  both against the STOCK tools (oracle/_ref/modutils, modmap): every file and every stable output line byte-identical;

  * modimizer_b200/modutils_gpu - a C modutils (our own command interpreter) in which the reference's seqio parses the
    files and libmodgpu does the rest - against the STOCK modutils (oracle/_ref/modutils) on the same files: every file
    they write and every stable line they print must be byte-identical, and each reads the other's .mod files;
  * modimizer_b200/modmap_gpu - a C modmap over libmodgpu (index build, .mod/.ref writer, Q / seed / M lines) - against
    the STOCK modmap: same output bytes, same .mod and .ref bytes, and the stock tool maps from the files written here;
  * modimizer_b200/libmodshim.so - the reference's own seqhash symbols (seqhashCreate, modRCiterator, modRCnext,
    seqString) served by the GPU - against the oracle, through the reference's own struct layouts.

Both are built in the container (they compile against the reference's headers in place) and travel to the box."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_CLI = os.path.join(ROOT, "modimizer_b200", "modutils_gpu")
GPU_MODMAP = os.path.join(ROOT, "modimizer_b200", "modmap_gpu")
TOOL = lambda name: os.path.join(ROOT, "modimizer_b200", name)
SHIM = os.path.join(ROOT, "modimizer_b200", "libmodshim.so")


def stable(text):
    """the lines of a tool's output that do not carry rusage numbers (utils.c:176-204)"""
    return [l for l in text.splitlines() if not l.startswith("user\t") and not l.startswith("total resources used")]


def run(tool, args, cwd):
    r = subprocess.run([tool] + args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, (tool, args, r.stdout[-500:], r.stderr[-500:])
    return r


def make_reads(rng, genome, n, lo, hi):
    reads = []
    for _ in range(n):
        s = int(rng.integers(0, len(genome) - hi)); L = int(rng.integers(lo, hi))
        r = genome[s:s + L].copy()
        if rng.integers(0, 2):
            r = (3 - r)[::-1]
        reads.append(r)
    return reads


@pytest.mark.parametrize("tool", ["modutils_gpu", "modutils_dropin", "modutils_shim"])
def test_c_modutils_matches_stock_modutils(tmp_path, tool):
    """modutils_gpu: own driver over the ABI; modutils_dropin: the reference's modutils.c with addSequenceFile swapped at
    build time; modutils_shim: the reference's modutils.c unmodified on libmodshim (one device call per k-mer: smaller
    input).  The last two also run the commands our own driver does not have (-rt, -m, -d): reference code on
    libmodshim's modset.h symbols."""
    stock = H.ref_cli("modutils")
    exe = TOOL(tool)
    if not stock or not os.path.exists(exe):
        pytest.skip("stock modutils / %s not built (no /root/reference in the build container)" % tool)
    ref_source = tool != "modutils_gpu"
    small = tool == "modutils_shim"
    rng = np.random.default_rng(21)
    genome = rng.integers(0, 4, 300000).astype(np.uint8)
    genome[200000:220000] = genome[20000:40000]                      # a duplication: copy-2 k-mers
    reads = make_reads(rng, genome, 1200 if small else 9000, 20, 2500) + [genome[:18], genome[:19], genome[5:5]]   # incl. len < k, == k, empty
    more = make_reads(rng, genome, 150 if small else 800, 100, 1500)
    d = str(tmp_path)
    H.write_fasta(os.path.join(d, "r.fa"), reads, width=0)
    H.write_fasta(os.path.join(d, "m.fa"), more, width=70)
    H.write_fasta(os.path.join(d, "g.fa"), [genome[:60000], genome[60000:150000]], names=["chrA", "chrB"], width=60)
    # N runs and lower case in the text (N -> a, modutils.c:39)
    txt = open(os.path.join(d, "r.fa")).read().split("\n")
    txt[1] = txt[1][:50] + "NNNNNNNNnnnn" + txt[1][62:].lower()
    open(os.path.join(d, "r.fa"), "w").write("\n".join(txt))
    rd = lambda name: open(os.path.join(d, name), "rb").read()

    for (k, w) in ((19, 31), (31, 64)):
        def script(p):
            return ["-o", p + ".out", "-c", "22", str(k), str(w), "17", "-a", "r.fa", "-H", p + ".his", "-wt", p + ".txt",
                    "-s", "3", "9", "14", "-wt", p + "2.txt", "-a", "m.fa", "-sM", "12", "-H", p + "2.his", "-w", p + ".mod",
                    "-p", "2", "30", "-wt", p + "3.txt"]
        a, b = script("a"), script("b")
        run(stock, a, d)
        run(exe, b, d)
        for f in ("his", "txt"):
            for n in ("", "2", "3") if f == "txt" else ("", "2"):
                fa, fb = os.path.join(d, "a%s.%s" % (n, f)), os.path.join(d, "b%s.%s" % (n, f))
                assert open(fa, "rb").read() == open(fb, "rb").read(), (k, w, fa)
        assert stable(open(os.path.join(d, "a.out")).read()) == stable(open(os.path.join(d, "b.out")).read()), (k, w)
        # each tool reads the other's .mod (gzip'd by fzopen, modutils.c:165) and prints the same set
        run(stock, ["-o", "c.out", "-r", "b.mod", "-H", "c.his", "-wt", "c.txt"], d)
        run(exe, ["-o", "d.out", "-r", "a.mod", "-H", "d.his", "-wt", "d.txt"], d)
        assert rd("c.his") == rd("d.his") == rd("a2.his")
        assert rd("c.txt") == rd("d.txt")
        assert stable(open(os.path.join(d, "c.out")).read()) == stable(open(os.path.join(d, "d.out")).read())
        # refpaint (-P, modutils.c:260-273) prints position and depth of every hit along a reference, to stdout
        pa = run(stock, ["-r", "a.mod", "-P", "g.fa"], d).stdout
        pb = run(exe, ["-r", "a.mod", "-P", "g.fa"], d).stdout
        assert stable(pa) == stable(pb) and any(l.startswith("  ") for l in pa.splitlines()), (k, w)
        if ref_source:
            # the whole .mod, index[] included (built on the device with the reference's probe rule); value[0] is
            # uninitialised heap in the reference (modset.c:27): masked
            import gzip
            ma, mb = bytearray(gzip.decompress(rd("a.mod"))), bytearray(gzip.decompress(rd("b.mod")))
            v0 = 8 + 4 + 4 + 8 + 80 + 4 * (1 << 22)
            ma[v0:v0 + 8] = bytes(8); mb[v0:v0 + 8] = bytes(8)
            assert ma == mb, (k, w, ".mod bytes")
            # commands only the reference's interpreter has: text round trip (modsetIndexFind (.., true) per line), merge
            # of a second set built from m.fa, per-mod depths in other sets - all on libmodshim's modset.h symbols
            # (the merge goes into a set made by -c: merging into an -rt / -r set reads uninitialised memory in the
            # reference - resize() does not zero the grown depth[] / info[], utils.h:54, modset.c:116-121)
            def script2(t, p):
                run(t, ["-c", "22", str(k), str(w), "17", "-a", "m.fa", "-w", p + "m.mod"], d)
                subprocess.run(["gunzip", "-c", p + "m.mod"], cwd=d, stdout=open(os.path.join(d, p + "m.raw"), "wb"), check=True)
                run(t, ["-o", p + "f.out", "-rt", "a3.txt", "-wt", p + "e.txt", "-s", "2", "5", "9", "-w", p + "f.mod"], d)
                return run(t, ["-o", p + "e.out", "-c", "22", str(k), str(w), "17", "-a", "r.fa", "-s", "3", "9", "14", "-m", p + "m.raw",
                               "-wt", p + "e2.txt", "-H", p + "e.his", "-d", p + "e.dep", p + "m.raw"], d)
            script2(stock, "a"); script2(exe, "b")
            for f in ("e.txt", "e2.txt", "e.his", "e.dep"):
                assert rd("a" + f) == rd("b" + f), (k, w, f)
            for o in ("e.out", "f.out"):
                assert stable(open(os.path.join(d, "a" + o)).read()) == stable(open(os.path.join(d, "b" + o)).read()), (k, w, o)


@pytest.mark.parametrize("tool", ["modmap_gpu", "modmap_dropin", "modmap_shim"])
def test_c_modmap_matches_stock_modmap(tmp_path, tool):
    stock = H.ref_cli("modmap")
    exe = TOOL(tool)
    if not stock or not os.path.exists(exe):
        pytest.skip("stock modmap / %s not built (no /root/reference in the build container)" % tool)
    H.modmap_case(str(tmp_path))
    H.modmap_driver_vs_stock(stock, exe, str(tmp_path), check_mod=True)
    if tool != "modmap_gpu":                         # -r is the reference's own referenceRead on libmodshim's modsetRead
        d = str(tmp_path)
        a = run(stock, ["-o", "ra.out", "-r", "a", "-q", "r.fa"], d)
        b = run(exe, ["-o", "rb.out", "-r", "a", "-q", "r.fa"], d)
        assert stable(open(os.path.join(d, "ra.out")).read()) == stable(open(os.path.join(d, "rb.out")).read())


def test_c_sharded_host(tmp_path):
    """modimizer_b200/sharded_host: the multi-GPU build as a C host (modgpuSharded* + NCCL, one process per GPU, no
    Python in the data path).  On a box with >= 2 GPUs: sharded over 2 GPUs == one modset on one GPU (entry totals and
    the whole depth histogram on a 48-Mbase sample per rank), a skewed group is skipped on every rank (transactional)
    and fits after modgpuShardedSetRobust.  On a 1-GPU box the same program runs with world 1."""
    import json
    import torch
    exe = TOOL("sharded_host")
    if not os.path.exists(exe):
        pytest.skip("sharded_host not built")
    n = 2 if torch.cuda.device_count() >= 2 else 1
    for (k, d) in ((31, 64), (19, 31)):
        args = ["--gpus", str(n), "--gbases", "0.4", "--k", str(k), "--d", str(d), "--bits", "26", "--steps", "2", "--warmup", "1", "--check", "48"]
        if n > 1:
            args.append("--skew")
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout[-800:], r.stderr[-800:])
        out = json.loads(r.stdout.strip().splitlines()[-1])
        assert out["n_gpus"] == n and out["parity"]["ok"] is True, out
        assert abs(out["selected"] - n * 0.4e9 / d) < 0.02 * n * 0.4e9 / d, out
        if n > 1:
            assert out["skew_transactional"] is True, out


class Seqhash(C.Structure):                      # reference seqhash.h:15-23
    _fields_ = [("seed", C.c_int), ("k", C.c_int), ("w", C.c_int), ("mask", C.c_uint64), ("shift1", C.c_int), ("shift2", C.c_int),
                ("factor1", C.c_uint64), ("factor2", C.c_uint64), ("patternRC", C.c_uint64 * 4)]


def test_shim_exports_the_reference_seqhash_api(orc):
    if not os.path.exists(SHIM):
        pytest.skip("libmodshim.so not built (no /root/reference in the build container)")
    lib = C.CDLL(SHIM)
    lib.seqhashCreate.restype = C.POINTER(Seqhash)
    lib.seqhashCreate.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.modRCiterator.restype = C.c_void_p
    lib.modRCiterator.argtypes = [C.POINTER(Seqhash), C.c_char_p, C.c_int]
    lib.modRCnext.restype = C.c_bool
    lib.modRCnext.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_bool)]
    lib.seqString.restype = C.c_char_p
    lib.seqString.argtypes = [C.c_uint64, C.c_int]
    libc = C.CDLL(None)
    assert C.sizeof(Seqhash) == 80
    rng = np.random.default_rng(9)
    for (k, w, seed) in ((19, 31, 17), (31, 64, 17), (5, 1, 3), (12, 8, 0), (31, 7, 99)):
        sh = lib.seqhashCreate(k, w, seed)
        o = orc.hasher(k, w, seed)
        assert (sh.contents.mask, sh.contents.shift1, sh.contents.factor1, sh.contents.factor2) == (o["mask"], o["shift"], o["factor1"], o["factor2"])
        assert [sh.contents.patternRC[b] for b in range(4)] == [(3 - b) << (2 * (k - 1)) for b in range(4)]
        for n in (0, k - 1, k, k + 1, 777, 40000):
            codes = rng.integers(0, 4, n).astype(np.uint8)
            buf = codes.tobytes()
            it = lib.modRCiterator(sh, buf, n)
            got = []
            km, pos, isf = C.c_uint64(), C.c_int(), C.c_bool()
            while lib.modRCnext(it, C.byref(km), C.byref(pos), C.byref(isf)):
                got.append((km.value, pos.value, int(isf.value)))
            assert not lib.modRCnext(it, None, None, None)                       # stays false; NULL out-pointers are fine
            ek, ep, ef = orc.mod_scan(k, w, seed, codes)
            assert got == list(zip([int(x) for x in ek], [int(x) for x in ep], [int(x) for x in ef])), (k, w, seed, n)
            # what the header-inline seqhashRCiteratorDestroy does (seqhash.h:54-55): hashBuf and fBuf sit at offsets 40 / 48
            hb, fb = C.c_void_p.from_address(it + 40).value, C.c_void_p.from_address(it + 48).value
            libc.free(C.c_void_p(hb)); libc.free(C.c_void_p(fb)); libc.free(C.c_void_p(it))
        assert lib.seqString(0x31b, 5) == b"tacgt"
        libc.free(C.cast(sh, C.c_void_p))
