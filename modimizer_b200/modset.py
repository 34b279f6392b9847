"""Seqhash / Modset / Reference: the reference's operator interface for the hot path.

Names and argument meaning follow the reference (`k`, `w` = modimizer modulus,
`seed`, table `bits`; indices are 1-based, 0 = absent; depth saturates at
65535; copy classes 0,1,2,3=M live in the low two info bits).  Errors that make
the reference die() raise ModgpuError here.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import ModgpuError, Hasher, check

U64MAX = 0xFFFFFFFFFFFFFFFF


def kmer_string(v, k):
    """seqString (reference seqhash.c:198-206): lower case, first base most significant"""
    v = int(v)
    return "".join("acgt"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def _as_bytes(seq):
    """bytes-like -> (contiguous uint8 array, isAscii).  Arrays of codes 0..3 or ASCII text."""
    if isinstance(seq, str):
        a = np.frombuffer(seq.encode(), np.uint8)
        return a, 1
    a = np.ascontiguousarray(seq, dtype=np.uint8)
    is_ascii = 1 if (a.size and a.max() > 3) else 0
    return a, is_ascii


def concat(seqs):
    """list of sequences -> (bytes, offsets, isAscii) in the batch layout of include/modgpu.h"""
    arrs, flags = [], []
    for s in seqs:
        a, f = _as_bytes(s)
        arrs.append(a)
        flags.append(f)
    offs = np.zeros(len(arrs) + 1, np.uint64)
    if arrs:
        offs[1:] = np.cumsum([len(a) for a in arrs])
    data = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    is_ascii = 1 if any(flags) else 0
    if is_ascii and not all(f or len(a) == 0 for a, f in zip(arrs, flags)):
        # mixing code arrays with ASCII text: bring the code arrays to ASCII
        data = np.concatenate([a if f else np.frombuffer(b"ACGT", np.uint8)[a] for a, f in zip(arrs, flags)])
    return np.ascontiguousarray(data, np.uint8), offs, is_ascii


def seqio_pack(codes, offsets):
    """sqioSeqPack (reference seqio.c:557-570) over a batch of code sequences: four bases per byte, first base in the
    top two bits, every sequence on its own bytes, a short last byte right-aligned.  Returns (packed, byte_offsets)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    nbytes = (lens + 3) // 4
    boffs = np.zeros(len(lens) + 1, np.uint64)
    boffs[1:] = np.cumsum(nbytes).astype(np.uint64)
    out = np.zeros(int(boffs[-1]), np.uint8)
    for r in range(len(lens)):
        s = codes[int(offsets[r]):int(offsets[r + 1])] & 3
        L = len(s)
        if not L:
            continue
        full = (L - 1) // 4 * 4 if L % 4 else L          # `while (len > 4)`: the last byte is the remainder loop's, even when it has four bases
        q = s[:full].reshape(-1, 4)
        b = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        o = int(boffs[r])
        out[o:o + len(b)] = b
        if L > full:
            v = 0
            for c in s[full:]:
                v = (v << 2) | int(c)
            out[o + len(b)] = v
    return out, boffs


class Seqhash:
    """seqhashCreate (reference seqhash.c:20-37): k, w (modulus), seed -> multiplier from libc random()"""

    def __init__(self, k, w, seed):
        lib = _lib.load()
        self._h = Hasher()
        check(lib.modgpuHasherInit(C.byref(self._h), k, w, seed), "seqhashCreate")

    k = property(lambda s: s._h.k)
    w = property(lambda s: s._h.w)
    seed = property(lambda s: s._h.seed)
    mask = property(lambda s: s._h.mask)
    shift1 = property(lambda s: s._h.shift1)
    factor1 = property(lambda s: s._h.factor1)

    def hash(self, kmer):
        """seqhash() (reference seqhash.h:58)"""
        return int(_lib.load().modgpuHash(C.byref(self._h), int(kmer)))

    def report(self):
        """seqhashReport (reference seqhash.c:55-56)"""
        return "SH k %d  w/m %d  s %d\n" % (self.k, self.w, self.seed)


class Modset:
    """modsetCreate (reference modset.c:15-31) on the device + the batched addSequence loop."""

    def __init__(self, bits, k=19, w=31, seed=17, exact_order=False, _handle=None, _owner=None):
        self._lib = _lib.require_device()
        self._owner = _owner
        if _handle is None:
            _handle = self._lib.modgpuModsetCreate(bits, k, w, seed)
            if not _handle:
                raise ModgpuError("modsetCreate: " + _lib.last_error())
        self._p = _handle
        self.bits = bits
        h = self._lib.modgpuModsetHasher(self._p).contents
        self.k, self.w, self.seed = h.k, h.w, h.seed
        self.factor1 = h.factor1
        self.total_hashes = 0
        if exact_order:
            self.set_exact_order(True)

    def close(self):
        if self._p and self._owner is None:
            self._lib.modgpuModsetDestroy(self._p)
        self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration -------------------------------------------------
    def set_stream(self, cuda_stream):
        check(self._lib.modgpuModsetSetStream(self._p, C.c_void_p(cuda_stream)), "set_stream")

    def set_flags(self, flags):
        check(self._lib.modgpuModsetSetFlags(self._p, flags), "set_flags")
        self._flags = int(flags)

    def set_exact_order(self, on=True):
        """number entries by first occurrence like the reference's index = ++max (modset.c:57)"""
        check(self._lib.modgpuModsetSetExactOrder(self._p, 1 if on else 0), "set_exact_order")

    # ---- the hot loop ---------------------------------------------------
    def add(self, data, offsets=None, is_ascii=None):
        """addSequence over a batch (reference modutils.c:19-31); returns the number of hashes.

        `data` is either a list of sequences, or one concatenated uint8 array with
        `offsets` (nSeq+1, uint64).  Host memory; copies happen inside the call."""
        if offsets is None:
            data, offsets, auto = concat(data)
        else:
            data = np.ascontiguousarray(data, np.uint8)
            # (a scan of the whole batch: only when the caller did not say what the bytes are)
            auto = (1 if (data.size and data.max() > 3) else 0) if is_ascii is None else 0
        if is_ascii is None:
            is_ascii = auto
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = self._lib.modgpuModsetAdd(self._p, data.ctypes.data if data.size else None, offsets.ctypes.data,
                                      len(offsets) - 1, int(is_ascii))
        if n == U64MAX:
            raise ModgpuError("modsetAdd: " + _lib.last_error())
        self.total_hashes += n
        return int(n)

    def add_pointers(self, host_ptr, offsets_ptr, nseq, is_ascii=0):
        """the same through raw host pointers (pinned buffers filled by a parser)"""
        n = self._lib.modgpuModsetAdd(self._p, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq, is_ascii)
        if n == U64MAX:
            raise ModgpuError("modsetAdd: " + _lib.last_error())
        self.total_hashes += n
        return int(n)

    def add_packed(self, packed, byte_offsets, offsets):
        """the batch in the reference's own 2-bit packing (sqioSeqPack, seqio.c:557-570; see seqio_pack): a quarter
        of the bytes cross PCIe, the device expands them"""
        packed = np.ascontiguousarray(packed, np.uint8)
        byte_offsets = np.ascontiguousarray(byte_offsets, np.uint64)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = self._lib.modgpuModsetAddPacked(self._p, packed.ctypes.data if packed.size else None, byte_offsets.ctypes.data,
                                            offsets.ctypes.data, len(offsets) - 1)
        if n == U64MAX:
            raise ModgpuError("modsetAddPacked: " + _lib.last_error())
        self.total_hashes += n
        return int(n)

    def add_packed_pointers(self, packed_ptr, byte_offsets_ptr, offsets_ptr, nseq):
        n = self._lib.modgpuModsetAddPacked(self._p, C.c_void_p(packed_ptr), C.c_void_p(byte_offsets_ptr), C.c_void_p(offsets_ptr), nseq)
        if n == U64MAX:
            raise ModgpuError("modsetAddPacked: " + _lib.last_error())
        self.total_hashes += n
        return int(n)

    def add_device(self, d_bases, d_offsets, nseq, nbases, is_ascii=0):
        """the batch is already resident in device memory (raw device pointers)"""
        n = self._lib.modgpuModsetAddDevice(self._p, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases, is_ascii)
        if n == U64MAX:
            raise ModgpuError("modsetAddDevice: " + _lib.last_error())
        self.total_hashes += n
        return int(n)

    # ---- whole-set operations ------------------------------------------
    @property
    def max(self):
        """ms->max: number of distinct modimizers"""
        m = self._lib.modgpuModsetMax(self._p)
        if m == 0xFFFFFFFF:
            raise ModgpuError("modsetMax: " + _lib.last_error())
        return int(m)

    def export(self):
        """(value, depth, info) for indices 1..max - the host arrays of the reference Modset"""
        n = self.max
        v = np.zeros(n, np.uint64)
        d = np.zeros(n, np.uint16)
        i = np.zeros(n, np.uint8)
        if n:
            check(self._lib.modgpuModsetExport(self._p, v.ctypes.data, d.ctypes.data, i.ctypes.data), "modsetExport")
        return v, d, i

    def sorted_dump(self):
        """(kmer, depth, info&3) sorted by k-mer: the parity key of the -wt dump"""
        v, d, i = self.export()
        o = np.argsort(v, kind="stable")
        return v[o], d[o], i[o] & 3

    def histogram(self):
        """depthHistogram bins (reference modutils.c:53-63)"""
        b = np.zeros(65536, np.uint32)
        check(self._lib.modgpuModsetHistogram(self._p, b.ctypes.data), "depthHistogram")
        return b

    def histogram_text(self):
        b = self.histogram()
        nz = np.nonzero(b)[0]
        return "".join("DP\t%u\t%u\n" % (i, b[i]) for i in nz)

    def set_copy(self, copy1min, copy2min, copyMmin):
        """-s (reference modutils.c:205-214); returns the copy0..copyM tallies"""
        c = np.zeros(4, np.uint32)
        check(self._lib.modgpuModsetSetCopy(self._p, copy1min, copy2min, copyMmin, c.ctypes.data), "setcopy")
        return c

    def set_copy_m(self, copyMmin):
        """-sM (reference modutils.c:215-219)"""
        c = np.zeros(4, np.uint32)
        check(self._lib.modgpuModsetSetCopyM(self._p, copyMmin, c.ctypes.data), "setcopyM")
        return c

    def find(self, kmers):
        """batched modsetIndexFind(ms, kmer, false) (reference modset.c:45-62): (index, copy)"""
        kmers = np.ascontiguousarray(kmers, np.uint64)
        idx = np.zeros(len(kmers), np.uint32)
        cp = np.zeros(len(kmers), np.uint8)
        if len(kmers):
            check(self._lib.modgpuModsetFind(self._p, kmers.ctypes.data, len(kmers), idx.ctypes.data, cp.ctypes.data), "modsetFind")
        return idx, cp

    def summary(self):
        """modsetSummary text (reference modset.c:130-153)"""
        buf = C.create_string_buffer(4096)
        n = self._lib.modgpuModsetSummary(self._p, buf, 4096)
        if n < 0:
            raise ModgpuError("modsetSummary: " + _lib.last_error())
        return buf.raw[:n].decode()

    def import_entries(self, value, depth=None, info=None):
        """upload a host modset (entries 1..max of a reference Modset, e.g. after modsetRead)"""
        value = np.ascontiguousarray(value, np.uint64)
        depth = None if depth is None else np.ascontiguousarray(depth, np.uint16)
        info = None if info is None else np.ascontiguousarray(info, np.uint8)
        check(self._lib.modgpuModsetImport(self._p, value.ctypes.data,
                                           depth.ctypes.data if depth is not None else None,
                                           info.ctypes.data if info is not None else None, len(value)), "modsetImport")

    def prune(self, min_depth, max_depth=0):
        """modsetDepthPrune (reference modset.c:64-77): keep min <= depth < max (0 = no upper bound)"""
        check(self._lib.modgpuModsetPrune(self._p, int(min_depth), int(max_depth)), "modsetDepthPrune")

    def merge(self, other):
        """modsetMerge (reference modset.c:106-128); False when the hashers are incompatible"""
        rc = self._lib.modgpuModsetMerge(self._p, other._p)
        if rc < 0:
            raise ModgpuError("modsetMerge: " + _lib.last_error())
        return bool(rc)

    def write_mod(self, path, gzip=False):
        """modsetWrite (reference modset.c:79-88): "MSHSTv2" file an unmodified modutils -r can load"""
        check(self._lib.modgpuModsetWriteMod(self._p, str(path).encode(), 1 if gzip else 0), "modsetWrite")

    @classmethod
    def read_mod(cls, path):
        """modsetRead (reference modset.c:90-104), plain or gzip'd"""
        lib = _lib.require_device()
        p = lib.modgpuModsetReadMod(str(path).encode())
        if not p:
            raise ModgpuError("modsetRead: " + _lib.last_error())
        return cls(int(lib.modgpuModsetBits(p)), _handle=p)

    def readset(self, data, offsets=None, is_ascii=None, reset_depth=True, cap=None):
        """hot loop of modasm's readsetFileRead (reference modasm.c:151-191)"""
        if offsets is None:
            data, offsets, auto = concat(data)
        else:
            data = np.ascontiguousarray(data, np.uint8)
            # (a scan of the whole batch: only when the caller did not say what the bytes are)
            auto = (1 if (data.size and data.max() > 3) else 0) if is_ascii is None else 0
        if is_ascii is None:
            is_ascii = auto
        offsets = np.ascontiguousarray(offsets, np.uint64)
        nseq = len(offsets) - 1
        if cap is None:
            cap = max(1, len(data))
        ho = np.zeros(nseq + 1, np.uint64)
        hit = np.zeros(cap, np.uint32)
        dx = np.zeros(cap, np.uint16)
        miss = np.zeros(max(nseq, 1), np.int32)
        n = self._lib.modgpuModsetReadset(self._p, data.ctypes.data if data.size else None, offsets.ctypes.data, nseq,
                                          int(is_ascii), 1 if reset_depth else 0, ho.ctypes.data, hit.ctypes.data,
                                          dx.ctypes.data, miss.ctypes.data, cap)
        if n == U64MAX:
            raise ModgpuError("readset: " + _lib.last_error())
        n = int(min(n, cap))
        return dict(hitOff=ho, hit=hit[:n], dx=dx[:n], nMiss=miss[:nseq])

    def write_text(self, path):
        """-wt dump (reference modutils.c:191-200)"""
        v, d, i = self.export()
        with open(path, "w") as f:
            f.write("modset bits %d size %d k %d w %d seed %d\n" % (self.bits, len(v) + 1, self.k, self.w, self.seed))
            for j in range(len(v)):
                f.write("%d\t%s\t%d\t%d\n" % (j + 1, kmer_string(v[j], self.k), d[j], i[j]))

    def set_accumulate(self, n_chunks):
        """deferred build: up to n_chunks device chunks share one region build (include/modgpu.h); a full table is
        then reported by flush() or by the first reader instead of by add*()"""
        check(self._lib.modgpuModsetSetAccumulate(self._p, int(n_chunks)), "modsetSetAccumulate")

    def flush(self):
        """apply the k-mers waiting in the buckets; raises when the table is over its capacity"""
        check(self._lib.modgpuModsetFlush(self._p), "modsetFlush")

    def clear(self):
        """empty the set (the table is rewritten lazily by the next bulk build)"""
        check(self._lib.modgpuModsetClear(self._p), "modsetClear")

    # ---- profiling -------------------------------------------------------
    def profile(self, on=True):
        check(self._lib.modgpuModsetProfile(self._p, 1 if on else 0), "profile")

    def times(self):
        ms = (C.c_double * 4)()
        ln = (C.c_uint64 * 4)()
        check(self._lib.modgpuModsetTimes(self._p, ms, ln), "times")
        names = ("pack", "select", "insert", "other")
        return {n: (ms[i], int(ln[i])) for i, n in enumerate(names)}


class Reference:
    """modmap's Reference (reference modmap.c:35-47): built by referenceFastaRead's loop, queried by queryProcess's."""

    def __init__(self, bits, k, w, seed, data, offsets=None, is_ascii=None):
        self._lib = _lib.require_device()
        if offsets is None:
            data, offsets, auto = concat(data)
        else:
            data = np.ascontiguousarray(data, np.uint8)
            # (a scan of the whole batch: only when the caller did not say what the bytes are)
            auto = (1 if (data.size and data.max() > 3) else 0) if is_ascii is None else 0
        if is_ascii is None:
            is_ascii = auto
        offsets = np.ascontiguousarray(offsets, np.uint64)
        counts = np.zeros(4, np.uint32)
        self._p = self._lib.modgpuReferenceBuild(bits, k, w, seed, data.ctypes.data if data.size else None,
                                                 offsets.ctypes.data, len(offsets) - 1, int(is_ascii), counts.ctypes.data)
        if not self._p:
            raise ModgpuError("referenceBuild: " + _lib.last_error())
        #: { nHashes, nCopy1, nCopy2, nMulti }: the two lines of modmap.c:123-130
        self.counts = counts
        self.nseq = len(offsets) - 1
        self.total_len = int(offsets[-1])
        self.ms = Modset(bits, _handle=self._lib.modgpuReferenceModset(self._p), _owner=self)

    def close(self):
        if self._p:
            self.ms._p = None
            self._lib.modgpuReferenceDestroy(self._p)
        self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max(self):
        return int(self._lib.modgpuReferenceMax(self._p))

    def build_report(self):
        """the two lines referenceFastaRead prints (reference modmap.c:123-130)"""
        c = self.counts
        return ("  %d hashes from %d reference sequences, total length %d\n" % (c[0], self.nseq, self.total_len) +
                "  %d copy 1, %d copy 2, %d multiple\n" % (c[1], c[2], c[3]))

    def export(self):
        n, m = self.max, self.ms.max + 1
        out = {k: np.zeros(max(n, 1), np.uint32) for k in ("index", "offset", "id", "rev")}
        out["depth"] = np.zeros(m, np.uint32)
        out["loc"] = np.zeros(m, np.uint32)
        check(self._lib.modgpuReferenceExport(self._p, out["index"].ctypes.data, out["offset"].ctypes.data,
                                              out["id"].ctypes.data, out["depth"].ctypes.data,
                                              out["rev"].ctypes.data, out["loc"].ctypes.data), "referenceExport")
        for k in ("index", "offset", "id", "rev"):
            out[k] = out[k][:n]
        return out

    def query(self, data, offsets=None, is_ascii=None, cap=None):
        """seed loop of queryProcess (reference modmap.c:196-231) for a batch of reads"""
        if offsets is None:
            data, offsets, auto = concat(data)
        else:
            data = np.ascontiguousarray(data, np.uint8)
            # (a scan of the whole batch: only when the caller did not say what the bytes are)
            auto = (1 if (data.size and data.max() > 3) else 0) if is_ascii is None else 0
        if is_ascii is None:
            is_ascii = auto
        offsets = np.ascontiguousarray(offsets, np.uint64)
        nseq = len(offsets) - 1
        if cap is None:
            cap = max(1, len(data))
        so = np.zeros(nseq + 1, np.uint64)
        si = np.zeros(cap, np.uint32)
        sp = np.zeros(cap, np.uint32)
        hid = np.zeros(2 * cap, np.uint32)
        hoff = np.zeros(2 * cap, np.uint32)
        ctr = np.zeros(4 * max(nseq, 1), np.int32)
        n = self._lib.modgpuReferenceQuery(self._p, data.ctypes.data if data.size else None, offsets.ctypes.data, nseq,
                                           int(is_ascii), so.ctypes.data, si.ctypes.data, sp.ctypes.data,
                                           hid.ctypes.data, hoff.ctypes.data, ctr.ctypes.data, cap)
        if n == U64MAX:
            raise ModgpuError("referenceQuery: " + _lib.last_error())
        n = int(min(n, cap))
        return dict(seedOff=so, index=si[:n], pos=sp[:n], hitId=hid[:2 * n].reshape(-1, 2),
                    hitOffset=hoff[:2 * n].reshape(-1, 2), counters=ctr[:4 * nseq].reshape(-1, 4))

    @staticmethod
    def format_query(res, names, lengths, ref_names, verbose=True):
        """the Q line and the -v seed lines of queryProcess (reference modmap.c:209-231)"""
        out = []
        so = res["seedOff"]
        for r in range(len(names)):
            a, b = int(so[r]), int(so[r + 1])
            miss, c1, c2, cm = (int(x) for x in res["counters"][r])
            nseed = b - a
            # 0 seeds: the reference divides 0 by 0.0, which glibc's printf shows as "-nan" on x86-64
            frac = ("%.2f" % ((nseed - miss) / float(nseed))) if nseed else "-nan"
            out.append("Q\t%s\t%d\t%d miss, %d copy1, %d copy2, %d multi, %s hit\n" % (names[r], lengths[r], miss, c1, c2, cm, frac))
            if verbose:
                for s in range(a, b):
                    h0, h1 = res["hitId"][s]
                    if h0 == 0xFFFFFFFF:
                        continue
                    if h1 == 0xFFFFFFFF:
                        out.append("  %6d\t%s %d\n" % (res["pos"][s], ref_names[h0], res["hitOffset"][s][0]))
                    else:
                        out.append("  %6d\t%s %d\t%s %d\n" % (res["pos"][s], ref_names[h0], res["hitOffset"][s][0],
                                                              ref_names[h1], res["hitOffset"][s][1]))
        return "".join(out)
