"""ctypes view of tests/host/libhostemul.so (TEST INFRASTRUCTURE): the kernels'
__host__ __device__ arithmetic and the synthetic generators compiled for the CPU."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")

_lib = None


class ReadSpec(C.Structure):
    _fields_ = [("genomeSeed", C.c_uint64), ("genomeLen", C.c_uint64), ("readSeed", C.c_uint64),
                ("readLen", C.c_uint32), ("subPPM", C.c_uint32), ("insPPM", C.c_uint32), ("delPPM", C.c_uint32),
                ("fragLen", C.c_uint32), ("pairMode", C.c_int32), ("dupMode", C.c_int32), ("pad_", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(HERE, "host"), "-s"], check=True)
        l = C.CDLL(os.path.join(HERE, "host", "libhostemul.so"))
        l.hm_select.restype = C.c_int64
        l.hm_select.argtypes = [C.c_int, C.c_int, C.c_uint64, u8p, C.c_uint64, C.c_int, u64p, C.c_uint64, C.c_int,
                                u64p, u32p, u8p, C.c_int64]
        l.hm_pack.restype = None
        l.hm_pack.argtypes = [u8p, C.c_uint64, C.c_int, u64p, C.c_uint64]
        l.hm_khasher.restype = None
        l.hm_khasher.argtypes = [C.c_int, C.c_int, C.c_uint64, u64p]
        l.hm_divisible.restype = C.c_int
        l.hm_divisible.argtypes = [C.c_int, C.c_uint64]
        l.hm_lut_check.restype = C.c_int64
        l.hm_lut_check.argtypes = [C.c_int, C.c_int, C.c_uint64, u8p, C.c_uint64, C.c_int, u64p]
        l.hm_slot_hash.restype = C.c_uint64
        l.hm_slot_hash.argtypes = [C.c_uint64, C.c_uint32]
        l.hm_owner.restype = C.c_uint32
        l.hm_owner.argtypes = [C.c_uint64, C.c_uint32]
        l.hs_genome.restype = None
        l.hs_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p]
        l.hs_reads.restype = None
        l.hs_reads.argtypes = [C.POINTER(ReadSpec), C.c_uint64, C.c_uint64, C.c_int, u8p]
        _lib = l
    return _lib


def select(k, d, factor1, data, offs, is_ascii=0, prefilter=-1):
    data = np.ascontiguousarray(data, np.uint8)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = int(offs[-1])
    cap = max(n, 1)
    km = np.zeros(cap, np.uint64)
    gp = np.zeros(cap, np.uint32)
    isf = np.zeros(cap, np.uint8)
    cnt = lib().hm_select(k, d, factor1, data if data.size else np.zeros(1, np.uint8), n, is_ascii, offs, len(offs) - 1,
                          prefilter, km, gp, isf, cap)
    return km[:cnt], gp[:cnt], isf[:cnt]


def lut_check(k, d, factor1, data, is_ascii=0):
    """(mismatching candidate masks, candidates) of the table-driven prefilter vs the arithmetic one; (-1, 0) if n/a"""
    data = np.ascontiguousarray(data, np.uint8)
    ncand = np.zeros(1, np.uint64)
    bad = lib().hm_lut_check(k, d, factor1, data, len(data), is_ascii, ncand)
    return int(bad), int(ncand[0])


def genome(seed, start, n, dup_mode=0):
    out = np.zeros(n, np.uint8)
    lib().hs_genome(seed, start, n, dup_mode, out)
    return out


def read_spec(genome_seed, genome_len, read_seed, read_len, sub_ppm=0, ins_ppm=0, del_ppm=0,
              frag_len=0, pair_mode=0, dup_mode=0):
    return ReadSpec(genome_seed, genome_len, read_seed, read_len, sub_ppm, ins_ppm, del_ppm, frag_len, pair_mode, dup_mode, 0)


def reads(spec, first, n, ont=False):
    out = np.zeros(n * spec.readLen, np.uint8)
    lib().hs_reads(C.byref(spec), first, n, 1 if ont else 0, out)
    return out
