"""ctypes binding of libmodgpu.so (include/modgpu.h).

The library is the product: if it is missing or no CUDA device is usable the
package raises - there is no Python or CPU fallback for any operation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MODGPU_LIB: another build of the same library (A/B runs of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("MODGPU_LIB") or os.path.join(_HERE, "libmodgpu.so")

u64 = C.c_uint64
u32 = C.c_uint32
vp = C.c_void_p


class ModgpuError(RuntimeError):
    pass


class Hasher(C.Structure):
    """ModgpuHasher (include/modgpu.h) == the reference Seqhash essentials (seqhash.h:15-23)"""
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("seed", C.c_int32), ("shift1", C.c_int32),
                ("mask", u64), ("factor1", u64), ("factor2", u64)]


class ReadSpec(C.Structure):
    """MgReadSpec (include/modgpu_synth.h)"""
    _fields_ = [("genomeSeed", u64), ("genomeLen", u64), ("readSeed", u64),
                ("readLen", u32), ("subPPM", u32), ("insPPM", u32), ("delPPM", u32),
                ("fragLen", u32), ("pairMode", C.c_int32), ("dupMode", C.c_int32), ("pad_", C.c_int32)]


_SIGS = {
    "modgpuLastError": (C.c_char_p, []),
    "modgpuDeviceCount": (C.c_int, []),
    "modgpuSetDevice": (C.c_int, [C.c_int]),
    "modgpuVersion": (C.c_char_p, []),
    "modgpuHasherInit": (C.c_int, [C.POINTER(Hasher), C.c_int, C.c_int, C.c_int]),
    "modgpuHasherFromSeqhash": (C.c_int, [C.POINTER(Hasher), vp]),
    "modgpuHash": (u64, [C.POINTER(Hasher), u64]),
    "modgpuPackedWords": (u64, [u64]),
    "modgpuEndsWords": (u64, [u64]),
    "modgpuPack2bit": (C.c_int, [vp, u64, C.c_int, vp, vp]),
    "modgpuMarkEnds": (C.c_int, [vp, u64, u64, vp, vp]),
    "modgpuHashSelectWorkspace": (u64, [u64]),
    "modgpuHashSelect": (C.c_int, [C.POINTER(Hasher), vp, vp, u64, vp, vp, u64, vp, vp, C.c_int, vp]),
    "modgpuLocate": (C.c_int, [vp, u64, vp, u64, vp, vp, vp]),
    "modgpuTableCreate": (vp, [C.c_int, vp]),
    "modgpuTableDestroy": (None, [vp]),
    "modgpuTableClear": (C.c_int, [vp, vp]),
    "modgpuTableSlots": (u64, [vp]),
    "modgpuTableDevicePtr": (vp, [vp]),
    "modgpuTableInsert": (C.c_int, [vp, vp, u64, vp, C.c_int, vp]),
    "modgpuTableEntries": (u64, [vp, vp]),
    "modgpuTableNumber": (C.c_int, [vp, vp, u64, vp, vp]),
    "modgpuTableLookup": (C.c_int, [vp, vp, u64, vp, vp]),
    "modgpuTableHistogram": (C.c_int, [vp, vp, vp]),
    "modgpuTableClassify": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "modgpuTableExport": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "modgpuTableImport": (C.c_int, [vp, vp, vp, vp, u64, vp]),
    "modgpuModsetCreate": (vp, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "modgpuModsetDestroy": (None, [vp]),
    "modgpuModsetHasher": (C.POINTER(Hasher), [vp]),
    "modgpuModsetTable": (vp, [vp]),
    "modgpuModsetBits": (C.c_int, [vp]),
    "modgpuModsetDevice": (C.c_int, [vp]),
    "modgpuModsetSetStream": (C.c_int, [vp, vp]),
    "modgpuModsetSetFlags": (C.c_int, [vp, C.c_int]),
    "modgpuModsetSetExactOrder": (C.c_int, [vp, C.c_int]),
    "modgpuModsetAdd": (u64, [vp, vp, vp, u64, C.c_int]),
    "modgpuModsetAddDevice": (u64, [vp, vp, vp, u64, u64, C.c_int]),
    "modgpuModsetAddPacked": (u64, [vp, vp, vp, vp, u64]),
    "modgpuModsetMax": (u32, [vp]),
    "modgpuModsetExport": (C.c_int, [vp, vp, vp, vp]),
    "modgpuModsetHistogram": (C.c_int, [vp, vp]),
    "modgpuModsetSetCopy": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    "modgpuModsetSetCopyM": (C.c_int, [vp, C.c_int, vp]),
    "modgpuModsetFind": (C.c_int, [vp, vp, u64, vp, vp]),
    "modgpuModsetSummary": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "modgpuModsetImport": (C.c_int, [vp, vp, vp, vp, u64]),
    "modgpuModsetSelectDevice": (C.c_int, [vp, vp, vp, u64, u64, C.c_int, C.POINTER(vp), C.POINTER(u64)]),
    "modgpuModsetSelectHost": (C.c_int, [vp, vp, vp, u64, C.c_int, C.POINTER(vp), C.POINTER(u64)]),
    "modgpuModsetSelectOwnersDevice": (C.c_int, [vp, vp, vp, u64, u64, C.c_int, u32, vp, u64, vp]),
    "modgpuModsetSelectOwnersHost": (C.c_int, [vp, vp, vp, u64, C.c_int, u32, vp, u64, vp]),
    "modgpuModsetRegions": (u32, [vp]),
    "modgpuModsetSelectBucketsDevice": (C.c_int, [vp, vp, vp, u64, u64, C.c_int, u32, vp, u32, vp, vp, u64, vp, vp]),
    "modgpuModsetSelectBucketsHost": (C.c_int, [vp, vp, vp, u64, C.c_int, u32, vp, u32, vp, vp, u64, vp, vp]),
    "modgpuModsetBuildFromBuckets": (C.c_int, [vp, vp, vp, u32, u32, vp, u64, vp]),
    "modgpuModsetBuildFromPeers": (C.c_int, [vp, vp, vp, u32, u32, vp, u64, vp]),
    "modgpuPeerAlloc": (vp, [C.c_size_t]),
    "modgpuPeerFree": (None, [vp]),
    "modgpuPeerExport": (C.c_int, [vp, vp]),
    "modgpuPeerOpen": (vp, [C.c_char_p]),
    "modgpuPeerClose": (C.c_int, [vp]),
    "modgpuCommFromNccl": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "modgpuCommNcclRelease": (None, [vp]),
    "modgpuShardedCreate": (vp, [C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "modgpuShardedDestroy": (None, [vp]),
    "modgpuShardedLocal": (vp, [vp]),
    "modgpuShardedSetStream": (C.c_int, [vp, vp]),
    "modgpuShardedReserve": (C.c_int, [vp, u64]),
    "modgpuShardedSetAccumulate": (C.c_int, [vp, C.c_int]),
    "modgpuShardedSetRobust": (C.c_int, [vp, C.c_int]),
    "modgpuShardedAdd": (C.c_int, [vp, vp, vp, u64, C.c_int]),
    "modgpuShardedAddDevice": (C.c_int, [vp, vp, vp, u64, u64, C.c_int]),
    "modgpuShardedFlush": (C.c_int, [vp]),
    "modgpuShardedSynchronize": (C.c_int, [vp, C.POINTER(u64)]),
    "modgpuShardedClear": (C.c_int, [vp]),
    "modgpuModsetInsertSegments": (C.c_int, [vp, vp, u32, u64, vp, u64]),
    "modgpuModsetInsertDevice": (C.c_int, [vp, vp, u64]),
    "modgpuModsetClear": (C.c_int, [vp]),
    "modgpuModsetSetAccumulate": (C.c_int, [vp, C.c_int]),
    "modgpuModsetFlush": (C.c_int, [vp]),
    "modgpuOwnerOf": (u32, [u64, u32]),
    "modgpuOwnerCount": (C.c_int, [vp, u64, u32, vp, vp]),
    "modgpuOwnerScatter": (C.c_int, [vp, u64, u32, vp, vp, vp]),
    "modgpuModsetPrune": (C.c_int, [vp, C.c_int, C.c_int]),
    "modgpuModsetMerge": (C.c_int, [vp, vp]),
    "modgpuModsetCreateWithHasher": (vp, [C.c_int, C.POINTER(Hasher)]),
    "modgpuModsetWriteMod": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "modgpuModsetReadMod": (vp, [C.c_char_p]),
    "modgpuModsetReadset": (u64, [vp, vp, vp, u64, C.c_int, C.c_int, vp, vp, vp, vp, u64]),
    "modgpuModsetIndexFindBatch": (C.c_int, [vp, vp, u64, C.c_int, vp]),
    "modgpuModsetSetDepthInfo": (C.c_int, [vp, vp, vp, u64]),
    "modgpuModsetReferenceIndex": (C.c_int, [vp, vp]),
    "modgpuModsetProfile": (C.c_int, [vp, C.c_int]),
    "modgpuModsetTimes": (C.c_int, [vp, vp, vp]),
    "modgpuReferenceBuild": (vp, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, u64, C.c_int, vp]),
    "modgpuReferenceDestroy": (None, [vp]),
    "modgpuReferenceModset": (vp, [vp]),
    "modgpuReferenceMax": (u32, [vp]),
    "modgpuReferenceExport": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "modgpuReferenceQuery": (u64, [vp, vp, vp, u64, C.c_int, vp, vp, vp, vp, vp, vp, u64]),
    "modgpuScannerCreate": (vp, [C.POINTER(Hasher)]),
    "modgpuScannerDestroy": (None, [vp]),
    "modgpuScannerScan": (u64, [vp, vp, vp, u64, C.c_int, vp, vp, vp, u64]),
    "modgpuHostAlloc": (vp, [C.c_size_t]),
    "modgpuHostFree": (None, [vp]),
    "modgpuSynthGenome": (C.c_int, [u64, u64, u64, C.c_int, vp, vp]),
    "modgpuSynthReads": (C.c_int, [vp, u64, u64, C.c_int, vp, vp]),
}

#: every symbol include/modgpu.h declares; tests check the .so exports all of them
SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """Load libmodgpu.so; raises ModgpuError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ModgpuError(
            "modimizer_b200/libmodgpu.so is missing: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C modimizer_b200/csrc`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().modgpuLastError().decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise ModgpuError("%s failed (%d): %s" % (what or "libmodgpu call", rc, last_error()))


def require_device():
    """Fail loudly when no CUDA device can be used (no CPU fallback)."""
    lib = load()
    if lib.modgpuDeviceCount() < 1:
        raise ModgpuError("no usable CUDA device: %s" % (last_error() or "cudaGetDeviceCount returned 0"))
    return lib
