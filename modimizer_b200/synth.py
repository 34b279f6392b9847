"""Synthetic genomes / readsets (include/modgpu_synth.h) generated on the device.

Used by bench.py and the GPU parity tests; the oracle side generates the same
bytes from the same header compiled for the host (tests/host/synth_host.c)."""
import ctypes as C

from . import _lib
from ._lib import ReadSpec, check


def genome_device(seed, start, n, dup_mode, d_ptr, stream=0):
    check(_lib.require_device().modgpuSynthGenome(seed, start, n, dup_mode, C.c_void_p(d_ptr), C.c_void_p(stream)), "synthGenome")


def reads_device(spec, first_read, n_reads, ont, d_ptr, stream=0):
    check(_lib.require_device().modgpuSynthReads(C.byref(spec), first_read, n_reads, 1 if ont else 0,
                                                  C.c_void_p(d_ptr), C.c_void_p(stream)), "synthReads")


def read_spec(genome_seed, genome_len, read_seed, read_len, sub_ppm=0, ins_ppm=0, del_ppm=0,
              frag_len=0, pair_mode=0, dup_mode=0):
    return ReadSpec(genome_seed, genome_len, read_seed, read_len, sub_ppm, ins_ppm, del_ppm, frag_len, pair_mode, dup_mode, 0)
