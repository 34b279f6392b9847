#!/usr/bin/env python
"""All five BASELINE.json configurations at FULL size on one B200 (not the bench line: bench.py measures
configs[1]; these are the other shapes, timed the same way so that DESIGN.md can quote them).

Inputs are generated on the device by include/modgpu_synth.h in chunks and are not part of the timed
region; every chunk goes through the public C ABI (modgpuModsetAddDevice / modgpuReferenceBuild /
modgpuReferenceQuery).  Prints one JSON line per configuration.

  python tools/configs_bench.py [--scale 1.0] [--only 0,2,3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import modimizer_b200 as mg
from modimizer_b200 import synth

dev = torch.device("cuda:0")


ACCUMULATE = 1


def count_reads(name, spec, n_reads, k, d, bits, chunk_reads, ont=False):
    """modset build + count of a synthetic readset, chunk by chunk from device memory"""
    L = spec.readLen
    ms = mg.Modset(bits, k, d, 17)
    if n_reads > chunk_reads:
        ms.set_accumulate(ACCUMULATE)      # > 1: deferred build, several chunks share one pass over the table
    buf = torch.empty(chunk_reads * L + 64, dtype=torch.uint8, device=dev)
    offs = (torch.arange(chunk_reads + 1, dtype=torch.int64, device=dev) * L)
    passes = []
    for rep in range(2):                       # the second pass reuses the set's buffers: steady state of a long-lived set
        if rep:
            ms.clear()
        tot, gpu_ms, bases = 0, 0.0, 0
        for first in range(0, n_reads, chunk_reads):
            n = min(chunk_reads, n_reads - first)
            synth.reads_device(spec, first, n, ont, buf.data_ptr())
            torch.cuda.synchronize()
            # the library works on its own stream: wall clock between two device-wide synchronisations, so that work a
            # call leaves in flight (deferred mode returns before a direct insert has finished) is inside the timed region
            t0 = time.perf_counter()
            tot += ms.add_device(buf.data_ptr(), offs.data_ptr(), n, n * L)
            torch.cuda.synchronize()
            gpu_ms += 1e3 * (time.perf_counter() - t0); bases += n * L
        t0 = time.perf_counter()
        ms.flush()                             # what is still waiting in the buckets: part of the timed work
        torch.cuda.synchronize()
        gpu_ms += 1e3 * (time.perf_counter() - t0)
        passes.append(gpu_ms)
    gpu_ms = passes[-1]
    t0 = time.perf_counter()
    h = ms.histogram()
    hist_ms = 1e3 * (time.perf_counter() - t0)
    out = {"config": name, "k": k, "d": d, "tableBits": bits, "bases": bases, "reads": n_reads, "hashes": int(tot),
           "distinct": int(ms.max), "ms": gpu_ms, "gbases_per_s": bases / gpu_ms / 1e6,
           "ms_first_pass_with_allocations": passes[0], "histogram_ms": hist_ms,
           "accumulate": ACCUMULATE if n_reads > chunk_reads else 1, "modal_depth": int(np.argmax(h[2:]) + 2)}
    ms.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="0,1,2,3,4")
    ap.add_argument("--accumulate", type=int, default=8, help="chunks per region build for the read sets (1 = build after every chunk)")
    a = ap.parse_args()
    global ACCUMULATE
    ACCUMULATE = a.accumulate
    only = {int(x) for x in a.only.split(",")}
    mg.require_device()
    G = int(3_100_000_000 * a.scale)

    if 0 in only:       # configs[0]: 10 Mb genome, 30x of 10 kb reads, k=19 d=31
        sp = synth.read_spec(12345, 10_000_000, 7, 10_000)
        count_reads("warm-up", sp, 30_000, 19, 31, 24, 30_000)           # allocations, module load
        print(json.dumps(count_reads("configs[0] modutils build+count, 10 Mb genome, 30x 10 kb reads", sp, 30_000, 19, 31, 24, 30_000)))

    ref = None
    if 1 in only or 4 in only:   # configs[1]: modmap reference index of a 3.1 Gb genome, k=31 d=64 (host buffers: PCIe inside)
        nb = G - G % 32
        d_g = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
        synth.genome_device(12345, 0, nb, 1, d_g.data_ptr())
        torch.cuda.synchronize()
        pinned = torch.empty(nb, dtype=torch.uint8, pin_memory=True)      # "seqio feeds pinned buffers"
        pinned.copy_(d_g[:nb]); torch.cuda.synchronize()
        genome = pinned.numpy()
        del d_g
        torch.cuda.empty_cache()
        cuts = (np.arange(25, dtype=np.float64) * (nb / 24)).astype(np.uint64); cuts[-1] = nb
        t0 = time.perf_counter()
        ref = mg.Reference(28, 31, 64, 17, genome, cuts, is_ascii=0)
        dt = time.perf_counter() - t0
        print(json.dumps({"config": "configs[1] modmap reference index, 3.1 Gb genome, 24 records (modgpuReferenceBuild, host buffers)",
                          "k": 31, "d": 64, "tableBits": 28, "bases": nb, "hits": int(ref.counts[0]), "copy1": int(ref.counts[1]),
                          "copy2": int(ref.counts[2]), "multi": int(ref.counts[3]), "ms": 1e3 * dt, "gbases_per_s": nb / dt / 1e9}))
        del genome, pinned

    if 2 in only:       # configs[2]: HiFi-like 30x of 3.1 Gb, 15 kb reads, 0.1 % errors (one GPU here; bench.py --gpus 8 shards it)
        sp = synth.read_spec(12345, G, 11, 15_000, sub_ppm=1000, dup_mode=1)
        n_reads = int(30 * G / 15_000)
        print(json.dumps(count_reads("configs[2] HiFi-like 30x of 3.1 Gb, 15 kb reads, 0.1 % errors (single GPU)", sp, n_reads, 31, 64, 30, 200_000)))

    if 3 in only:       # configs[3]: Illumina-like 2 x 150 bp, 40x, k=19 d=31: histogram + single-copy classes
        sp = synth.read_spec(12345, G, 13, 150, sub_ppm=3000, frag_len=400, pair_mode=1, dup_mode=1)
        n_reads = 2 * int(40 * G / 300)
        r = count_reads("configs[3] Illumina-like 2x150 bp 40x of 3.1 Gb, 0.3 % errors", sp, n_reads, 19, 31, 31, 20_000_000)
        print(json.dumps(r))

    if 4 in only and ref is not None:   # configs[4]: 1 M ONT-like reads (10 % errors) against the index
        sp = synth.read_spec(12345, G, 5, 10_000, 30_000, 30_000, 40_000, dup_mode=1)
        n_reads, chunk = int(1_000_000 * a.scale), 50_000
        buf = torch.empty(chunk * 10_000 + 64, dtype=torch.uint8, device=dev)
        pin = torch.empty(chunk * 10_000, dtype=torch.uint8, pin_memory=True)
        offs = np.arange(chunk + 1, dtype=np.uint64) * np.uint64(10_000)
        seeds, hits1, dt, bases = 0, 0, 0.0, 0
        # output arrays allocated once (the C ABI fills caller-owned buffers, like the reference's seeds[] array)
        import ctypes as C
        from modimizer_b200 import _lib
        lib = _lib.load()
        cap = chunk * 400
        # in pinned memory (modgpuHostAlloc / cudaMallocHost in a C caller): the seeds come back by DMA
        def pinned(n, dt):
            return torch.zeros(n, dtype=dt, pin_memory=True).numpy()
        so = pinned(chunk + 1, torch.int64).view(np.uint64); si = pinned(cap, torch.int32).view(np.uint32); spos = pinned(cap, torch.int32).view(np.uint32)
        hid = pinned(2 * cap, torch.int32).view(np.uint32); hoff = pinned(2 * cap, torch.int32).view(np.uint32); ctr = pinned(4 * chunk, torch.int32)
        for first in range(0, n_reads, chunk):
            n = min(chunk, n_reads - first)
            synth.reads_device(sp, first, n, True, buf.data_ptr())
            torch.cuda.synchronize()
            pin[:n * 10_000].copy_(buf[:n * 10_000]); torch.cuda.synchronize()
            t0 = time.perf_counter()
            ns = lib.modgpuReferenceQuery(ref._p, pin.data_ptr(), offs.ctypes.data, n, 0, so.ctypes.data, si.ctypes.data, spos.ctypes.data,
                                          hid.ctypes.data, hoff.ctypes.data, ctr.ctypes.data, cap)
            dt += time.perf_counter() - t0
            assert ns != 0xFFFFFFFFFFFFFFFF and ns <= cap, _lib.last_error()
            seeds += int(ns); hits1 += int(ctr[:4 * n].reshape(-1, 4)[:, 1].sum()); bases += n * 10_000
        print(json.dumps({"config": "configs[4] modmap matching of 1 M ONT-like 10 kb reads (10 % errors) against the 3.1 Gb index (host buffers)",
                          "reads": n_reads, "bases": bases, "seeds": seeds, "copy1_hits": hits1, "ms": 1e3 * dt,
                          "gbases_per_s": bases / dt / 1e9}))
    if ref is not None:
        ref.close()


if __name__ == "__main__":
    main()
