#!/usr/bin/env python
"""BASELINE configs[2] as it is specified: HiFi-like 30x of a 3.1 Gb genome (93 Gbases, 15 kb reads, 0.1 % errors),
k=31 d=64, counted into ONE modset whose table is hash-sharded over the GPUs of the node.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \\
         tools/config2_sharded.py [--scale 1.0] [--bits 27] [--check]

Every rank generates ITS reads on its device (not timed), adds them in 3-Gbase chunks through ShardedModset (select ->
per-(owner, region) buckets in its own HBM -> the owner's region build reads them over NVLink), and the totals are
reduced at the end: hashes, distinct modimizers, the depth histogram.  Counts are commutative sums, so the result must
equal the single-GPU run of tools/configs_bench.py on the same reads bit for bit (--check repeats it on rank 0 at
reduced scale; at full scale the reference values are the single-GPU ones recorded in profiles/configs_r01_final.jsonl).
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import modimizer_b200 as mg
from modimizer_b200 import synth
from modimizer_b200.dist import ShardedModset


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--bits", type=int, default=27, help="per-GPU table bits")
    ap.add_argument("--check", action="store_true", help="rank 0 repeats the whole set on one GPU and compares")
    ap.add_argument("--accumulate", type=int, default=1, help="batches per count exchange + peer build (deferred build)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mg.require_device()
    G = int(3_100_000_000 * a.scale)
    L = 15_000
    sp = synth.read_spec(12345, G, 11, L, sub_ppm=1000, dup_mode=1)
    n_reads = int(30 * G / L)
    per_rank = n_reads // world                                  # (the remainder of a scaled run is dropped on every rank alike)
    n_chunks = max(1, -(-per_rank // 200_000))
    chunk = -(-per_rank // n_chunks)
    buf = torch.empty(chunk * L + 64, dtype=torch.uint8, device=dev)
    offs = torch.arange(chunk + 1, dtype=torch.int64, device=dev) * L
    sm = ShardedModset(a.bits, 31, 64, 17)
    if a.accumulate > 1:
        sm.set_accumulate(a.accumulate)

    def one_pass():
        t_add = 0.0
        for c in range(n_chunks):
            first = rank * per_rank + c * chunk
            n = min(chunk, per_rank - c * chunk)
            synth.reads_device(sp, first, n, False, buf.data_ptr())
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            sm.add_device(buf.data_ptr(), offs.data_ptr(), n, n * L)
            torch.cuda.synchronize()
            t_add += time.perf_counter() - t0
        t0 = time.perf_counter()
        hashes = sm.synchronize()
        entries = sm.local.max
        torch.cuda.synchronize()
        t_add += time.perf_counter() - t0
        return t_add, hashes, entries

    passes = []
    for rep in range(2):                                         # second pass: buffers and peer mappings exist
        if rep:
            sm.clear()
        t, hashes, entries = one_pass()
        passes.append(t)
    hist = torch.from_numpy(sm.local.histogram().astype(np.int64)).to(dev)
    tot = torch.tensor([hashes, entries], dtype=torch.int64, device=dev)
    tmax = torch.tensor(passes, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(hist); dist.all_reduce(tot); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    hist = hist.cpu().numpy(); tot = tot.cpu().tolist(); tmax = tmax.cpu().tolist()
    out = None
    if rank == 0:
        bases = per_rank * world * L
        out = {"config": "configs[2] HiFi-like 30x of 3.1 Gb, 15 kb reads, 0.1 %% errors, table hash-sharded over %d GPUs" % world,
               "k": 31, "d": 64, "tableBits_per_gpu": a.bits, "n_gpus": world, "accumulate": a.accumulate, "bases": bases, "reads": per_rank * world,
               "hashes": tot[0], "distinct": tot[1], "ms": 1e3 * tmax[1], "gbases_per_s": bases / tmax[1] / 1e9,
               "ms_first_pass_with_allocations": 1e3 * tmax[0], "modal_depth": int(np.argmax(hist[2:]) + 2),
               "timing": "wall clock between device-wide synchronisations around every add (+ the final count readback), max over ranks"}
        if a.check:
            ms = mg.Modset(min(a.bits + 3, 31), 31, 64, 17)
            h1 = 0
            for r in range(world):
                for c in range(n_chunks):
                    n = min(chunk, per_rank - c * chunk)
                    synth.reads_device(sp, r * per_rank + c * chunk, n, False, buf.data_ptr())
                    torch.cuda.synchronize()
                    h1 += ms.add_device(buf.data_ptr(), offs.data_ptr(), n, n * L)
            same = bool(h1 == tot[0] and ms.max == tot[1] and np.array_equal(ms.histogram().astype(np.int64), hist))
            out["single_gpu_identical"] = same
            ms.close()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
    sm.close()
    if world > 1:
        dist.destroy_process_group()
    if out is not None and out.get("single_gpu_identical") is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
