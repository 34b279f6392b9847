#!/usr/bin/env python
"""File-fed end to end, once ("also report the stock CLI wall time on a FASTA of the same data"):
the stock reference tools (oracle/_ref/modutils, modmap: CPU, single thread) against the same tools on the GPU path -
modutils_dropin / modmap_dropin (the reference's own sources with the caller loops swapped at build time, linked against
libmodshim.so) and modutils_gpu / modmap_gpu (own C drivers over the ABI) - on the same FASTA files, wall clock of
the whole process (parse + compute + print), outputs compared line by line.

  python tools/filefed_bench.py [--gbases 3.1] [--reads 50000] [--dir /tmp/filefed]

FASTA parsing stays on the host in every variant (seqio: ~0.17 Gbases/s per core), so the GPU variants are parse-bound
by construction; the double-buffered feeder of libmodshim overlaps the GPU work with it.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from modimizer_b200 import synth


def stable(text):
    return [l for l in text.splitlines() if not l.startswith("user\t") and not l.startswith("total resources used")]


def write_fasta(path, d_codes, offs, names, width=0):
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=d_codes.device)
    with open(path, "wb") as f:
        for r in range(len(offs) - 1):
            a, b = int(offs[r]), int(offs[r + 1])
            f.write((">%s\n" % names[r]).encode())
            step = 1 << 28
            for x in range(a, b, step):
                y = min(b, x + step)
                f.write(lut[d_codes[x:y].long()].cpu().numpy().tobytes())
            f.write(b"\n")


def run(tool, args, cwd):
    t0 = time.perf_counter()
    r = subprocess.run([tool] + args, cwd=cwd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode:
        raise RuntimeError("%s %s failed: %s" % (tool, args, r.stderr[-500:]))
    return dt, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbases", type=float, default=3.1)
    ap.add_argument("--reads", type=int, default=50000)
    ap.add_argument("--dir", default="/tmp/filefed")
    ap.add_argument("--skip-stock", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    dev = torch.device("cuda:0")
    nb = int(a.gbases * 1e9); nb -= nb % 32
    d = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
    synth.genome_device(12345, 0, nb, 1, d.data_ptr())
    torch.cuda.synchronize()
    offs = (np.arange(25, dtype=np.float64) * (nb / 24)).astype(np.uint64); offs[-1] = nb
    write_fasta(os.path.join(a.dir, "g.fa"), d, offs, ["chr%d" % (i + 1) for i in range(24)])
    del d
    sp = synth.read_spec(12345, nb, 5, 10_000, 30_000, 30_000, 40_000, dup_mode=1)     # ONT-like: 10 % errors
    rd = torch.empty(a.reads * 10_000 + 64, dtype=torch.uint8, device=dev)
    synth.reads_device(sp, 0, a.reads, True, rd.data_ptr())
    torch.cuda.synchronize()
    roffs = np.arange(a.reads + 1, dtype=np.uint64) * np.uint64(10_000)
    write_fasta(os.path.join(a.dir, "r.fa"), rd, roffs, ["read%d" % i for i in range(a.reads)])
    del rd
    torch.cuda.empty_cache()

    tool = lambda n: os.path.join(ROOT, "modimizer_b200", n)
    stock = lambda n: os.path.join(ROOT, "oracle", "_ref", n)
    out = {"genome_bases": nb, "reads": a.reads, "read_bases": a.reads * 10_000, "host_cpus": os.cpu_count(), "runs": []}
    # modutils: -c 28 31 64 17 -a g.fa -H x.his     (configs[1] shape through the CLI)
    ref_out = None
    for name, exe in (("stock modutils", stock("modutils")), ("modutils_dropin", tool("modutils_dropin")), ("modutils_gpu", tool("modutils_gpu"))):
        if not os.path.exists(exe) or (a.skip_stock and name.startswith("stock")):
            continue
        tag = name.replace(" ", "_")
        dt, r = run(exe, ["-o", tag + ".out", "-c", "28", "31", "64", "17", "-a", "g.fa", "-H", tag + ".his"], a.dir)
        lines = stable(open(os.path.join(a.dir, tag + ".out")).read())
        his = open(os.path.join(a.dir, tag + ".his")).read()
        same = None
        if ref_out is None:
            ref_out = (lines, his)
        else:
            same = (lines, his) == ref_out
        out["runs"].append({"tool": name, "command": "-c 28 31 64 17 -a g.fa -H", "wall_s": dt, "gbases_per_s": nb / dt / 1e9,
                            "identical_to_stock": same, "summary": [l for l in lines if l.startswith("added")][:1]})
        print(json.dumps(out["runs"][-1]), flush=True)
    # modmap: -K 31 -W 64 -B 28 -f g.fa -q r.fa
    ref_out = None
    for name, exe in (("stock modmap", stock("modmap")), ("modmap_dropin", tool("modmap_dropin")), ("modmap_gpu", tool("modmap_gpu"))):
        if not os.path.exists(exe) or (a.skip_stock and name.startswith("stock")):
            continue
        tag = name.replace(" ", "_")
        dt, r = run(exe, ["-o", tag + ".out", "-K", "31", "-W", "64", "-B", "28", "-f", "g.fa", "-q", "r.fa"], a.dir)
        lines = stable(open(os.path.join(a.dir, tag + ".out")).read())
        same = None
        if ref_out is None:
            ref_out = lines
        else:
            same = lines == ref_out
        out["runs"].append({"tool": name, "command": "-K 31 -W 64 -B 28 -f g.fa -q r.fa", "wall_s": dt,
                            "gbases_per_s": (nb + a.reads * 10_000) / dt / 1e9, "identical_to_stock": same,
                            "lines": len(lines), "M_lines": sum(1 for l in lines if l.startswith("M\t"))})
        print(json.dumps(out["runs"][-1]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
