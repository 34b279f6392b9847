#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json from the UNMODIFIED reference
(oracle/_ref/libmodref.so, built by oracle/Makefile from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference ships no tests or vectors of its own; these known
answers are what pins the oracle and the CUDA path on boxes without the reference.
Inputs come from the deterministic generators of include/modgpu_synth.h (via
tests/hostemul.py) or are stored literally; outputs are stored in full when small and
as SHA-256 of the little-endian array bytes otherwise."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H       # noqa: E402
import hostemul as he     # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cases():
    """shared with tests/test_golden.py: name -> inputs"""
    c = {}
    c["kat131"] = dict(kind="scan", k=19, w=31, seed=17, ascii=(
        "ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGTACGATCGTAGCTAGCTAGCTGATCGATGCATGCATCGATCGTAGCTAGCTAGCTAGCATCGATGCATGCAAATTTGGGCCCATATCGCGATATCGC"))
    c["kat_short"] = dict(kind="scan", k=19, w=31, seed=17, ascii="ACGTTGCATGCCGATAGC")
    c["kat_k5"] = dict(kind="scan", k=5, w=1, seed=17, ascii="ACGTAC")
    c["kat_pal"] = dict(kind="scan", k=4, w=1, seed=17, ascii="ACGTACGTAATT")
    c["kat_n"] = dict(kind="scan", k=11, w=2, seed=3, ascii="ACGTNNNNACGTTGCAnnacgtACGTTTGACCA")
    for (k, w, seed) in ((19, 31, 17), (31, 64, 17), (16, 32, 0), (1, 1, 5), (24, 7, 99)):
        c["scan_g_k%d_w%d" % (k, w)] = dict(kind="scan", k=k, w=w, seed=seed, genome=dict(seed=777, start=100, n=3000, dup=0))
    c["modset_c1"] = dict(kind="modset", bits=20, k=19, w=31, seed=17,
                          reads=dict(genomeSeed=12345, genomeLen=60000, readSeed=99, readLen=2000, n=600, sub=0),
                          setcopy=[10, 30, 50], setcopyM=35)
    c["modset_c4"] = dict(kind="modset", bits=20, k=19, w=31, seed=17,
                          reads=dict(genomeSeed=12345, genomeLen=30000, readSeed=5, readLen=150, n=6000, sub=3000, frag=400, pair=1),
                          setcopy=[5, 45, 80], setcopyM=60)
    c["modset_k31"] = dict(kind="modset", bits=20, k=31, w=64, seed=17,
                           reads=dict(genomeSeed=4242, genomeLen=80000, readSeed=1, readLen=3000, n=300, sub=1000),
                           setcopy=[3, 20, 40], setcopyM=25)
    c["ref_c2"] = dict(kind="ref", bits=22, k=31, w=64, seed=17, genome=dict(seed=4242, start=0, n=400000, dup=1),
                       seqlens=[200000, 120000, 0, 79969, 31],
                       plant=[[150000, 10000, 10000], [310000, 10000, 2000], [390000, 100000, 5000]],
                       query=dict(readLen=1500, n=150, sub=30000, ins=30000, dele=40000, ont=1, readSeed=8))
    c["ref_k19"] = dict(kind="ref", bits=22, k=19, w=31, seed=17, genome=dict(seed=4242, start=0, n=400000, dup=1),
                        seqlens=[200000, 120000, 0, 79969, 31],
                        plant=[[150000, 10000, 10000], [310000, 10000, 2000], [390000, 100000, 5000]],
                        query=dict(readLen=1000, n=200, sub=10000, ins=0, dele=0, ont=0, readSeed=9))
    return c


def build_inputs(case):
    if "ascii" in case:
        return H.codes_from_ascii(case["ascii"])
    g = case["genome"]
    return he.genome(g["seed"], g["start"], g["n"], g["dup"])


def build_reads(r, ont=False, genomeSeed=None, genomeLen=None):
    sp = he.read_spec(genomeSeed if genomeSeed is not None else r["genomeSeed"],
                      genomeLen if genomeLen is not None else r["genomeLen"], r["readSeed"], r["readLen"],
                      r.get("sub", 0), r.get("ins", 0), r.get("dele", 0), r.get("frag", 0), r.get("pair", 0), r.get("dup", 0))
    data = he.reads(sp, 0, r["n"], ont)
    offs = np.arange(r["n"] + 1, dtype=np.uint64) * np.uint64(r["readLen"])
    return data, offs


def ref_genome(case):
    g = build_inputs(case)
    for (dst, src, n) in case["plant"]:
        g[dst:dst + n] = g[src:src + n]
    offs = np.concatenate([[0], np.cumsum(case["seqlens"])]).astype(np.uint64)
    assert int(offs[-1]) == len(g)
    return g, offs


def evaluate(chk, name, case):
    """run one case on a checker (the reference here, the oracle in the tests)"""
    out = {}
    if case["kind"] == "scan":
        codes = build_inputs(case)
        k, p, f = chk.mod_scan(case["k"], case["w"], case["seed"], codes)
        out = dict(input_sha=sha(codes), n=len(k), kmer=[int(x) for x in k], pos=[int(x) for x in p], isF=[int(x) for x in f])
        out["hasher"] = chk.hasher(case["k"], case["w"], case["seed"])
    elif case["kind"] == "modset":
        data, offs = build_reads(case["reads"])
        ms = chk.modset_new(case["bits"], case["k"], case["w"], case["seed"])
        out["input_sha"] = sha(data)
        out["total_hashes"] = chk.modset_add(ms, data, offs)
        out["max"] = int(chk._modset_max(ms))
        v, d, i = chk.modset_export(ms)
        out["value_sha"], out["depth_sha"] = sha(v), sha(d)
        out["value_head"] = [int(x) for x in v[:8]]
        sv, sd, si = chk.modset_sorted(ms)
        out["sorted_value_sha"], out["sorted_depth_sha"] = sha(sv), sha(sd)
        h = chk.modset_hist(ms)
        out["hist"] = {int(j): int(h[j]) for j in np.nonzero(h)[0]}
        out["summary"] = chk.modset_summary(ms)
        chk._modset_setcopy(ms, *case["setcopy"])
        out["summary_setcopy"] = chk.modset_summary(ms)
        out["info_sha_setcopy"] = sha(chk.modset_sorted(ms)[2])
        chk._modset_setcopyM(ms, case["setcopyM"])
        out["summary_setcopyM"] = chk.modset_summary(ms)
        out["info_sha_setcopyM"] = sha(chk.modset_sorted(ms)[2])
        chk._modset_free(ms)
    elif case["kind"] == "ref":
        g, offs = ref_genome(case)
        r, counts = chk.ref_build(case["bits"], case["k"], case["w"], case["seed"], g, offs)
        out["input_sha"] = sha(g)
        out["counts"] = [int(x) for x in counts]
        e = chk.ref_export(r)
        for key in ("index", "offset", "id", "depth", "rev", "loc"):
            out[key + "_sha"] = sha(e[key])
        v, d, i = chk.modset_export(chk._ref_modset(r))
        out["value_sha"], out["info_sha"], out["depth_all_zero"] = sha(v), sha(i), bool((d == 0).all())
        q = case["query"]
        rd, roffs = build_reads(q, bool(q["ont"]), genomeSeed=case["genome"]["seed"], genomeLen=case["seqlens"][0])
        res = chk.ref_query(r, rd, roffs)
        out["query_input_sha"] = sha(rd)
        for key in ("seedOff", "index", "pos", "hitId", "hitOffset", "counters"):
            out["q_" + key + "_sha"] = sha(res[key])
        out["q_nseeds"] = int(len(res["index"]))
        out["q_counter_sums"] = [int(x) for x in res["counters"].sum(axis=0)]
        chk._ref_free(r)
    return out


def main():
    ref = H.reference()
    if ref is None:
        sys.exit("oracle/_ref/libmodref.so not available: run `make -C oracle` where /root/reference exists")
    golden = {"_generated_by": "tests/golden/make_golden.py from the unmodified reference (oracle/_ref)",
              "_libc_note": "factor1 comes from glibc random(); seed 17 -> 0x49308bb9003cb3ad"}
    for name, case in cases().items():
        golden[name] = evaluate(ref, name, case)
        print(name, "ok")
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
