#!/usr/bin/env python
"""bench.py - Gbases/s hashed + counted into the modset on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] - synthetic 3.1 Gb genome, 24
records, k=31 d=64, table bits 28: one STEP = one modset build + count over the
whole genome from an empty table: clear -> hash_count_kernel (K1 pack2bit fused
into the tile loader + K2 hash/select + the bucket scatter of K3) -> region build
(K3 insert/count in shared memory) -> entries readback.

  value     whole-job Gbases/s with the bases already resident in HBM
  e2e       the same through the reference-facing C-ABI call on HOST buffers
            (modgpuModsetAdd: pinned host memory, H2D inside the timed region,
            D2H read of ms->max)
  roofline  dominant kernel by device time: algorithmic bytes /
            CUDA-event duration measured live on the launching stream, against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU path (oracle/_ref when built, else the
            oracle port) on a bounded sample of the same genome, 1 core

N > 1 (torchrun, one rank per GPU): weak scaling - every rank builds from its own
3.1 Gb shard, the table is sharded by k-mer hash, and the selected k-mers reach
their owner GPU through peer memory: they are scattered into per-(owner, region)
buckets in the selecting rank's HBM and the owner's region build reads them over
NVLink (two small NCCL all-to-alls carry the fill counts and act as the barrier).

--impl reference: the reference's CPU implementation of the same path on the
host cores (all threads it can use: independent modsets on disjoint chunks, the
reference's only parallel recipe), same metric/config, bounded sample per step.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_SEED = 12345
K, D, HSEED = 31, 64, 17


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gbases", type=float, default=3.1, help="genome size per GPU in Gbases")
    ap.add_argument("--records", type=int, default=24)
    ap.add_argument("--bits", type=int, default=28)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-mbases", type=float, default=400.0, help="bounded CPU-baseline sample")
    ap.add_argument("--flags", type=int, default=0, help="MODGPU_SEL_* flags for A/B runs")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smmax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def record_offsets(nbases, nrec):
    import numpy as np
    cuts = (np.arange(nrec + 1, dtype=np.float64) * (nbases / nrec)).astype(np.uint64)
    cuts[-1] = nbases
    return cuts


# ------------------------------------------------------------ reference arm --
def cpu_checker():
    import harness
    ref = harness.reference()
    if ref is not None:
        return ref, "reference"
    return harness.oracle(), "port"


def cpu_build(chk, codes, offs, bits):
    """the reference's addSequence loop (modutils.c:19-31) over one batch; returns (seconds, hashes, distinct)"""
    ms = chk.modset_new(bits, K, D, HSEED)
    t0 = time.perf_counter()
    tot = chk.modset_add(ms, codes, offs)
    dt = time.perf_counter() - t0
    mx = chk._modset_max(ms)
    chk._modset_free(ms)
    return dt, tot, mx


def run_reference(args):
    """--impl reference: CPU path, all host threads, bounded sample per step"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import hostemul as he
    chk, kind = cpu_checker()
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(ncpu, 32))
    per_thread = int(32e6)                                   # bases per thread per step
    bits = 24                                                # 4.2 M-entry capacity >> 0.5 M selected per chunk
    chunks = []
    for t in range(threads):
        codes = he.genome(GENOME_SEED, t * per_thread, per_thread, 1)
        chunks.append((codes, np.array([0, per_thread], np.uint64)))

    def one_step():
        res = [None] * threads

        def work(i):
            res[i] = cpu_build(chk, chunks[i][0], chunks[i][1], bits)
        th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    bases = threads * per_thread * args.steps
    val = bases / total / 1e9
    sample = ("%d threads x %.0f Mbases of the same synthetic genome per step, independent modsets (bits %d) on "
              "disjoint chunks, no merge" % (threads, per_thread / 1e6, bits))
    out = {"impl": "reference", "metric": "Gbases/s hashed+counted into modset", "value": val, "unit": "Gbases/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": workload_config(args, per_gpu_bases=threads * per_thread),
           "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def workload_config(args, per_gpu_bases):
    return {"workload": "modset build+count, synthetic genome, k=%d d=%d seed=%d tableBits=%d (BASELINE configs[1])" % (K, D, HSEED, args.bits),
            "bases_per_gpu": int(per_gpu_bases), "records_per_gpu": args.records,
            "l2": "inputs (%.1f GB codes per step) larger than the 126 MB L2" % (per_gpu_bases / 1e9),
            "parallelism": ("table sharded by k-mer hash, reads by input chunk; peer-memory exchange: the owner's region build "
                            "reads every rank's buckets over NVLink, NCCL only for the fill counts") if args.gpus > 1 else "single GPU"}


def bind_to_gpu_numa_node(torch, index):
    """run this rank (and allocate its pinned staging buffer) on the CPUs next to its GPU: with 4 or 8 ranks
    feeding their GPUs from host memory, buffers on the wrong socket halve the PCIe rate of the e2e path"""
    try:
        pr = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


# ------------------------------------------------------------------ our arm --
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import modimizer_b200 as mg
    from modimizer_b200 import synth, _lib
    from modimizer_b200.dist import ShardedModset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    mg.require_device()
    lib = _lib.load()
    _lib.check(lib.modgpuSetDevice(local_rank if world > 1 else 0))
    numa = bind_to_gpu_numa_node(torch, local_rank if world > 1 else 0)
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream()

    nb = int(args.gbases * 1e9)
    nb -= nb % 32
    offs = record_offsets(nb, args.records)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_bases = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
    synth.genome_device(GENOME_SEED, rank * nb, nb, 1, d_bases.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()

    sm = ShardedModset(args.bits, K, D, HSEED)
    sm.local.set_flags(args.flags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a group of records must stay below 2^32 bases per select call
    groups = []
    r0 = 0
    for r in range(1, args.records + 1):
        if r == args.records or offs[r + 1] - offs[r0] > (1 << 32) - (1 << 20):
            groups.append((r0, r)); r0 = r
    d_goffs = [torch.from_numpy((offs[a:b + 1] - offs[a]).view(np.int64)).to(dev) for a, b in groups]

    state = {}

    def step_device():
        sm.clear()
        tot = 0
        for (a, b), go in zip(groups, d_goffs):
            tot += sm.add_device(d_bases.data_ptr() + int(offs[a]), go.data_ptr(), b - a, int(offs[b] - offs[a]))
        state["entries"] = sm.local.max                     # D2H readback of ms->max (syncs)
        state["hashes"] = tot if world == 1 else sm.synchronize()

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * nb / (ms_step * 1e-3) / 1e9
    hashes, entries = state["hashes"], state["entries"]

    # ---- per-kernel device time (CUDA events on the launching stream) for the roofline
    sm.local.profile(True)
    for _ in range(max(2, min(args.steps, 3))):
        step_device()
    prof_steps = max(2, min(args.steps, 3))
    times = sm.local.times()
    sm.local.profile(False)
    peak, peak_src = measured_peak()
    n_ins = hashes                                          # inserts this rank issued per step (world 1)
    # algorithmic bytes: K1 is fused into K2's tile loader, so the select pass reads the raw bytes
    # (1 B/base) and writes the selected k-mers (8/d B/base): 1 + 8/d; the "pack" scope only holds the sparse
    # end-flag marking (two offsets read, one flag word touched, per sequence); insert = 20 B per selected k-mer
    alg = {"pack": 2.0 * args.records * 20.0, "select": nb * (1.0 + 8.0 / D), "insert": n_ins * 20.0}
    kern = {}
    for name in ("pack", "select", "insert"):
        ms_k = times[name][0] / prof_steps
        kern[name] = {"ms_per_step": ms_k, "alg_bytes_per_step": alg[name],
                      "achieved_gbs": (alg[name] / (ms_k * 1e-3) / 1e9) if ms_k > 0 else None}
    kern["pack"]["what"] = "mark_ends_kernel + unmark_ends_kernel: one flag per sequence set and cleared again (K1 itself runs inside the select kernel)"
    kern["select"]["what"] = "lut_build_kernel + hash_count_kernel: fused K1 pack2bit + K2 hash/select + K3 bucket scatter"
    kern["insert"]["what"] = ("region_build_pipe_kernel: persistent, pipelined region build in shared memory (+ overflow inserts)" +
                              (", reading every rank's buckets over NVLink" if world > 1 else "; writes the 2 GiB table once"))
    dom = max(("pack", "select", "insert"), key=lambda n: kern[n]["ms_per_step"])
    launches_per_step = sum(times[n][1] for n in times) // prof_steps
    traffic = None
    kname = {"pack": "mark_ends_kernel", "select": "hash_count_kernel",
             "insert": "region_build_pipe_kernel"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
        if abs(nb - 3.1e9) < 1e8 and world == 1 and args.flags == 0:
            traffic = tr.get(kname[dom])    # dram bytes per launch from the committed ncu --set full capture of this config
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname[dom],
                "achieved": kern[dom]["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (kern[dom]["achieved_gbs"] / peak) if kern[dom]["achieved_gbs"] else None,
                "traffic": traffic, "kernels": kern,
                "note": "hash_count_kernel (fused pack + hash/select + scatter) is bound by instruction issue and the shared-memory "
                        "pipe, not by HBM (ncu r01 v6: issue 74 %, ALU pipe 62 %, shared-memory pipe ~65 %, DRAM 34 %, 12.8 inst/base); "
                        "1 + 8/d algorithmic bytes per base; see DESIGN.md section 3 and profiles/"}

    # ---- e2e: host buffers through the C ABI, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        h_ptr = lib.modgpuHostAlloc(nb + 64)
        if not h_ptr:
            raise RuntimeError("pinned allocation failed: " + _lib.last_error())
        import ctypes
        # D2H of the generated shard into the pinned buffer (outside the timed region)
        hview = (ctypes.c_uint8 * nb).from_address(h_ptr)
        h_np = np.frombuffer(hview, dtype=np.uint8)
        chunk = 1 << 28
        for a in range(0, nb, chunk):
            b = min(nb, a + chunk)
            h_np[a:b] = d_bases[a:b].cpu().numpy()
        h_goffs = [np.ascontiguousarray(offs[a:b + 1] - offs[a]) for a, b in groups]

        def step_host():
            sm.clear()
            tot = 0
            for (a, b), go in zip(groups, h_goffs):
                if world == 1:
                    tot += sm.local.add_pointers(h_ptr + int(offs[a]), go.ctypes.data, b - a, 0)
                else:
                    tot += sm.add_pointers(h_ptr + int(offs[a]), go.ctypes.data, b - a, 0)
            state["entries_e2e"] = sm.local.max
            state["hashes_e2e"] = tot if world == 1 else sm.synchronize()

        for _ in range(max(1, args.warmup - 1)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_host()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt.item())
        e2e = {"value": world * nb * args.steps / wall / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(nb + 8 * (args.records + len(groups))), "d2h_bytes_per_step": 24 * len(groups) + 16,
               "ms_per_step": 1e3 * wall / args.steps, "host_cpus": numa,
               "timing": "host wall clock around the C-ABI calls (they synchronise), max over ranks"}
        assert state["hashes_e2e"] == hashes and state["entries_e2e"] == entries, "host path and device path disagree"
        lib.modgpuHostFree(h_ptr)

    clocks = sampler.stop() if rank == 0 else None       # sampled across both timed regions (value and e2e)

    # ---- CPU baseline (rank 0, N = 1): bounded sample of the same genome
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        chk, kind = cpu_checker()
        sample_b = int(min(nb, args.cpu_mbases * 1e6))
        sample_b -= sample_b % 32
        codes = d_bases[:sample_b].cpu().numpy()
        soffs = np.array([0, sample_b], np.uint64)
        dt, tot, mx = cpu_build(chk, codes, soffs, args.bits)
        # parity spot check of the very same sample on the GPU
        gms = mg.Modset(args.bits, K, D, HSEED)
        gtot = gms.add(codes, soffs, is_ascii=0)
        gmax = gms.max
        gms.close()
        assert (gtot, gmax) == (tot, mx), "GPU and CPU disagree on the baseline sample: %s vs %s" % ((gtot, gmax), (tot, mx))
        cpu = {"value": sample_b / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": kind,
               "sample": "first %.0f Mbases of the same genome as one record, tableBits %d, 1 thread (the reference is single-threaded); "
                         "hashes %d, distinct %d identical on GPU" % (sample_b / 1e6, args.bits, tot, mx)}

    if rank == 0:
        out = {"metric": "Gbases/s hashed+counted into modset", "value": value, "unit": "Gbases/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
               "config": workload_config(args, nb), "hashes_per_step_rank0": int(hashes), "entries_rank0": int(entries),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
               "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(out))
    sm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
