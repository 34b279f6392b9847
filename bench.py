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

Also on the N = 1 line: "configs" - BASELINE configs[0] (the reference's default
k=19 d=31, 30x 10 kb reads of a 10 Mb genome, full size) and the same k=19 d=31
over the bench genome, each with device-resident rate, e2e and the roofline of the
full-scan select kernel; "e2e.roofline" - the e2e rate against the pinned H2D
bandwidth measured on this box in this run; "e2e_packed" - the same build fed
through modgpuModsetAddPacked (the reference's own 2-bit sqioSeqPack layout).
On the N > 1 lines: "parity" - a bounded sample of every rank's shard built into a
second sharded set, summed over the ranks and compared with oracle/_ref on rank 0.

--impl reference: the reference's CPU implementation of the same path on the
host cores (all threads it can use: independent modsets on disjoint chunks, the
reference's only parallel recipe, followed by its modsetMerge), same
metric/config (tableBits included), bounded sample per step.
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
    ap.add_argument("--no-configs", action="store_true", help="skip the extra configurations on the N = 1 line")
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
    """--impl reference: CPU path, all host threads, bounded sample per step.  Same hasher and the same tableBits as the
    GPU arm; the reference is single-threaded, so "all cores" is its own recipe: one modset per chunk of the input
    (modutils.c:101-103), then modsetMerge (modset.c:106-128) - both figures are reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import hostemul as he
    chk, kind = cpu_checker()
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(ncpu, 32))
    bits = args.bits
    # a tableBits-28 modset touches ~ one 4 KiB page of index[] per new k-mer: budget the threads against free memory
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_thread = int(32e6)                                   # bases per thread per step
    need = (4 << bits) + (per_thread // D) * 11 * 2 + (1 << 26)   # index[] (touched) + value/depth/info of one instance
    threads = max(1, min(threads, int(0.6 * avail / need)))
    chunks = []
    for t in range(threads):
        codes = he.genome(GENOME_SEED, t * per_thread, per_thread, 1)
        chunks.append((codes, np.array([0, per_thread], np.uint64)))

    def one_step():
        # the tables are created and their pages touched OUTSIDE the timed region: a 32-Mbase sample would otherwise pay
        # the page faults of a whole 1 GiB index[] that the full 3.1-Gbase job pays once (< 1 % of its time)
        sets = [chk.modset_new(bits, K, D, HSEED) for _ in range(threads)]
        for ms in sets:
            chk._modset_prefault(ms)

        def work(i):
            chk.modset_add(sets[i], chunks[i][0], chunks[i][1])
        th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(1, threads):                          # modutils -m: the union into the first set
            assert chk._modset_merge(sets[0], sets[i]) == 1
        t_merge = time.perf_counter() - t0
        mx = chk._modset_max(sets[0])
        for ms in sets:
            chk._modset_free(ms)
        return t_build, t_merge, mx

    for _ in range(min(args.warmup, 1)):
        one_step()
    res = [one_step() for _ in range(args.steps)]
    t_build = sum(r[0] for r in res); t_merge = sum(r[1] for r in res)
    bases = threads * per_thread * args.steps
    val = bases / t_build / 1e9
    val_merged = bases / (t_build + t_merge) / 1e9
    sample = ("%d threads x %.0f Mbases of the same synthetic genome per step, one modset (tableBits %d, k=%d d=%d) per thread on "
              "disjoint chunks = the reference's own parallel recipe (modutils.c:101-103); value excludes, value_with_merge includes "
              "the modsetMerge of the %d sets into one (modset.c:106-128: %.2f s per step, %d distinct)"
              % (threads, per_thread / 1e6, bits, K, D, threads, t_merge / args.steps, res[-1][2]))
    nb = int(args.gbases * 1e9); nb -= nb % 32
    out = {"impl": "reference", "metric": "Gbases/s hashed+counted into modset", "value": val, "unit": "Gbases/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_build / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": workload_config(args, per_gpu_bases=nb),
           "sample_bases_per_step": threads * per_thread,
           "cpu_baseline": {"value": val, "value_with_merge": val_merged, "merge_s_per_step": t_merge / args.steps,
                            "unit": "Gbases/s", "cores": threads, "kind": kind, "sample": sample},
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


def h2d_ceiling(torch, dev, barrier, nbytes=1 << 30):
    """pinned host -> device bandwidth on this box, every rank copying at the same time (GB/s of THIS rank)"""
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src.zero_()
    best = 0.0
    for it in range(4):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del src, dst
    return best


def seqio_pack_device(torch, d_codes, offs):
    """sqioSeqPack (reference seqio.c:557-570) of the records of a device-resident batch, computed with torch on the
    device (bench set-up, not timed); returns pinned (packed, byteOffs)"""
    import numpy as np
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    nbytes = (lens + 3) // 4
    boffs = np.zeros(len(lens) + 1, np.uint64)
    boffs[1:] = np.cumsum(nbytes).astype(np.uint64)
    out = torch.empty(int(boffs[-1]), dtype=torch.uint8, pin_memory=True)
    for r in range(len(lens)):
        L = int(lens[r])
        if not L:
            continue
        c = d_codes[int(offs[r]):int(offs[r]) + L] & 3
        full = L // 4 * 4
        q = c[:full].view(-1, 4)
        b = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        o = int(boffs[r])
        out[o:o + b.numel()].copy_(b)
        if L > full:                                          # the short last byte is right-aligned
            v = 0
            for x in c[full:].cpu().tolist():
                v = (v << 2) | int(x)
            out[o + b.numel()] = v
    torch.cuda.synchronize()
    return out, boffs


def time_modset(torch, ms, step, steps, warmup):
    """ms/step of `step` (device time on torch's current stream, which the modset was given) and per-kernel times"""
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    ms.profile(True)
    for _ in range(2):
        step()
    t = ms.times()
    ms.profile(False)
    return ms_step, {k: (v[0] / 2, v[1] // 2) for k, v in t.items()}


def bench_config(torch, mg, lib, name, k, d, bits, d_bases, d_offs, h_offs, nseq, nb, peak, steps, warmup):
    """one more configuration through the same public calls: device-resident rate, e2e from pinned host memory, and the
    roofline of its select kernel (for k=19 d=31 the full scan of every window: hash_count_kernel<0,...>)"""
    import ctypes
    import numpy as np
    ms = mg.Modset(bits, k, d, HSEED)
    ms.set_stream(torch.cuda.current_stream().cuda_stream)
    st = {}

    def step():
        ms.clear()
        st["hashes"] = ms.add_device(d_bases.data_ptr(), d_offs.data_ptr(), nseq, nb)
        st["entries"] = ms.max
    ms_step, kern = time_modset(torch, ms, step, steps, warmup)
    sel_ms = kern["select"][0]
    alg = nb * (1.0 + 8.0 / d)
    out = {"config": name, "k": k, "d": d, "tableBits": bits, "bases": int(nb), "sequences": int(nseq), "hashes": int(st["hashes"]),
           "entries": int(st["entries"]), "ms_per_step": ms_step, "value": nb / (ms_step * 1e-3) / 1e9, "unit": "Gbases/s",
           "roofline": {"bound": "hbm", "kernel": "hash_count_kernel (%s)" % ("table-driven candidates" if (k >= 30 and d & (d - 1) == 0) else "full scan of every window"),
                        "ms_per_step": sel_ms, "alg_bytes_per_step": alg, "achieved": alg / (sel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (sel_ms * 1e-3) / 1e9 / peak},
           "insert_ms_per_step": kern["insert"][0]}
    # e2e: the same batch from pinned host memory through modgpuModsetAdd
    h_ptr = lib.modgpuHostAlloc(nb + 64)
    hview = np.frombuffer((ctypes.c_uint8 * nb).from_address(h_ptr), dtype=np.uint8)
    for a in range(0, nb, 1 << 28):
        b = min(nb, a + (1 << 28))
        hview[a:b] = d_bases[a:b].cpu().numpy()

    def step_host():
        ms.clear()
        st["hashes_e2e"] = ms.add_pointers(h_ptr, h_offs.ctypes.data, nseq, 0)
        st["entries_e2e"] = ms.max
    for _ in range(2):
        step_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    assert (st["hashes_e2e"], st["entries_e2e"]) == (st["hashes"], st["entries"])
    out["e2e"] = {"value": nb / wall / 1e9, "unit": "Gbases/s", "ms_per_step": 1e3 * wall, "h2d_bytes_per_step": int(nb + 8 * (nseq + 1))}
    lib.modgpuHostFree(h_ptr)
    ms.close()
    return out


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
    kern["select"]["what"] = "lut_build_kernel + hash_count2_kernel: fused K1 pack2bit + K2 hash/select + K3 bucket scatter"
    kern["insert"]["what"] = (("region_build_tma_kernel: persistent region build in shared memory (+ overflow inserts), every rank's buckets "
                               "pulled over NVLink by TMA bulk copies into a shared-memory ring") if world > 1 else
                              "region_build_pipe_kernel: persistent, pipelined region build in shared memory (+ overflow inserts); writes the 2 GiB table once")
    dom = max(("pack", "select", "insert"), key=lambda n: kern[n]["ms_per_step"])
    launches_per_step = sum(times[n][1] for n in times) // prof_steps
    kname = {"pack": "mark_ends_kernel", "select": "hash_count2_kernel",
             "insert": "region_build_tma_kernel" if world > 1 else "region_build_pipe_kernel"}
    # dram__bytes_read + dram__bytes_write of the dominant kernel: NOT measured by this run (ncu cannot ride along a timed
    # run); it is the committed `ncu --set full` capture of the same kernel at this same configuration, named beside it
    traffic, traffic_src = None, None
    for tf in ("traffic_r02.json", "traffic_r01.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if abs(nb - 3.1e9) < 1e8 and world == 1 and args.flags == 0 and kname[dom] in tr:
                traffic, traffic_src = tr[kname[dom]], "profiles/%s: %s" % (tf, tr.get("_note", ""))
                break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": kname[dom],
                "achieved": kern[dom]["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (kern[dom]["achieved_gbs"] / peak) if kern[dom]["achieved_gbs"] else None,
                "traffic": traffic, "traffic_source": traffic_src, "kernels": kern,
                "note": "hash_count2_kernel (fused pack + hash/select + scatter) is bound by instruction issue and the shared-memory "
                        "pipe, not by HBM (ncu r02, profiles/ncu_summary_r02_lut.json: issue 79 %, shared-memory pipe 69 %, ALU pipe 66 %, "
                        "10.3 thread instructions per base, DRAM traffic 1.05x the algorithmic bytes); 1 + 8/d algorithmic bytes per base; "
                        "see DESIGN.md section 3 and profiles/"}

    # ---- e2e: host buffers through the C ABI, H2D inside the timed region
    e2e, e2e_packed = None, None
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
        # what bounds it: the pinned H2D rate of this box, measured now with every rank copying at once
        ceil = h2d_ceiling(torch, dev, barrier)
        e2e["roofline"] = {"bound": "pcie_h2d", "peak": ceil, "unit": "GB/s", "peak_source": "1 GiB pinned cudaMemcpyAsync on this box in this run, all %d ranks at once, best of 3" % world,
                           "achieved": e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9,
                           "frac": e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9 / ceil if ceil else None}
        if world == 1:
            # the same build fed with the reference's own 2-bit packing (sqioSeqPack, seqio.c:557-570): 0.25 B/base over PCIe
            pk, boffs = seqio_pack_device(torch, d_bases, offs)
            h_boffs = [np.ascontiguousarray(boffs[a:b + 1]) for a, b in groups]      # absolute byte offsets: the base pointer is shared

            def step_packed():
                sm.clear()
                tot = 0
                for (a, b), go, bo in zip(groups, h_goffs, h_boffs):
                    tot += sm.local.add_packed_pointers(pk.data_ptr(), bo.ctypes.data, go.ctypes.data, b - a)
                state["entries_pk"] = sm.local.max
                state["hashes_pk"] = tot
            for _ in range(2):
                step_packed()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_packed()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / args.steps
            assert state["hashes_pk"] == hashes and state["entries_pk"] == entries, "packed path and device path disagree"
            e2e_packed = {"value": nb / wall / 1e9, "unit": "Gbases/s", "ms_per_step": 1e3 * wall,
                          "h2d_bytes_per_step": int(pk.numel() + 16 * (args.records + len(groups))),
                          "what": "modgpuModsetAddPacked: host buffers in the reference's sqioSeqPack layout, expanded on the device"}
            del pk

    clocks = sampler.stop() if rank == 0 else None       # sampled across both timed regions (value and e2e)

    # ---- N > 1: the sharded result against the reference, on a bounded sample of every rank's shard
    parity = None
    if world > 1:
        sample = int(min(nb, 64e6)); sample -= sample % 32
        sm2 = ShardedModset(26, K, D, HSEED)
        go = torch.tensor([0, sample], dtype=torch.int64, device=dev)
        sm2.add_device(d_bases.data_ptr(), go.data_ptr(), 1, sample)
        sel = torch.tensor([sm2.synchronize()], dtype=torch.int64, device=dev)
        dist.all_reduce(sel)
        g_hist, g_entries = sm2.histogram(), sm2.global_max()
        sm2.close()
        if rank == 0:
            import hostemul as he
            chk, kind = cpu_checker()
            cms = chk.modset_new(26, K, D, HSEED)
            tot = 0
            for r in range(world):                           # every rank's sample, regenerated on the host
                tot += chk.modset_add(cms, he.genome(GENOME_SEED, r * nb, sample, 1), np.array([0, sample], np.uint64))
            ok = bool(tot == int(sel.item()) and chk._modset_max(cms) == g_entries and np.array_equal(chk.modset_hist(cms), g_hist))
            parity = {"n_gpus": world, "bases": world * sample, "hashes": int(tot), "entries": int(g_entries), "against": kind,
                      "what": "first %d Mbases of every rank's shard into a second sharded set (tableBits 26 per GPU): total hashes, "
                              "distinct entries and the 65536-bin depth histogram summed over the ranks == oracle/_ref on the "
                              "concatenated samples" % (sample // 1000000), "ok": ok}
            chk._modset_free(cms)
            assert ok, "sharded modset differs from the CPU reference: %s" % parity

    # ---- N = 1: the other shapes the driver should see (BASELINE configs[0]; the reference's default k=19 d=31)
    configs = None
    if world == 1 and not args.no_configs:
        configs = []
        sp = synth.read_spec(GENOME_SEED, 10_000_000, 7, 10_000)
        n_reads, L = 30_000, 10_000
        rbuf = torch.empty(n_reads * L + 64, dtype=torch.uint8, device=dev)
        roffs = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L)
        d_roffs = torch.from_numpy(roffs.view(np.int64)).to(dev)
        synth.reads_device(sp, 0, n_reads, False, rbuf.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        configs.append(bench_config(torch, mg, lib, "BASELINE configs[0]: modutils build+count, synthetic 10 Mb genome, 30x 10 kb reads, k=19 d=31 (full size)",
                                    19, 31, 24, rbuf, d_roffs, roffs, n_reads, n_reads * L, peak, max(3, args.steps), 3))
        del rbuf
        if abs(nb - 3.1e9) < 1e8:
            configs.append(bench_config(torch, mg, lib, "the bench genome (3.1 Gbases, 24 records) at the reference's default k=19 d=31: the full-scan select kernel at scale",
                                        19, 31, 29, d_bases, d_offs, offs, args.records, nb, peak, 3, 2))

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
        if e2e_packed:
            out["e2e_packed"] = e2e_packed
        if parity:
            out["parity"] = parity
        if configs:
            out["configs"] = configs
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
