#!/usr/bin/env python
"""ncu report -> the handful of numbers DESIGN.md / bench.py quote (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/r2/prof.ncu-rep profiles/ncu_summary_r02_x.json [units-per-launch for inst/unit]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    f = lambda k: float(d[k]) if d.get(k) not in (None, "", "no data") else None
    unit = lambda k: rows[1][hdr.index(k)] if k in hdr else ""
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    rd = f("dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1.0)
    wr = f("dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1.0)
    tscale = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1.0)
    s = {"kernel": d.get("Kernel Name"), "gpu_time_ms": f("gpu__time_duration.sum") * tscale,
         "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_traffic_bytes": rd + wr,
         "warp_instructions": f("smsp__inst_executed.sum"),
         "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
         "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "pipe_fmaheavy_pct": f("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
         "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
         "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
         "shared_bank_conflict_wavefronts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
         "shared_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
         "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "registers_per_thread": f("launch__registers_per_thread"),
         "achieved_occupancy_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "local_load_requests": f("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"),
         "local_store_requests": f("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
         "stalls_per_issue": {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(v), 3)
                              for h, v in zip(hdr, r) if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h and v not in ("", "no data") and float(v) >= 0.1}}
    if units and s["warp_instructions"]:
        s["thread_instructions_per_unit"] = s["warp_instructions"] * 32 / units
    res.append(s)
json.dump({"report": rep, "command": "ncu --set full --clock-control none --import-source on", "launches": res}, open(out, "w"), indent=1)
for s in res:
    print(s["kernel"][:60], "%.3f ms" % s["gpu_time_ms"], "traffic %.3f GB" % (s["dram_traffic_bytes"] / 1e9), "issue %.1f%%" % s["issue_active_pct"])
