#!/usr/bin/env python3
"""Reduce an .ncu-rep (ncu --set full) to the few numbers DESIGN.md / bench.py quote, as JSON.
usage: ncu_summary.py report.ncu-rep out.json [spp_per_launch] ["capture note"]
Every launch also records the hash of the kernel sources it was captured from (bench.kernel_source_hash): bench.py
prints `stale: true` beside the instruction counts it takes from here once those sources have changed."""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rep, out = sys.argv[1], sys.argv[2]
spp = int(sys.argv[3]) if len(sys.argv) > 3 else None
note = sys.argv[4] if len(sys.argv) > 4 else None
try:
    import bench
    src_hash = bench.kernel_source_hash()
except Exception:  # noqa: BLE001
    src_hash = None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
want = {
    "gpu__time_duration.sum": "duration", "launch__registers_per_thread": "registers_per_thread", "launch__grid_size": "grid",
    "launch__block_size": "block", "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "avg_active_lanes_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__thread_inst_executed.sum": "thread_instructions", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct", "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
}
res = []
for r in data:
    d = {"kernel": r[h.index("Kernel Name")] if "Kernel Name" in h else ""}
    for k, n in want.items():
        if k in h:
            i = h.index(k)
            try:
                v = float(r[i])
            except ValueError:
                continue
            d[n] = v
            d[n + "_unit"] = units[i]
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if "dram_read" in d:
        d["dram_bytes_per_launch"] = d["dram_read"] * to_bytes.get(d.get("dram_read_unit", "byte"), 1) + \
            d["dram_write"] * to_bytes.get(d.get("dram_write_unit", "byte"), 1)
    if "issue_slot_utilisation_pct" in d and "avg_active_lanes_per_instruction" in d:
        d["issue_roofline_frac"] = round(d["issue_slot_utilisation_pct"] / 100 * d["avg_active_lanes_per_instruction"] / 32, 4)
    d["kernel_source_sha256"] = src_hash
    if spp is not None:
        d["spp_per_launch"] = spp
    if note:
        d["capture"] = note
    res.append(d)
json.dump({"report": rep.split("/")[-1], "launches": res}, open(out, "w"), indent=1)
for d in res:
    print({k: v for k, v in d.items() if not k.endswith("_unit")})
