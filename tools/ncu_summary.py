"""Key counters of an .ncu-rep (one or more kernel launches) as text + a JSON entry for profiles/ncu_summary.json.
usage: ncu_summary.py <rep.ncu-rep> <out.txt> [json_key]"""
import csv, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "sm__sass_inst_executed_op_shared_ld.sum"]
lines, js = [], []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append("=" * 100)
    for k in want:
        for h in hdr:
            if h == k or h.startswith(k + "."):
                lines.append(f"{h:75s} {d.get(h, ''):>22s} {units[hdr.index(h)]}")
    lines.append("-- warp stall reasons (average warps stalled per issue-active cycle) --")
    st = [(h, float(d[h])) for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d.get(h) not in (None, "")]
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        lines.append(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.3f}")
    try:
        js.append({"kernel": d["Kernel Name"][:80], "gpu_time_ms": float(d["gpu__time_duration.sum"]) * (1e-6 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1.0 if units[hdr.index("gpu__time_duration.sum")] == "ms" else 1e-3),
                   "dram_bytes_per_launch": (float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]] +
                                             float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]),
                   "fp64_pipe_pct_of_active": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
                   "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                   "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                   "registers": int(float(d["launch__registers_per_thread"]))})
    except Exception as e:
        lines.append(f"(json summary failed: {e})")
open(out, "w").write(f"ncu --set full --clock-control none --import-source on, report {os.path.basename(rep)} (ncu -i ... --page raw --csv)\n" + "\n".join(lines) + "\n")
if key and js:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[key] = js[0] if len(js) == 1 else js
    json.dump(cur, open(p, "w"), indent=1)
print("\n".join(lines[:60]))
