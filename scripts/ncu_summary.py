"""ncu report -> compact JSON rows under profiles/ (one row per captured launch).  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json [commit]"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
commit = sys.argv[3] if len(sys.argv) > 3 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hdr, units, drows = rr[0], rr[1], rr[2:]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("smsp__inst_executed.sum", "warp_insts"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"), ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"), ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"), ("launch__shared_mem_per_block_dynamic", "smem_dynamic"), ("launch__shared_mem_per_block_static", "smem_static")]
rows = []
for r in drows:
    d = {}
    for hname, k in want:
        if hname not in hdr: continue
        i = hdr.index(hname); v, u = r[i], units[i]
        if k in ("dram_read", "dram_write"): d[k + "_bytes"] = float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        elif k == "time": d["time_ms"] = float(v.replace(",", "")) * {"us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}.get(u, 1)
        elif k == "kernel": d[k] = v.split("(")[0].replace("void pt::", "").replace("pt::", "").replace("void ", "")
        else:
            try: d[k] = float(v.replace(",", ""))
            except ValueError: d[k] = v
    rows.append(d)
json.dump({"commit": commit, "source": rep, "rows": rows}, open(out, "w"), indent=1)
for d in rows: print("%-44s %9.3f ms  issue %5.1f%%  occ %5.1f%%  regs %3d  dram %5.1f%%  warp-inst %8.2f M" % (d["kernel"][:44], d["time_ms"], d.get("issue_active_pct", 0), d.get("occupancy_pct", 0), int(d.get("regs", 0)), d.get("dram_pct", 0), d.get("warp_insts", 0) / 1e6))
