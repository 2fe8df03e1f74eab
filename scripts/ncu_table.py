"""Per-kernel table out of an `ncu --set full` report: `python scripts/ncu_table.py report.ncu-rep > profiles/xxx.txt`.
First launch of every kernel name: duration, DRAM bytes and throughput, SM / issue utilisation, occupancy, registers, instructions."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
want = {"gpu__time_duration.sum": "dur_us", "dram__bytes_read.sum": "dram_rd_MB", "dram__bytes_write.sum": "dram_wr_MB", "dram__bytes.sum.per_second": "dram_GBps",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
        "launch__registers_per_thread": "regs", "smsp__inst_executed.sum": "warp_inst", "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conflicts",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio": "stall_barrier", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb"}
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
units = rows[1]
ki = h.index("Kernel Name")
cols = {m: h.index(m) for m in want if m in h}
seen = collections.OrderedDict()
for r in rows[2:]:
    if len(r) <= ki: continue
    name = r[ki].split("(")[0].replace("mrslam::", "")
    a = seen.setdefault(name, {"n": 0})
    a["n"] += 1
    if a["n"] > 1: continue
    for m, c in cols.items():
        try: v = float(r[c].replace(",", ""))
        except ValueError: continue
        u = units[c]
        if m == "gpu__time_duration.sum": v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        if m == "dram__bytes.sum.per_second": v = {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}.get(u, 1e-9) * v
        elif m.startswith("dram__bytes"): v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * v
        a[want[m]] = v
keys = ["dur_us", "dram_GBps", "dram_pct", "sm_pct", "issue_pct", "occ_pct", "fp64_pct", "regs", "warp_inst", "grid", "block", "n"]
print("%-34s" % "kernel" + "".join("%12s" % k for k in keys))
for name, a in seen.items():
    print("%-34s" % name[:34] + "".join(("%12.4g" % a[k]) if isinstance(a.get(k), float) else ("%12s" % a.get(k, "-")) for k in keys))
