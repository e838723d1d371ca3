"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and the full .ncu-rep captures into the
small text summaries committed under profiles/.   python tools/summarize_profiles.py <tag>"""
import csv, os, re, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
dst = os.path.join(os.path.dirname(src), "profiles")
os.makedirs(dst, exist_ok=True)
# ---- launch list ----
lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp, errors="replace") if not l.startswith("==")) if r]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi: continue
        name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void ", "", name)
        v = float(r[vi].replace(",", "")); u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
    allus = sum(t[1] for t in tot.values())
    with open(os.path.join(dst, f"{tag}_launches_32768x4096.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python tools/prof_one.py   (one qr! of 32768x4096, nb=128)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / allus:6.1f}%\n")
        f.write(f"{'TOTAL':60s} {sum(t[0] for t in tot.values()):8d} {allus:12.1f}\n")
    print(open(os.path.join(dst, f"{tag}_launches_32768x4096.txt")).read())
# ---- full captures ----
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    out = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    with open(os.path.join(dst, f"{tag}_{rep.replace('.ncu-rep', '')}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on (one launch per row); source report: gpurun_out/{rep}\n")
        for vals in rows[2:]:
            ni = hdr.index("Kernel Name")
            f.write(f"kernel: {vals[ni][:150]}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w); f.write(f"  {w:90s} {vals[i]:>18s} {units[i]}\n")
    print("wrote", f"{tag}_{rep.replace('.ncu-rep', '')}_ncu.txt")
