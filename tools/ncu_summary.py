"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_sfs.ncu-rep > profiles/rNN_sfs_full.txt"""
import csv, subprocess, sys, collections, io, re

def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    per = collections.OrderedDict()
    tot = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "")
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"): ns *= 1e3
        if r["Metric Unit"] in ("ms", "msecond"): ns *= 1e6
        d = per.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += ns; tot += ns
        rows.append((r["ID"], name, ns, r["Grid Size"], r["Block Size"]))
    print("# ncu --metrics gpu__time_duration.sum launch list (cold-cache, serialised: compare SHARES)")
    print("# kernel, launches, total_ms, share")
    for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %5d %12.3f %6.1f%%" % (k[:60], n, ns / 1e6, 100 * ns / tot))
    print("# total_ms %.3f over %d launches" % (tot / 1e6, len(rows)))
    print("# first launches:")
    for r in rows[:40]:
        print("%4s %-60s %12.3f ms grid=%s block=%s" % (r[0], r[1][:60], r[2] / 1e6, r[3], r[4]))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none, raw page excerpt of", path)
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("## kernel:", d.get("Kernel Name", ("", "?"))[1], "grid", d.get("Grid Size", ("", "?"))[1], "block", d.get("Block Size", ("", "?"))[1])
        for k in WANT:
            if k in d:
                print("%-80s %-16s %s" % (k, d[k][0], d[k][1]))
        for k in hdr:
            if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued") and k in d:
                print("%-80s %-16s %s" % (k, d[k][0], d[k][1]))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
