"""Turn ncu output (read in the build container, no GPU needed) into small text files for profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv  profiles/<name>.csv
    python tools/summarize_ncu.py full     gpurun_out/<tag>_prof.ncu-rep  profiles/<name>.csv
"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def short(name):
    name = name.replace("void ", "").replace("ccqp::", "")
    return name if len(name) < 90 else name[:87] + "..."


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    with open(dst, "w") as f:
        f.write("# per-launch device time from: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised)\n")
        f.write("id,kernel,grid,block,ms\n")
        for r in rows:
            ms = float(r[14]) / 1e6
            k = short(r[4])
            f.write("%s,\"%s\",\"%s\",\"%s\",%.4f\n" % (r[0], k, r[8], r[7], ms))
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1; a[1] += ms
        tot = sum(a[1] for a in agg.values())
        f.write("# ---- share of all kernel time in the command (includes problem generation by torch)\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("# %6.2f%%  %10.3f ms  x%-4d %s\n" % (100 * ms / tot, ms, c, k))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w") as f:
        f.write("# from ncu --set full --clock-control none; one column per captured launch\n")
        f.write("metric,unit," + ",".join("\"%s %s\"" % (r[0], short(r[4])) for r in data) + "\n")
        for i in idx:
            f.write("%s,%s,%s\n" % (hdr[i], units[i], ",".join(r[i] for r in data)))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
