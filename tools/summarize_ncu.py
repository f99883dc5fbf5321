#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  launches <launches.csv> <out.txt>      per-kernel share of a `--metrics gpu__time_duration.sum` launch list
  full <report.ncu-rep> <out.txt>        key metrics per captured kernel from a `--set full` report
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[start + 2:]:
        if len(r) < len(hdr):
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])[:100]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache,\n"
                "# serialised: compare SHARES, not absolutes).  source: %s\n" % path)
        f.write("total %.1f us over %d launches\n" % (tot, sum(a[0] for a in agg.values())))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write("%6.2f%% %10.1f us %6d  %s\n" % (100 * t / tot, t, n, k))


def full(path, out):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode(errors="replace")
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# key metrics from `ncu --set full --clock-control none` (source: %s)\n" % path)
        for r in rows[2:]:
            f.write("---- %s\n" % r[idx["Kernel Name"]][:160])
            for k in KEYS:
                if k in idx:
                    f.write("  %-72s %s %s\n" % (k, r[idx[k]], units[idx[k]]))
            stalls = [(h, float(r[idx[h]].replace(",", "") or 0)) for h in hdr
                      if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
            for h, v in sorted(stalls, key=lambda t: -t[1])[:6]:
                f.write("  stall %-66s %.2f\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
