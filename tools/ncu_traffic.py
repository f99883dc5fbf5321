#!/usr/bin/env python
"""ncu `--set full --page raw --csv` export -> (1) profiles/roofline_traffic.json: measured DRAM traffic per launch
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of each kernel family, keyed by
the C-ABI entry point bench.py accounts the kernel under) and (2) a per-launch text summary.

    python tools/ncu_traffic.py gpurun_out/full_raw.csv profiles/roofline_traffic.json profiles/rNN_ncu_full.txt
"""
import csv
import json
import re
import sys

FAMILY = [("stem_fwd_rows_kernel<(bool)1>", "i2v_conv_stem_fwd_pool_f32"), ("stem_fwd_rows_kernel", "i2v_conv_stem_fwd_f32"), ("stem_dgrad_pool_kernel", "i2v_conv_stem_dgrad_pool_f32"),
          ("stem_dgrad_direct_kernel", "i2v_conv_stem_dgrad_f32"),
          ("conv_tc_persist_kernel", "i2v_conv_tc_f32"), ("conv_tc_pair_kernel", "i2v_conv_tc_f32"),
          ("conv3x3_halo_kernel", "i2v_conv_tc_f32"), ("conv_tc_kernel", "i2v_conv_tc_f32"),
          # (the first-layer entry points are composites — im2col / col2im pass + a conv_tc_persist GEMM whose launches
          # cannot be told apart from the other convolutions' by name — so they get no per-launch traffic figure)
          ("maxpool_fwd", "i2v_maxpool_fwd_f32"), ("maxpool_bwd", "i2v_maxpool_bwd_f32"),
          ("cosine_loss_grad", "i2v_cosine_loss_grad_f32"), ("adam_compose", "i2v_adam_compose_table_f32")]
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    src, out_json, out_txt = sys.argv[1:4]
    rows = list(csv.reader(open(src, errors="replace")))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    cols = [c for c in COLS if c in ix]
    fam = {}
    lines = ["# ncu --set full --clock-control none; per launch: " + ", ".join(c.split(".")[0] for c in cols),
             "# time in us, dram bytes in MB; source: " + src]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("i2v::", "")
        vals = {}
        for c in cols:
            try:
                v = float(r[ix[c]].replace(",", ""))
            except ValueError:
                continue
            vals[c] = v * SCALE.get(units[ix[c]], 1.0)
        lines.append("%-44s " % name[:44] + " ".join(
            "%9.1f" % (vals[c] / 1e6 if "bytes" in c else vals[c]) if c in vals else "        -" for c in cols))
        for key, entry in FAMILY:
            if key in name:
                f = fam.setdefault(entry, {"launches": 0, "bytes": 0.0, "us": 0.0})
                f["launches"] += 1
                f["bytes"] += vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)
                f["us"] += vals.get("gpu__time_duration.sum", 0.0)
                break
    traffic = {k: v["bytes"] / v["launches"] for k, v in fam.items()}
    json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)
    lines.append("# average DRAM traffic per launch (-> %s):" % out_json)
    for k, v in sorted(fam.items()):
        lines.append("#   %-32s %4d launches  %8.1f MB/launch  %7.1f us/launch" % (k, v["launches"], v["bytes"] / v["launches"] / 1e6,
                                                                                v["us"] / v["launches"]))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-len(fam) - 1:]))


if __name__ == "__main__":
    main()
