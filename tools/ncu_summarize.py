#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brings back into the committed summaries under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv  profiles/r01_ncu_launch_shares.txt
  python tools/ncu_summarize.py full     gpurun_out/conv_tc.ncu-rep profiles/r01_ncu_conv_tc_layers.txt [labels,...]
  python tools/ncu_summarize.py traffic  gpurun_out/conv_tc.ncu-rep profiles/ncu_traffic.json conv_tc_row "<1, 1, 1" "<workload>"

`launches`: per-kernel totals/shares of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised:
compare shares, not absolutes).  `traffic`: average dram__bytes_read.sum + dram__bytes_write.sum per captured launch
of the kernels whose name contains the filter, stored under a key of profiles/ncu_traffic.json together with the
workload string of the capture — bench.py copies it into `roofline.traffic` when the workload matches.  `full`: one row per captured launch of an `ncu --set full` report with the
numbers the roofline keys of bench.py quote (DRAM bytes read+written = `traffic`, tensor-pipe activity, clocks).
"""
import csv
import collections
import io
import subprocess
import sys

FULL = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_elapsed_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3,
         "hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    k_name, k_metric, k_val, k_unit = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[k_metric] != "gpu__time_duration.sum":
            continue
        us = float(r[k_val].replace(",", "")) * SCALE.get(r[k_unit], 1.0) / (1.0 if r[k_unit] in ("us", "ns", "ms") else 1.0)
        name = r[k_name].split("(")[0]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    out = io.StringIO()
    out.write("(launch list of `ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised under"
              " ncu: compare SHARES, not absolutes)\n\n")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("%-60s n=%4d  total %10.1f us  avg %8.1f us  share %5.1f%%\n" % (name[:60], n, t, t / n, 100 * t / total))
    open(dst, "w").write(out.getvalue())
    sys.stdout.write(out.getvalue())


def full(src, dst, labels):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = io.StringIO()
    out.write("(one row per launch captured by `ncu --set full --clock-control none --import-source on`; report: %s)\n" % src)
    cols = [(hdr.index(m), short) for m, short in FULL if m in hdr]
    out.write("%-22s %-34s " % ("launch", "kernel") + " ".join("%12s" % s for _, s in cols) + "   traffic_MB\n")
    for i, r in enumerate(rows[2:]):
        vals = {}
        for c, short in cols:
            try:
                vals[short] = float(r[c].replace(",", "")) * SCALE.get(units[c], 1.0)
            except ValueError:
                vals[short] = float("nan")
        label = labels[i] if i < len(labels) else "#%d" % i
        kern = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        out.write("%-22s %-34s " % (label, kern[:34]) + " ".join("%12.3f" % vals[s] for _, s in cols)
                  + "   %10.1f\n" % (vals.get("dram_rd_MB", 0) + vals.get("dram_wr_MB", 0)))
    open(dst, "w").write(out.getvalue())
    sys.stdout.write(out.getvalue())


def traffic(src, dst, key, name_filter, workload):
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr, tm = (hdr.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    tot, us, n = 0.0, 0.0, 0
    for r in rows[2:]:
        if name_filter not in r[kn]:
            continue
        tot += (float(r[rd]) * SCALE[units[rd]] + float(r[wr]) * SCALE[units[wr]]) * 1e6
        us += float(r[tm]) * SCALE[units[tm]]
        n += 1
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data[key] = {"bytes_per_launch": tot / max(n, 1), "launches_captured": n, "avg_launch_us_under_ncu": us / max(n, 1),
                 "kernel_filter": name_filter, "workload": workload,
                 "source": "ncu --set full --clock-control none (%s)" % os.path.basename(src)}
    json.dump(data, open(dst, "w"), indent=1)
    print(key, data[key])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4].split(",") if len(sys.argv) > 4 else [])
