#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
  python tools/ncu_summary.py full gpurun_out/prof_conv_r1.ncu-rep profiles/conv_r1.md
"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
           "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("b200asr::", "").replace("tc::", "")
        t = float(r[vi].replace(",", ""))
        t = t / 1e3 if r[ui] == "ns" else (t * 1e3 if r[ui] == "ms" else t)   # -> microseconds
        n, s = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, s + t)
        total += t
    with open(dst, "w") as f:
        f.write("# per-launch device times (ncu --metrics gpu__time_duration.sum --clock-control none), one training step\n")
        f.write("# cold-cache, serialised: compare SHARES with bench.py's event timings, not absolutes\n")
        f.write("# source: %s ; %d launches, %.2f ms total\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (src, len(rows) - 1, total / 1e3))
        for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (name, n, s, 100 * s / total))
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none capture: %s\n\n" % src)
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[hdr.index("Kernel Name")][:110])
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
