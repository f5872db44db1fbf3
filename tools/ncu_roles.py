#!/usr/bin/env python
"""Per-role stall samples of a warp-specialised kernel from an ncu report (source page, SASS view).

  python tools/ncu_roles.py gpurun_out/prof_ConvPolicy_r1b.ncu-rep [kernel-index] [--top N]

Prints the hottest SASS lines with their stall reasons.  Samples per warp ~ total / warps-per-CTA, so a role whose lines
add up to about that much was busy (or spinning) for the whole kernel.
"""
import csv
import subprocess
import sys


def main():
    src = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else -1
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 5:
            cur["rows"].append(r)
    k = kern[which]
    h = k["hdr"]
    si, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    stall = [x for x in h if x.startswith("stall_") and "Not" not in x]
    tot = sum(int(r[si]) for r in k["rows"])
    print("kernel:", k["name"][:120], "| samples", tot)
    hot = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][si]))[:top]
    for i in sorted(hot):
        r = k["rows"][i]
        st = {x[6:]: int(r[h.index(x)]) for x in stall if int(r[h.index(x)]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5d %5.1f%% %9s  %-58s %s" % (i, 100.0 * int(r[si]) / tot, r[ie], r[1].strip()[:58], st))


if __name__ == "__main__":
    main()
