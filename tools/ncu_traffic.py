#!/usr/bin/env python
"""profiles/r02_traffic.json from an .ncu-rep of tools/prof_ogs.py 64 (run here, on the CPU box):
   python tools/ncu_traffic.py gpurun_out/prof_ogs_r2.ncu-rep
dram__bytes_read.sum + dram__bytes_write.sum per launch of the gather and gather+scatter tile kernels — what bench.py
reports as roofline.traffic (it is a property of the kernel + mesh, measured under ncu, not of a timed run)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = {"source": os.path.basename(rep), "how": "ncu --set full --clock-control none, one launch each, 64^3 degree-7 hex box"}
    for r in data:
        name = r[hdr.index("Kernel Name")]
        def val(k):
            i = hdr.index(k)
            return float(r[i].replace(",", "")) * UNIT[units[i]]
        total = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        if "ogs_tile_pf_kernel" in name:  # template arguments: T, OP, MODE (0 gather, 2 gather+scatter), ...
            mode = name.split("<")[1].split(",")[2].strip()
            res["bs8_gather_scatter_64_p7" if mode == "2" else "bs6_gather_64_p7"] = int(total)
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(res)


main()
