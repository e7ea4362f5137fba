"""profiles/ncu_kernels.json from `ncu --set full` reports: per kernel symbol the DRAM traffic and the executed
warp instructions of ONE launch, with the capture it came from.  bench.py reads it for `roofline.traffic` (only when the
symbol it timed is the one captured, on the same grid) and `alu_ceiling`.

    python scripts/ncu_kernels_json.py CELLS report1.ncu-rep [report2.ncu-rep ...] > profiles/ncu_kernels.json

CELLS = cells of the tile the captured step ran on (512 * 40 * 512 = 10485760 for the benchmark workload)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

cells = int(sys.argv[1])
out = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    col = {n: i for i, n in enumerate(h)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        sym = re.sub(r"\(.*", "", name)
        sym = re.sub(r"^void ", "", sym)
        sym = re.sub(r"^.*::", "", sym).strip()
        f = lambda k: float(r[col[k]].replace(",", "")) if k in col and r[col[k]] not in ("", "n/a") else None
        rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
        unit_r, unit_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        inst = f("smsp__inst_executed.sum")
        dur = f("gpu__time_duration.sum")
        dur_unit = rows[1][col["gpu__time_duration.sum"]]
        out[sym] = {
            "capture": os.path.basename(rep),
            # the committed text summary of that capture (scripts/ncu_summary.py): <tag>_top_<what>.ncu-rep ->
            # profiles/<tag>_top_kernels_<what>.txt
            "summary": "profiles/" + re.sub(r"_top_(\w+?)(_v\d+)?\.ncu-rep$", r"_top_kernels_\1.txt", os.path.basename(rep)),
            "grid_size": r[col["Grid Size"]] if "Grid Size" in col else None,
            "block_size": r[col["Block Size"]] if "Block Size" in col else None,
            "tile_cells": cells,
            "dram_bytes_per_launch": rd * scale[unit_r] + wr * scale[unit_w],
            "warp_instructions_per_launch": inst,
            "lane_instructions_per_cell": inst * 32.0 / cells,
            "duration_us_under_ncu": dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(dur_unit, 1.0),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        }
print(json.dumps(out, indent=1, sort_keys=True))
