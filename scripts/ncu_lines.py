"""Hottest source lines of one kernel of an `ncu --set full --import-source on` report.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_id [top]"""
import collections, csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; per = collections.Counter(); samp = collections.Counter(); src = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isx = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (cur, int(r[0])); per[key] += int(r[ie] or 0); samp[key] += int(r[isx] or 0); src[key] = r[1][:120]
tot = sum(per.values()); ts = sum(samp.values())
print("total warp-inst", tot, "samples", ts)
it = sorted(((samp[k], per[k], k) for k in per), reverse=True)
for s, n, k in it[:top]:
    print(f"{s / ts * 100:5.2f}% smp {n / tot * 100:5.2f}% inst  {k[0]}:{k[1]}: {src[k]}")
