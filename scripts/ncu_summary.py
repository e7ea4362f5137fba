"""Text summary of one kernel from an `ncu --set full` report: key raw metrics + SASS opcode mix.
usage: python scripts/ncu_summary.py report.ncu-rep cells_processed > profiles/xxx.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def source_mix(rep, kernel_id, cells):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kernel_id}"],
                         capture_output=True, text=True).stdout
    src = list(csv.reader(io.StringIO(out)))
    hi = next((n for n, r in enumerate(src) if "Source" in r and "Instructions Executed" in r), None)
    if hi is None:
        return
    h = src[hi]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    ops = collections.Counter()
    for r in src[hi + 1:]:
        if len(r) <= ie or not (r[ie] or "0").isdigit():
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
        if m:
            ops[m.group(2).split(".")[0]] += int(r[ie] or 0)
    tot = sum(ops.values())
    if not tot:
        return
    print("opcode mix:")
    for op, n in ops.most_common(16):
        print(f"  {op:10s} {n / tot * 100:5.1f}%")
    tma = [op for op in ops if op.startswith("UTMA") or op.startswith("SYNCS")]
    if tma:
        print("TMA / mbarrier opcodes present:", ", ".join(f"{t} x{ops[t]}" for t in tma))


def main():
    rep = sys.argv[1]
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    print(f"# {rep}")
    for n, vals in enumerate(raw[2:], 1):
        d = dict(zip(hdr, zip(vals, units)))
        print(f"\n== launch {n}: {d.get('Kernel Name', ('?', ''))[0][:100]}")
        for k in KEYS:
            if k in d:
                print(f"{k:80s} {d[k][0]:>16s} {d[k][1]}")
        stalls = sorted(((float(v[0]), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")),
                        reverse=True)
        print("top warp stall reasons (warps per issue-active cycle):")
        for v, k in stalls[:6]:
            print(f"  {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:6.2f}")
        if cells and "smsp__inst_executed.sum" in d:
            inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
            print(f"warp instructions executed: {inst:.0f}   lane-instructions per cell: {inst * 32 / cells:.1f}")
        source_mix(rep, n, cells)


if __name__ == "__main__":
    main()
