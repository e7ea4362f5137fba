"""Static facts of every kernel in the built objects: registers, stack frame, spill bytes (ptxas -v logs the Makefile
keeps under icar_b200/csrc/_obj/*.ptxas.log) and SASS instruction count (cuobjdump -sass).  No GPU needed.

    python scripts/static_kernel_table.py > profiles/<round>_static_kernel_table.txt

Used to compare candidates that have not run on hardware yet with the kernels they would replace; it says nothing
about time on its own (an issue-bound kernel follows its DYNAMIC instruction count)."""
import os
import re
import subprocess

OBJ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "icar_b200", "csrc", "_obj")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for d in out:
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void icar::", "", d)
        d = re.sub(r"\(.*$", "", d)
        short.append(d)
    return dict(zip(names, short))


def main():
    rows = []
    for log in sorted(f for f in os.listdir(OBJ) if f.endswith(".ptxas.log")):
        text = open(os.path.join(OBJ, log)).read()
        info = {}
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n.*?Used (\d+) registers", text):
            info[m.group(1)] = (int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)))
        obj = os.path.join(OBJ, log.replace(".ptxas.log", ".o"))
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        count, name = {}, None
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                name = m.group(1)
                count[name] = 0
            elif name and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
                count[name] += 1
        names = demangle(list(info))
        for k, (regs, stack, st, ld) in info.items():
            rows.append((log.replace(".ptxas.log", ""), names[k], regs, stack, st, ld, count.get(k, 0)))
    print(f"{'file':16s} {'kernel':64s} {'regs':>5s} {'stack':>6s} {'spill st':>8s} {'spill ld':>8s} {'SASS':>6s}")
    for r in sorted(rows):
        print(f"{r[0]:16s} {r[1][:64]:64s} {r[2]:5d} {r[3]:6d} {r[4]:8d} {r[5]:8d} {r[6]:6d}")


if __name__ == "__main__":
    main()
