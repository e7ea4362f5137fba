"""Which kernels of this tree's objects have the same SASS as those of another build of the library?

    git archive <commit> icar_b200/csrc include | tar -x -C /tmp/val && make -C /tmp/val/icar_b200/csrc -j8
    python scripts/sass_same_as.py /tmp/val/icar_b200/csrc/_obj

Used when kernels are added or code is moved after the last run on hardware: a kernel whose SASS is identical
to the one that passed the GPU tests needs no new validation (addresses and the per-file hash in the names of
anonymous-namespace kernels are normalised; everything else, encodings included, must match)."""
import os
import re
import subprocess
import sys

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "icar_b200", "csrc", "_obj")


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out = re.sub(r"_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_0-9]+_cu_[0-9a-f]+", "ANON", out)
    found, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            found[cur] = []
        elif cur and ln.strip():
            found[cur].append(" ".join(ln.split()))
    return found


def main(other):
    bad = 0
    for name in sorted(f for f in os.listdir(other) if f.endswith(".o")):
        mine = os.path.join(HERE, name)
        if not os.path.exists(mine):
            print(f"{name}: not built here")
            continue
        a, b = kernels(os.path.join(other, name)), kernels(mine)
        changed = [k for k in a if a[k] != b.get(k)]
        print(f"{name}: {len(a)} kernels there, {len(b)} here, {len(a) - len(changed)} identical, "
              f"{len(b) - len([k for k in b if k in a])} new")
        for k in changed:
            print("   CHANGED", k)
        bad += len(changed)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
