"""Static opcode mix of a kernel's main loop, from `cuobjdump -sass` of an object file or the library.

    python scripts/sass_loop_mix.py icar_b200/csrc/_obj/mpdata_fused.o k_corrector_fused_p2ILi40 [top_n]

Picks the largest backward branch as the loop, leaves out the blocks a forward branch jumps over when they
contain UTMALDG (the TMA issue block only thread 0 executes), and prints the instruction count of the rest --
for an unrolled loop divide by the unroll factor.  Used for DESIGN.md section 3.1 (issue-bound kernels: the
static count of the hot path predicts the time)."""
import re, sys, collections, subprocess
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name: continue
    ins = []
    for ln in b.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", ln)
        if m: ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    # backward branches = loops
    loops = []
    for a, op, rest in ins:
        if op.startswith("BRA") and "0x" in rest:
            t = int(re.findall(r"0x([0-9a-f]+)", rest)[-1], 16)
            if t < a: loops.append((t, a))
    big = max(loops, key=lambda l: l[1] - l[0])
    body = [(a, op, rest) for a, op, rest in ins if big[0] <= a <= big[1]]
    # drop the TMA issue block: from first UTMALDG-8 .. last UTMALDG+3 inside body, approximated by the guarded region
    ut = [a for a, op, _ in body if op.startswith("UTMALDG")]
    cnt = collections.Counter(op for a, op, _ in body)
    n_issue = 0
    skips = []
    for a, op, rest in body:
        if op.startswith("BRA") and "0x" in rest:
            t = int(re.findall(r"0x([0-9a-f]+)", rest)[-1], 16)
            if t > a and any(a < u < t for u in ut):
                skips.append((a + 16, t - 16))
    kept = [(a, op) for a, op, _ in body if not any(s0 <= a <= s1 for s0, s1 in skips)]
    n_issue = len(body) - len(kept)
    cnt = collections.Counter(op for a, op in kept)
    n = sum(cnt.values())
    print(name[-70:], "total insns", len(ins), "loop", hex(big[0]), hex(big[1]), "body", len(body), "issue block", n_issue, "hot", n)
    for op, c in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 14): print(f"   {op:26s} {c:4d} {100*c/n:5.1f}%")
