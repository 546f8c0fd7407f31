#!/usr/bin/env python
"""Lists the innermost data-pass loops of a kernel from `cuobjdump -sass` output: instruction mix, FP64 count,
sum of the static stall counts, local-memory traffic, and the loop body itself. This is where the per-row
instruction counts quoted in DESIGN.md / abn_b200.cu (fp64_per_row) come from.
usage: cuobjdump -sass -fun <mangled name> <obj> | tools/sass_loops.py [--body]"""
import re
import sys
from collections import Counter


def main():
    body = "--body" in sys.argv
    lines = sys.stdin.read().split("\n")
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF))
                i += 2
                continue
        i += 1
    at = {a: k for k, (a, _, _) in enumerate(ins)}
    loops = []
    for k, (a, t, _) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in at and k - at[tgt] < 600:
                loops.append((at[tgt], k))
    # innermost only
    loops = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    for lo, hi in loops:
        sel = ins[lo:hi + 1]
        ops = [(t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t, _ in sel]
        c = Counter(ops)
        fp64 = sum(c[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        wide = sum(1 for _, t, _ in sel if ".128" in t)
        if wide < 2 or fp64 < 16:
            continue
        local = sum(1 for _, t, _ in sel if "STL" in t or "LDL" in t)
        print(f"loop {sel[0][0]:05x}-{sel[-1][0]:05x}: {len(sel)} instructions, {fp64} FP64, {len(sel) - fp64} other, "
              f"static stall sum {sum(s for _, _, s in sel)}, local ld/st {local}")
        print("   " + ", ".join(f"{o} {n}" for o, n in c.most_common()))
        if body:
            for a, t, s in sel:
                print(f"   {a:05x} st={s:2d}  {t}")


if __name__ == "__main__":
    main()
