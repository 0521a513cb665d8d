#!/usr/bin/env python
"""Estimate FP64-pipe cycles of a kernel's hot loop from its SASS, with the register-file read model measured on
B200 by scripts/ubench/fp64_ubench.cu:
    cycles(FP64 warp-instruction) = max(2, #distinct 64-bit REGISTER source operands not served by the
                                         operand-reuse cache)          (3 distinct regs -> 3.06 cycles measured,
                                         3 regs with one .reuse-shared -> 2.23, <=2 regs -> 2.06)
usage: sass_fp64_cost.py <lib.so> <mangled-substring> [pairs_per_iteration]
Finds the innermost loop (largest backward branch body containing MUFU.RSQ64H) and reports cycles per pair."""
import re
import subprocess
import sys


def load(lib, key):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = key in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def fp64_cost(ins, prev_reuse):
    """ins: text of a DFMA/DMUL/DADD; prev_reuse: dict slot->reg kept by the previous instruction."""
    op, rest = ins.split(None, 1)
    ops = [o.strip() for o in rest.split(",")]
    srcs = ops[1:]
    regs, keep = [], {}
    for slot, o in enumerate(srcs):
        m = re.match(r"[-|]*R(\d+)(\.reuse)?", o)
        if not m:
            continue  # immediate / constant / RZ
        r = int(m.group(1))
        if m.group(2):
            keep[slot] = r
        if prev_reuse.get(slot) == r:
            continue  # served by the reuse cache
        regs.append(r)
    return max(2, len(set(regs))), keep


def main():
    lib, key = sys.argv[1], sys.argv[2]
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    code = load(lib, key)
    addr = {a: i for i, (a, _) in enumerate(code)}
    best = None
    for i, (a, t) in enumerate(code):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = code[addr[tgt]:i + 1]
                n_mufu = sum("MUFU.RSQ64H" in x for _, x in body)
                if n_mufu and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        sys.exit("no loop with MUFU.RSQ64H found")
    cyc, n64, n3, reuse = 0, 0, 0, {}
    others = 0
    for _, t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        if re.match(r"(DFMA|DMUL|DADD)\b", t):
            c, reuse = fp64_cost(t, reuse)
            cyc += c
            n64 += 1
            n3 += c >= 3
        else:
            others += 1
            reuse = {}  # conservative: another instruction in between drops the reuse entries
    print(f"loop: {len(best)} instr, {n64} FP64 ({n64 / pairs:.2f}/pair), {others} other; "
          f"{n3} FP64 instr cost 3 cycles; FP64-pipe cycles/pair = {cyc / pairs:.2f} "
          f"(floor {2 * n64 / pairs:.1f}); issue slots/pair = {len(best) / pairs:.1f}")


if __name__ == "__main__":
    main()
