#!/usr/bin/env python
"""Instruction mix of one kernel of a built library: tools/sass_mix.py <lib.so> <substring of mangled name>.

Prints the opcode histogram of the whole function and of its hottest (largest backward-branch) loop."""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0].split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        print("==", name, len(ins), "instructions")
        hist = collections.Counter(opcode(i) for _, i in ins)
        print("  all :", dict(hist.most_common(14)))
        # innermost hot loop = backward branch spanning the most FP64 instructions
        best = None
        for addr, i in ins:
            m = re.search(r"BRA\S*\s+(?:\S+\s+)?`\(\.L_x_\d+\)|BRA\S*.*?0x([0-9a-f]+)", i)
            m2 = re.search(r"0x([0-9a-f]+)", i) if opcode(i) == "BRA" else None
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < addr:
                    body = [x for a, x in ins if tgt <= a <= addr]
                    if best is None or len(body) > len(best):
                        best = body
        if best:
            h = collections.Counter(opcode(i) for i in best)
            print("  loop:", len(best), "instr", dict(h.most_common(14)))


if __name__ == "__main__":
    main()
