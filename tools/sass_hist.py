#!/usr/bin/env python
"""SASS opcode histogram per kernel of libquadx_b200.so (cuobjdump -sass), for profiles/ and quick checks.
Usage: python tools/sass_hist.py [substring-of-kernel-name ...] [--loop]   (--loop: only the hottest backward-branch loop body)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc", "libquadx_b200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, fn = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fn[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return fn


def opcode(ins):
    t = ins.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def largest_loop(code):
    """(start, end) addresses of the backward branch spanning the most instructions that contains no other backward branch target
    outside itself -- good enough to find the sub-step loop."""
    best = None
    for addr, ins in code:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    return best


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    loop = "--loop" in sys.argv
    for name, code in functions().items():
        if args and not any(a in name for a in args):
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        sel = code
        note = ""
        if loop:
            lp = largest_loop(code)
            if lp:
                sel = [(a, i) for a, i in code if lp[0] <= a <= lp[1]]
                note = f" loop 0x{lp[0]:x}..0x{lp[1]:x}"
        h = collections.Counter(opcode(i) for _, i in sel)
        print(f"## {dem}\n{len(sel)} instructions{note}")
        print("  " + "  ".join(f"{k} {v}" for k, v in h.most_common(40)))
