#!/usr/bin/env python
"""Opcode mix of the sub-step loop (the backward-branch loop of 300-700 instructions with the most FFMA) of the kernels whose
mangled name matches a regex.  Usage: python tools/sass_loop.py 'hot_kernelILb1ELi4E(NS_2S1E|f)Lb1'"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc", "libquadx_b200.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "quadx_step")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); fn[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and cur:
        fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
for k, v in fn.items():
    if not pat.search(k):
        continue
    best = None
    want_packed = any("FFMA2" in i for _, i in v)
    for addr, ins in v:
        m = re.search(r"BRA(?:\.\w+)* (?:\S+, )?`?\(?(0x[0-9a-f]+)\)?", ins)
        if m and int(m.group(1), 16) < addr:
            a = int(m.group(1), 16)
            body = [i for ad, i in v if a <= ad <= addr]
            if 300 <= len(body) <= 700:
                c = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]).split(".")[0] for i in body)
                if want_packed and c["FFMA2"] == 0:
                    continue
                if best is None or c["FFMA"] + 2 * c["FFMA2"] > best[1]["FFMA"] + 2 * best[1]["FFMA2"]:
                    best = (len(body), c, sum(1 for i in body if "LDL" in i or "STL" in i))
    if best:
        print(k)
        print("  loop instructions:", best[0], " local-memory ops:", best[2])
        print("  ", ", ".join(f"{o} {n}" for o, n in best[1].most_common(28)))
