#!/usr/bin/env python
"""Where a kernel's executed instructions come from: joins the per-SASS-instruction execution counts of an ncu report
(`ncu -i <rep> --page source --csv --print-source sass`) with the line table of the same kernel in the built library
(`nvdisasm -g`), by instruction index.  Prints (a) the split into instructions every warp executes once per launch, N times
(the sub-step loop) and under divergence, (b) the once-per-launch part grouped by function, (c) the top source lines.

    python tools/ncu_source_breakdown.py gpurun_out/k1_r2i.ncu-rep '_ZN2qx21quadx_step_hot_kernelILb1ELi4ENS_2S1ELb0EEE'

CPU only (reads the report and the .so).  The library must be the build the report was taken from."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc", "libquadx_b200.so")
CSRC = os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc")
GROUPS = [  # (file, first line pattern, last line pattern, label) resolved against the current sources
    ("qx_model.cuh", "QX_DI float fast_atan2f", "^// Analytic camera", "euler <-> quaternion (atan2 / asin / sincos)"),
    ("qx_model.cuh", "void vision\\(const Env", "^// vision_mode = 1", "vision()"),
    ("qx_model.cuh", "QX_DI float4 ldg4", "^QX_DI void pack_core\\(", "state load / store"),
    ("qx_kernels.cu", "void build_obs", "^// ---- compute_term_trunc_reward", "build_obs"),
    ("qx_kernels.cu", "bool reward_and_flags", "^// ---- SB3 VecEnv auto-reset", "reward_and_flags"),
    ("qx_kernels.cu", "void write_obs", "^// MODE_STEP_INLINE", "write_obs"),
]


def line_ranges():
    out = []
    for f, a, b, label in GROUPS:
        src = open(os.path.join(CSRC, f)).read().split("\n")
        la = next((i + 1 for i, l in enumerate(src) if re.search(a, l)), None)
        lb = next((i + 1 for i, l in enumerate(src) if la and i + 1 > la and re.search(b, l)), None)
        if la and lb:
            out.append((f, la, lb, label))
    return out


def main():
    rep, mangled = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hdr, R = rows[1], []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        R.append(r)
    iI, iT = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, capture_output=True)
        cubin = os.path.join(d, "qx_kernels.sm_100a.cubin")
        elf = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
        idx = next(l.split()[0] for l in elf.split("\n") if l.split() and l.split()[-1] == mangled and l.split()[0].startswith("0x"))
        dis = subprocess.run(["nvdisasm", "-g", "-fun", idx, cubin], capture_output=True, text=True, errors="ignore").stdout
    ins, cur = [], None
    for line in dis.split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+.*?;", line):
            ins.append(cur)
    n = min(len(R), len(ins))
    counts = [int(R[k][iI]) for k in range(n)]
    tot = sum(counts)
    from collections import Counter

    freq = Counter(c for c in counts if c > 0)
    base = min(c for c, k in freq.items() if k >= 200)  # warps of the launch = executions of a once-per-launch instruction
    print(f"{n} SASS instructions, {tot} warp-instructions executed, {base} warps")
    cls = defaultdict(lambda: [0, 0])
    for c in counts:
        k = "never" if c == 0 else (f"{c // base}x every warp" if c % base == 0 else "divergent / partial")
        cls[k][0] += 1
        cls[k][1] += c
    for k, v in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:22s} static {v[0]:5d}   executed {100 * v[1] / tot:5.1f} %   = {v[1] / base:7.1f} per warp")
    rng = line_ranges()
    grp = defaultdict(int)
    for k in range(n):
        if counts[k] == base and ins[k]:
            f, l = ins[k]
            grp[next((lab for ff, a, b, lab in rng if ff == f and a <= l < b), f)] += 1
    print("once-per-launch instructions by function:")
    for k, v in sorted(grp.items(), key=lambda kv: -kv[1]):
        print(f"  {v:5d}  {k}")
    by = defaultdict(int)
    for k in range(n):
        by[ins[k]] += counts[k]
    print("top source lines (executed warp-instructions per warp):")
    for fl, v in sorted(by.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {v / base:7.1f}  {fl[0]}:{fl[1]}" if fl else f"  {v / base:7.1f}  ?")


if __name__ == "__main__":
    main()
