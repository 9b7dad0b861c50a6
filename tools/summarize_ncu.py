#!/usr/bin/env python
"""Summarise an ncu capture (gpurun_out/*.ncu-rep + launches csv) into profiles/<tag>.md and profiles/k1_traffic.json.
Usage: python tools/summarize_ncu.py <tag> [envs]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
prefix = sys.argv[3] if len(sys.argv) > 3 else "k1"
rep = os.path.join(ROOT, "gpurun_out", f"{prefix}_{tag}.ncu-rep")
launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv") if prefix == "k1" else "/nonexistent"
SEL = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
out = [f"# ncu summary `{tag}` ({envs} envs)\n"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
seen = set()
traffic = None
def _dur(r):
    try:
        return float(r[hdr.index("gpu__time_duration.sum")])
    except ValueError:
        return 0.0
for r in sorted(rows[2:], key=lambda r: -_dur(r)):  # the longest instance of every kernel
    name = r[hdr.index("Kernel Name")]
    if name in seen:
        continue
    seen.add(name)
    out.append(f"\n## `{name}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
    out.append("| metric | value | unit |\n|---|---|---|")
    vals = {}
    for k in SEL:
        if k in hdr:
            vals[k] = r[hdr.index(k)]
            out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
    if "dram__bytes_read.sum" in vals and traffic is None and ("quadx_step_kernel<1" in name or "quadx_step_hot_kernel" in name or "policy_forward" in name):
        def tobytes(v, u):
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = tobytes(vals["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = tobytes(vals["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
        traffic = {"kernel": name, "envs": envs, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "per_env_step": (rd + wr) / envs, "source": f"ncu --set full, profiles/{prefix}_{tag}.md"}
        if "smsp__inst_executed.sum" in vals:  # the compute side of the ridge: what actually bounds the hover step
            traffic["warp_instructions_per_launch"] = float(vals["smsp__inst_executed.sum"])
            traffic["thread_instructions_per_env_step"] = float(vals["smsp__inst_executed.sum"]) * 32 / envs
            traffic["issue_active_pct"] = float(vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan"))
            traffic["kernel_us_under_ncu"] = float(vals["gpu__time_duration.sum"])
if os.path.exists(launches):
    out.append("\n## launch list (gpu__time_duration, cold-cache, serialised)\n")
    tot = {}
    for line in csv.reader(open(launches)):
        if len(line) > 5 and line[-1].replace(".", "").isdigit() and "gpu__time_duration" in line[-3]:
            tot.setdefault(line[4][:90], []).append(float(line[-1]))
    s = sum(sum(v) for v in tot.values())
    out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / s:.1f} % |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{prefix}_{tag}.md"), "w").write("\n".join(out) + "\n")
if traffic and "quadx" in traffic["kernel"]:
    import hashlib

    hsh = hashlib.sha256()
    for f in ("qx_kernels.cu", "qx_model.cuh", "qx_lanes.cuh", "qx_ref_constants.cuh"):  # bench.py compares this with the sources it runs
        hsh.update(open(os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc", f), "rb").read())
    traffic["kernel_sources_sha16"] = hsh.hexdigest()[:16]
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w"), indent=1)
print("\n".join(out[:60]))
