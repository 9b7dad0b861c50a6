#!/usr/bin/env python
"""Warp-stall samples of one kernel of an ncu report, summed per stall reason, and the most-sampled SASS instructions with
their top reasons (`ncu -i <rep> --page source --csv --print-source sass --kernel-name regex:<name>`).  CPU only.

    python tools/ncu_stall_samples.py gpurun_out/k1_r2k.ncu-rep quadx_step_hot [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, name = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{name}"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    b = blocks[0]
    ix = {k: i for i, k in enumerate(b["hdr"])}
    stalls = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    tot = collections.Counter()
    for r in b["rows"]:
        for k in stalls:
            try:
                tot[k] += int(r[ix[k]])
            except (ValueError, IndexError):
                pass
    s = sum(tot.values())
    print(f"{b['name']}\n{len(b['rows'])} SASS instructions, {s} warp samples (first captured launch)\n")
    print("| stall reason | samples | share |\n|---|---|---|")
    for k, v in tot.most_common(12):
        print(f"| {k} | {v} | {100 * v / s:.1f} % |")
    rows = []
    for i, r in enumerate(b["rows"]):
        try:
            rows.append((int(r[ix["# Samples"]]), i, r))
        except (ValueError, IndexError):
            pass
    rows.sort(key=lambda t: -t[0])
    print("\n| samples | # | SASS | executed | top reasons |\n|---|---|---|---|---|")
    for n, i, r in rows[:top]:
        st = sorted(((k, int(r[ix[k]] or 0)) for k in stalls), key=lambda x: -x[1])[:2]
        print(f"| {n} | {i} | `{r[ix['Source']].strip()[:60]}` | {r[ix['Instructions Executed']]} | {', '.join(f'{k[6:]} {v}' for k, v in st)} |")


if __name__ == "__main__":
    main()
