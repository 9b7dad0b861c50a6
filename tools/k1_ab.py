#!/usr/bin/env python
"""A/B of tuning builds of the library on one GPU box: runs tools/perstep.py once per library (QX_LIB) and prints the median
step-launch time of steps 3..14 (nothing finishes there) per workload, twice, interleaved, so that box drift shows.

    QX_LIB_OUT=.../libquadx_b200_x.so QX_NVCC_EXTRA=-DQX_... python fpv-drone-rl-agent_b200/csrc/build.py --force   (here)
    python tools/k1_ab.py fpv-drone-rl-agent_b200/csrc/libquadx_b200_*.so                                         (GPU box)"""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:]
for rep in range(2):
    for lib in libs:
        env = dict(os.environ, QX_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "perstep.py"), "1048576", "30"], env=env, capture_output=True, text=True).stdout
        row = {}
        for line in out.splitlines():
            name, _, js = line.partition(" ")
            if js.startswith("{"):
                steps = json.loads(js)["step_us|reset_us|n_done per step"]
                row[name] = (round(statistics.median(s[0] for s in steps[3:15]), 1), round(statistics.median(s[0] + s[1] for s in steps[16:30]), 1))
        print(os.path.basename(lib), rep, json.dumps(row), flush=True)
