"""CPU: the compiled kernels stay inside their register / spill / size budgets (read from ptxas -v in csrc/build.log and
from the SASS of the built library).  A spill or a bloated loop in the issue-bound env-step kernel is a performance bug
that no parity test notices; this one does, without a GPU."""
import os
import re
import subprocess

import pytest


@pytest.fixture(scope="module")
def ptxas():
    import __graft_entry__ as ge

    ge.build()
    from fpv_drone_rl_agent_b200 import _lib

    log = os.path.join(os.path.dirname(_lib.LIB_PATH), "build.log")
    if not os.path.exists(log):  # library was already up to date and the log is not in the checkout: rebuild once
        import importlib.util

        spec = importlib.util.spec_from_file_location("qx_build", os.path.join(os.path.dirname(_lib.LIB_PATH), "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(force=True)
    text = open(log).read()
    info = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers", text, re.S):
        info[m.group(1)] = dict(stack=int(m.group(2)), spill_st=int(m.group(3)), spill_ld=int(m.group(4)), regs=int(m.group(5)))
    assert info, "no ptxas -v output in build.log"
    return info, _lib.LIB_PATH


def _find(info, *needles):
    hits = [k for k in info if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return info[hits[0]]


def test_env_step_kernels_do_not_spill(ptxas):
    info, _ = ptxas
    # quadx_step_kernel<MODE, TASK, CASC, REF>: mangled template arguments ILi<MODE>ELi<TASK>ELb<CASC>ELb<REF>E
    for mode in (1, 3):  # the step launch and the reset-queue launch of qx_step
        for ref in (0, 1):
            k = _find(info, f"quadx_step_kernelILi{mode}ELi0ELb0ELb{ref}E")
            assert k["spill_st"] == 0 and k["spill_ld"] == 0 and k["regs"] <= 128, (mode, ref, k)  # 128 registers = 4 blocks of 128 threads per SM
    yaw = _find(info, "quadx_step_kernelILi1ELi1ELb0ELb0E")
    assert yaw["spill_st"] == 0 and yaw["regs"] <= 128


def test_policy_kernel_fits_three_slots(ptxas):
    info, _ = ptxas
    k = _find(info, "policy_forward_kernel")
    assert k["spill_st"] == 0 and k["spill_ld"] == 0, k
    assert k["regs"] <= 80, k  # 768 threads per CTA: 65 536 / 768 = 85 registers, allocated in units of 8


def test_reference_constant_kernel_is_the_smaller_one(ptxas):
    """Static SASS size of the hover step kernel: the reference-constant instantiation must stay clearly below the generic
    one (it was 2 432 vs 2 736 instructions when it was introduced; 3 328 vs 3 800 since -fmad=false spells every product-sum of
    the once-per-step epilogue as two instructions and the raster-vision call was added) -- if it does not, the literals stopped
    folding."""
    _, lib = ptxas
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=300).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    counts, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            counts[cur] = 0
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            counts[cur] += 1
    gen = [v for k, v in counts.items() if "quadx_step_kernelILi1ELi0ELb0ELb0E" in k][0]
    ref = [v for k, v in counts.items() if "quadx_step_kernelILi1ELi0ELb0ELb1E" in k][0]
    assert ref <= 3500 and gen <= 4000 and ref <= 0.92 * gen, (ref, gen)
