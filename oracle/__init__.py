"""CPU oracle for the QuadX hover / yaw env step.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or
as the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status (see DESIGN.md "Oracle"):
  * env layer (action scaling, obs packing, reward, termination, truncation,
    reset protocol, rectangle detector): PINNED -- checked bit-for-bit in
    float64 against the reference's own ``simulation/hover.py`` imported with
    stub modules (tests/golden/make_golden.py, tests/test_oracle_vs_reference.py).
  * PyFlyt 0.21.0 / pybullet 3.2.7 arithmetic underneath (motor model, PID,
    drag, Bullet integration, camera pose): **PARITY UNPINNED** -- neither
    package is vendored in /root/reference nor installable here; this is a
    restatement of their published algorithms with every uncertain choice
    exposed as a switch in ``QuadXParams``.
"""
