"""The op-counting build of the oracle (algorithmic FP64 operations per trajectory).

`oracle/rko_count.cpp` compiles the same restatement of rsruckig with `double` replaced by a counting scalar. The count is
only meaningful if that build takes exactly the branches of the plain one: the tool asserts status and duration bit-equal.
CPU only.
"""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("count_algorithmic_ops", os.path.join(ROOT, "tools", "count_algorithmic_ops.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_counting_build_follows_the_plain_oracle_and_counts_are_stable():
    from rsruckig_b200 import workloads
    m = _tool()
    n = 4096
    per, status = m.count(workloads.bench_inputs(n, 7, seed=3), n, 7)  # asserts status + duration bit-equal inside
    assert (status >= 0).mean() > 0.9
    assert abs(per["total"] - (per["add_sub"] + per["mul"] + per["div"] + per["sqrt"] + per["libm"])) < 1e-6
    assert abs(per["total"] - sum(per["by_phase"].values())) < 0.1  # every operation is attributed to exactly one phase
    assert per["by_phase"]["shrink_interval (all callers)"] > 0.15 * per["total"]  # the Newton bracketing of Step-2 VEL is the largest part
    # the committed figure bench.py uses was produced by the same tool on 262144 trajectories: a 4096 sample lands within 5 %
    ref = json.load(open(os.path.join(ROOT, "profiles", "algorithmic_ops_oracle.json")))
    assert abs(per["total"] / ref["flop_per_trajectory_7dof"] - 1.0) < 0.05
    # a single-DoF problem has no Step 2: far fewer operations than one seventh of a 7-DoF trajectory's Step 1 + Step 2
    per1, _ = m.count(workloads.bench_inputs(n, 1, seed=3), n, 1)
    assert per1["total"] < per["total"] / 7
    assert per1["by_phase"]["shrink_interval (all callers)"] == 0 and per1["by_phase"]["step2 VEL (without Newton and check)"] == 0


def test_velocity_interface_is_cheap():
    """configs[2]: velocity interface + Phase synchronisation needs a small fraction of the position interface's work."""
    from rsruckig_b200 import _cabi as cabi
    from rsruckig_b200 import workloads
    m = _tool()
    n = 2048
    per, _ = m.count(workloads.config3_inputs(n, seed=1), n, 3, control_interface=cabi.RK_VELOCITY, synchronization=cabi.RK_SYNC_PHASE)
    assert 0 < per["total"] < 5000
