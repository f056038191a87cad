"""Host-side checks of bench.py (no GPU): every workload describes a valid problem, the attended-pair counts the
FLOP metric is built on agree between the closed forms, the oracle and the library, the strong-scaling split keeps
whole heads together for the channel-last workload, and the reference arm never loads the product library."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import bench
from tf_flash_attention_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", sorted(n for n, w in bench.WORKLOADS.items() if not w.get("layout")))
def test_workload_is_a_valid_problem_with_consistent_flops(name):
    w = bench.WORKLOADS[name]
    code = {"float16": 0, "float32": 1, "float64": 2}[w["dtype"]]
    units = int(np.prod(w["batch"]))
    p = _capi.make_problem(code, w["seq_dims"], w["rule"], w["sync"], (units, w["d"]) + w["q"], (units, w["d"]) + w["k"],
                           (units, w["v_d"]) + w["k"], w["w"], w["s"], w["c"])
    nq, nk = int(np.prod(w["q"])), int(np.prod(w["k"]))
    if nq * nk <= 1 << 28:                 # the library counts by enumeration; C5 (131072^2) has its closed form below
        assert _capi.count_attended(p) == bench.attended_pairs(w)
    else:
        assert w["rule"] == "causal" and nq == nk and bench.attended_pairs(w) == nq * (nq + 1) // 2
    fwd, bwd = bench.flops_of(w, bench.attended_pairs(w))
    assert fwd == 2.0 * bench.attended_pairs(w) * (w["d"] + w["v_d"]) * units
    assert bwd == 2.0 * bench.attended_pairs(w) * (3 * w["d"] + 2 * w["v_d"]) * units
    if w.get("channel_last_heads"):
        for world in (1, 2, 4, 8):          # every rank's slice keeps whole (outer, all heads) groups
            assert (units // world) % w["channel_last_heads"] == 0
        p.layout, p.heads = _capi.FA_LAYOUT_CHANNEL_LAST, w["channel_last_heads"]
        assert _capi.dispatch_path(p) == ("tcgen05_f16", "") and _capi.dispatch_path(p, backward=True)[0] == "tcgen05_f16"
    elif not w.get("ring"):
        assert _capi.dispatch_path(p)[0] != "generic_simt", "a bench workload must not fall back to the SIMT kernels"
        assert _capi.dispatch_path(p, backward=True)[0] != "generic_simt"
    assert _capi.lib.fa_workspace_bytes(C.byref(p), 1) >= _capi.lib.fa_workspace_bytes(C.byref(p), 0) or code == 1


def test_reference_arm_runs_without_the_product_library():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) prints one JSON line with the contract's
    keys and must not import tf_flash_attention_b200."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'C1', '--steps', '1', "
            "'--warmup', '1', '--cpu-heads', '1']; runpy.run_path('bench.py', run_name='__main__'); "
            "assert not any(m.startswith('tf_flash_attention_b200') for m in sys.modules), 'product library imported'")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config"):
        assert key in line
