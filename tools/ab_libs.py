#!/usr/bin/env python
"""Developer A/B of alternative builds of the library (tools/build_variant.sh): per library, in its own process,
bit-equality of the forward outputs against the committed build on two shapes and the forward time of C2 / C4 / S1.

  gpurun -- 'python tools/ab_libs.py build_ab/libfa_x.so ...'
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os, torch
sys.path.insert(0, %(root)r)
from tf_flash_attention_b200 import _capi
from tf_flash_attention_b200 import flash_attention as fa
g = torch.Generator(device="cuda").manual_seed(1)
def u(*shape):
    return (torch.rand(shape, generator=g, device="cuda") * 4 - 2).half()
out = {"lib": os.environ.get("FA_B200_LIB", "default")}
sums = {}
for name, (b, d, s) in {"small": (4, 128, 1024), "ragged": (3, 128, 1000), "d64": (5, 64, 776)}.items():
    Q, K, V = u(b, d, s), u(b, d, s), u(b, d, s)
    O, l, m = fa.causal_1d(Q, K, V, "none_front", returning_l_m=True)
    torch.cuda.synchronize()
    sums[name] = [float(O.double().sum()), float(l.double().sum()), float(m.double().sum()), int(O.view(torch.int16).long().sum())]
out["checks"] = sums
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
Q, K, V = u(256, 128, 8192), u(256, 128, 8192), u(256, 128, 8192)
import subprocess, threading, time
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
timeit(lambda: fa.causal_1d(Q, K, V, "none_front"), 100)
t0 = time.time()
ms = timeit(lambda: fa.causal_1d(Q, K, V, "none_front"), 150)
t1 = time.time()
proc.terminate()
sel = [r.split(",") for t, r in rows if t0 + 0.05 <= t <= t1]
if sel:
    out["c2_clock_mhz"] = sorted(float(x[0]) for x in sel)[len(sel) // 2]
    out["c2_power_w"] = sorted(float(x[1]) for x in sel)[len(sel) // 2]
    out["c2_power_cap_active"] = sum("Active" in x[2] and "Not" not in x[2] for x in sel) / len(sel)
out["c2_fwd_ms"] = round(ms, 4); out["c2_fwd_tflops"] = round(4.398583382016e12 / (ms * 1e-3) / 1e12, 1)
del Q, K, V
q, k, v = u(256, 64, 1024), u(256, 64, 8192), u(256, 64, 8192)
out["c4_fwd_ms"] = round(timeit(lambda: fa.full_1d(q, k, v, "scale_end")), 4)
q, k, v = u(16384, 64, 256), u(16384, 64, 256), u(16384, 64, 256)
out["s1_fwd_ms"] = round(timeit(lambda: fa.causal_1d(q, k, v, "none_front")), 4)
print(json.dumps(out))
'''


def main():
    libs = [None] + sys.argv[1:] + [None]
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["FA_B200_LIB"] = os.path.abspath(lib)
        else:
            env.pop("FA_B200_LIB", None)
        try:
            r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True,
                               timeout=240, env=env)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:]
        except subprocess.TimeoutExpired:
            line = json.dumps({"lib": lib, "error": "timeout (hung kernel?)"})
        print(line, flush=True)


if __name__ == "__main__":
    main()
