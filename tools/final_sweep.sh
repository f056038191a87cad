#!/bin/bash
# Round-end verification on one B200: smoke, the GPU test suite, the default bench line, the reference arm, every other
# workload, the reference's own benchmark, and the ncu launch list of the bench command. Outputs under gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/r2_final_smoke.log 2>&1; tail -1 $O/r2_final_smoke.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2_final_gpu_tests.log 2>&1; tail -1 $O/r2_final_gpu_tests.log
timeout 400 python bench.py > $O/r2_final_c2_n1.json 2> $O/r2_final_c2_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_reference_arm.json 2>/dev/null; echo "reference rc=$?"
for w in C1 C2cl C3 C4 C4f32 C4f64 S1 S2 T1; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-refkernel > $O/r2_final_${w}_n1.json 2>/dev/null
  echo "$w rc=$?"
done
timeout 200 python bench.py --workload C1 --graph --steps 20 --warmup 3 --no-cpu-baseline --no-refkernel --no-e2e > $O/r2_final_C1_n1_graph.json 2>/dev/null
timeout 300 python -m tf_flash_attention_b200.tests.test_1d TestGroup.benchmark > $O/r2_final_package_bench_1d.log 2>&1
timeout 300 python -m tf_flash_attention_b200.tests.test_2d TestGroup.benchmark > $O/r2_final_package_bench_2d.log 2>&1
timeout 120 python tools/bench_channel_last.py > $O/r2_final_channel_last.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-refkernel --no-e2e > $O/r2_final_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final_*_n1*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        e = d.get("e2e", {})
        print(f.split("r2_final_")[1], round(d["value"], 2), d["unit"], round(d["ms_per_step"], 3), "ms", "frac", round(r.get("frac", 0), 3),
              "e2e", round(e.get("value", 0), 1) if e else None, d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
