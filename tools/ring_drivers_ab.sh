#!/bin/bash
# Native (fa_ring_causal_*) against Python driver of the causal K/V ring on N GPUs: parity check + C5 bench lines.
# Usage (on a box with N GPUs): bash tools/ring_drivers_ab.sh N   -> gpurun_out/ring_ab_*.{log,json}
N=${1:-2}
mkdir -p gpurun_out
for drv in native python; do
  FA_RING_DRIVER=$drv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29533 tools/ring_check.py > gpurun_out/ring_ab_check_${drv}_n$N.log 2>&1
  grep -E "RING_CHECK" gpurun_out/ring_ab_check_${drv}_n$N.log | tail -2
  for flag in "" "--ring-bwd"; do
    tag=fwd; [ -n "$flag" ] && tag=fwdbwd
    FA_RING_DRIVER=$drv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29534 bench.py --gpus $N --workload C5 $flag --steps 5 --warmup 3 \
      > gpurun_out/ring_ab_${drv}_${tag}_n$N.json 2> gpurun_out/ring_ab_${drv}_${tag}_n$N.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ring_ab_${drv}_${tag}_n$N.json").read().strip().splitlines()[-1])
    print("$drv $tag", round(d["value"], 1), "TFLOPS", round(d["ms_per_step"], 3), "ms")
except Exception as e:
    print("$drv $tag failed", e)
PY
  done
done
