#!/bin/bash
# developer A/B: run the fwd-only C2 bench against alternative builds of the library
for lib in "$@"; do
  echo "== $lib"
  FA_B200_LIB=$lib python bench.py --fwd-only --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-refkernel 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'TFLOPS', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
