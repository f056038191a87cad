#!/bin/bash
# developer A/B: full fwd+bwd C2 bench against alternative builds of the library
for lib in "$@"; do
  echo "== $lib"
  FA_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-refkernel 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'TFLOPS', round(d['ms_per_step'],3), 'ms', {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])"
done
