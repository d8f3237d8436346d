#!/bin/bash
# development: process-level A/B of stream priorities for the bench loop (one box, alternating)
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-pass --no-fp32-side 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
for r in 1 2; do
  run default
  MM3D_BENCH_MAIN_PRIO=-1 run main_high
  MM3D_BENCH_MAIN_PRIO=-2 MM3D_BENCH_PREP_PRIO=0 run main_higher
done
