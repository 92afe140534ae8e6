#!/bin/bash
# Runs the headline bench (kernel-only) once per library variant built by tools/ab_variants.sh, twice round-robin.
for rep in 1 2; do for v in "$@"; do
  PV_B200_LIB=phase-vocoder_b200/build/variants/libpv_b200_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-c5 ${AB_ARGS:-} 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,2))"
done; done
