#!/bin/bash
# tuning sweep of the numeric-pass knobs (experiments only): "TILE UNROLL MINBLOCKS CARVEOUT MIDGATHER" ("-" = default)
for cfg in "$@"; do set -- $cfg; 
  env_args=""
  [ "${1:--}" != "-" ] && env_args="$env_args FEMX_TILE=$1"
  [ "${2:--}" != "-" ] && env_args="$env_args FEMX_UNROLL=$2"
  [ "${3:--}" != "-" ] && env_args="$env_args FEMX_MINBLOCKS=$3"
  [ "${4:--}" != "-" ] && env_args="$env_args FEMX_CARVEOUT=$4"
  [ "${5:--}" != "-" ] && env_args="$env_args FEMX_MIDGATHER=$5"
  env $env_args python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 --workload ${WL:-cfg2} 2>&1 | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${WL:-cfg2} tile/unroll/minb/carve/mid=$cfg', round(j['ms_per_step'],4), round(j['roofline']['frac'],4))"; done
