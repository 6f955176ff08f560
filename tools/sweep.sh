#!/bin/bash
# tuning sweep of the numeric-pass knobs (experiments only): TILE UNROLL MINBLOCKS CARVEOUT
for cfg in "$@"; do set -- $cfg; FEMX_TILE=$1 FEMX_UNROLL=$2 FEMX_MINBLOCKS=$3 FEMX_CARVEOUT=$4 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 --workload ${WL:-cfg2} 2>&1 | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile/unroll/minb/carve=$cfg', round(j['ms_per_step'],4), round(j['roofline']['frac'],4))"; done
