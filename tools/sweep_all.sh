#!/bin/bash
# usage: tools/sweep_all.sh cfgN "ENV1=a ENV2=b" "..."   — measure_all on one config under env settings
cfg=$1; shift
for e in "$@"; do env $e python tools/measure_all.py $cfg | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('$cfg [$e]', 'csr_ms', round(j['csr_ms'],4), round(j['roofline_frac'],4))"; done
