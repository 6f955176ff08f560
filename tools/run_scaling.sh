#!/bin/bash
# One strong-scaling point of the headline benchmark under torchrun:  bash tools/run_scaling.sh N   (N = 2, 4, 8; e.g. gpurun --gpus N)
# writes gpurun_out/r02_bench_cfg3_nN.json and prints the headline fields
cd /root/repo
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
timeout 500 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_cfg3_n$N.json 2> gpurun_out/r02_bench_cfg3_n$N.err
python - <<PY
import json
s=open('gpurun_out/r02_bench_cfg3_n$N.json').read(); d=json.loads(s[s.index('{"metric'):])
c=d.get('cg',{}); print('n$N', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['all_ranks_ok'], c.get('its_per_s'), c.get('spmv_ms'), c.get('residual@100'), c.get('error'))
PY
