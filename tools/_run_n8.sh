cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 $TR tools/dist_check.py 64 2>&1 | grep -E "^\{|Error|error" | tail -3
timeout 500 $TR bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r02_bench_cfg3_n8.json 2> gpurun_out/r02_bench_cfg3_n8.err
python - <<PY
import json
s=open('gpurun_out/r02_bench_cfg3_n8.json').read(); d=json.loads(s[s.index('{"metric'):])
c=d.get('cg',{}); print('n8', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['all_ranks_ok'], c.get('its_per_s'), c.get('spmv_ms'), c.get('residual@100'), c.get('error'))
PY
