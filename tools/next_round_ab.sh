#!/bin/bash
# A/B list prepared at the end of round 1 (no GPU minutes were left to time it): knobs that are implemented,
# verified on the host (tests/test_stencil_host.py) and off by default.  One gpurun call:
#   gpurun --timeout 900 -- 'bash tools/next_round_ab.sh | tee gpurun_out/ab.txt'
# Every line: config [knobs] csr_ms roofline_frac (tools/measure_all.py, mean of 10 launches).
cd "$(dirname "$0")/.."
bash tools/sweep_all.sh cfg3 "A=baseline" "FEMX_ROWSUM=1" "FEMX_RCP3=1" "FEMX_ROWSUM=1 FEMX_RCP3=1" \
     "FEMX_ROWSUM=1 FEMX_RCP3=1 FEMX_SPEC_AHEAD=1" "FEMX_ROWSUM=1 FEMX_RCP3=1 FEMX_SPEC_AHEAD=3" \
     "FEMX_CHAINORDER=1" "FEMX_CHAINORDER=1 FEMX_ROWSUM=1 FEMX_RCP3=1" "FEMX_ROWSUM=1 FEMX_RCP3=1 FEMX_MINBLOCKS=5" "FEMX_ROWSUM=1 FEMX_RCP3=1 FEMX_TILE=64 FEMX_MINBLOCKS=8"
bash tools/sweep_all.sh cfg2 "A=baseline" "FEMX_MINBLOCKS=7" "FEMX_MINBLOCKS=9" "FEMX_SPEC_AHEAD=2" "FEMX_TILE=64 FEMX_MINBLOCKS=16"
# fp64 issue rates incl. the shared-operand (.reuse) and DMUL variants (build first: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_peak tools/micro/fp64_peak.cu)
[ -x tools/micro/fp64_peak ] && tools/micro/fp64_peak | grep 'mix=[245]'
# the bitwise spec == generic tests must hold with the knobs on as well
FEMX_ROWSUM=1 FEMX_RCP3=1 python -m pytest tests/test_stencil.py tests/test_gpu_parity.py -m gpu -x -q | tail -2
