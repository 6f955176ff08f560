#!/bin/bash
# ncu evidence for every kernel of the hot path (run on the GPU box, one GPU): one `--set full` capture per kernel,
# summarised by profiles/ncu_summary.py (DRAM throughput, sectors/request, pipe utilisation, warp stall reasons), plus the
# launch list of a bench.py run.  Writes gpurun_out/r02_ncu_*.txt and gpurun_out/r02_launches_*.csv; copy what is to
# be kept into profiles/.   usage: bash tools/profile_kernels.sh
set -u
mkdir -p gpurun_out/jit
export FEMX_JIT_DUMP=gpurun_out/jit
NCU="ncu --set full --import-source on --clock-control none -f"
prof() {  # name, kernel regex, launch-skip, command...
  local name=$1 kern=$2 skip=$3; shift 3
  $NCU -k "regex:$kern" --launch-skip "$skip" -c 1 -o gpurun_out/r02_ncu_$name "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  python profiles/ncu_summary.py gpurun_out/r02_ncu_$name.ncu-rep gpurun_out/r02_ncu_$name.txt > /dev/null 2>&1 || echo "summary of $name failed"
}
# numeric pass: cfg3 (element-once lattice pass + boundary-row kernel), cfg2 (stencil-class pass), cfg4 (generic pass, elasticity)
prof csr_lattice_cfg3 '^femx_csr$' 3 python tools/measure_one.py cfg3 4
prof rowlist_cfg3 '^femx_rowlist$' 3 python tools/measure_one.py cfg3 4
prof csr_stencil_cfg2 '^femx_csr$' 3 python tools/measure_one.py cfg2 4
prof csr_generic_cfg4 '^femx_csr$' 3 python tools/measure_one.py cfg4 4
prof coo_cfg2 '^femx_coo$' 1 python tools/measure_one.py cfg2 2 coo
# symbolic pass (lattice-templated and general) and the validation kernels
prof pattern_fill_cfg3 'lat_row_fill_k' 1 python tools/measure_one.py cfg3 1
prof pattern_general_rowfill_cfg2 'row_fill' 1 env FEMX_LATTICE_PATTERN=0 python tools/measure_one.py cfg2 1
prof spmv_cfg3 'spmv_tile_k' 2 python tools/measure_one.py cfg3 1 cg
prof cg_update_cfg3 'cg_update_k' 5 python tools/measure_one.py cfg3 1 cg
# launch list of one symbolic pass + a few numeric passes (which kernels the one-time setup spends its time in)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_setup_cfg3.csv \
  python tools/measure_one.py cfg3 2 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches_setup_cfg3.csv gpurun_out/r02_launches_setup_cfg3_summary.txt "python tools/measure_one.py cfg3 2" > /dev/null 2>&1
# launch list of the benchmark command itself (shares, not absolutes: cold caches, serialised launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg3.csv \
  python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_launches_cfg3.log 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches_cfg3.csv gpurun_out/r02_launches_cfg3_summary.txt "python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline" > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_*.txt gpurun_out/r02_launches_cfg3_summary.txt
