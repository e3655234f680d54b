#!/bin/bash
# 1-GPU validation: full parity suite + the default bench line (200k loci) + 2-lane multi-"device" [0,0] long-scale smoke
TAG=${1:-g}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_$TAG.json').read().strip().splitlines()[-1])
print('value M nt/s', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e6,2), round(d['e2e']['ms_per_step'],1), 'fill', round(d['roofline']['kernel_ms'],1), 'frac', round(d['roofline']['frac'],4), d['stage_ms_serial_pass'], d.get('parity_in_run',{}).get('equal'), d.get('cpu_baseline',{}).get('value'), d.get('drop_in'))
PY
timeout 600 python tools/long_scale.py 4000 24 > gpurun_out/r02_long_scale_smoke_$TAG.log 2>&1; tail -2 gpurun_out/r02_long_scale_smoke_$TAG.log
