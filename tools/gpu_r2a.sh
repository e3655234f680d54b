#!/bin/bash
# round 2, first GPU call: parity tests, the new bench line (200k loci through one context), A/B of the lanes
TAG=${1:-a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_$TAG.err
MIRFOLD_SERIAL=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-sha --no-dropin > gpurun_out/r02_bench_serial_$TAG.json 2> gpurun_out/r02_bench_serial_$TAG.err
timeout 600 python bench.py --workload parity --steps 5 --warmup 3 --no-cpu --no-dropin > gpurun_out/r02_bench_parity_$TAG.json 2> gpurun_out/r02_bench_parity_$TAG.err
MIRFOLD_HOST_TIMING=1 timeout 300 python bench.py --workload parity --steps 1 --warmup 1 --no-cpu --no-dropin --no-sha > /dev/null 2> gpurun_out/r02_hosttiming_$TAG.err
python - <<PY
import json
for f in ("r02_bench_$TAG","r02_bench_serial_$TAG","r02_bench_parity_$TAG"):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],2), 'wall', round(d['wall_ms_per_step_resident'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'fill', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4), 'serial', d['stage_ms_serial_pass'], d.get('parity_in_run'), d.get('cpu_baseline'), d.get('drop_in'))
    except Exception as e:
        print(f, 'ERR', e)
PY
