#!/bin/bash
# One GPU-box round: parity tests, bench, ncu launch list, ncu --set full of the fill kernels (full 10k batch).
# usage (from the repo root, through gpurun): bash tools/gpu_round.sh [tag] [skip_ncu]
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('bench ms/step', d['ms_per_step'], 'fill ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'cpu', d.get('cpu_baseline'))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
if [ -z "$2" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu_$TAG.log 2>&1
# the two fill launches (stride-608 and stride-352 buckets) of the timed step, on the bench's own 10k-locus batch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16 -s 2 -c 2 -o gpurun_out/prof_fill_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/prof_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traceback -s 1 -c 1 -o gpurun_out/prof_tb_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/prof_tb_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
