#!/bin/bash
# round 2: no odd-aligned fML16 copy + bulk-async fML row stores -- parity tests, A/B, ncu traffic
TAG=${1:-d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_$TAG.log
Q="--steps 3 --warmup 2 --no-cpu --no-sha --no-dropin"
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 600 python bench.py $Q $ARGS > gpurun_out/r02_ab_${name}_$TAG.json 2> gpurun_out/r02_ab_${name}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ab_${name}_$TAG.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'fill(serial)', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4), d['stage_ms_serial_pass'])
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
ARGS="--loci 40000"
run arab_bulk X=1
run arab_nobulk MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_nb.so
ARGS="--workload parity"
run par_bulk X=1
run par_nobulk MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_nb.so
export MIRFOLD_CHUNK_CELLS=1e12 MIRFOLD_SERIAL=1
B="python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu --no-sha --no-dropin"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16 -s 1 -c 1 -o gpurun_out/r02_prof_fill352_$TAG -f $B > gpurun_out/r02_prof_fill_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
