#!/bin/bash
# round 2: uniform-warp fill kernel (k_fill_u16) -- parity tests, A/B against the round-1 schedule, cost-weight variants
TAG=${1:-c}
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
    print('${name}', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'fill(serial)', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
ARGS="--loci 40000"
run arab_new X=1
run arab_old MIRFOLD_OPTS=8
for v in wa wb wc wd; do run arab_$v MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_$v.so; done
ARGS="--workload parity"
run par_new X=1
run par_old MIRFOLD_OPTS=8
for v in wa wb wc wd; do run par_$v MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_$v.so; done
