#!/bin/bash
# same-box A/B: current build / without bulk row stores / first round-2 commit (round-1 kernels), arabidopsis 40k, two repeats
TAG=${1:-h}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 600 python -m pytest tests -m gpu -x -q -k "matrices or sha256_pins or golden or edge or arabidopsis_length or span_sweep or long_loci" > gpurun_out/r02_pytest_$TAG.log 2>&1; tail -2 gpurun_out/r02_pytest_$TAG.log
Q="--steps 3 --warmup 2 --no-cpu --no-sha --no-dropin --loci 40000"
run() { name=$1; shift
  env "$@" timeout 600 python bench.py $Q > gpurun_out/r02_ab_${name}_$TAG.json 2> gpurun_out/r02_ab_${name}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ab_${name}_$TAG.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['stage_ms_serial_pass'])
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
for rep in 1 2; do
run cur$rep X=1
run nobulk$rep MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_nb.so
run r1_$rep MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_r1.so
done
