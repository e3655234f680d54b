#!/bin/bash
# same-box A/B of library builds: bash tools/gpu_ab.sh TAG "name=path name2=path2 ..." (cur = the in-tree library); parity subset first
TAG=${1:-ab}
VARIANTS=${2:-}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-golden or sha256 or matrices or wide or 16bit or edge or tiled or randomized or long_loci or span_sweep or text_path}" > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_$TAG.log
Q="--steps 3 --warmup 2 --no-cpu --no-sha --no-dropin"
run() { name=$1; shift
  env "$@" timeout 600 python bench.py $Q $ARGS > gpurun_out/r02_ab_${name}_$TAG.json 2> gpurun_out/r02_ab_${name}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ab_${name}_$TAG.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['stage_ms_serial_pass'])
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
for W in "--loci 40000" "--workload parity" ${AB_EXTRA:+"$AB_EXTRA"}; do
  ARGS="$W"; n=$(echo $W | tr -d ' -' | cut -c1-12)
  for rep in 1 2; do
    run ${n}_cur$rep X=1
    for v in $VARIANTS; do run ${n}_${v%%=*}$rep MIRFOLD_LIB_PATH=$PWD/${v#*=}; done
  done
done
