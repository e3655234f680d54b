#!/bin/bash
# round 2: fML16 row-pair layouts A/B -- (ocopy) round-1 layout with the second, odd-aligned copy; (shfl) single copy, neighbour
# word by SHFL; (default) single copy, second LDG.  Fill time = serial-pass CUDA events.
TAG=${1:-e}
mkdir -p gpurun_out
Q="--steps 3 --warmup 2 --no-cpu --no-sha --no-dropin"
run() { name=$1; shift
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
for W in "--loci 40000" "--workload parity" "--workload long --loci 700" "--workload sweep --loci 4000 --span 500"; do
  ARGS="$W"; n=$(echo $W | tr -d ' -' | cut -c1-14)
  run ${n}_ldg X=1
  run ${n}_shfl MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_shfl.so
  run ${n}_ocopy MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_ocopy.so
done
timeout 600 python -m pytest tests -m gpu -x -q -k "matrices or sha256 or tiled or wide or 16bit or randomized" > gpurun_out/r02_pytest_$TAG.log 2>&1; tail -2 gpurun_out/r02_pytest_$TAG.log
