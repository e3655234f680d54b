#!/bin/bash
# last GPU call of the round: parity subset over everything the dynamic int32-strip switch touches, then three A/B lines
TAG=${1:-z}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 75 python -m pytest tests -m gpu -x -q -k "16bit or tiled or big_tile or span_sweep or sha256 or wide or matrices or long_loci or randomized" > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$? t=$SECONDS"; tail -4 gpurun_out/r02_pytest_$TAG.log
for cfg in "sweep 5000 500" "parity 10000 300" "long 1400 300"; do
  timeout 40 python tools/ab_span.py $cfg 2>&1 | tail -1
done
echo "t=$SECONDS"
