#!/bin/bash
# 1-GPU pass for the stride-864 bucket: full parity suite at the default threshold and with the bucket forced on for every span,
# then same-box A/B of the threshold on the workloads with loci > 608 nt (bash tools/gpu_bigtile.sh TAG through gpurun)
TAG=${1:-w}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest(default) rc=$?"; tail -4 gpurun_out/r02_pytest_$TAG.log
MIRFOLD_BIG_TILE_MIN_SPAN=0 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_big0_$TAG.log 2>&1
echo "pytest(big tiles at every span) rc=$?"; tail -4 gpurun_out/r02_pytest_big0_$TAG.log
: > gpurun_out/r02_ab_bigtile_$TAG.jsonl
for cfg in "sweep 5000 500" "sweep 5000 300" "long 1400 300" "long 1400 500"; do
  for thr in 100000 0; do
    MIRFOLD_BIG_TILE_MIN_SPAN=$thr timeout 300 python tools/ab_span.py $cfg >> gpurun_out/r02_ab_bigtile_$TAG.jsonl 2>> gpurun_out/r02_ab_bigtile_$TAG.err
  done
done
cat gpurun_out/r02_ab_bigtile_$TAG.jsonl; tail -5 gpurun_out/r02_ab_bigtile_$TAG.err
# 768-thread build of the 864 bucket (24 warps, 80 registers): make VARIANT=nt768 EXTRA=-DMF_B864_NT=768
if [ -f mir_prefer_b200/libmirfold_nt768.so ]; then
  for cfg in "sweep 5000 500" "long 1400 300"; do
    MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_nt768.so MIRFOLD_BIG_TILE_MIN_SPAN=0 timeout 300 python tools/ab_span.py $cfg | sed 's/^{/{"variant": "nt768", /'
  done
fi
