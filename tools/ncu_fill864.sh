#!/bin/bash
# ncu --set full of the four fill buckets of one sweep-law fold at L=500 (launch order per chunk: 864, 608, 352, 160); the first fold is skipped
TAG=${1:-y}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache MIRFOLD_CHUNK_CELLS=1e12 MIRFOLD_SERIAL=1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16 -s 4 -c 4 -o gpurun_out/r02_prof_fill864_$TAG -f python tools/ab_span.py sweep 3000 500 1 > gpurun_out/r02_prof_fill864_$TAG.log 2>&1
echo "ncu rc=$? t=$SECONDS"; ls -la gpurun_out/r02_prof_fill864_$TAG.ncu-rep; tail -3 gpurun_out/r02_prof_fill864_$TAG.log
