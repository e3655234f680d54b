#!/bin/bash
# ncu --set full capture of one fill-kernel launch (stride-608 bucket of the timed step) on a 2000-locus batch.
# usage (through gpurun): bash tools/ncu_fill.sh <tag> <kernel-regex> [MIRFOLD_OPTS]
TAG=$1; KREG=$2; export MIRFOLD_OPTS=${3:-0}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 2 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 1 --loci 2000 --no-cpu > gpurun_out/prof_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
