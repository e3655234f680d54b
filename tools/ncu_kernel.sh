#!/bin/bash
# ncu --set full capture of one launch of a named kernel on a 2000-locus batch:
# usage (through gpurun): bash tools/ncu_kernel.sh <tag> <kernel-regex> [skip]
TAG=$1; KREG=$2; SKIP=${3:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREG -s $SKIP -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 1 --loci 2000 --no-cpu > gpurun_out/prof_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
