#!/bin/bash
# round 2: ncu --set full of the stride-352 fill kernel, k_traceback and k_f3 on the arabidopsis law (one chunk, serial),
# the launch list of a short bench, and the per-warp phase timeline of one CTA (MF_TIMELINE build)
TAG=${1:-b}
mkdir -p gpurun_out
export MIRFOLD_CHUNK_CELLS=1e12 MIRFOLD_SERIAL=1
B="python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu --no-sha --no-dropin"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16 -s 1 -c 1 -o gpurun_out/r02_prof_fill352_$TAG -f $B > gpurun_out/r02_prof_fill_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traceback -s 1 -c 1 -o gpurun_out/r02_prof_tb_$TAG -f $B > gpurun_out/r02_prof_tb_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_f3 -s 1 -c 1 -o gpurun_out/r02_prof_f3_$TAG -f $B > gpurun_out/r02_prof_f3_$TAG.log 2>&1
unset MIRFOLD_CHUNK_CELLS MIRFOLD_SERIAL
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_$TAG.csv python bench.py --loci 20000 --steps 2 --warmup 1 --no-cpu --no-sha --no-dropin > gpurun_out/r02_b_ncu_$TAG.log 2>&1
MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_tl.so MIRFOLD_CHUNK_CELLS=1e12 timeout 300 python - > gpurun_out/r02_timeline_$TAG.log 2>&1 <<PY
import sys; sys.path.insert(0,'.')
import mir_prefer_b200 as mp
from mir_prefer_b200.corpus import synth_loci
seqs = synth_loci(1002, 2000, "arabidopsis")
with mp.MirFold() as mf:
    mf.fold(seqs, 300).close()
PY
tail -20 gpurun_out/r02_timeline_$TAG.log
ls -la gpurun_out/*.ncu-rep
