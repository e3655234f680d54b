#!/bin/bash
# per-warp phase timeline of one CTA (MF_TIMELINE build) for the arabidopsis and parity laws
mkdir -p gpurun_out
for law in arabidopsis parity; do
MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_tl.so MIRFOLD_CHUNK_CELLS=1e12 timeout 300 python - > gpurun_out/r02_timeline_$law.log 2>&1 <<PY
import sys; sys.path.insert(0,'.')
import mir_prefer_b200 as mp
from mir_prefer_b200.corpus import synth_loci
seqs = synth_loci(1002, 2000, "$law")
with mp.MirFold() as mf:
    mf.fold(seqs, 300).close()
PY
grep TL gpurun_out/r02_timeline_$law.log | sort -k5 -n | head -20
done
