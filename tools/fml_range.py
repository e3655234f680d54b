"""How far fML drops on ordinary sequence, by span (CPU, oracle matrices): the basis of MF_DYNW_MIN_SPAN (host picks the
k_fill_s16 instantiation with the int32-strip switch from there on).  Output of the run of record: profiles/r02_v7_bigtile_ab.txt."""
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import numpy as np
import oracle as O
O.build()
from mir_prefer_b200.corpus import synth_loci
seqs = synth_loci(1004, 400, "sweep")
for L in (300, 400, 500):
    nl=0; fl=0; mins=[]
    for s in seqs:
        if len(s) < 420 or len(s) > 900: continue
        o = O.fold(s, L, matrices=True)
        m = o["m"]; mm = int(m[m < 500000].min()) if (m < 500000).any() else 0
        cm = int(o["c"][o["c"] < 500000].min())
        nl += 1; fl += (mm < -13600) or (cm < -32000); mins.append(mm)
        if nl >= 40: break
    print(L, "loci", nl, "flagged", fl, "median min fML", int(np.median(mins)), "min", min(mins))
