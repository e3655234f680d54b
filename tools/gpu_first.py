import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/oracle')
import numpy as np
import oracle as O
from mir_prefer_b200.corpus import synth_loci, lcg_records
import mir_prefer_b200 as mp
mf = mp.MirFold()
print(mp._lib.load().mirfold_version())
def check(seqs, L, tag):
    t=time.time()
    with mf.fold(seqs, L) as res:
        dt=time.time()-t
        bad=0
        for r,s in enumerate(seqs):
            o = O.fold(s, L)
            g = res.hits(r)
            if o['hits']!=g or o['total']!=res.total(r):
                bad+=1
                if bad<=3:
                    print('MISMATCH', tag, r, len(s), 'total', o['total'], res.total(r), 'nh', len(o['hits']), len(g))
                    for a,b in zip(o['hits'], g):
                        if a!=b: print('  o', a, '\n  g', b); break
        print(tag, 'L',L,'n',len(seqs),'bad',bad,'time %.3f'%dt, {k:(round(v,3) if isinstance(v,float) else v) for k,v in res.stats.items()})
    return bad
# matrix check first
for s in ['GGGGAAAACCCC', 'GGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC', synth_loci(5,1,(60,60))[0], synth_loci(6,1,(320,320))[0]]:
    for L in (30,300):
        o=O.fold(s,L,matrices=True); c,m,f3=mf.debug_matrices(s,L)
        n=len(s)
        okc=(o['c']==c).all(); okm=(np.minimum(o['m'],1000000)==np.minimum(m,1000000)).all(); okf=(o['f3'][:n+3]==f3[:n+3]).all()
        print('matrices n',n,'L',L,'c',okc,'m',okm,'f3',okf)
        if not okc:
            idx=np.argwhere(o['c']!=c); print(' c diff at', idx[:5], o['c'][tuple(idx[0])], c[tuple(idx[0])])
        if not okm:
            idx=np.argwhere(np.minimum(o['m'],1000000)!=np.minimum(m,1000000)); print(' m diff at', idx[:5], o['m'][tuple(idx[0])], m[tuple(idx[0])])
check(['GGGGAAAACCCC','AAAAAAAAAAAAAAAAAAAAA','ACGUNNNACGU','ACG','','GGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC'], 30, 'tiny')
check([s for _,s in lcg_records(3,300,20,200)], 40, 'pin3')
check(synth_loci(11, 64, 'parity'), 300, 'parity64')
check(synth_loci(12, 16, (1000,2000)), 300, 'long')
check(synth_loci(13, 32, 'sweep'), 150, 'sweep150')
check(synth_loci(13, 32, 'sweep'), 500, 'sweep500')
seqs=synth_loci(1001, 2000, 'parity')
for it in range(3):
    t=time.time()
    with mf.fold(seqs,300) as res:
        print('2000 loci: %.3f s'%(time.time()-t), {k:(round(v,3) if isinstance(v,float) else v) for k,v in res.stats.items()})
