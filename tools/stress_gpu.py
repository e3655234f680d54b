import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import oracle as O
from mir_prefer_b200.corpus import synth_loci
import mir_prefer_b200 as mp
O.build()
mf = mp.MirFold()
# (a) many tiny loci
rng = np.random.default_rng(1)
tiny = ["".join(rng.choice(list("ACGU"), size=int(rng.integers(0, 64)))) for _ in range(60000)]
t0 = time.time()
with mf.fold(tiny, 300) as res:
    print("tiny: %d loci, %d hits, %.1f ms device, %.2f s wall" % (len(tiny), res.nhits, res.stats["ms_device"], time.time() - t0))
    for r in range(0, len(tiny), 997):
        o = O.fold(tiny[r], 300)
        assert res.hits(r) == o["hits"] and res.total(r) == o["total"], r
# (b) one very long locus + a few others
big = synth_loci(5, 1, (20000, 20000))[0]
seqs = [big] + synth_loci(6, 3, (300, 500))
t0 = time.time()
with mf.fold(seqs, 300) as res:
    print("long: n=%d, units=%d, %d hits, %.1f ms device" % (len(big), res.stats["fill_units"], res.nhits, res.stats["ms_device"]))
    t1 = time.time()
    o = O.fold(big, 300)
    print("oracle %.1f s" % (time.time() - t1))
    assert res.hits(0) == o["hits"] and res.total(0) == o["total"]
    for r in (1, 2, 3):
        o = O.fold(seqs[r], 300)
        assert res.hits(r) == o["hits"] and res.total(r) == o["total"]
# (c) large span through the generic kernel
s = synth_loci(7, 1, (1500, 1500))[0]
with mf.fold([s], 1000) as res:
    t1 = time.time(); o = O.fold(s, 1000); print("oracle L=1000 %.1f s" % (time.time() - t1))
    assert res.hits(0) == o["hits"] and res.total(0) == o["total"]
    print("generic L=1000 ok, %d hits, fill %.1f ms" % (res.nhits, res.stats["ms_fill"]))
print("stress ok")
