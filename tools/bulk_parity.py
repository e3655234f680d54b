#!/usr/bin/env python3
"""Full-size parity run (GPU box): a whole synthetic config through libmirfold AND through the reference's
own RNALfold binary (oracle/_ref/RNALfold, one process per host core like the reference's fold stage),
byte-for-byte comparison of the complete RNALfold-format text plus sha256 of both.
usage: python tools/bulk_parity.py [law] [nloci] [span] [seed]      (default: parity 10000 300 1001 = BASELINE configs[1])"""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mir_prefer_b200.corpus import synth_loci  # noqa: E402
import mir_prefer_b200 as mp  # noqa: E402

law = sys.argv[1] if len(sys.argv) > 1 else "parity"
nloci = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
span = int(sys.argv[3]) if len(sys.argv) > 3 else 300
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1001
rlf = os.path.join(ROOT, "oracle", "_ref", "RNALfold")
assert os.access(rlf, os.X_OK), "oracle/_ref/RNALfold not staged"

seqs = synth_loci(seed, nloci, law)
recs = [">locus%d:%d-%d + 1-22 0 1,22,+\n%s\n" % (k, 1, len(s) + 1, s) for k, s in enumerate(seqs)]
text = "".join(recs)
cores = os.cpu_count() or 1
per = (len(recs) + cores - 1) // cores
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
t0 = time.time()
procs = []
for k in range(cores):
    part = recs[k * per:(k + 1) * per]
    if not part:
        continue
    fn = os.path.join(tmp, "p%d.fa" % k)
    open(fn, "w").write("".join(part))
    procs.append((fn, subprocess.Popen([rlf, "-L", str(span)], stdin=open(fn), stdout=open(fn + ".out", "w"))))
want = []
for fn, p in procs:
    assert p.wait() == 0
    want.append(open(fn + ".out").read())
want = "".join(want)
t_ref = time.time() - t0
t0 = time.time()
with mp.MirFold() as mf:
    got = mf.fold_text(text, span)
    t_gpu = time.time() - t0
    t0 = time.time()
    again = mf.fold_text_bytes(text, span)       # warm: buffers allocated, context up
    t_warm = time.time() - t0
    assert again.decode() == got
nt = sum(len(s) for s in seqs)
print("config: law=%s nloci=%d span=%d seed=%d  (%d nt)" % (law, nloci, span, seed, nt))
print("reference RNALfold on %d cores: %.1f s; libmirfold text in -> text out: %.2f s first call (context + allocations), %.2f s warm"
      % (len(procs), t_ref, t_gpu, t_warm))
print("bytes  ref %d  ours %d" % (len(want), len(got)))
print("sha256 ref  %s" % hashlib.sha256(want.encode()).hexdigest())
print("sha256 ours %s" % hashlib.sha256(got.encode()).hexdigest())
print("IDENTICAL" if got == want else "MISMATCH")
sys.exit(0 if got == want else 1)
