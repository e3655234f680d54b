"""Smallest possible GPU check of the fill-kernel selection (no torch, no oracle run on the box): three loci at spans that
take the 864/608 buckets through both instantiations (int32-strip switch on/off), hits hashed against values computed
with the oracle on the CPU (python tools/gpu_mini_check.py --make prints them)."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mir_prefer_b200.corpus import synth_loci  # noqa: E402

CASES = [(900, 300), (900, 500), (560, 450), (560, 300), (1500, 420)]
WANT = {(900, 300): '19e7611f5782251f', (900, 500): '273f2baddbb5d63c', (560, 450): 'fe53eb98de390548', (560, 300): 'd72cae519ccd55e6',
        (1500, 420): '33ee84c7b7207024'}


def digest(hits, total):
    return hashlib.sha256(repr((hits, total)).encode()).hexdigest()[:16]


def main():
    seqs = {n: synth_loci(700 + n, 1, (n, n))[0] for n in {c[0] for c in CASES}}
    if "--make" in sys.argv:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        import oracle as O
        O.build()
        print({c: digest(O.fold(seqs[c[0]], c[1])["hits"], O.fold(seqs[c[0]], c[1])["total"]) for c in CASES})
        return
    import mir_prefer_b200 as mp
    bad = 0
    with mp.MirFold() as mf:
        for n, L in CASES:
            with mf.fold([seqs[n]], L) as r:
                ok = digest(r.hits(0), r.total(0)) == WANT[(n, L)]
                bad += not ok
                print(n, L, "ok" if ok else "MISMATCH", r.stats["fill_units"])
    print("mini check:", "PASS" if not bad else "FAIL")


if __name__ == "__main__":
    main()
