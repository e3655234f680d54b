"""Seeded corpora shared by tests, the oracle checker and bench.py (no file I/O, no reference access).

* lcg_records: the Python-RNG-free golden-pin generator of SURVEY.md Appendix C.
* synth_loci:  the synthetic locus generator of SURVEY.md section 8(d) (plant-like GC, embedded
  hairpins, N runs, lowercase masking).
"""
import numpy as np


def lcg_records(seed, nrec, lo, span):
    x = seed
    recs = []

    def r():
        nonlocal x
        x = (1103515245 * x + 12345) % (1 << 31)
        return x >> 16

    for k in range(nrec):
        n = lo + r() % span
        seq = "".join("ACGT"[r() & 3] for _ in range(n))
        recs.append(("r%d" % k, seq))
    return recs


def records_to_fasta(recs):
    return "".join(">%s\n%s\n" % (h, s) for h, s in recs)


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def _lengths(rng, nrec, law):
    if law == "parity":            # cfg-2: n ~ U{300..600}
        return rng.integers(300, 601, size=nrec)
    if law == "arabidopsis":       # cfg-3: extend_region law, MP:1272-1300
        u = rng.random(nrec)
        n = np.where(u < 0.70, 300, np.where(u < 0.75, 299, np.where(u < 0.95, 325, 0)))
        extra = rng.integers(301, 351, size=nrec)
        return np.where(n == 0, extra, n)
    if law == "long":              # cfg-4: log-uniform [2000, 10000]
        return np.exp(rng.uniform(np.log(2000), np.log(10000), size=nrec)).astype(np.int64)
    if law == "sweep":             # cfg-5: lognormal(ln 400, 0.6) clipped [60, 5000]
        return np.clip(np.exp(rng.normal(np.log(400), 0.6, size=nrec)), 60, 5000).astype(np.int64)
    if isinstance(law, tuple):     # (lo, hi) inclusive uniform
        return rng.integers(law[0], law[1] + 1, size=nrec)
    raise ValueError(law)


def synth_loci(seed, nrec, law="parity", plain=False):
    """Return list[str] of DNA loci (alphabet as `samtools faidx` would give)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = _lengths(rng, nrec, law)
    out = []
    alphabet = np.array(list("ACGT"))
    p = np.array([0.30, 0.20, 0.20, 0.30])
    for n in lens:
        n = int(n)
        s = alphabet[rng.choice(4, size=n, p=p)]
        if not plain and rng.random() < 0.5 and n >= 120:
            a = int(rng.integers(18, 31)); l = int(rng.integers(4, 41))
            arm5 = list(alphabet[rng.choice(4, size=a, p=p)])
            arm3 = [_COMP[c] for c in reversed(arm5)]
            for _ in range(int(rng.integers(0, 5))):
                arm3[int(rng.integers(0, len(arm3)))] = "ACGT"[int(rng.integers(0, 4))]
            for _ in range(int(rng.integers(0, 3))):
                del arm3[int(rng.integers(0, len(arm3)))]
            loop = list(alphabet[rng.choice(4, size=l, p=p)])
            hp = arm5 + loop + arm3
            off = int(rng.integers(0, n - len(hp) + 1))
            s[off:off + len(hp)] = hp
        s = "".join(s)
        if not plain:
            if rng.random() < 0.01:
                k = int(rng.integers(1, 21)); off = int(rng.integers(0, n - k + 1))
                s = s[:off] + "N" * k + s[off + k:]
            if rng.random() < 0.05:
                a, b = sorted(int(v) for v in rng.integers(0, n + 1, size=2))
                s = s[:a] + s[a:b].lower() + s[b:]
        out.append(s)
    return out


def cells(n, L):
    """DP cells the reference iterates for one locus."""
    Ls = min(L, n)
    tot = 0
    for i in range(1, n - 3):
        tot += min(n, i + Ls) - i - 3
    return tot
