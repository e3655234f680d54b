#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ by running the REFERENCE's own RNALfold binary
(/root/reference/dependency/Linux/x64/RNALfold, ViennaRNA 1.8.5) on small seeded inputs.
Run in the build container only (the reference does not exist on the GPU box); outputs are committed.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mir_prefer_b200.corpus import lcg_records, records_to_fasta, synth_loci  # noqa: E402

RLF = "/root/reference/dependency/Linux/x64/RNALfold"


def rlf(text, L):
    return subprocess.run([RLF, "-L", str(L)], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.decode()


def emit(name, text, spans):
    with open(os.path.join(HERE, name + ".in"), "w") as f:
        f.write(text)
    for L in spans:
        with open(os.path.join(HERE, "%s.L%d.out" % (name, L)), "w") as f:
            f.write(rlf(text, L))


def main():
    # hand-checkable vectors of SURVEY.md Appendix C + program-I/O edge cases (A.6)
    edge = (">poly\nGGGGAAAACCCC\n>polyA\nAAAAAAAAAAAAAAAAAAAAA\n>withN\nACGUNNNACGU\n"
            ">three\nGGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC\n>empty\n\n>n4\nACGU\n>n5\nGCGCA\n*comment line\n"
            "GGGGAAAACCCC trailing tokens ignored\n>lead\n  GGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC\n"
            ">Chr1:5000-5114 + 5020-5041 0 5020,5041,+ M:5019-5040/+/120\n"
            "AACTATACGTTAGCATTTGGATTGAAGGGAGCTCTACATCTTCTTGCTAGATTCATCAGTTAACCTAGCAAGAAGAAGTAGAGCTCCCGTCAATCCAAATTATACCGATAACTA\n")
    emit("edge", edge, (30, 300))
    emit("pin3_head", records_to_fasta(lcg_records(3, 60, 20, 200)), (40,))
    loci = synth_loci(2001, 8, "parity")
    hdr = [">Chr%d:%d-%d + %d-%d 0 %d,%d,+ M:%d-%d/+/%d" % (k + 1, 1000 * k, 1000 * k + len(s), 1000 * k + 50, 1000 * k + 71,
                                                            1000 * k + 50, 1000 * k + 71, 1000 * k + 50, 1000 * k + 71, 10 + k)
           for k, s in enumerate(loci)]
    emit("synth8", "".join("%s\n%s\n" % (h, s) for h, s in zip(hdr, loci)), (300,))
    rng = np.random.default_rng(5)
    recs = [("a%d" % k, "".join(rng.choice(list("ACGUTacgutNnKXIRYkxi"), size=int(rng.integers(5, 120))))) for k in range(40)]
    emit("alphabet", records_to_fasta(recs), (40,))
    sweep = synth_loci(2004, 6, "sweep")
    emit("sweep6", records_to_fasta([("w%d" % k, s) for k, s in enumerate(sweep)]), (150, 500))
    pins = []
    for seed, n, lo, span, L in ((1, 1000, 50, 351, 300), (2, 200, 300, 301, 300), (3, 300, 20, 200, 40)):
        text = records_to_fasta(lcg_records(seed, n, lo, span))
        out = rlf(text, L)
        pins.append({"seed": seed, "nrec": n, "lo": lo, "span": span, "L": L,
                     "sha256_stdin": hashlib.sha256(text.encode()).hexdigest(),
                     "sha256_stdout": hashlib.sha256(out.encode()).hexdigest(), "stdout_bytes": len(out)})
    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump({"binary_sha256": hashlib.sha256(open(RLF, "rb").read()).hexdigest(), "pins": pins}, f, indent=1)


if __name__ == "__main__":
    main()
