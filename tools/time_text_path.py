import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
from mir_prefer_b200.corpus import synth_loci
import mir_prefer_b200 as mp
seqs = synth_loci(1001, 10000, "parity")
text = "".join(">locus%d:%d-%d + 1-22 0 1,22,+\n%s\n" % (k, 1, len(s) + 1, s) for k, s in enumerate(seqs))
with mp.MirFold() as mf:
    mf.fold_text_bytes(text, 300)
    for _ in range(2):
        t0 = time.time(); items = mp.parse_rnalfold_input(text); t1 = time.time()
        toks = [tok for kind, tok in items if kind == "seq"]
        res = mf.fold(toks, 300); t2 = time.time()
        data, offs = res.record_blocks(); t3 = time.time()
        res.close()
        t4 = time.time(); b = mf.fold_text_bytes(text, 300); t5 = time.time()
        print("parse %.3f fold(pack+gpu) %.3f format %.3f | fold_text_bytes total %.3f (%d MB)" % (t1 - t0, t2 - t1, t3 - t2, t5 - t4, len(b) >> 20))
    # stage 1 (candidate structures) for the same batch: Python rules vs mirfold_classify
    from mir_prefer_b200 import structures as S
    headers = [">c:%d-%d + 1-22 0 1,22,+" % (k, k + len(s)) for k, s in enumerate(seqs)]
    with mf.fold(seqs, 300) as res:
        t0 = time.time(); nat = list(S.structures_from_result_native(headers, res, 55)); t1 = time.time()
        py = list(S.structures_from_result(headers, res, 55)); t2 = time.time()
        print("stage-1 structures: native %.3f s, python %.3f s, %d structures, equal=%s" % (t1 - t0, t2 - t1, sum(len(x[2]) for x in nat), nat == py))
