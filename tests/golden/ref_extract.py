"""AST-extract pure functions of the reference (/root/reference/miR_PREFeR.py, Python-2 script that
still parses under Python 3) and exec them, so that tests/golden fixtures come from the reference's
OWN code.  Only usable in the build container (the reference is absent on the GPU box)."""
import ast
import re

REF = "/root/reference/miR_PREFeR.py"
WANTED = ["get_structures_next_extendregion", "is_stem_loop", "has_one_good_bifurcation", "filter_ss",
          "pos_genome_2_local", "pos_local_2_genome", "stat_duplex", "pass_stat_duplex", "get_maturestar_info"]


def load():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"re": re}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            code = compile(ast.Module(body=[node], type_ignores=[]), REF, "exec")
            exec(code, ns)
    missing = [w for w in WANTED if w not in ns]
    assert not missing, missing
    return ns
