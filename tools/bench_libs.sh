#!/bin/bash
# A/B variant builds of libmirfold: tools/bench_libs.sh "m4 m5" [extra bench args]   ("" = the default library)
mkdir -p gpurun_out
for v in $1; do
  if [ "$v" = "base" ]; then unset MIRFOLD_LIB_PATH; else export MIRFOLD_LIB_PATH=$PWD/mir_prefer_b200/libmirfold_$v.so; fi
  timeout 280 python bench.py --no-cpu --steps 3 --warmup 2 $2 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_v$v.json").read().strip().splitlines()[-1])
    print("lib $v: step %.1f ms fill %.2f ms f3 %.2f trace %.2f frac %.4f" % (d["ms_per_step"], d["stage_ms"]["ms_fill"], d["stage_ms"]["ms_f3"], d["stage_ms"]["ms_trace"], d["roofline"]["frac"]))
except Exception as e:
    print("lib $v failed", e); print(open("gpurun_out/bench_v$v.err").read()[-600:])
PY
done
