#!/bin/bash
# A/B the fill kernel under several MIRFOLD_OPTS values: tools/bench_opts.sh "0 256 768" [extra bench args]
mkdir -p gpurun_out
for o in $1; do
  MIRFOLD_OPTS=$o timeout 280 python bench.py --no-cpu --steps 3 --warmup 2 $2 > gpurun_out/bench_o$o.json 2> gpurun_out/bench_o$o.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_o$o.json").read().strip().splitlines()[-1])
    print("opts $o: step %.1f ms fill %.2f ms f3 %.2f trace %.2f frac %.4f" % (d["ms_per_step"], d["stage_ms"]["ms_fill"], d["stage_ms"]["ms_f3"], d["stage_ms"]["ms_trace"], d["roofline"]["frac"]))
except Exception as e:
    print("opts $o failed", e); print(open("gpurun_out/bench_o$o.err").read()[-600:])
PY
done
