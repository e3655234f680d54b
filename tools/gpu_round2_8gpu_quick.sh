#!/bin/bash
# 8-GPU box, quick: e2e / resident at N = 8 only (torchrun) + host phase log
TAG=${1:-t}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or stream or overflow or serial" > gpurun_out/r02_pytest_8gpu_$TAG.log 2>&1; tail -2 gpurun_out/r02_pytest_8gpu_$TAG.log
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no-weak > gpurun_out/r02_bench_${N}gpu_$TAG.json 2> gpurun_out/r02_bench_${N}gpu_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_8gpu_$TAG.json').read().strip().splitlines()[-1])
print(8, 'value M nt/s', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'e2e M nt/s', round(d['e2e']['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['e2e']['last_step_split'], d['parity_in_run']['equal'])
PY
