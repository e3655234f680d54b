#!/bin/bash
# 8-GPU box: strong-scaling bench lines at N = 8/4/2 launched like the driver does, host phase log of one N = 8 call
TAG=${1:-i}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu_$TAG.json 2> gpurun_out/r02_bench_${N}gpu_$TAG.err
  echo "bench N=$N rc=$?"
done
python - <<PY
import json
for N in (2,4,8):
    try:
        d=json.loads(open('gpurun_out/r02_bench_%dgpu_$TAG.json'%N).read().strip().splitlines()[-1])
        print(N, 'value M nt/s', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'wall', round(d['wall_ms_per_step_resident'],1), 'e2e M nt/s', round(d['e2e']['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['sharding'], d.get('parity_in_run',{}).get('equal'), d['stage_ms_serial_pass'], d.get('weak'))
    except Exception as e:
        print(N, 'ERR', e)
PY
MIRFOLD_HOST_TIMING=1 timeout 300 python - > gpurun_out/r02_hosttiming_8gpu_$TAG.log 2>&1 <<PY
import sys, time; sys.path.insert(0,'.')
import bench, mir_prefer_b200 as mp
buf, off = bench.workload_packed(0, 200000)
with mp.MirFold(devices=list(range(8))) as mf:
    for k in range(3):
        t0=time.perf_counter(); mf.fold_packed(buf, off, 300).close(); print("call %d: %.1f ms" % (k, 1e3*(time.perf_counter()-t0)), file=sys.stderr)
PY
grep -E "call|shard plan|result buffers|devices|publish" gpurun_out/r02_hosttiming_8gpu_$TAG.log | tail -12
