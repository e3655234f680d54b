#!/bin/bash
# round 2, 8-GPU box: the product's multi-device path -- tests, strong-scaling bench lines at N = 8/4/2/1 launched exactly
# like the driver does (torchrun, one rank per GPU; rank 0 drives ONE context over all N devices), configs[3] at scale.
TAG=${1:-f}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
nvidia-smi -L | head -8
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or overflow or stream or batch or fused or serial or outlive" > gpurun_out/r02_pytest_8gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_8gpu_$TAG.log
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu_$TAG.json 2> gpurun_out/r02_bench_${N}gpu_$TAG.err
  echo "bench N=$N rc=$?"
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-dropin > gpurun_out/r02_bench_1gpu_$TAG.json 2> gpurun_out/r02_bench_1gpu_$TAG.err
python - <<PY
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/r02_bench_%dgpu_$TAG.json'%N).read().strip().splitlines()[-1])
        print(N, 'value M nt/s', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'wall', round(d['wall_ms_per_step_resident'],1), 'e2e M nt/s', round(d['e2e']['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['sharding'], d.get('parity_in_run',{}).get('equal'), d.get('weak'))
    except Exception as e:
        print(N, 'ERR', e)
PY
timeout 1200 python tools/long_scale.py ${LONG_N:-2000000} 160 > gpurun_out/r02_long_scale_$TAG.log 2>&1
echo "long rc=$?"; tail -3 gpurun_out/r02_long_scale_$TAG.log
