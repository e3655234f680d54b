#!/bin/bash
# final 1-GPU pass of the round (bash tools/gpu_final.sh TAG through gpurun): parity subset around the tiled paths, smoke(), the bench
# line of record, the secondary workloads, and -- if time is left -- one ncu --set full capture of the stride-864 kernel
TAG=${1:-x}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 240 python -m pytest tests -m gpu -q -k "tiled or big_tile or golden or sha256 or long_loci or span_sweep or randomized or multi_device" > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$? t=$SECONDS"; tail -4 gpurun_out/r02_pytest_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "t=$SECONDS"
timeout 300 python bench.py > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err
echo "bench rc=$? t=$SECONDS"; tail -c 300 gpurun_out/r02_bench_$TAG.err
Q="--steps 5 --warmup 3 --no-cpu --no-dropin --no-sha"
timeout 120 python bench.py --workload long --loci 1400 $Q > gpurun_out/r02_bench_long_$TAG.json 2> gpurun_out/r02_bench_long_$TAG.err
for L in 500 300 150; do
  timeout 120 python bench.py --workload sweep --loci 5000 --span $L $Q > gpurun_out/r02_bench_sweep_L${L}_$TAG.json 2> gpurun_out/r02_bench_sweep_L${L}_$TAG.err
done
echo "t=$SECONDS"
python - <<PY
import json
for f in ("bench","bench_long","bench_sweep_L150","bench_sweep_L300","bench_sweep_L500"):
    try:
        d=json.loads(open('gpurun_out/r02_%s_$TAG.json'%f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,2), round(d['e2e']['ms_per_step'],2), 'cells/s G', round(d['dp_cells_per_s']/1e9,2), 'fill', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4), d['stage_ms_serial_pass'], d.get('parity_in_run',{}).get('equal'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
if [ $SECONDS -lt 175 ]; then
  MIRFOLD_CHUNK_CELLS=1e12 MIRFOLD_SERIAL=1 timeout 80 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16ILi864 -s 1 -c 1 -o gpurun_out/r02_prof_fill864_$TAG -f python tools/ab_span.py sweep 3000 500 1 > gpurun_out/r02_prof_fill864_$TAG.log 2>&1
  echo "ncu rc=$? t=$SECONDS"; ls -la gpurun_out/r02_prof_fill864_$TAG.ncu-rep
fi
