#!/bin/bash
# round 2, 1-GPU pass (bash tools/gpu_round2.sh TAG through gpurun): parity suite, the bench line of record + reference arm, secondary workloads, span sweep with ncu,
# launch list and ncu --set full of the dominant kernel
TAG=${1:-k}
mkdir -p gpurun_out
export MIRFOLD_CORPUS_DIR=$PWD/.corpus_cache
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_$TAG.json 2> gpurun_out/r02_bench_reference_$TAG.err
Q="--steps 5 --warmup 3 --no-cpu --no-dropin"
timeout 600 python bench.py --workload parity $Q > gpurun_out/r02_bench_parity_$TAG.json 2> gpurun_out/r02_bench_parity_$TAG.err
timeout 600 python bench.py --workload long --loci 1400 $Q --no-sha > gpurun_out/r02_bench_long_$TAG.json 2> gpurun_out/r02_bench_long_$TAG.err
for L in 150 300 500; do
  timeout 600 python bench.py --workload sweep --loci 5000 --span $L $Q --no-sha > gpurun_out/r02_bench_sweep_L${L}_$TAG.json 2> gpurun_out/r02_bench_sweep_L${L}_$TAG.err
done
python - <<PY
import json
for f in ("bench","bench_parity","bench_long","bench_sweep_L150","bench_sweep_L300","bench_sweep_L500"):
    try:
        d=json.loads(open('gpurun_out/r02_%s_$TAG.json'%f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,2), round(d['e2e']['ms_per_step'],2), 'cells/s G', round(d['dp_cells_per_s']/1e9,2), 'fill', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4), d['stage_ms_serial_pass'], d.get('parity_in_run',{}).get('equal'), d.get('cpu_baseline',{}).get('value'), d.get('drop_in'))
    except Exception as e:
        print(f, 'ERR', e)
PY
# ncu: span sweep (smem vs L2 throughput of the fill kernel per span), then the launch list and the full capture of the bench's own kernel
export MIRFOLD_CHUNK_CELLS=1e12 MIRFOLD_SERIAL=1
if [ -n "$SWEEP_NCU" ]; then
for L in 150 300 500; do
  timeout 600 ncu --set full --clock-control none -k regex:k_fill_s16 -s 2 -c 2 -o gpurun_out/r02_prof_sweep_L${L}_$TAG -f python bench.py --workload sweep --loci 3000 --span $L --steps 1 --warmup 1 --no-cpu --no-sha --no-dropin > gpurun_out/r02_prof_sweep_L${L}_$TAG.log 2>&1
done
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traceback -s 1 -c 1 -o gpurun_out/r02_prof_tb_$TAG -f python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu --no-sha --no-dropin > gpurun_out/r02_prof_tb_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_s16 -s 1 -c 1 -o gpurun_out/r02_prof_fill352_$TAG -f python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu --no-sha --no-dropin > gpurun_out/r02_prof_fill_$TAG.log 2>&1
unset MIRFOLD_CHUNK_CELLS MIRFOLD_SERIAL
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_$TAG.csv python bench.py --loci 40000 --steps 2 --warmup 1 --no-cpu --no-sha --no-dropin > gpurun_out/r02_b_ncu_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
