# r02 call 7: GPU suite; queue kernel with cooperative rests (4 / 8 / 16 lanes per item); ncu.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02f_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --density-windows 0"
for lib in default c4 c16; do
    L=""; [ "$lib" != default ] && L="$lib"
    PAVGPU_NO_GRAPH=1 PAVGPU_LIB=$L PAVGPU_HOMOLOGY=queue timeout 90 python bench.py $LEAN > gpurun_out/r02f_bench_${lib}_queue.json 2> gpurun_out/r02f_bench_${lib}_queue.err; echo "$lib rc=$?"
done
PAVGPU_HOMOLOGY=queue timeout 90 python bench.py $LEAN > gpurun_out/r02f_bench_default_queue_graph.json 2> gpurun_out/r02f_bench_default_queue_graph.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02f_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r02f_bench_')[1], 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], j['roofline']['per_kernel_ms'], 'parity', j['oracle_spot_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
PAVGPU_NO_GRAPH=1 PAVGPU_HOMOLOGY=queue timeout 200 ncu --set full --clock-control none --import-source on -k regex:homology -s 4 -c 1 -o gpurun_out/r02f_hom_queue python bench.py $LEAN --steps 3 --warmup 2 > gpurun_out/r02f_ncu_queue.log 2>&1; echo "ncu rc=$?"
