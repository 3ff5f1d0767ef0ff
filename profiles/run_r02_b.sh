# r02 call 3: parity of everything (device row counts, CUDA graph, queue kernel), then A/B gather / queue with and without the graph.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 2 --density-windows 0"
for k in gather queue; do
  for g in 0 1; do
    PAVGPU_NO_GRAPH=$g PAVGPU_HOMOLOGY=$k timeout 120 python bench.py $LEAN > gpurun_out/r02b_bench_${k}_nograph$g.json 2> gpurun_out/r02b_bench_${k}_nograph$g.err; echo "$k nograph=$g rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r02b_bench_')[1], 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], 'wall', '%.4f' % j['wall_ms_per_step_incl_flush'], j['roofline']['per_kernel_ms'], 'parity', j['oracle_spot_check'], 'cabi ms', j['e2e_cabi']['ms_per_step'], 'e2e', j['e2e']['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
PAVGPU_NO_GRAPH=1 PAVGPU_HOMOLOGY=queue timeout 200 ncu --set full --clock-control none --import-source on -k regex:homology -s 4 -c 1 -o gpurun_out/r02b_hom_queue python bench.py --no-cpu-baseline --e2e-steps 0 --density-windows 0 --steps 3 --warmup 2 > gpurun_out/r02b_ncu_queue.log 2>&1; echo "ncu queue rc=$?"
PAVGPU_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on -s 0 -c 12 -o gpurun_out/r02b_density python profiles/run_density_c5.py 296 1 > gpurun_out/r02b_ncu_density.log 2>&1; echo "ncu density rc=$?"
