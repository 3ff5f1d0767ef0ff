# r02 call 6: GPU suite (split kernels, fetch_runs, lazy windows); A/B gather / queue / split, graph on; density e2e trace.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02e_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --density-windows 0"
for k in gather queue split; do
  for g in 0 1; do
    PAVGPU_NO_GRAPH=$g PAVGPU_HOMOLOGY=$k timeout 90 python bench.py $LEAN > gpurun_out/r02e_bench_${k}_nograph$g.json 2> gpurun_out/r02e_bench_${k}_nograph$g.err; echo "$k nograph=$g rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02e_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r02e_bench_')[1], 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], j['roofline']['per_kernel_ms'], 'parity', j['oracle_spot_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 120 python profiles/run_density_e2e_trace.py > gpurun_out/r02e_density_e2e.log 2>&1; echo "density e2e rc=$?"; tail -12 gpurun_out/r02e_density_e2e.log
PAVGPU_NO_GRAPH=1 PAVGPU_HOMOLOGY=split timeout 200 ncu --set full --clock-control none --import-source on -k regex:homology -s 12 -c 3 -o gpurun_out/r02e_hom_split python bench.py $LEAN --steps 3 --warmup 2 > gpurun_out/r02e_ncu_split.log 2>&1; echo "ncu split rc=$?"
