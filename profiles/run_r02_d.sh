# r02 call 5: GPU suite with the N summaries; A/B nsum on/off, batched loads, occupancy; density step; ncu of the queue kernel.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02d_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --density-windows 296"
for lib in default nosum batch q3; do
  for k in gather queue; do
    L=""; [ "$lib" != default ] && L="$lib"
    PAVGPU_NO_GRAPH=1 PAVGPU_LIB=$L PAVGPU_HOMOLOGY=$k timeout 90 python bench.py $LEAN > gpurun_out/r02d_bench_${lib}_${k}.json 2> gpurun_out/r02d_bench_${lib}_${k}.err; echo "$lib $k rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02d_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        s = j.get('secondary') or {}
        print(f.split('r02d_bench_')[1], 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], j['roofline']['per_kernel_ms'], 'parity', j['oracle_spot_check'], 'density', s.get('ms_per_step'), s.get('ms_kmer'), s.get('ms_kde'), s.get('oracle_spot_check'))
    except Exception as e:
        print(f, 'ERR', e)
PY
PAVGPU_NO_GRAPH=1 PAVGPU_HOMOLOGY=queue timeout 200 ncu --set full --clock-control none --import-source on -k regex:homology -s 4 -c 1 -o gpurun_out/r02d_hom_queue python bench.py --no-cpu-baseline --e2e-steps 0 --density-windows 0 --steps 3 --warmup 2 > gpurun_out/r02d_ncu_queue.log 2>&1; echo "ncu queue rc=$?"
