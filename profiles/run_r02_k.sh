# r02 call 12: suite; density step; queue kernel with / without L2 prefetch; inv batch timing.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02k_pytest.log
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02_kern_error.json'))
print({k: v for k, v in r.items() if k.startswith('_')})
PY
timeout 100 python profiles/run_density_c5.py 296 3; echo "c5 rc=$?"
LEAN="--no-cpu-baseline --e2e-steps 0 --c5-windows 0 --scale 0.01 --steps 20"
for lib in default pf; do
    L=""; [ "$lib" != default ] && L="$lib"
    PAVGPU_LIB=$L timeout 200 python bench.py $LEAN > gpurun_out/r02k_bench_${lib}.json 2> gpurun_out/r02k_bench_${lib}.err; echo "$lib rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02k_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        c = j['c2']
        print(f, 'C2 value %.3e' % c['value'], 'ms %.4f' % c['ms_per_step'], c['roofline']['per_kernel_ms'], c['oracle_spot_check'], 'C3small', j['value'], j['oracle_spot_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python profiles/run_inv_batch.py 1024 --out gpurun_out/r02k_inv_batch_1024.json 2>&1 | tail -1 | cut -c1-700
