# First GPU call of the next round: the GPU suite, then A/B lines for the switches that were wired in after round 1's GPU budget
# was spent.     gpurun --timeout 420 -- 'bash profiles/run_next_ab.sh'
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/next_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/next_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --density-windows 296"
timeout 120 python bench.py > gpurun_out/next_bench_default.json 2> gpurun_out/next_bench_default.err; echo "default rc=$?"
PAVGPU_L2_FETCH=32 timeout 90 python bench.py $LEAN > gpurun_out/next_bench_l2fetch32.json 2> gpurun_out/next_bench_l2fetch32.err; echo "l2fetch32 rc=$?"
PAVGPU_L2_FETCH=32 PAVGPU_HOMOLOGY_NBR=1 timeout 90 python bench.py $LEAN > gpurun_out/next_bench_l2fetch32_nbr.json 2> gpurun_out/next_bench_l2fetch32_nbr.err; echo "l2fetch32+nbr rc=$?"
PAVGPU_L2_FETCH=128 timeout 90 python bench.py $LEAN > gpurun_out/next_bench_l2fetch128.json 2> gpurun_out/next_bench_l2fetch128.err; echo "l2fetch128 rc=$?"
python - <<'PY'
import json
for f in ('default', 'l2fetch32', 'l2fetch32_nbr', 'l2fetch128'):
    try:
        j = json.loads(open('gpurun_out/next_bench_%s.json' % f).read().strip().splitlines()[-1])
        s = j.get('secondary') or {}
        print(f, 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], j['roofline']['per_kernel_ms'], 'density ms', s.get('ms_per_step'), s.get('ms_kmer'), s.get('ms_kde'))
    except Exception as e:
        print(f, 'ERR', e)
PY
