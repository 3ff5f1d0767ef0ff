# r02 call 8: full suite; default bench (C3 + C5 + C2 + CPU baseline) at N=1 with wall time; reference arm, short.
set -x
mkdir -p gpurun_out
nproc; free -g | head -2; df -h /tmp | tail -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02g_pytest.log
( time timeout 900 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err ) 2>&1 | tail -4; echo "bench rc=$?"
tail -25 gpurun_out/r02g_bench_n1.err
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02g_bench_n1.json').read().strip().splitlines()[-1])
    print('value %.3e' % j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'parity', j['oracle_spot_check'], 'roof', j['roofline']['per_kernel_ms'], j['roofline']['frac'])
    print('cpu', j['cpu_baseline'])
    s = j['secondary']; print('C5 value', s.get('value'), 'e2e', s.get('e2e', {}).get('value'), s.get('oracle_spot_check'), s.get('cpu_baseline'))
    c = j['c2']; print('C2', c.get('value'), c.get('ms_per_step'), c.get('e2e', {}).get('ms_per_step'), c.get('e2e_cabi', {}).get('ms_per_step'), c.get('roofline', {}).get('per_kernel_ms'), c.get('error'))
except Exception as e:
    print('ERR', e)
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err ) 2>&1 | tail -4
tail -8 gpurun_out/r02g_bench_ref.err; head -c 1500 gpurun_out/r02g_bench_ref.json
