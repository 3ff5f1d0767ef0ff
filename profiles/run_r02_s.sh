#!/bin/bash
# r02 call S: full GPU suite (incl. the rule-level drop-in tests), bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -6 gpurun_out/s_pytest.log
python bench.py > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err; echo "bench rc=$?"
grep "e2e step" gpurun_out/s_bench_n1.err | cut -c1-400
python - <<'PY'
import json
j = json.loads(open('gpurun_out/s_bench_n1.json').read().strip().splitlines()[-1])
print('value %.3e ms %.4f e2e %.3e (%.0f ms)' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']))
s = j['secondary']; print('C5 value', s['value'], 'ms', s['ms_per_step'], 'e2e', s['e2e']['value'])
c = j['c2']; print('C2', c['ms_per_step'], c['e2e']['value'], c['e2e_cabi']['value'])
PY
