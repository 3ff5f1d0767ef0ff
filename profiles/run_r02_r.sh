#!/bin/bash
# r02 call R: full GPU suite, call_inv_batch at 1,024 loci, bench N=1 (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log; tail -4 gpurun_out/r_pytest.log
timeout 300 python profiles/run_inv_batch.py 1024 --out gpurun_out/r_inv_batch_1024.json 2>&1 | grep -v "^INV Found" | tail -1 | cut -c1-400
python bench.py > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r_bench_ref.json 2> gpurun_out/r_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r_bench_n1.json').read().strip().splitlines()[-1])
print('value %.3e ms %.4f e2e %.3e (%.0f ms)' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']), j['roofline']['per_kernel_ms'], 'frac', j['roofline']['frac'], 'gather', j['roofline']['gather'])
s = j['secondary']; print('C5 value', s['value'], 'ms', s['ms_per_step'], 'e2e', s['e2e']['value'], s['e2e']['phase_seconds_last_step_rank0'], s['roofline']['kmer_part']['frac'])
c = j['c2']; print('C2', c['ms_per_step'], c['roofline']['per_kernel_ms'], c['e2e']['value'], c['e2e_cabi']['value'], c['roofline']['gather'], c['roofline']['traffic'])
print('cpu_baseline', j['cpu_baseline']['value'], s['cpu_baseline']['value'])
r = json.loads(open('gpurun_out/r_bench_ref.json').read().strip().splitlines()[-1])
print('ref arm', r['value'], r['cpu_baseline']['cores'], r['secondary']['value'])
PY
