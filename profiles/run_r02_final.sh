#!/bin/bash
# r02 final: full GPU suite, both bench arms, launch lists, ncu --set full of the Path-A kernels on C2, call_inv_batch at 1,024 loci
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -4 gpurun_out/r02_pytest_gpu_final.log
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r02_bench_n1_reference_arm_final.json 2> gpurun_out/r02_bench_n1_reference_arm_final.err; echo "ref rc=$?"
timeout 300 python profiles/run_inv_batch.py 1024 --out gpurun_out/r02_inv_batch_1024.json 2>&1 | grep -v "^INV Found" | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python profiles/run_walk_c2.py 3 > /dev/null 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cigar_count_kernel|cigar_walk_kernel|homology_queue_kernel" -s 6 -c 3 -o gpurun_out/r02_cigar_final -f python profiles/run_walk_c2.py 4 > gpurun_out/r02_ncu_cigar_final.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[-1])
print('value %.3e ms %.4f e2e %.3e (%.0f ms)' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']), j['roofline']['per_kernel_ms'], 'frac', j['roofline']['frac'], 'gather frac', j['roofline']['gather']['frac'])
s = j['secondary']; print('C5 value', s['value'], 'ms', s['ms_per_step'], 'e2e', s['e2e']['value'], s['roofline']['kmer_part']['frac'])
c = j['c2']; print('C2', c['ms_per_step'], c['roofline']['per_kernel_ms'], c['e2e']['value'], c['e2e_cabi']['value'], c['roofline']['gather']['frac'], c['roofline']['traffic'], c['roofline']['frac'])
r = json.loads(open('gpurun_out/r02_bench_n1_reference_arm_final.json').read().strip().splitlines()[-1])
print('ref arm', r['value'], r['cpu_baseline']['cores'], r['secondary']['value'])
PY
