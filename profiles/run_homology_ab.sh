# A/B of the homology kernels on one box (gathers = default, PAVGPU_HOMOLOGY_TILED=1, PAVGPU_HOMOLOGY_NBR=1), after the GPU suite.
#   gpurun --timeout 300 -- 'bash profiles/run_homology_ab.sh'
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/s17_pytest.log
timeout 120 python bench.py > gpurun_out/s17_bench_default.json 2> gpurun_out/s17_bench_default.err; echo "bench rc=$?"
PAVGPU_HOMOLOGY_NBR=1 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 0 --density-windows 0 > gpurun_out/s17_bench_nbr.json 2> gpurun_out/s17_bench_nbr.err; echo "bench nbr rc=$?"
python - <<'PY'
import json
for f in ('default', 'nbr'):
    try:
        j = json.loads(open('gpurun_out/s17_bench_%s.json' % f).read().strip().splitlines()[-1])
        print(f, j['value'], j['ms_per_step'], j['roofline']['per_kernel_ms'], j['roofline']['kernel'], j['roofline']['frac'], 'e2e', j['e2e']['value'], 'cabi', j['e2e_cabi']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
