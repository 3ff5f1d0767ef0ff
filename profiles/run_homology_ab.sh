set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/s16_pytest.log
PAVGPU_HOMOLOGY_TILED=1 timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_cigar_gpu.py -q -k "golden_gpu_tiled or kernels_agree" > gpurun_out/s16_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/s16_memcheck.log
timeout 120 python bench.py > gpurun_out/s16_bench_auto.json 2> gpurun_out/s16_bench_auto.err; echo "bench rc=$?"
PAVGPU_HOMOLOGY_TILED=0 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 0 --density-windows 0 > gpurun_out/s16_bench_gather.json 2> gpurun_out/s16_bench_gather.err; echo "bench gather rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:homology_tiled -s 3 -c 1 -f -o gpurun_out/s16_homtiled python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --density-windows 0 > gpurun_out/s16_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ('auto','gather'):
    try:
        j=json.loads(open('gpurun_out/s16_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, j['value'], j['ms_per_step'], j['roofline']['per_kernel_ms'], j['roofline']['kernel'], j['roofline']['frac'], 'e2e', j['e2e']['value'], 'cabi', j['e2e_cabi']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
