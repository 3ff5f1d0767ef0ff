# r02: A/B of the homology kernels (gather / bulk / nbr) x scoring bodies (default lib = score_indel2, variant v1 = score_indel),
# CTA sizes of the bulk kernel, after the parity tests of all four kernels; one ncu capture of gather + bulk.
#   gpurun --timeout 600 -- 'bash profiles/run_r02_hom_ab.sh'
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cigar_gpu.py tests/test_random_gpu.py -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --density-windows 0"
for lib in default v1 b256 b64; do
  for k in gather bulk nbr; do
    if [ "$lib" != default ] && [ "$lib" != v1 ] && [ "$k" != bulk ]; then continue; fi
    L=""; [ "$lib" != default ] && L="$lib"
    PAVGPU_LIB=$L PAVGPU_HOMOLOGY=$k timeout 90 python bench.py $LEAN > gpurun_out/r02a_bench_${lib}_${k}.json 2> gpurun_out/r02a_bench_${lib}_${k}.err; echo "$lib $k rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02a_bench_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r02a_bench_')[1], 'value %.3e' % j['value'], 'ms %.4f' % j['ms_per_step'], j['roofline']['per_kernel_ms'], 'parity', j['oracle_spot_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
for k in gather bulk; do
  PAVGPU_HOMOLOGY=$k timeout 200 ncu --set full --clock-control none --import-source on -k regex:homology -s 4 -c 1 -o gpurun_out/r02a_hom_$k python bench.py $LEAN --steps 3 --warmup 2 > gpurun_out/r02a_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
