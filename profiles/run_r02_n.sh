#!/bin/bash
# r02 call N: on-chip k-mer tables (kmer_window_kernel), 4-row finish kernel, fused exp in the table scan.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -15 gpurun_out/n_pytest.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_density_gpu.py -q -x -k "onchip or batch_vs_oracle" > gpurun_out/n_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/n_memcheck.log
for oc in 1 0; do
  echo "== PAVGPU_DENSITY_ONCHIP=$oc" >> gpurun_out/n_density.log
  PAVGPU_DENSITY_ONCHIP=$oc timeout 100 python profiles/run_density_c5.py 296 4 >> gpurun_out/n_density.log 2>&1
done
cat gpurun_out/n_density.log | cut -c1-400
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2 -c 14 --csv --log-file gpurun_out/n_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/n_density_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows:
    agg.setdefault((r[0], r[4][:40]), {})[r[-3]] = r[-1]
for (i, k), v in agg.items():
    print(i, k, v)
PY
