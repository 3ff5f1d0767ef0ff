#!/bin/bash
# r02 call P: rows per thread of finish_rows_kernel (default 4; variants 1, 2, 8); compact_kernel with its loads hoisted
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_density_gpu.py tests/test_inv_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in "" fin1 fin2 fin8; do
  echo "== variant '$v'"
  if [ -n "$v" ]; then export PAVGPU_LIB=$PWD/pav_b200/libpavgpu.$v.so; else unset PAVGPU_LIB; fi
  timeout 100 python profiles/run_density_c5.py 296 4 2>&1 | tail -1 | cut -c1-330
done
unset PAVGPU_LIB
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2 -c 12 --csv --log-file gpurun_out/p_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/p_density_launches.csv')):
    if len(r) > 10 and r[0].isdigit(): print(r[0], r[4][:40], r[-1])
PY
