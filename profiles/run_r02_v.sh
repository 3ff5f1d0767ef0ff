#!/bin/bash
# r02 call V: rolling k-mers in kmer_window_kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_density_gpu.py tests/test_inv_gpu.py tests/test_flag_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_density_gpu.py -q -x -k "onchip" 2>&1 | tail -3
timeout 100 python profiles/run_density_c5.py 296 4 2>&1 | tail -1 | cut -c1-330
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2 -c 12 --csv --log-file gpurun_out/v_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/v_density_launches.csv')):
    if len(r) > 10 and r[0].isdigit(): print(r[0], r[4][:40], r[-1])
PY
