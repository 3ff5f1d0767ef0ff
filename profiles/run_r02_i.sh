# r02 call 10: suite (suffix-sum KDE table, exact KERN goldens, threaded staging, single-copy runs); density step + e2e; kernel list.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02i_pytest.log
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02_kern_error.json'))
print({k: v for k, v in r.items() if k.startswith('_')})
PY
timeout 100 python profiles/run_density_c5.py 296 3; echo "c5 rc=$?"
timeout 200 python profiles/run_density_e2e_trace.py 2048 3 > gpurun_out/r02i_density_e2e.log 2>&1; echo "density e2e rc=$?"; grep "^call" gpurun_out/r02i_density_e2e.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2 -c 14 --csv --log-file gpurun_out/r02i_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02i_density_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows:
    agg.setdefault((r[0], r[4][:40]), {})[r[-3]] = r[-1]
for (i, k), v in agg.items():
    print(i, k, v)
PY
