# r02 call 14: suite; density step + launch list (ref_insert with one atomic per distinct k-mer, sliced runs_stats, compact prefix).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02l_pytest.log
timeout 100 python profiles/run_density_c5.py 296 3; echo "c5 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2 -c 14 --csv --log-file gpurun_out/r02l_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02l_density_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows:
    agg.setdefault((r[0], r[4][:40]), {})[r[-3]] = r[-1]
for (i, k), v in agg.items():
    print(i, k, v)
PY
timeout 300 python profiles/run_inv_batch.py 1024 --out gpurun_out/r02l_inv_batch_1024.json 2>&1 | tail -1 | cut -c1-300
