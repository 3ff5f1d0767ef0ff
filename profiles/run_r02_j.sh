# r02 call 11: debug small_rev_cluster; suite; density step (cluster k-mer tables) + e2e + launch list; inv batch profile.
set -x
mkdir -p gpurun_out
timeout 100 python profiles/dbg_small_rev_cluster.py 2>&1 | tail -25
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02j_pytest.log
timeout 100 python profiles/run_density_c5.py 296 3; echo "c5 rc=$?"
PAVGPU_DENSITY_GLOBAL_TABLES=1 timeout 100 python profiles/run_density_c5.py 296 3; echo "c5 global rc=$?"
timeout 200 python profiles/run_density_e2e_trace.py 2048 3 > gpurun_out/r02j_density_e2e.log 2>&1; echo "density e2e rc=$?"; grep "^call" gpurun_out/r02j_density_e2e.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2 -c 14 --csv --log-file gpurun_out/r02j_density_launches.csv python profiles/run_density_c5.py 296 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02j_density_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows:
    agg.setdefault((r[0], r[4][:40]), {})[r[-3]] = r[-1]
for (i, k), v in agg.items():
    print(i, k, v)
PY
PROFILE=1 timeout 300 python profiles/run_inv_batch.py 512 > gpurun_out/r02j_inv_batch.log 2>&1; echo "inv rc=$?"; head -60 gpurun_out/r02j_inv_batch.log | cut -c1-200; tail -2 gpurun_out/r02j_inv_batch.log | cut -c1-600
