# r02 call 9: cigar suite; density e2e profile at 2048 windows; C3 e2e host-settings sweep.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02h_pytest.log
cat gpurun_out/r02_kern_error.json | tail -5
PROFILE=1 timeout 200 python profiles/run_density_e2e_trace.py 2048 3 > gpurun_out/r02h_density_e2e.log 2>&1; echo "density e2e rc=$?"; grep -v "^---" gpurun_out/r02h_density_e2e.log | head -50
timeout 400 python profiles/run_e2e_c3_sweep.py > gpurun_out/r02h_c3_sweep.log 2>&1; echo "sweep rc=$?"; grep readers gpurun_out/r02h_c3_sweep.log
