#!/bin/bash
# r02 call M: full GPU suite after the fused finish kernels, then A/B of table waves for the density step, then bench N=1.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -3 gpurun_out/m_pytest.log
for mb in 0 24 48 96; do
  echo "== PAVGPU_DENSITY_WAVE_MB=$mb" >> gpurun_out/m_density.log
  PAVGPU_DENSITY_WAVE_MB=$mb python profiles/run_density_c5.py 296 4 >> gpurun_out/m_density.log 2>&1
done
grep -E "==|ms_total|'ms" gpurun_out/m_density.log | cut -c1-600
python bench.py > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err; tail -c 3000 gpurun_out/m_bench_n1.json
