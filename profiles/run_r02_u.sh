#!/bin/bash
# r02 call U: count kernel with four chunks per warp
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cigar_gpu.py tests/test_random_gpu.py tests/test_fullsize_gpu.py tests/test_sidecar_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_cigar_gpu.py -q -x -k "golden and not tiled and not nbr" 2>&1 | tail -3
python bench.py --c5-windows 296 --e2e-steps 0 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1])
print('  C3 ms %.4f' % j['ms_per_step'], {k: round(v, 4) for k, v in j['roofline']['per_kernel_ms'].items()}, ' C2 ms %.4f' % j['c2']['ms_per_step'], {k: round(v, 4) for k, v in j['c2']['roofline']['per_kernel_ms'].items()})
PY
