#!/bin/bash
# r02: bench at N GPUs (torchrun, one rank per GPU):  bash profiles/run_r02_nN.sh N
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err ) 2>&1 | tail -4; echo "bench rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_bench_n$N.err | grep "e2e step 2\|e2e_dist step 2" | cut -c1-700
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    j = json.loads(open(f'gpurun_out/r02_bench_n{N}.json').read().strip().splitlines()[-1])
    print("value %.3e" % j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], "e2e_dist", (j.get("e2e_dist") or {}).get("value"), (j.get("e2e_dist") or {}).get("ms_per_step"), "parity", j["oracle_spot_check"], "planes", j["planes_verified_on_every_rank"], "bcast ms", j["ref_broadcast_ms"])
    for p in j['per_rank']: print({k: p[k] for k in ('rank', 'records', 'rows', 'ms_per_step', 'ms_count_scan', 'ms_walk', 'ms_homology', 'planes_checksum_equal_rank0', 'oracle_spot_check')})
    s = j['secondary']; print('C5 value', s.get('value'), 'e2e', s.get('e2e', {}).get('value'), s.get('oracle_spot_check'), [(p['rank'], round(p['ms'], 2), p['windows']) for p in s['per_rank']])
except Exception as e:
    print('ERR', e)
PY
