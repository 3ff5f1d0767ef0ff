# r02: two GPUs -- the sharded product path: 2-GPU test (no skip), then the default bench at N = 2.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rs > gpurun_out/r02_n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_n2_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ) 2>&1 | tail -4; echo "bench rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_bench_n2.err | tail -12 | cut -c1-400
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
    print("value %.3e" % j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], "e2e_dist", (j.get("e2e_dist") or {}).get("value"), (j.get("e2e_dist") or {}).get("ms_per_step"), "parity", j["oracle_spot_check"], "planes", j["planes_verified_on_every_rank"], "bcast ms", j["ref_broadcast_ms"])
    for p in j['per_rank']: print(p)
    s = j['secondary']; print('C5 value', s.get('value'), 'e2e', s.get('e2e', {}).get('value'), s.get('oracle_spot_check'), [(p['rank'], p['ms'], p['windows']) for p in s['per_rank']])
except Exception as e:
    print('ERR', e)
PY
