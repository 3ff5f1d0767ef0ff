#!/bin/bash
# r02 call Q: scattered-read microbenchmark; host profile of call_inv_batch
mkdir -p gpurun_out
timeout 120 profiles/microbench/gather_rate.bin > gpurun_out/q_gather_rate.jsonl 2>&1; cat gpurun_out/q_gather_rate.jsonl
PROFILE=1 timeout 300 python profiles/run_inv_batch.py 1024 --out gpurun_out/q_inv.json 2>&1 | grep -v "^INV Found" > gpurun_out/q_inv_prof.txt
grep -n "cumulative" -A48 gpurun_out/q_inv_prof.txt | cut -c1-170 | head -70
