#!/bin/bash
# single-GPU: field loop / blast workloads at 512^3 (configs[2], [3] at N=1), counters with the final defaults, adapter test
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_adapter.py tests/test_gpu_parity.py -q -m gpu -k "adapter or error_paths" > gpurun_out/r2_t12.log 2>&1; tail -n 3 gpurun_out/r2_t12.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 --no-extra --no-sustained"
for w in field_loop blast; do
timeout 600 $B --workload $w > gpurun_out/r2_b12_$w.json 2> gpurun_out/r2_b12_$w.err
python -c "
import json
try:
    j=json.load(open('gpurun_out/r2_b12_$w.json')); print('$w', round(j['value'],1), round(j['ms_per_step'],2), j['sim'], j['per_kernel_ms'])
except Exception as e: print('$w failed', e)
"; tail -n 2 gpurun_out/r2_b12_$w.err
done
bash profiles/r2/counters.sh 256 unfused
bash profiles/r2/counters.sh 512 ordered
bash profiles/r2/counters.sh 256 tiled
