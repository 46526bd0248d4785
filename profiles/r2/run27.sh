#!/bin/bash
# A/B: merged plane-group work item inside k_riemann_all (PPK_RALL_MERGE) at 512^3 ordered and 256^3 ordered / tiled
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=27
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
for cfg in "512 ordered 0" "512 ordered 1" "256 ordered 0" "256 ordered 1" "256 unfused 1" "512 unfused 1"; do
set -- $cfg
PPK_RALL_MERGE=$3 timeout 600 $B --size $1 --pipeline $2 > gpurun_out/r2_b${T}.json 2>> gpurun_out/r2_b$T.err
python -c "
import json; j=json.load(open('gpurun_out/r2_b${T}.json')); pk=j['per_kernel_ms']
print('$1 $2 merge $3', round(j['value'],1), round(j['ms_per_step'],3), {k: pk[k] for k in pk if k.startswith(('flux','emf','riemann'))}, j['clocks']['sm_mhz'])"
done
tail -3 gpurun_out/r2_b$T.err
