#!/bin/bash
# 512^3: unfused (+ plane group) against ordered (+ merged work item), alternating, same box
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained --size 512"
for p in unfused ordered unfused ordered; do
timeout 600 $B --pipeline $p > gpurun_out/r2_b28.json 2>> gpurun_out/r2_b28.err
python -c "
import json; j=json.load(open('gpurun_out/r2_b28.json')); pk=j['per_kernel_ms']
print('512 $p', round(j['value'],1), round(j['ms_per_step'],3), {k: pk[k] for k in pk if k.startswith(('flux','emf','riemann'))}, j['clocks']['sm_mhz'], j['clocks']['power_w_max'])"
done
