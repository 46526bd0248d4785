#!/bin/bash
# A/B: x-faces + y-faces + z-edges on one staged tile (k_plane_group, PPK_PGROUP=1) against the three separate TMA kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=26
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload_128 or (exact_mode_bit_identical and unfused) or (fast_mode and unfused)" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained --size 256 --pipeline unfused"
for on in 0 1 0 1; do
PPK_PGROUP=$on timeout 600 $B > gpurun_out/r2_b${T}_$on.json 2>> gpurun_out/r2_b$T.err
python -c "
import json; j=json.load(open('gpurun_out/r2_b${T}_$on.json')); pk=j['per_kernel_ms']
print('pgroup $on', round(j['value'],1), round(j['ms_per_step'],3), {k: pk[k] for k in pk if k.startswith(('flux','emf'))}, j['clocks']['sm_mhz'])"
done
tail -3 gpurun_out/r2_b$T.err
