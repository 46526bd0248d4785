#!/bin/bash
# x-z group kernel (mhd_xzgroup.inc, PPK_XZGROUP=1): the whole GPU suite with it on, then A/B against the two separate kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time PPK_XZGROUP=1 timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_xz_pytest.log 2>&1
tail -4 gpurun_out/r2_xz_pytest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
show() {
python -c "
import json,sys; j=json.load(open('gpurun_out/r2_b30.json')); pk=j['per_kernel_ms']
print('$1', round(j['value'],1), round(j['ms_per_step'],3), {k: round(pk[k],3) for k in pk if k.startswith(('flux','emf'))}, j['clocks']['sm_mhz'], j['clocks']['power_w_max'])" 2>&1 | tail -1
}
for size in 256 512; do
for xz in 0 1 0 1; do
PPK_XZGROUP=$xz timeout 300 $B --size $size > gpurun_out/r2_b30.json 2>> gpurun_out/r2_b30.err
show "$size xz=$xz"
done
done
for mb in 4 6; do
PPK_XZGROUP=1 PPK_XZ_MINB=$mb timeout 300 $B --size 512 > gpurun_out/r2_b30.json 2>> gpurun_out/r2_b30.err
show "512 xz=1 minb=$mb"
done
for slab in 24 96; do
PPK_XZGROUP=1 PPK_XZ_SLAB_MB=$slab timeout 300 $B --size 512 > gpurun_out/r2_b30.json 2>> gpurun_out/r2_b30.err
show "512 xz=1 slab_mb=$slab"
done
tail -5 gpurun_out/r2_b30.err
