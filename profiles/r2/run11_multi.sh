#!/bin/bash
# multi-GPU session (N GPUs of one box): z-slab bit-identity tests (log kept), weak scaling at 512^3 per GPU, the
# configs[3] strong-scaling field loop, per-kernel breakdown of the decomposed step
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NG=${1:-4}
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > gpurun_out/r2_multi_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "${TESTS:-z_slabs or blocks}" -v > gpurun_out/r2_t11_zslabs.log 2>&1
tail -n 14 gpurun_out/r2_t11_zslabs.log | cut -c1-400
run() { # n_gpus n extra-args tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --size $2 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 --no-extra $3 2>gpurun_out/r2_scale_err_$4.log > gpurun_out/r2_scale_$4.json
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_scale_$4.json')); print('$4', d['n_gpus'], round(d['value']), 'Mcell/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value']), 'sust', d['sustained'] and round(d['sustained']['value']), d['per_kernel_ms'], 'divB', d['sim']['max_divB_per_rank'])
except Exception as e: print('$4 failed', e)
"
}
if [ "${RUNS:-all}" = all ]; then
run 1 512 "" w512_N1
run 2 512 "" w512_N2
run $NG 512 "" w512_N$NG
run 2 256 "" w256_N2
fi
run $NG 256 "" w256_N$NG
run 1 512 "--workload field_loop --strong" fl512_N1
run 2 512 "--workload field_loop --strong" fl512_N2
run $NG 512 "--workload field_loop --strong" fl512_N$NG
for f in gpurun_out/r2_scale_err_*.log; do tail -n 2 $f; done | tail -n 20
