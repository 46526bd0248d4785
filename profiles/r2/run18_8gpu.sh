#!/bin/bash
# 8-GPU session: 2x2x2 block bit-identity, weak scaling 512^3 per GPU at N=8 (and N=2 again with the sampler fix), strong field loop N=8
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > gpurun_out/r2_8gpu_box.txt; free -g | head -2 >> gpurun_out/r2_8gpu_box.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shape3" -v > gpurun_out/r2_t18_blocks8.log 2>&1
tail -n 6 gpurun_out/r2_t18_blocks8.log | cut -c1-300
run() { # n_gpus n extra-args tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --size $2 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 --no-extra $3 2>gpurun_out/r2_scale8_err_$4.log > gpurun_out/r2_scale8_$4.json
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_scale8_$4.json')); print('$4', d['n_gpus'], round(d['value']), 'Mcell/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value']), 'sust', d['sustained'] and round(d['sustained']['value']), d['per_kernel_ms'], 'divB', max(d['sim']['max_divB_per_rank']))
except Exception as e: print('$4 failed', e)
"
}
run 8 512 "" w512_N8
run 2 512 "" w512_N2
run 8 512 "--workload field_loop --strong" fl512_N8
run 8 256 "" w256_N8
for f in gpurun_out/r2_scale8_err_*.log; do grep -i "error\|exceeds\|Traceback" $f | head -3; done
