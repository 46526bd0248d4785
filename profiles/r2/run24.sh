#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hdf5_mock.py tests/test_host_layer.py -q -x > gpurun_out/r2_t24.log 2>&1; tail -n 5 gpurun_out/r2_t24.log
timeout 600 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 > gpurun_out/r2_b24.json 2> gpurun_out/r2_b24.err
python -c "
import json; j=json.load(open('gpurun_out/r2_b24.json')); print(round(j['value'],1), j['config'], j['details'], j['roofline']['per_kernel_measured_dram'])"
tail -n 3 gpurun_out/r2_b24.err
