#!/bin/bash
# round 2, first GPU run: parity of the tiled pipeline + A/B of its two halves at 256^3
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(nproc; free -g; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/r2_host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates" > gpurun_out/r2_t01.log 2>&1
tail -5 gpurun_out/r2_t01.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --e2e-depth 1"
$B --pipeline unfused > gpurun_out/r2_b01_unfused.json 2> gpurun_out/r2_b01.err
$B --pipeline tiled > gpurun_out/r2_b01_tiled.json 2>> gpurun_out/r2_b01.err
PPK_RALL=0 $B --pipeline tiled > gpurun_out/r2_b01_tiled_norall.json 2>> gpurun_out/r2_b01.err
python - <<'PY'
import json
for n in ("unfused","tiled","tiled_norall"):
    try:
        j=json.load(open(f"gpurun_out/r2_b01_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"])
    except Exception as e:
        print(n, "failed", e)
PY
