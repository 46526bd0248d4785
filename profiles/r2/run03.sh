#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=${1:-03}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --e2e-depth 1"
$B --pipeline tiled > gpurun_out/r2_b${T}_tiled.json 2> gpurun_out/r2_b$T.err
python - <<PY
import json
for n in ("tiled",):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"])
    except Exception as e:
        print(n, "failed", e)
PY
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_producer -s 2 -c 1 -o gpurun_out/r2_prof${T}_producer -f python profiles/r2/mini.py 256 4 tiled > gpurun_out/r2_ncu_producer.log 2>&1
$NCU -k regex:k_riemann_all -s 2 -c 1 -o gpurun_out/r2_prof${T}_rall -f python profiles/r2/mini.py 256 4 tiled > gpurun_out/r2_ncu_rall.log 2>&1
