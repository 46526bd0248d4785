#!/bin/bash
# A/B: TMA boxes issued by lane 0 of every warp (new) -- per-kernel times of unfused 256^3, ordered 512^3, + parity subset
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=16
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload_128" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
timeout 600 $B --size 256 > gpurun_out/r2_b${T}_256.json 2>> gpurun_out/r2_b$T.err
timeout 600 $B --size 256 --pipeline ordered > gpurun_out/r2_b${T}_256o.json 2>> gpurun_out/r2_b$T.err
timeout 600 $B --size 512 > gpurun_out/r2_b${T}_512.json 2>> gpurun_out/r2_b$T.err
python - <<PY
import json
for n in ("256","256o","512"):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, j["details"]["pipeline"], round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2_b$T.err
