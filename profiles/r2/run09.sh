#!/bin/bash
# producer with one-instruction store addresses; e2e with the fourth rotating array
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=09
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload_256" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 8 --no-extra --no-sustained"
PPK_RALL_PERS=0 timeout 600 $B --n 256 --pipeline tiled > gpurun_out/r2_b${T}_256_tiled.json 2>> gpurun_out/r2_b$T.err
PPK_RALL_PERS=0 timeout 600 $B --n 256 --pipeline unfused > gpurun_out/r2_b${T}_256_unfused.json 2>> gpurun_out/r2_b$T.err
PPK_RALL_PERS=0 timeout 600 $B --n 512 --pipeline tiled > gpurun_out/r2_b${T}_512_tiled.json 2>> gpurun_out/r2_b$T.err
python - <<PY
import json
for n in ("256_tiled","256_unfused","512_tiled"):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"], "e2e", round(j["e2e"]["value"],1), j["e2e"].get("pcie_GBs_per_gpu_each_way"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2_b$T.err
