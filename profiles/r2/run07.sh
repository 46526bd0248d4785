#!/bin/bash
# A/B of the persistent Riemann kernel v2 (mbarrier hand-off, shared solver bodies): single-shot | 5 CTAs/SM | 4 CTAs/SM
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=07
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or golden" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 6 --no-extra --no-sustained --pipeline tiled --n 256"
PPK_RALL_PERS=0 timeout 600 $B > gpurun_out/r2_b${T}_pers0.json 2>> gpurun_out/r2_b$T.err
PPK_RALL_PERS=1 PPK_RALL_CTAS=5 timeout 600 $B > gpurun_out/r2_b${T}_pers5.json 2>> gpurun_out/r2_b$T.err
PPK_RALL_PERS=1 PPK_RALL_CTAS=4 timeout 600 $B > gpurun_out/r2_b${T}_pers4.json 2>> gpurun_out/r2_b$T.err
python - <<PY
import json
for n in ("pers0","pers5","pers4"):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"], "e2e", round(j["e2e"]["value"],1))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2_b$T.err
