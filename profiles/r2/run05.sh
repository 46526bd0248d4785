#!/bin/bash
# A/B of the persistent Riemann kernel (PPK_RALL_PERS) on the tiled pipeline, 256^3 and 512^3
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=${1:-05}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or golden or fast_mode" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 --no-extra --no-sustained --pipeline tiled"
for n in 256 512; do
for pers in 0 1; do
PPK_RALL_PERS=$pers timeout 600 $B --n $n > gpurun_out/r2_b${T}_${n}_pers$pers.json 2>> gpurun_out/r2_b$T.err
done; done
python - <<PY
import json
for n in ("256_pers0","256_pers1","512_pers0","512_pers1"):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"], "e2e", round(j["e2e"]["value"],1))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2_b$T.err
