#!/bin/bash
# A/B: TMA-staged trace kernel (PPK_TRACE_TMA=1, new) against the plain-load k_trace; parity subset first
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T=20
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload_128 or (exact_mode_bit_identical and unfused)" > gpurun_out/r2_t$T.log 2>&1
tail -n 3 gpurun_out/r2_t$T.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
for tm in 0 1; do
PPK_TRACE_TMA=$tm timeout 600 $B --size 256 > gpurun_out/r2_b${T}_256_$tm.json 2>> gpurun_out/r2_b$T.err
PPK_TRACE_TMA=$tm timeout 600 $B --size 512 > gpurun_out/r2_b${T}_512_$tm.json 2>> gpurun_out/r2_b$T.err
done
python - <<PY
import json
for n in ("256_0","256_1","512_0","512_1"):
    try:
        j=json.load(open(f"gpurun_out/r2_b${T}_{n}.json"))
        print(n, j["details"]["pipeline"], round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms trace", j["per_kernel_ms"]["trace"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2_b$T.err
