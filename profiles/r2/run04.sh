#!/bin/bash
# round-2 checkpoint: whole GPU suite + bench at 256^3 (tiled / unfused) and the default 512^3 line
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t04.log 2>&1
tail -n 3 gpurun_out/r2_t04.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 --no-extra"
timeout 600 $B --n 256 --pipeline tiled > gpurun_out/r2_b04_256_tiled.json 2> gpurun_out/r2_b04.err
timeout 600 $B --n 256 --pipeline unfused > gpurun_out/r2_b04_256_unfused.json 2>> gpurun_out/r2_b04.err
timeout 900 $B --n 512 --pipeline tiled > gpurun_out/r2_b04_512_tiled.json 2>> gpurun_out/r2_b04.err
timeout 900 $B --n 512 --pipeline unfused > gpurun_out/r2_b04_512_unfused.json 2>> gpurun_out/r2_b04.err
python - <<PY
import json
for n in ("256_tiled","256_unfused","512_tiled","512_unfused"):
    try:
        j=json.load(open(f"gpurun_out/r2_b04_{n}.json"))
        print(n, round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["per_kernel_ms"], "sust", j.get("sustained",{}) and round(j["sustained"]["value"],1), "e2e", round(j["e2e"]["value"],1))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2_b04.err
