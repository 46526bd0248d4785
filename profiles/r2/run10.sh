#!/bin/bash
# checkpoint: whole GPU suite, hardware counters of the default pipelines, the default bench line (512^3) as the driver runs it
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t10.log 2>&1
tail -n 3 gpurun_out/r2_t10.log
bash profiles/r2/counters.sh 256 unfused
bash profiles/r2/counters.sh 512 ordered
bash profiles/r2/counters.sh 256 tiled
( time python bench.py > gpurun_out/r2_b10_default.json 2> gpurun_out/r2_b10.err ) 2>&1 | grep real
python - <<PY
import json
j=json.load(open("gpurun_out/r2_b10_default.json"))
print(round(j["value"],1), "Mcell/s", round(j["ms_per_step"],3), "ms", j["details"]["pipeline"], j["per_kernel_ms"])
print("sustained", j["sustained"]); print("e2e", j["e2e"]["value"], "extra", j["extra"], "cpu", j["cpu_baseline"], "refcuda", j["reference_cuda_baseline"])
print("roofline", j["roofline"]); print("clocks", j["clocks"])
PY
tail -5 gpurun_out/r2_b10.err
