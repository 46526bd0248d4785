#!/bin/bash
# what the driver runs at round end, on one GPU: smoke, the whole -m gpu suite, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1; tail -n 6 gpurun_out/r2_final_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_final_pytest.log 2>&1; tail -n 6 gpurun_out/r2_final_pytest.log
( time python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err ) 2>&1 | grep real
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final2_ref.json 2> gpurun_out/r2_final2_ref.err ) 2>&1 | grep real
python - <<PY
import json
j=json.load(open("gpurun_out/r2_final2_bench.json")); r=json.load(open("gpurun_out/r2_final2_ref.json"))
print("ours", round(j["value"],1), j["unit"], round(j["ms_per_step"],2), "ms", j["details"]["pipeline"], "sust", round(j["sustained"]["value"],1), "e2e", round(j["e2e"]["value"],1), "n256", round(j["extra"]["n256"]["value"],1))
print("roofline", {k: j["roofline"][k] for k in ("kernel","achieved","frac","traffic")}, j["roofline"]["whole_step"], j["roofline"]["per_kernel_measured_dram"])
print("fp64", j["roofline"]["fp64_pipe"]["frac"], j["roofline"]["fp64_pipe"]["issue_frac"], "cpu", j["cpu_baseline"]["value"], "refcuda", j["reference_cuda_baseline"] and j["reference_cuda_baseline"]["value"], "launches", j["gpu_launches"], j["clocks"])
print("reference arm", round(r["value"],2), r["unit"], r["config"]["workload"][:60], r["cpu_baseline"]["sample"][:120], "ratio e2e", round(j["e2e"]["value"]/r["value"],1), "resident", round(j["value"]/r["value"],1))
PY
