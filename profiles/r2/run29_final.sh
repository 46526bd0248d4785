#!/bin/bash
# final state of the round: whole GPU suite + smoke, counters of the default schedule, launch list / ncu --set full / clocks / bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1; tail -n 4 gpurun_out/r2_final_smoke.log | head -2
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_final_pytest.log 2>&1; tail -n 5 gpurun_out/r2_final_pytest.log | head -2
bash profiles/r2/counters.sh 256 unfused
bash profiles/r2/counters.sh 512 unfused
bash profiles/r2/capture_final.sh
