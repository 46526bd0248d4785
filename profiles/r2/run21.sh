#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1; tail -n 6 gpurun_out/r2_final_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_final_pytest.log 2>&1; tail -n 6 gpurun_out/r2_final_pytest.log
