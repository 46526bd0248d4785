#!/bin/bash
# the whole GPU suite on the final binary of the round (both merged Riemann kernels at their final register targets)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 108 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2_final_pytest.log 2>&1
tail -n 6 gpurun_out/r2_final_pytest.log
