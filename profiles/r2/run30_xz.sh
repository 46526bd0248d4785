#!/bin/bash
# x-z group kernel (mhd_xzgroup.inc, PPK_XZGROUP=1): the whole GPU suite with it on, then the in-process A/B (ab_xz.py)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time PPK_XZGROUP=1 timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_xz_pytest.log 2>&1
tail -4 gpurun_out/r2_xz_pytest.log
timeout 200 python profiles/r2/ab_xz.py 256 40 2>&1 | tee gpurun_out/r2_xz_ab256.log
timeout 300 python profiles/r2/ab_xz.py 512 20 2>&1 | tee gpurun_out/r2_xz_ab512.log
