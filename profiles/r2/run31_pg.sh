#!/bin/bash
# plane group at 4 CTAs per SM / 128 registers against 5 / 96 (in-process A/B, x-z group on = the new default)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 200 python profiles/r2/ab_xz.py 256 40 pg 2>&1 | tee gpurun_out/r2_pg_ab256.log
timeout 300 python profiles/r2/ab_xz.py 512 20 pg 2>&1 | tee gpurun_out/r2_pg_ab512.log
