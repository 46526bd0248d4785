#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_producer -s 2 -c 1 -o gpurun_out/r2_prof_producer -f python profiles/r2/mini.py 256 4 tiled > gpurun_out/r2_ncu_producer.log 2>&1
$NCU -k regex:k_riemann_all -s 2 -c 1 -o gpurun_out/r2_prof_rall -f python profiles/r2/mini.py 256 4 tiled > gpurun_out/r2_ncu_rall.log 2>&1
tail -3 gpurun_out/r2_ncu_producer.log gpurun_out/r2_ncu_rall.log
