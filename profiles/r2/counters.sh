#!/bin/bash
# Per-kernel hardware counters of one step (ncu, 3 steps captured, averaged): DRAM bytes, executed warp instructions,
# FP64-pipe warp instructions, duration. Usage: counters.sh <n> <pipeline> ; writes gpurun_out/r2_counters_<n>_<pipeline>.csv
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-256}; P=${2:-auto}
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/r2_counters_${N}_${P}.csv python profiles/r2/mini.py $N 3 $P > gpurun_out/r2_counters_${N}_${P}.log 2>&1
tail -2 gpurun_out/r2_counters_${N}_${P}.log
