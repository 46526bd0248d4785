#!/bin/bash
# one ncu --set full capture of a kernel of the tiled pipeline: ncu1.sh <kernel-regex> <out-name> [n] [env...]
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
K=$1; O=$2; N=${3:-256}
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/$O -f python profiles/r2/mini.py $N 4 tiled > gpurun_out/$O.log 2>&1
tail -3 gpurun_out/$O.log
