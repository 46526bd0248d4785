#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for n in 256 512; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/r2/timeline.py $n > gpurun_out/r2_timeline_N2_$n.md 2> gpurun_out/r2_timeline_N2_$n.err
head -50 gpurun_out/r2_timeline_N2_$n.md
done
tail -3 gpurun_out/r2_timeline_N2_512.err
