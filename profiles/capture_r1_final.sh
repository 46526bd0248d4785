# ncu evidence for the state at the end of round 1 (run under gpurun, one GPU):
#   1. launch list of the bench command (cold-cache, serialised per-launch times: the SHARE of each kernel must agree with
#      bench.py's CUDA-event numbers, not the absolute)
#   2. --set full capture of one step's kernels (DRAM bytes, FP64-pipe %, stalls)
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --e2e-depth 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_05_launches.csv $CMD > gpurun_out/r1_05_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(prim_dt|elec_dbf|trace|flux_tma|emf_tma|update)" --launch-skip 40 -c 10 -f -o gpurun_out/prof_r1_05 $CMD > gpurun_out/r1_05_full.log 2>&1
tail -2 gpurun_out/r1_05_full.log
python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_final.json | python profiles/pk.py
