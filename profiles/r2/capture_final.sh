#!/bin/bash
# Round-2 evidence for the default bench command (512^3, ordered pipeline): launch list of `python bench.py` (short form),
# ncu --set full of every kernel of one step at 512^3, the nvidia-smi clocks line during a bench run.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/r2_final_launches.log 2>&1
tail -n 2 gpurun_out/r2_final_launches.log | cut -c1-300
# one step at 512^3, every kernel, full set (3 steps run, the kernels of the 2nd step are captured: skip 12 launches, take 12)
ncu --set full --clock-control none -s 15 -c 15 -o gpurun_out/r2_final_512 -f python profiles/r2/mini.py 512 3 auto > gpurun_out/r2_final_512.log 2>&1
tail -n 2 gpurun_out/r2_final_512.log
python profiles/summarize.py gpurun_out/r2_final_512.ncu-rep > gpurun_out/r2_final_512.md
ncu --set full --clock-control none -s 15 -c 15 -o gpurun_out/r2_final_256 -f python profiles/r2/mini.py 256 3 auto > gpurun_out/r2_final_256.log 2>&1
tail -n 2 gpurun_out/r2_final_256.log
python profiles/summarize.py gpurun_out/r2_final_256.ncu-rep > gpurun_out/r2_final_256.md
ls -la gpurun_out/*.ncu-rep; find gpurun_out -name '*.ncu-rep' -size +24M -delete
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_final_clocks.csv &
SMI=$!
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
kill $SMI
python -c "
import json; j=json.load(open('gpurun_out/r2_final_bench.json')); print(round(j['value'],1), round(j['ms_per_step'],2), j['details']['pipeline'], 'sust', round(j['sustained']['value'],1), 'e2e', round(j['e2e']['value'],1), 'n256', j['extra'])"
