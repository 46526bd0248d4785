#!/bin/bash
# Final captures of the round with the x-z group in the default schedule: counters (-> counters.json on the box, so that the
# bench line's roofline.traffic / fp64_pipe refer to THIS schedule), the default bench command untraced with the clocks line,
# its ncu launch list, ncu --set full of every kernel of one 512^3 step, smoke.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intermediates or benchmarked_workload_128 or split_phase" ) > gpurun_out/r2_final_pytest_subset.log 2>&1; tail -n 5 gpurun_out/r2_final_pytest_subset.log | head -2
timeout 150 bash profiles/r2/counters.sh 512 unfused
python profiles/r2/counters.py gpurun_out/r2_counters_512_unfused.csv 512 unfused 3 | tee gpurun_out/r2_counters_512_summary.txt
cp profiles/r2/counters.json gpurun_out/r2_counters.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_final_clocks.csv &
SMI=$!
( time timeout 400 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err ) 2>&1 | grep real
kill $SMI
python -c "
import json; j=json.load(open('gpurun_out/r2_final_bench.json')); print(round(j['value'],1), round(j['ms_per_step'],2), j['details']['pipeline'], 'sust', round(j['sustained']['value'],1), 'e2e', round(j['e2e']['value'],1), 'n256', j['extra'], 'traffic', j['roofline']['traffic'], j['roofline']['frac'])"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --e2e-steps 2 --no-extra --no-sustained"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/r2_final_launches.log 2>&1
tail -n 1 gpurun_out/r2_final_launches.log | cut -c1-200
# one step at 512^3, every kernel, full set (3 steps run; one step = 12 launches)
timeout 240 ncu --set full --clock-control none --import-source on -s 12 -c 13 -o gpurun_out/r2_final_512 -f python profiles/r2/mini.py 512 3 auto > gpurun_out/r2_final_512.log 2>&1
tail -n 1 gpurun_out/r2_final_512.log
python profiles/summarize.py gpurun_out/r2_final_512.ncu-rep > gpurun_out/r2_final_512.md
ls -la gpurun_out/*.ncu-rep; find gpurun_out -name '*.ncu-rep' -size +40M -delete
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1; tail -n 4 gpurun_out/r2_final_smoke.log | head -2
