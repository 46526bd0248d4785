# compute-sanitizer passes over a slice of the parity tests (bounded: the tools slow kernels down 10-100x)
export PYTHONUNBUFFERED=1
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 --log-file gpurun_out/racecheck.log python -m pytest tests -m gpu -x -q -k "exact_mode_bit_identical and ot_16x12x8" > gpurun_out/racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/racecheck_pytest.log; tail -5 gpurun_out/racecheck.log
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 --log-file gpurun_out/memcheck.log python -m pytest tests -m gpu -x -q -k "intermediates and field_loop" > gpurun_out/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/memcheck_pytest.log; tail -5 gpurun_out/memcheck.log
