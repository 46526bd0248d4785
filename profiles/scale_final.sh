# final round-1 weak-scaling points on one 8-GPU box: N = 8 and N = 1 with the default schedule
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --e2e-steps 3 2>gpurun_out/scale_err_$1.log | tee gpurun_out/scale_final_N$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['per_kernel_ms'])"; }
run 8


