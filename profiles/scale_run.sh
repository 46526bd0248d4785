set -x
nvidia-smi -L | head -8
python -m pytest tests -m gpu -x -q -k "z_slabs" 2>&1 | tail -3
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_err_$N.log | tee gpurun_out/scale_r1_N$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['per_kernel_ms'])"
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/scale_r1_N1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
