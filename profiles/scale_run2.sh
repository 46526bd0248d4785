run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/scale_err_2.log | python -c "import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG'), d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['per_kernel_ms'])"; }
TAG=eh1_dd1_pr1 PPK_EARLY_HALO=1 PPK_DEFER_DT=1 PPK_COMM_PRIORITY=1 run
TAG=eh0_dd1_pr1 PPK_EARLY_HALO=0 PPK_DEFER_DT=1 PPK_COMM_PRIORITY=1 run
TAG=eh0_dd0_pr1 PPK_EARLY_HALO=0 PPK_DEFER_DT=0 PPK_COMM_PRIORITY=1 run
TAG=eh0_dd0_pr0 PPK_EARLY_HALO=0 PPK_DEFER_DT=0 PPK_COMM_PRIORITY=0 run
TAG=eh1_dd0_pr1 PPK_EARLY_HALO=1 PPK_DEFER_DT=0 PPK_COMM_PRIORITY=1 run
