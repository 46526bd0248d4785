"""CUDA-event timeline of the decomposed step (ppk_mhd3d_kernel_timeline): launched with torchrun on N GPUs, prints rank 0's
launches of the last profiled step -- which stream, start, end -- so that what runs under what is visible.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/r2/timeline.py 256"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ppkmhd_b200 as ppk  # noqa: E402
from bench import make_ini  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ini = make_ini(n, world, 10 ** 9)
p, t_end, _ = ppk.params_from_ini(ini, rank_z=rank, device=local, exact=False)
s = ppk.Mhd3d(p)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(ppk.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    s.comm_init(bytes(idt.cpu().tolist()), world, rank)
s.upload(ppk.init_condition_from_ini(ini, rank_z=rank))
s.set_time(0.0, t_end, 0)
s.run(5)
s.synchronize()
if world > 1:
    dist.barrier()
s.profile(True)
s.run(3)
s.synchronize()
tl = s.kernel_timeline()
s.profile(False)
if rank == 0:
    import bench

    # from the last-but-one trace launch to the end: one whole cycle trace ... update | ghost fill, primitives, E ... trace ... update
    traces = [i for i, e in enumerate(tl) if e[0] == "trace"]
    lo = traces[-2] if len(traces) >= 2 else 0
    t0 = tl[lo][2]
    out = [f"# rank 0 of {world}, Orszag-Tang {n}^3 per GPU, pipeline {s.pipeline()}: CUDA-event timeline from the trace of one step to the end of the next (ms)",
           "", "| launch | stream | start | end | ms |", "|---|---|---|---|---|"]
    for name, comm, a, b in tl[lo:]:
        out.append(f"| {name} | {'comm' if comm else 'compute'} | {a - t0:8.3f} | {b - t0:8.3f} | {b - a:6.3f} |")
    os.write(bench._REAL_STDOUT, ("\n".join(out) + "\n").encode())  # (importing bench.py points fd 1 at stderr)
s.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
