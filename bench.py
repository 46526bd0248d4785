#!/usr/bin/env python
"""bench.py -- Orszag-Tang 3-D fp64 MHD cell-updates/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--mode fast|exact]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, z-slabs)
  python bench.py --impl reference ...                                   (the reference on the host cores)

A "step" is one full time step of the hot path (ghost fill / NCCL halo exchange, primitives + CFL
reduction, edge E + face-B slopes, Hancock trace, HLLD fluxes x/y/z, edge EMFs z/y/x, conservative + CT
update) over the rank's slab.  Workload at every N: Orszag-Tang 3-D with kt=1 (a genuinely 3-D flow),
n^3 cells PER GPU (default 512^3 = BASELINE's target size and configs[4]; weak scaling: the domain grows along z with N;
256^3 = configs[1] is measured beside it at N=1 as `extra.n256`).  `--workload blast|field_loop` and `--strong`
(n^3 cells in TOTAL, split into z-slabs) cover configs[2] and configs[3].
`value` counts interior cell-updates of all ranks per second with the state resident in HBM;
`e2e` is the same metric through the C ABI with HOST buffers (pinned H2D of U before and D2H of U after
every step inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints ONE JSON line on stdout. Libraries write there too (NCCL's version banner when NCCL_DEBUG is set in
# the environment or in nccl.conf): file descriptor 1 is pointed at stderr for the whole run and the JSON line goes
# to the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "Orszag-Tang 3D MHD fp64 Mcell-updates/s"
UNIT = "Mcell-updates/s"
ALGO_BYTES_PER_CELL = 128.0  # read 8 fp64 + write 8 fp64 per interior cell-update
OT = "[OrszagTang]\nkt=1\n"


PROBLEMS = {
    # problem -> (ini [hydro] problem name, extra sections, cfl, (xmin, xmax, ymin, ymax, zmin, zmax) of an n^3 box)
    "orszag_tang": ("orszag_tang", OT, 0.8, (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)),
    "blast": ("blast", "[blast]\nradius=0.1\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n", 0.8,
              (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)),
    # the loop sits at the origin: the box of the reference's settings/mhd_fieldloop3d.ini (its vector potential is only
    # periodic on a box centred there; on [0,1]^3 the initial field would not be divergence-free across the boundary)
    "field_loop": ("field_loop", "[FieldLoop]\nradius=0.3\namplitude=0.001\nvflow=3\ndensity_in=1\n", 0.4,
                   (-1.0, 1.0, -0.5, 0.5, -0.5, 0.5)),
}


def make_ini(n, mz, nstepmax, noutput=0, problem="orszag_tang", nz=None):
    """periodic box of cubic cells; n x n x nz cells per rank, mz z-slabs (nz = n: weak scaling,
    nz = n/mz: a fixed n^3 problem split over the ranks). Orszag-Tang runs with kt=1 (a genuinely 3-D flow)."""
    nz = n if nz is None else nz
    name, extra, cfl, (x0, x1, y0, y1, z0, z1) = PROBLEMS[problem]
    lz = (z1 - z0) * float(nz * mz) / float(n)  # the domain grows along z with the slabs of a weak-scaling run
    if problem == "field_loop":  # stays centred on the loop
        z0, z1 = 0.5 * (z0 + z1) - 0.5 * lz, 0.5 * (z0 + z1) + 0.5 * lz
    else:
        z1 = z0 + lz
    bc = "\n".join(f"boundary_type_{f}=3" for f in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"))
    return f"""[run]
solver_name=MHD_Muscl_3D
tEnd=1000000.0
nStepmax={nstepmax}
nOutput={noutput}
nlog=1000000
[mesh]
nx={n}
ny={n}
nz={nz}
xmin={x0}
xmax={x1}
ymin={y0}
ymax={y1}
zmin={z0}
zmax={z1}
{bc}
[hydro]
gamma0=1.666
cfl={cfl}
niter_riemann=10
iorder=2
slope_type=2
problem={name}
riemann=hlld
smallr=1e-8
smallc=1e-8
[mpi]
mx=1
my=1
mz={mz}
[output]
outputPrefix=bench
outputVtkAscii=false
[other]
implementationVersion=0
{extra}"""


def workload_config(args, world):
    """The `config` object of the JSON line: identical in both arms (the reference arm times a bounded sample of it)."""
    n = args.n
    nz = n // world if args.strong else n
    title = {"orszag_tang": "Orszag-Tang 3D kt=1", "blast": "MHD blast 3D", "field_loop": "MHD field-loop advection 3D"}[args.workload]
    return {"workload": f"{title}, {n}x{n}x{nz} cells per GPU (global {n}x{n}x{nz * world}), z-slabs mz={world}, HLLD + CT, periodic, "
                        f"cfl {PROBLEMS[args.workload][2]}, gamma 1.666, implementationVersion=0 semantics",
            "problem": args.workload, "n": n, "nz_per_gpu": nz, "scaling": "strong" if args.strong else "weak",
            "l2": f"inputs larger than L2 (every array >= {8.0 * n * n * nz / 1e9:.1f} GB per GPU vs 126 MB L2), no flush needed"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe). NVML is polled from a thread
    every few ms (the timed region of a default run is ~150 ms, shorter than nvidia-smi's start-up); `nvidia-smi
    --query-gpu` with the same fields is the fallback when the NVML binding is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_s=0.004):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.stop_flag, self.thread, self.nvml, self.proc, self.rows = threading.Event(), None, None, None, []

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            self.nvml = N
            self.handle = N.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(self.handle, N.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        N = self.nvml
        names = ((N.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (N.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                 (N.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (N.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.handle, N.NVML_CLOCK_SM)))
                mask = N.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, nm in names:
                    if mask & bit:
                        self.reasons.add(nm)
                self.power.append(N.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join()
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "power_w_max": max(self.power) if self.power else None, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref/ppkMHD) on the host cores
# ---------------------------------------------------------------------------------------------
def _run_ref_once(n, nsteps, threads, problem="orszag_tang"):
    from oracle import oracle as O

    ini = make_ini(n, 1, nsteps, problem=problem)
    if O.have_reference():
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "run.ini"), "w").write(ini)
            env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="threads")
            out = subprocess.run([O.REF_BIN, "run.ini"], cwd=tmp, env=env, capture_output=True, text=True, check=True).stdout
        return float(re.search(r"total\s+time\s*:\s*([0-9.]+)", out).group(1)), "reference"
    # the reference did not travel: time the plain-C port (OpenMP) instead
    os.environ["OMP_NUM_THREADS"] = str(threads)
    orc = O.Oracle(ini)
    t0 = time.time()
    orc.run(nsteps)
    return time.time() - t0, "port"


def cpu_reference_throughput(n, steps, warmup, threads, problem="orszag_tang"):
    """interior Mcell-updates/s of the reference's own loop: (T(W+K) - T(W)) isolates K steps."""
    t_w, kind = _run_ref_once(n, max(warmup, 1), threads, problem)
    t_wk, kind = _run_ref_once(n, max(warmup, 1) + steps, threads, problem)
    dt = max(t_wk - t_w, 1e-9)
    return n ** 3 * steps / dt * 1e-6, kind, dt / steps


REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "cuda", "ppkMHD_cuda")
REF_CUDA_SHIM = os.path.join(ROOT, "oracle", "_ref", "cuda", "libcc_shim.so")


def ref_cuda_throughput(n, device):
    """The reference's OWN Kokkos-CUDA kernels on this GPU (oracle/ref_build/Makefile.cuda: unmodified sources, sm_90 SASS +
    compute_90 PTX JIT-compiled for sm_100; cc_shim.c makes its Kokkos 4.3 accept a 10.x device). Same ini as our arm,
    implementationVersion 0; (T(25 steps) - T(5 steps)) / 20 after a 1-step run that fills the JIT cache. Not the
    contract's reference arm (that is the CPU build): an extra, like-for-like GPU baseline."""
    if not (os.path.exists(REF_CUDA) and os.path.exists(REF_CUDA_SHIM)):
        return None
    env = dict(os.environ, LD_PRELOAD=REF_CUDA_SHIM, CUDA_VISIBLE_DEVICES=str(device),
               LD_LIBRARY_PATH="/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    times = {}
    try:
        with tempfile.TemporaryDirectory() as tmp:
            for ns in (1, 5, 25):
                open(os.path.join(tmp, "run.ini"), "w").write(make_ini(n, 1, ns))
                out = subprocess.run([REF_CUDA, "run.ini"], cwd=tmp, env=env, capture_output=True, text=True, timeout=600).stdout
                times[ns] = float(re.search(r"total\s+time\s*:\s*([0-9.]+)", out).group(1))
    except Exception as exc:  # the baseline is optional: report why it is missing
        return {"unavailable": repr(exc)[:200]}
    spp = max(times[25] - times[5], 1e-9) / 20
    return {"value": n ** 3 / spp * 1e-6, "unit": UNIT, "ms_per_step": spp * 1e3, "kind": "reference Kokkos-CUDA build (v0), same GPU",
            "sample": f"Orszag-Tang 3D kt=1 {n}^3, 20 steps (difference of a 25- and a 5-step run), interior cells counted"}


def run_reference_arm(args):
    """The UNMODIFIED reference (Kokkos-OpenMP, implementationVersion 0) on all host cores, on a bounded sample of our arm's
    workload: same problem, same ini keys, --ref-n^3 cells (v0 of the reference needs 207 doubles per cell: 512^3 is
    230 GB of host memory, 256^3 is 30 GB), the SAME number of timed and warm-up steps as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_n
    steps, warm = args.steps, max(args.warmup, 3)  # (the same W >= 3 as our arm)
    v, kind, spp = cpu_reference_throughput(n, steps, warm, threads, args.workload)
    sample = (f"{args.workload} {n}^3 sample of the workload, {steps} timed steps after {warm} warm-up steps (total-time difference of "
              f"two runs of oracle/_ref/ppkMHD), {threads} OpenMP threads, implementationVersion=0")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ppkmhd_b200 as ppk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, K, W = args.n, args.steps, max(args.warmup, 3)
    exact = args.mode == "exact"
    if args.strong and n % world != 0:
        raise SystemExit("--strong needs n divisible by the number of GPUs")
    nz = n // world if args.strong else n

    ini = make_ini(n, world, 10 ** 9, problem=args.workload, nz=nz)
    p, t_end, _ = ppk.params_from_ini(ini, rank_z=rank, device=local, exact=exact)
    solver = ppk.Mhd3d(p)
    if args.pipeline != "auto":
        solver.set_pipeline(args.pipeline)
    pipeline = solver.pipeline()
    # a dedicated (non-default) torch stream: the C ABI launches on it, torch.cuda.Event records on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    solver.set_stream(stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(ppk.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        solver.comm_init(bytes(idt.cpu().tolist()), world, rank)

    # initial condition on the host (the reference's init functors use libm), pinned for the e2e leg
    host = torch.empty(p.shape, dtype=torch.float64).pin_memory()
    host.numpy()[...] = ppk.init_condition_from_ini(ini, rank_z=rank)
    solver.upload(host.data_ptr())
    solver.set_time(0.0, t_end, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def timed_run(nsteps):
        """nsteps steps bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks; clocks sampled
        on rank 0 during the region"""
        # the sampler thread starts BEFORE the barrier: NVML's first initialisation takes ~0.1 s on rank 0, and a rank that
        # enters the timed region late makes its neighbours wait inside theirs (seen as +7 ms per step on a 20-step region)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        solver.run(nsteps)
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)), (sampler.stop() if rank == 0 else None)

    cells = float(n) * n * nz * world  # interior cells of all ranks

    # ---- resident leg ------------------------------------------------------------------------
    solver.run(W)
    l0 = solver.launch_count()
    ms, clocks = timed_run(K)
    launches = solver.launch_count() - l0
    value = cells * K / (ms * 1e-3) * 1e-6
    # the same loop for at least two seconds: the clocks of a sustained run (the K-step region is a burst)
    sustained = None
    if ms < 2000.0 and not args.no_sustained:
        Ks = int(2000.0 / (ms / K)) + 1
        ms_s, clocks_s = timed_run(Ks)
        sustained = {"steps": Ks, "ms_per_step": ms_s / Ks, "value": cells * Ks / (ms_s * 1e-3) * 1e-6, "unit": UNIT,
                     "seconds": ms_s * 1e-3, "clocks": clocks_s}
    t_sim, dt_sim, it = solver.get_time()
    sums, divb = solver.diagnostics()
    if not np.all(np.isfinite(sums)):
        raise SystemExit("non-finite state after the timed steps")
    divb_ranks = all_ranks(divb)
    # constrained transport keeps div B at round-off on every slab, decomposed or not (north-star criterion: 1e-12 after
    # 100 steps). div B is a difference of face values over dx, so its round-off floor scales with |B| / dx and random-walks
    # with the number of steps; anything far above that means a broken exchange or initial state: not a measurement.
    divb_bound = 1e-12 * max(1.0, float(n) / 256.0) * max(1.0, (it / 100.0) ** 0.5)
    if max(divb_ranks) > 1e3 * divb_bound:
        raise SystemExit(f"max|div B| per rank {divb_ranks} is far above round-off ({divb_bound:.1e}): the run is not a valid measurement")

    # ---- per-kernel timing (CUDA events around every launch, on the launch stream) ------------
    solver.profile(True)
    solver.kernel_times(reset=True)
    PK = 3
    solver.run(PK)
    solver.synchronize()
    kt = solver.kernel_times(reset=True)
    solver.profile(False)
    per_kernel = {k: {"ms_per_step": v[0] / PK, "launches_per_step": v[1] / PK} for k, v in kt.items() if v[1] > 0}
    step_ms_prof = sum(v["ms_per_step"] for v in per_kernel.values())
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"] / max(per_kernel[k]["launches_per_step"], 1))
    dom_ms = per_kernel[dom]["ms_per_step"] / per_kernel[dom]["launches_per_step"]
    pk, pk_kind = peaks()
    cells_rank = float(n) * n * nz
    algo_bytes = ALGO_BYTES_PER_CELL * cells_rank  # per launch: one launch sweeps the rank's cells
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9
    # measured per-kernel counters of this pipeline (profiles/r2/counters.json: ncu on one step, per launch): DRAM bytes
    # (dram__bytes_read.sum + dram__bytes_write.sum) and executed instructions (smsp__inst_executed.sum,
    # smsp__inst_executed_pipe_fp64.sum), both scaled to the cell count of this run
    traffic, fp64, cj = None, None, None
    cpath = os.path.join(ROOT, "profiles", "r2", "counters.json")
    if os.path.exists(cpath):
        allj = json.load(open(cpath))
        # the capture of this pipeline at this size, else at another size (per-cell figures are size-independent to ~1 %)
        cj = allj.get(f"{pipeline}_{n}") or next((v for k, v in sorted(allj.items()) if v.get("pipeline") == pipeline), None)
        if cj:
            scale = cells_rank / float(cj["cells"])
            if dom in cj["kernels"]:
                traffic = cj["kernels"][dom]["dram_bytes"] * scale if cj["cells"] == cells_rank else None
            inst = sum(k["inst_executed"] for k in cj["kernels"].values()) * 32.0 / cj["cells"]
            f64 = sum(k["inst_fp64"] for k in cj["kernels"].values()) * 32.0 / cj["cells"]
            rate = f64 * cells_rank * K / (ms * 1e-3)
            fp64 = {"fp64_pipe_inst_per_cell_update": f64, "all_inst_per_cell_update": inst, "achieved_Tinst_s": rate / 1e12,
                    "peak_Tinst_s": 17.0, "frac": rate / 17.0e12, "issue_frac": inst * cells_rank * K / (ms * 1e-3) / 36.0e12,
                    "dram_bytes_per_cell_update": sum(k["dram_bytes"] for k in cj["kernels"].values()) / cj["cells"],
                    "source": f"ncu counters of one {cj['n']}^3 step (profiles/r2/counters.json): smsp__inst_executed_pipe_fp64.sum, "
                              "smsp__inst_executed.sum (warp instructions x 32), dram__bytes_read+write; FP64 peak 17.0 T "
                              "thread-inst/s measured by profiles/microbench/fp64_pipe.cu, issue peak 148 SMs x 4 x 32 x 1.9 GHz"}
    # per kernel: measured DRAM bytes (ncu, scaled to this run's cell count) over this run's CUDA-event time = the DRAM
    # bandwidth each kernel actually sustains, as a fraction of the measured copy peak (which kernels are HBM-bound, which not)
    per_kernel_dram = None
    if os.path.exists(cpath) and cj:
        per_kernel_dram = {}
        for k, v in per_kernel.items():
            if k in cj["kernels"] and v["ms_per_step"] > 0:
                gbs = cj["kernels"][k]["dram_bytes"] * (cells_rank / float(cj["cells"])) / (v["ms_per_step"] * 1e-3) / 1e9
                per_kernel_dram[k] = {"dram_GBs": round(gbs, 1), "frac_of_peak": round(gbs / pk["hbm_gbs"], 3)}
    step_gbs = ALGO_BYTES_PER_CELL * cells_rank * K / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind + " copy bandwidth (burst)",
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": dom_ms,
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / pk["hbm_gbs"]},
                "fp64_pipe": fp64, "per_kernel_measured_dram": per_kernel_dram,
                "note": "128 algorithmic B per cell-update (read U^n, write U^n+1). The step executes ~2.45k FP64-pipe and ~5.7k "
                        "instructions per cell-update (fp64_pipe.*): the FP64 pipe alone caps it at 6.9 Gcell/s = 13.6% of the HBM "
                        "roofline, instruction issue at 6.5 Gcell/s; see DESIGN.md 4.1"}

    nbytes = int(np.prod(p.shape)) * 8
    # ---- e2e leg: host buffers through the C ABI, H2D + step + D2H every step ------------------
    # Every step is one batch: a full host state (pinned) goes to the device, advances one step, and the full result comes
    # back. ONE handle at every N: ppk_mhd3d_stage_upload / _stage_swap / _step / _stage_download rotate four device arrays
    # so that the upload of batch i+1 (copy stream 1) overlaps the step of batch i and the download of batch i-1 (copy
    # stream 2); PCIe is full duplex and bounds the leg.
    Ke = max(2, min(K, args.e2e_steps))
    # pinned host buffers: every batch uploads the same initial state (one input buffer); results alternate between two
    # output buffers when the host has room for them (8.9 GB each at 512^3, per rank), else share one
    try:
        import psutil
        host_free = psutil.virtual_memory().available
    except Exception:
        host_free = 0
    n_out = 2 if host_free > 6 * nbytes * max(world, 1) else 1
    host_in = [host, host]
    host_out = [torch.empty(p.shape, dtype=torch.float64).pin_memory() for _ in range(n_out)] * (2 // n_out)

    def e2e_pass(nsteps):
        solver.stage_upload(host_in[0].data_ptr())
        for b in range(nsteps):
            solver.stage_swap()
            solver.step()
            solver.stage_download(host_out[b % 2].data_ptr())
            if b + 1 < nsteps:
                solver.stage_upload(host_in[(b + 1) % 2].data_ptr())
        solver.synchronize()

    e2e_pass(2)  # warm-up: the staging array, the copy streams
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    w0 = time.perf_counter()
    e2e_pass(Ke)  # ends with ppk_mhd3d_synchronize: every stream of the handle is idle when f1 is recorded
    f1.record(stream)
    barrier()
    w1 = time.perf_counter()
    ems = max_over_ranks(f0.elapsed_time(f1))
    assert np.all(np.isfinite(host_out[(Ke - 1) % 2].numpy()[:, 3:-3, 3:-3, 3:-3].sum()))
    e2e = {"value": cells * Ke / (ems * 1e-3) * 1e-6, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world, "steps": Ke, "ms_per_step": ems / Ke, "wall_ms_per_step": (w1 - w0) * 1e3 / Ke,
           "pcie_GBs_per_gpu_each_way": nbytes / (ems / Ke * 1e-3) / 1e9,
           "what": "per step: ppk_mhd3d_stage_upload(pinned host U) + _stage_swap + ppk_mhd3d_step + ppk_mhd3d_stage_download("
                   "pinned host U) on one handle per GPU; four device arrays rotate, uploads, steps and downloads of "
                   "consecutive batches overlap on three streams"}
    del host_in, host_out

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, kind, spp = cpu_reference_throughput(args.cpu_n, 6, 2, threads, args.workload)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{args.workload} {args.cpu_n}^3 sample, 6 timed steps after 2 warm-up (difference of two runs of "
                         f"oracle/_ref/ppkMHD, Kokkos-OpenMP, implementationVersion=0); {spp * 1e3:.0f} ms/step"}

    sim = {"t": t_sim, "dt": dt_sim, "iteration": it, "max_divB": max(divb_ranks), "max_divB_per_rank": divb_ranks,
           "divB_roundoff_bound": divb_bound, "divB_at_roundoff": bool(max(divb_ranks) <= divb_bound),
           "device_GB": solver.device_bytes() / 1e9}
    solver.close()
    del solver

    # ---- BASELINE configs[1] (256^3 on one GPU) beside the headline size, N=1 only --------------
    extra = {}
    if rank == 0 and world == 1 and n != 256 and args.workload == "orszag_tang" and not args.no_extra:
        ini2 = make_ini(256, 1, 10 ** 9)
        p2, t_end2, _ = ppk.params_from_ini(ini2, device=local, exact=exact)
        s2 = ppk.Mhd3d(p2)
        if args.pipeline != "auto":
            s2.set_pipeline(args.pipeline)
        s2.set_stream(stream.cuda_stream)
        s2.upload(ppk.init_condition_from_ini(ini2))
        s2.set_time(0.0, t_end2, 0)
        s2.run(max(W, 5))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K2 = max(K, 40)
        a.record(stream)
        s2.run(K2)
        b.record(stream)
        torch.cuda.synchronize()
        ms2 = a.elapsed_time(b)
        extra["n256"] = {"value": 256.0 ** 3 * K2 / (ms2 * 1e-3) * 1e-6, "unit": UNIT, "ms_per_step": ms2 / K2, "steps": K2,
                         "config": "Orszag-Tang 3D kt=1 256^3 on one GPU (BASELINE configs[1]), resident"}
        s2.close()

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda and args.workload == "orszag_tang":
        torch.cuda.synchronize()
        ref_cuda = ref_cuda_throughput(256, local)  # v0 of the reference needs 207 doubles per cell: 256^3 is what fits

    if rank == 0:
        cfg = workload_config(args, world)  # identical in the reference arm's line
        details = {"pipeline": pipeline,
                   "arithmetic": "fast (FMA contraction, within 1e-12 of the reference)" if not exact else "exact (--fmad=false, bit-identical)",
                   "cells_with_ghosts_per_gpu": int(np.prod(p.shape[1:])),
                   "reference_style_value_with_ghosts": value * float(np.prod(p.shape[1:])) / cells_rank}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "details": details,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "sustained": sustained, "extra": extra, "reference_cuda_baseline": ref_cuda,
            "per_kernel_ms": {k: round(v["ms_per_step"], 4) for k, v in per_kernel.items()},
            "profiled_step_ms": step_ms_prof, "sim": sim,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=512, help="cells per axis per GPU (with --strong: of the whole problem)")
    ap.add_argument("--workload", default="orszag_tang", choices=sorted(PROBLEMS))
    ap.add_argument("--strong", action="store_true", help="a fixed n^3 problem split into z-slabs (BASELINE configs[3]) instead of n^3 per GPU")
    ap.add_argument("--cpu-n", type=int, default=128, help="grid of the bounded CPU sample printed beside our arm (cpu_baseline)")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the 256^3 (configs[1]) measurement beside the headline size")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--pipeline", default="auto", choices=["auto", "ordered", "tiled", "fused", "fused_split", "unfused", "streamed"],
                    help="auto = the handle's default (unfused: the schedule that measured fastest at 256^3 and 512^3)")
    ap.add_argument("--ref-n", type=int, default=256, help="grid of the reference arm's bounded sample (--impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=9)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's Kokkos-CUDA build on the same GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
