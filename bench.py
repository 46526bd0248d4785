#!/usr/bin/env python
"""bench.py -- Orszag-Tang 3-D fp64 MHD cell-updates/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--mode fast|exact]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, z-slabs)
  python bench.py --impl reference ...                                   (the reference on the host cores)

A "step" is one full time step of the hot path (ghost fill / NCCL halo exchange, primitives + CFL
reduction, edge E + face-B slopes, Hancock trace, HLLD fluxes x/y/z, edge EMFs z/y/x, conservative + CT
update) over the rank's slab.  Workload at every N: Orszag-Tang 3-D with kt=1 (a genuinely 3-D flow),
n^3 cells PER GPU (default 256^3 = BASELINE configs[1]; weak scaling: the domain grows along z with N).
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


def make_ini(n, mz, nstepmax, noutput=0):
    """Orszag-Tang, kt=1, periodic, cubic cells, domain [0,1]x[0,1]x[0,mz]."""
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
nz={n}
xmin=0.0
xmax=1.0
ymin=0.0
ymax=1.0
zmin=0.0
zmax={float(mz)}
{bc}
[hydro]
gamma0=1.666
cfl=0.8
niter_riemann=10
iorder=2
slope_type=2
problem=orszag_tang
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
{OT}"""


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
def _run_ref_once(n, nsteps, threads):
    from oracle import oracle as O

    ini = make_ini(n, 1, nsteps)
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


def cpu_reference_throughput(n, steps, warmup, threads):
    """interior Mcell-updates/s of the reference's own loop: (T(W+K) - T(W)) isolates K steps."""
    t_w, kind = _run_ref_once(n, max(warmup, 1), threads)
    t_wk, kind = _run_ref_once(n, max(warmup, 1) + steps, threads)
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_n
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    v, kind, spp = cpu_reference_throughput(n, steps, warm, threads)
    sample = f"Orszag-Tang 3D kt=1 {n}^3, {steps} timed steps after {warm} warm-up steps (total-time difference of two runs), implementationVersion=0"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"Orszag-Tang 3D kt=1 (bounded CPU sample {n}^3; GPU arm runs {args.n}^3 per GPU)",
                                        "hlld": True, "implementationVersion": 0},
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

    ini = make_ini(n, world, 10 ** 9)
    p, t_end, _ = ppk.params_from_ini(ini, rank_z=rank, device=local, exact=exact)
    solver = ppk.Mhd3d(p)
    if args.pipeline != "auto":
        solver.set_pipeline(args.pipeline)
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

    # ---- resident leg ------------------------------------------------------------------------
    solver.run(W)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = solver.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solver.run(K)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = solver.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    cells = float(n) ** 3 * world
    value = cells * K / (ms * 1e-3) * 1e-6
    t_sim, dt_sim, it = solver.get_time()
    sums, divb = solver.diagnostics()
    if not np.all(np.isfinite(sums)):
        raise SystemExit("non-finite state after the timed steps")

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
    algo_bytes = ALGO_BYTES_PER_CELL * float(n) ** 3  # per launch: one launch sweeps the rank's n^3 cells
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes of the dominant kernel from the committed ncu --set full capture (same n), per launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "kernel_dram_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("n") == n:
            traffic = tj.get("dram_bytes_per_launch", {}).get(dom)
    # second roofline of this path: the FP64 pipe. Instructions per cell-update from the SASS of the
    # kernels timed here (profiles/sass_counts.json), pipe peak measured on this pool's B200 by
    # profiles/microbench/fp64_pipe.cu (17.0 T thread-instructions/s = 34 TFLOP/s)
    fp64 = None
    spath = os.path.join(ROOT, "profiles", "sass_counts.json")
    if os.path.exists(spath):
        per_cell = json.load(open(spath))["per_cell_update"]
        rate = per_cell["fp64_pipe"] * cells / world * K / (ms * 1e-3)
        fp64 = {"fp64_pipe_inst_per_cell_update": per_cell["fp64_pipe"], "all_inst_per_cell_update": per_cell["instructions"],
                "achieved_Tinst_s": rate / 1e12, "peak_Tinst_s": 17.0, "frac": rate / 17.0e12,
                "peak_source": "profiles/microbench/fp64_pipe.cu measured on B200 (DFMA, 16 warps/SM, ILP 4)"}
    step_gbs = ALGO_BYTES_PER_CELL * cells / world * K / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind + " copy bandwidth (burst)",
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": dom_ms,
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / pk["hbm_gbs"]},
                "fp64_pipe": fp64,
                "note": "128 algorithmic B per cell-update (read U^n, write U^n+1). The step needs ~3.0k FP64-pipe "
                        "instructions per cell-update, so the FP64 pipe bounds it at ~5.7 Gcell/s = 11% of the HBM roofline; "
                        "see DESIGN.md"}

    # ---- e2e leg: host buffers through the C ABI, H2D + step + D2H every step ------------------
    # Every step is one batch: upload a full host state (pinned), advance it one step, download the full result.
    # N=1: `depth` solver handles are in flight, each on its own stream, so that the download of batch i overlaps the
    # upload of batch i+1 and the step of the batch between them (PCIe is full duplex; the copies bound the leg).
    # N>1: one handle (one NCCL communicator per rank), the three phases run back to back.
    Ke = max(2, min(K, args.e2e_steps))
    nbytes = int(np.prod(p.shape)) * 8
    depth = args.e2e_depth if world == 1 else 1
    e2e_solvers, e2e_streams = [solver], [stream]
    for _ in range(depth - 1):
        s2 = ppk.Mhd3d(p)
        if args.pipeline != "auto":
            s2.set_pipeline(args.pipeline)
        st2 = torch.cuda.Stream()
        s2.set_stream(st2.cuda_stream)
        s2.set_time(0.0, t_end, 0)
        e2e_solvers.append(s2)
        e2e_streams.append(st2)
    host_in = [host] + [torch.empty(p.shape, dtype=torch.float64).pin_memory() for _ in range(depth - 1)]
    host_out = [torch.empty(p.shape, dtype=torch.float64).pin_memory() for _ in range(depth)]
    for hb in host_in[1:]:
        hb.copy_(host)

    def e2e_pass(nsteps):
        for it in range(nsteps):
            sv = e2e_solvers[it % depth]
            sv.synchronize()  # batch it-depth is back on the host: its buffers are free again
            sv.upload(host_in[it % depth].data_ptr())
            sv.step()
            sv.download_async(host_out[it % depth].data_ptr())
        for sv in e2e_solvers:
            sv.synchronize()

    e2e_pass(depth)  # warm the extra handles (first-touch allocations, flux arrays)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Stream()
    f0.record(join)
    for st_ in e2e_streams:
        st_.wait_stream(join)  # nothing of the timed region starts before f0
    w0 = time.perf_counter()
    e2e_pass(Ke)
    for st_ in e2e_streams:
        join.wait_stream(st_)
    f1.record(join)
    barrier()
    w1 = time.perf_counter()
    ems = max_over_ranks(f0.elapsed_time(f1))
    assert np.all(np.isfinite(host_out[0].numpy()[:, 3:-3, 3:-3, 3:-3].sum()))
    e2e = {"value": cells * Ke / (ems * 1e-3) * 1e-6, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world, "steps": Ke, "ms_per_step": ems / Ke, "wall_ms_per_step": (w1 - w0) * 1e3 / Ke,
           "handles_in_flight": depth,
           "what": "per step: ppk_mhd3d_upload(pinned host U) + ppk_mhd3d_step + ppk_mhd3d_download_async(pinned host U); "
                   f"{depth} independent batches in flight on {depth} streams, a batch's buffers are reused after "
                   "ppk_mhd3d_synchronize"}
    for s2 in e2e_solvers[1:]:
        s2.close()
    torch.cuda.set_stream(stream)

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, kind, spp = cpu_reference_throughput(args.ref_n, 6, 2, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"Orszag-Tang 3D kt=1 {args.ref_n}^3, 6 timed steps after 2 warm-up (difference of two runs of oracle/_ref/ppkMHD, "
                         f"Kokkos-OpenMP, implementationVersion=0); {spp * 1e3:.0f} ms/step"}

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        solver.synchronize()
        ref_cuda = ref_cuda_throughput(n, local)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"Orszag-Tang 3D kt=1, {n}^3 cells per GPU (global {n}x{n}x{n * world}), z-slabs mz={world}, "
                                   f"HLLD + CT, periodic, cfl 0.8, gamma 1.666, implementationVersion=0 semantics",
                       "pipeline": args.pipeline,
                       "arithmetic": "fast (FMA contraction, within 1e-12 of the reference)" if not exact else "exact (--fmad=false, bit-identical)",
                       "l2": "inputs larger than L2 (every array >= 1.1 GB vs 126 MB L2)",
                       "cells_with_ghosts_per_gpu": int(np.prod(p.shape[1:])),
                       "reference_style_value_with_ghosts": value * float(np.prod(p.shape[1:])) / float(n) ** 3},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "reference_cuda_baseline": ref_cuda,
            "per_kernel_ms": {k: round(v["ms_per_step"], 4) for k, v in per_kernel.items()},
            "profiled_step_ms": step_ms_prof,
            "sim": {"t": t_sim, "dt": dt_sim, "iteration": it, "max_divB": divb, "device_GB": solver.device_bytes() / 1e9},
        }
        emit(line)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="cells per axis per GPU")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--pipeline", default="auto", choices=["auto", "tiled", "fused", "fused_split", "unfused", "streamed"],
                    help="auto = the handle's default (tiled on one slab, unfused on decomposed runs)")
    ap.add_argument("--ref-n", type=int, default=128, help="grid of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=9)
    ap.add_argument("--e2e-depth", type=int, default=3, help="solver handles (batches) in flight in the e2e leg at N=1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's Kokkos-CUDA build on the same GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
