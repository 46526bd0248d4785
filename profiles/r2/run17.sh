#!/bin/bash
# upper bound of update / Riemann overlap (PPK_XOVERLAP=1: timing only, wrong results)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for cfg in "256 unfused 0" "256 unfused 1" "256 ordered 0" "256 ordered 1" "512 ordered 0" "512 ordered 1"; do
set -- $cfg
PPK_XOVERLAP=$3 python - <<PY
import sys, time; sys.path.insert(0,'.')
import torch
import ppkmhd_b200 as ppk
from bench import make_ini
n=$1
ini = make_ini(n, 1, 10**9)
p, t_end, _ = ppk.params_from_ini(ini, exact=False)
s = ppk.Mhd3d(p); s.set_pipeline("$2")
s.upload(ppk.init_condition_from_ini(ini)); s.set_time(0.0, t_end, 0)
s.run(3); s.synchronize()
K = 20 if n == 256 else 8
t0=time.perf_counter(); s.run(K); s.synchronize(); t1=time.perf_counter()
print("n $1 $2 xoverlap $3: %.3f ms/step" % ((t1-t0)*1e3/K))
PY
done
