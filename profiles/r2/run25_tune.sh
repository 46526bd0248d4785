#!/bin/bash
# L2-blocking knobs at 512^3 (ordered): y-slab size of the one-launch Riemann stage, slab rows of the plane sweeps
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
cat > /tmp/tune.py <<'PY'
import sys, os; sys.path.insert(0,'.')
import ppkmhd_b200 as ppk
from bench import make_ini
n=512
ini = make_ini(n, 1, 10**9)
p, t_end, _ = ppk.params_from_ini(ini, exact=False)
s = ppk.Mhd3d(p)
s.upload(ppk.init_condition_from_ini(ini)); s.set_time(0.0, t_end, 0)
s.run(2); s.synchronize(); s.profile(True); s.run(4); s.synchronize()
kt = s.kernel_times()
print(os.environ.get("TAG"), {k: round(v[0]/max(v[1],1),2) for k,v in kt.items() if v[1] and k in ("riemann_all","trace","update","elec_dbf","prim_dt")})
PY
for mb in 16 24 40 64 96; do TAG="RALL_SLAB_MB=$mb" PPK_RALL_SLAB_MB=$mb python /tmp/tune.py 2>&1 | grep SLAB; done
for rows in 32 64 128 256 512; do TAG="SLAB_ROWS=$rows" PPK_SLAB_ROWS=$rows python /tmp/tune.py 2>&1 | grep SLAB; done
