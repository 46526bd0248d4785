#!/bin/bash
# where does the Riemann kernel's time go: TMA staging alone vs the solves alone (persistent kernel, 5 and 4 CTAs/SM)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for cfg in "5 0" "5 1" "5 2" "5 3" "4 1" "4 2" "3 1" "3 2" "2 1"; do
set -- $cfg
PPK_RALL_PERS=1 PPK_RALL_CTAS=$1 PPK_RALL_XMODE=$2 python - <<PY
import sys; sys.path.insert(0,'.')
import ppkmhd_b200 as ppk
from bench import make_ini
ini = make_ini(256, 1, 10**9)
p, t_end, _ = ppk.params_from_ini(ini, exact=False)
s = ppk.Mhd3d(p); s.set_pipeline("tiled")
s.upload(ppk.init_condition_from_ini(ini)); s.set_time(0.0, t_end, 0)
s.run(2); s.profile(True); s.run(3); s.synchronize()
kt = s.kernel_times()
print("ctas $1 xmode $2", {k: round(v[0] / max(v[1], 1), 3) for k, v in kt.items() if k in ("riemann_all", "producer")})
PY
done
