"""In-process A/B of launcher knobs that are read at every launch (PPK_XZGROUP / PPK_XZ_MINB / PPK_XZ_SLAB_MB / PPK_PGROUP_MINB): Orszag-Tang kt=1
at n^3 through the C ABI, fast build, default schedule; per setting 3 warm-up + K timed steps, CUDA events around every launch."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ppkmhd_b200 as ppk  # noqa: E402
from bench import make_ini  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ini = make_ini(n, 1, 10 ** 9)
p, t_end, _ = ppk.params_from_ini(ini, exact=False)
s = ppk.Mhd3d(p)
s.upload(ppk.init_condition_from_ini(ini))
s.set_time(0.0, t_end, 0)
s.run(5)
s.synchronize()
SETS = {
    # run30: the x-z group against the two kernels it replaces, its register targets and y-slab sizes
    "xz": [{"PPK_XZGROUP": "0"}, {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "5"}, {"PPK_XZGROUP": "0"}, {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "5"},
           {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "4"}, {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "6"},
           {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "5", "PPK_XZ_SLAB_MB": "24"}, {"PPK_XZGROUP": "1", "PPK_XZ_MINB": "5", "PPK_XZ_SLAB_MB": "1000"}],
    # run31: register target of the plane group (5 CTAs per SM at 96 registers with spills | 4 at 128 without)
    "pg": [{"PPK_PGROUP_MINB": "5"}, {"PPK_PGROUP_MINB": "4"}, {"PPK_PGROUP_MINB": "5"}, {"PPK_PGROUP_MINB": "4"}],
}
settings = SETS[sys.argv[3] if len(sys.argv) > 3 else "xz"]
KNOBS = sorted({k for v in SETS.values() for st in v for k in st})
out = []
for st in settings:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(st)
    s.run(3)
    s.synchronize()
    t0 = time.perf_counter()
    s.run(K)
    s.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / K
    s.profile(True)
    s.kernel_times(reset=True)
    s.run(3)
    s.synchronize()
    kt = s.kernel_times(reset=True)
    s.profile(False)
    pk = {k: round(v[0] / 3, 4) for k, v in kt.items() if v[1] > 0 and v[0] / 3 > 0.002}
    rec = {"n": n, "setting": st, "ms_per_step": round(wall, 4), "Mcell_s": round(n ** 3 / wall * 1e-3, 1), "per_kernel_ms": pk}
    out.append(rec)
    print(json.dumps(rec), flush=True)
sums, divb = s.diagnostics()
print("max|div B|", divb, "t", s.get_time())
s.close()
