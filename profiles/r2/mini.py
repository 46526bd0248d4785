"""Tiny driver for ncu captures: N steps of Orszag-Tang kt=1 at n^3 through the C ABI (no torch import)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ppkmhd_b200 as ppk  # noqa: E402

sys.path.insert(0, ROOT)
from bench import make_ini  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pipeline = sys.argv[3] if len(sys.argv) > 3 else "auto"
exact = len(sys.argv) > 4 and sys.argv[4] == "exact"
ini = make_ini(n, 1, 10 ** 9)
p, t_end, _ = ppk.params_from_ini(ini, exact=exact)
s = ppk.Mhd3d(p)
if pipeline != "auto":
    s.set_pipeline(pipeline)
s.upload(ppk.init_condition_from_ini(ini))
s.set_time(0.0, t_end, 0)
s.run(steps)
s.synchronize()
print(s.get_time())
s.close()
