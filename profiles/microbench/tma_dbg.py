import sys
import os; sys.path.insert(0, os.getcwd())
import numpy as np, ppkmhd_b200 as ppk
from oracle import oracle as O
ini = O.make_ini("orszag_tang", (64, 40, 36), nstepmax=2, extra="[OrszagTang]\nkt=1\n", tend=10.0)
p,t_end,n = ppk.params_from_ini(ini, exact=True)
s=ppk.Mhd3d(p); s.set_pipeline("unfused")
s.upload(ppk.init_condition_from_ini(ini)); s.set_time(0.0,t_end,0)
s.step(); s.synchronize()
print(s.get_time())
orc=O.Oracle(ini); orc.step()
print('equal', np.array_equal(s.interior(), orc.interior()))
