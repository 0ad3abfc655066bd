import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D
n = 10 ** 6
G = D.CenteredDifference(2, 2, 1.0 / (n + 1), n) * D.Dirichlet0BC(np.float64)
plan = D.build_plans(G, (n,), (n,), np.float64)[0][0]
u = D.DeviceArray.from_host(np.random.default_rng(0).uniform(-1, 1, n))
du = D.DeviceArray((n,), np.float64)
for _ in range(5): plan.apply(du, u)
D.sync()
ms = min(plan.time(du, u, 1000) for _ in range(10))
print(os.environ.get("DEO_LINE_BLOCK", "256"), plan.info[0], f"{ms*1e3:.3f} us/apply  {n/ms/1e6:.1f} Gpts/s")
