import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D
shape = tuple(int(v) for v in sys.argv[1].split("x")); a = int(sys.argv[2])
h = tuple(1.0 / (s + 1) for s in shape)
A = D.CenteredDifference[1](2, a, h[0], shape[0])
for ax in range(2, len(shape) + 1):
    A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1])
Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape))
plan = D.build_plans(A * Q, shape, shape, np.float64)[0][0]
u = D.DeviceArray.from_host(np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, shape)))
du = D.DeviceArray(shape, np.float64)
for _ in range(4):
    plan.apply(du, u)
D.sync()
print(plan.info, plan.time(du, u, 5))
