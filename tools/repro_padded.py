import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, subprocess
import deo_b200 as D
def run(d,a,n,m):
    L = D.CenteredDifference(d,a,0.1,n)
    M = np.asfortranarray(np.random.default_rng(0).uniform(-1,1,(n+2,m)))
    got = L*M
    Lg = D.CenteredDifference(d,a,0.1,n)
    du = np.zeros((n,m),order="F"); D.mul_(du, Lg, M, flags=D._lib.DEO_FLAG_FORCE_GENERIC)
    print(d,a,n,m, "maxdiff", np.abs(got-du).max(), flush=True)
case = sys.argv[1:]
run(*[int(v) for v in case])
