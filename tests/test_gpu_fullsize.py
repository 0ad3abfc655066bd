"""BASELINE.json's configurations at their FULL sizes, every point against the CPU oracle (SURVEY 8d "Values / seeds").

The oracle runs on all host threads, slab by slab along the last axis (tools/workloads.oracle_rows: a sub-problem with a
discarded margin next to each cut reproduces the global problem's rows; bit-identical for uniform grids, within 1e-15 for
the non-uniform config whose upwind weights are built from global cumulative coordinates), so memory stays bounded at
1024^3.  Gates are north_star's: max|gpu - oracle| / max|oracle| <= 1e-13 (Float64) / 1e-5 (Float32) over ALL points; the
boundary-only maximum (points within 4 rows of a face) is reported and gated too.  Chunk scheduling, 32-bit index paths and
the many-wave grids only exist at these sizes."""
import os

import numpy as np
import pytest

from tests.helpers import TOL
from tools import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import deo_b200
    deo_b200.load_library()
    return deo_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _check_full(D, O, name, slab):
    shape, _, dtype, _ = W.WORKLOADS[name]
    n = shape[-1]
    u = W.field_planes(shape, dtype, 0, n)
    G = W.build_operator(D, name, shape, dtype)
    plan = D.build_plans(G, shape, shape, dtype)[0][0]
    ud = D.DeviceArray.from_host(u)
    dud = D.DeviceArray(shape, dtype)
    plan.apply(dud, ud)
    got = dud.to_host()
    del ud, dud
    kernel = plan.info[0]
    assert kernel != "generic", f"{name} must run on a tiled kernel, got {kernel}"
    nthreads = os.cpu_count() or 1
    u_of = lambda a, b: np.asfortranarray(u[..., a:b])
    err = ref = berr = 0.0
    for z0 in range(0, n, slab):
        z1 = min(z0 + slab, n)
        e, r, b = W.check_rows(O, name, shape, dtype, got[..., z0:z1], z0, z1, u_of, nthreads)
        err, ref, berr = max(err, e), max(ref, r), max(berr, b)
    tol = TOL[np.dtype(dtype)]
    print(f"[full-size] {name} {shape} {np.dtype(dtype).name} kernel={kernel}: max rel err {err / ref:.3e} (boundary-only {berr / ref:.3e}), gate {tol:.0e}")
    assert err / ref <= tol, f"{name}: {err / ref:.3e}"
    assert berr / ref <= tol


def test_full_size_config1_every_point(D, O):
    _check_full(D, O, "C1", 10 ** 6)


def test_full_size_config2_every_point(D, O):
    _check_full(D, O, "C2", 1024)


@pytest.mark.parametrize("name", ["C3", "C3f32"])
def test_full_size_config3_every_point(D, O, name):
    _check_full(D, O, name, 128)


def test_full_size_config4_every_point(D, O):
    _check_full(D, O, "C4", 128)


@pytest.mark.timeout(1800)
def test_full_size_config5_every_point(D, O):
    _check_full(D, O, "C5", 128)
