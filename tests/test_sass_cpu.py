"""The shipped library really contains the Blackwell path (no GPU needed: cuobjdump on diffeqoperators.jl_b200/libdeo_b200.so):
every cubin is sm_100a, the persistent tiled kernel of the bench workload issues TMA loads (UTMALDG), mbarrier operations
(SYNCS) and the setmaxnreg register split (USETMAXREG), and its arithmetic is DFMA / packed FFMA2 -- not a library call."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffeqoperators.jl_b200", "libdeo_b200.so")
K_C5 = "_ZN3deo7k_star2IdLi2ELb1ELi7ELb0EEEv14CUtensorMap_stNS_10StarParamsIT_XT0_EEENS_11Star2LaunchEPKS3_PS3_"
K_C3F32 = "_ZN3deo7k_star2IfLi3ELb1ELi7ELb0EEEv14CUtensorMap_stNS_10StarParamsIT_XT0_EEENS_11Star2LaunchEPKS3_PS3_"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")


def _ops(kernel):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, LIB], capture_output=True, text=True, timeout=300).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", out, flags=re.M)
    assert ops, f"{kernel} not found in the library"
    return ops


def test_every_cubin_is_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, timeout=300).stdout
    cubins = re.findall(r"ELF file\s+\d+: (\S+)", out)
    assert len(cubins) >= 10
    assert all(".sm_100a." in c for c in cubins), [c for c in cubins if ".sm_100a." not in c]


def test_bench_kernel_uses_tma_mbarriers_and_the_register_split():
    ops = _ops(K_C5)
    count = lambda prefix: sum(o.startswith(prefix) for o in ops)
    assert count("UTMALDG") >= 1            # cp.async.bulk.tensor (the producer warp's plane loads)
    assert count("SYNCS") >= 30             # mbarrier init / expect_tx / arrive / try_wait
    assert count("USETMAXREG") == 2         # setmaxnreg.dec (helper warpgroup) + setmaxnreg.inc (compute warpgroups)
    assert count("DFMA") >= 500             # the unrolled stencil arithmetic
    assert count("STG") >= 1


def test_float32_kernel_uses_packed_fma():
    ops = _ops(K_C3F32)
    assert sum(o == "FFMA2" for o in ops) >= 100
