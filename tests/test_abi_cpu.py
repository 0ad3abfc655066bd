"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/deo_b200.h declares, its struct images match the header, host-only entry points work, and the
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "deo_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint32_t\s+(deo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    assert len(names) >= 30
    for must in ("deo_plan_create", "deo_plan_apply", "deo_plan_apply_host", "deo_buffer_create", "deo_dist_plan_apply",
                 "deo_plan_update_coefficients", "deo_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import deo_b200 as D
    from deo_b200 import _lib
    L = D.load_library()
    declared = _declared_functions()
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/deo_b200.h but not exported by libdeo_b200.so"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.EXPORTS) == declared
    assert L.deo_abi_version() == 1


def test_struct_images_match_the_header_layout():
    from deo_b200 import _lib
    # deo_op_desc: 10 int32 + 4 pointers; deo_bc_desc: 4 int32 + 4 pointers; deo_plan_desc per the header
    assert C.sizeof(_lib.OpDesc) == 10 * 4 + 4 * 8
    assert C.sizeof(_lib.BcDesc) == 4 * 4 + 4 * 8
    assert _lib.PlanDesc.dims.offset == 8 and _lib.PlanDesc.padded.offset == 8 + 24
    assert _lib.PlanDesc.ops.offset == 8 + 24 + 12 + 8 + 4      # nops, accumulate, then 4 bytes of alignment padding
    assert C.sizeof(_lib.PlanDesc) == _lib.PlanDesc.bc.offset + 3 * C.sizeof(_lib.BcDesc) + 8


def test_slab_partition_is_a_partition():
    from deo_b200.dist import slab_bounds
    for n, p in [(1024, 8), (1000, 3), (17, 4), (8, 8)]:
        nxt = 0
        for r in range(p):
            s, c = slab_bounds(n, p, r)
            assert s == nxt and c in (n // p, n // p + 1)
            nxt = s + c
        assert nxt == n
    from deo_b200 import DeoError
    with pytest.raises(DeoError):
        slab_bounds(3, 4, 0)


def test_no_cpu_fallback_compute_fails_loudly_without_a_device():
    import deo_b200 as D
    L = D.load_library()
    n = C.c_int32(-1)
    rc = L.deo_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    A = D.CenteredDifference(2, 2, 0.1, 16) * D.Dirichlet0BC(np.float64)
    u = np.zeros(16)
    du = np.zeros(16)
    with pytest.raises(D.DeoError) as e:
        D.mul_(du, A, u)
    assert e.value.code == 2          # DEO_ERR_CUDA
    with pytest.raises(D.DeoError):
        D.DeviceArray((16,), np.float64)


def test_host_mirror_builds_the_reference_operand_bundles(golden):
    """The fields that cross the ABI equal the oracle's (Fornberg weights, boundary rows, BC stencils)."""
    import deo_b200 as D
    from oracle import oracle as O
    for T in (np.float64, np.float32):
        for (d, a) in [(1, 2), (2, 2), (2, 4), (2, 6), (3, 4), (4, 4)]:
            p, o = D.CenteredDifference(d, a, T(0.1), 24, dtype=T), O.CenteredDifference(d, a, T(0.1), 24, dtype=T)
            for f in ("stencil_coefs", "low_boundary_coefs", "high_boundary_coefs", "coefficients"):
                assert np.array_equal(np.asarray(getattr(p, f), dtype=T).ravel(), np.asarray(getattr(o, f), dtype=T).ravel()), (T, d, a, f)
            assert (p.stencil_length, p.boundary_stencil_length, p.boundary_point_count) == \
                   (o.stencil_length, o.boundary_stencil_length, o.boundary_point_count)
        for (d, a, off) in [(1, 1, 0), (1, 2, 0), (2, 2, 0), (2, 3, 1), (1, 3, 1)]:
            c = np.linspace(-1, 1, 24).astype(T)
            p, o = D.UpwindDifference(d, a, T(0.1), 24, c, offside=off, dtype=T), O.UpwindDifference(d, a, T(0.1), 24, c, offside=off, dtype=T)
            for f in ("stencil_coefs", "low_boundary_coefs", "high_boundary_coefs", "coefficients"):
                assert np.array_equal(np.asarray(getattr(p, f), dtype=T).ravel(), np.asarray(getattr(o, f), dtype=T).ravel()), (T, d, a, off, f)
        dx = (0.1 * (1 + 0.3 * np.sin(np.arange(1, 26)))).astype(T)
        p, o = D.CenteredDifference(2, 4, dx, 24, dtype=T), O.CenteredDifference(2, 4, dx, 24, dtype=T)
        assert np.array_equal(np.asarray(p.stencil_coefs, dtype=T).ravel(), np.asarray(o.stencil_coefs, dtype=T).ravel())
        q, r = D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 0.1, 3, dtype=T), O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 0.1, 3, T)
        assert np.array_equal(q.a_l, r.a_l) and np.array_equal(q.a_r, r.a_r) and q.b_l == r.b_l and q.b_r == r.b_r


def test_constructor_error_behaviour_matches_the_reference():
    import deo_b200 as D
    with pytest.raises(AssertionError):
        D.CenteredDifference(2, 1, 0.1, 16)            # derivative_operator.jl:84  @assert approximation_order > 1
    A = D.CenteredDifference[1](2, 2, 0.1, 16)
    with pytest.raises(Exception):
        D.build_plans(A, (16, 8), (16, 8), np.float64)   # derivative_operator_functions.jl:40: padded dim or BC required


def test_header_is_valid_c_and_links_from_plain_c(tmp_path):
    """include/deo_b200.h consumed by a C11 translation unit (gcc, no C++), linked against the shipped library."""
    import shutil
    import subprocess
    import deo_b200 as D
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.dirname(D.LIB_PATH)
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_driver.c"), "-o", exe, "-L", libdir, "-l:libdeo_b200.so",
                           f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def _struct_fields(text, name, lang):
    """Field names of struct `name` in the C header / the Julia glue, in declaration order."""
    if lang == "c":
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                m = re.search(r"\*?\s*([A-Za-z_][A-Za-z0-9_]*)\s*(\[[^\]]*\])?\s*$", part.strip())
                names.append(m.group(1))
        return names
    body = re.search(r"struct %s\n(.*?)\nend" % name, text, flags=re.S).group(1)
    return [m.group(1) for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_]*)::", body)]


def test_julia_glue_mirrors_the_header_field_for_field():
    """The Julia binding cannot be executed here, so its struct images are checked statically against the header and
    the ctypes binding (same names, same order), and every ccall names an exported symbol."""
    from deo_b200 import _lib
    hdr = open(HEADER).read()
    jl = open(os.path.join(ROOT, "diffeqoperators.jl_b200", "julia", "DiffEqOperatorsB200.jl")).read()
    for cname, jname, ctype in [("deo_op_desc", "OpDesc", _lib.OpDesc), ("deo_bc_desc", "BcDesc", _lib.BcDesc), ("deo_plan_desc", "PlanDesc", _lib.PlanDesc)]:
        c_fields = _struct_fields(hdr, cname, "c")
        j_fields = _struct_fields(jl, jname, "jl")
        py_fields = [f[0] for f in ctype._fields_]
        assert c_fields == py_fields, (cname, c_fields, py_fields)
        assert c_fields == j_fields, (cname, c_fields, j_fields)
    called = set(re.findall(r"ccall\(\(:(deo_[a-z0-9_]+), libdeo\)", jl))
    assert called and called <= set(_declared_functions()), called - set(_declared_functions())
