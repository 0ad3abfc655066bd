"""Host-side tiling decisions of the tiled kernel (csrc/tiling_host.hpp: tile-origin shifts, march-axis chunk length) swept on the
CPU: tests/tiling_driver.cpp compiles the very header the library is built from and checks, for every extent / radius /
vector length / chunk bound, that first and last tiles can hold their faces, that the shift is the smallest one with the fewest
tiles (against brute force), that chunk lists never cut a face's one-sided rows and respect the TABLE variants' cap."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_shifts_and_chunks_hold_their_invariants(tmp_path):
    exe = str(tmp_path / "tiling_driver")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "tiling_driver.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout[-2000:] + r.stderr[-2000:]
    assert int(r.stdout.split()[1]) > 100000
