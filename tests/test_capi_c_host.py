"""The C ABI from a plain C host: tests/capi_smoke.c includes include/fg_b200.h only.  CPU: it compiles as C99 and
links against libfg_b200.so.  GPU: it runs two train steps + the L-net calls and exits 0."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "face_generator_b200")


def build(tmp_path):
    exe = str(tmp_path / "capi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "capi_smoke.c"), "-L" + LIBDIR, "-lfg_b200", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_library_links(tmp_path):
    build(tmp_path)


@pytest.mark.gpu
def test_c_host_runs_train_steps(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "capi_smoke OK" in r.stdout and "step 2:" in r.stdout
