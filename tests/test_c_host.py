"""A plain-C host (no Python, no CUDA headers) against include/rst_b200.h + librst_b200.so: the drop-in boundary as a Rust / C
host would use it.  tests/c_host/host_smoke.c is compiled as strict C99 -- which also checks that
the header is valid C -- linked against the in-tree library and run."""
import os
import subprocess

import pytest

from rustedscithe_b200 import build as rbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    lib = rbuild.build()
    exe = str(tmp_path / "host_smoke")
    libdir = os.path.dirname(lib)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "c_host", "host_smoke.c"), "-L", libdir, "-lrst_b200", "-lm", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_host_compiles_links_and_is_refused_without_a_device(tmp_path):
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("a device is present: covered by the gpu test")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.startswith("NO DEVICE"), res.stdout + res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n", [20000, 1000000])
def test_c_host_solves_the_linear_bvp(tmp_path, n):
    exe = _build(tmp_path)
    res = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.startswith("OK"), res.stdout + res.stderr
    print(res.stdout.strip())
