"""include/pgf_b200.h is usable from plain C: examples/q6_abi.c compiles with gcc -std=c11 against the
header, links against libpgf_b200.so and (on a GPU) runs the Q6 shape through the C ABI only."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp):
    exe = os.path.join(tmp, "q6_abi")
    lib = os.path.join(ROOT, "pg_fusion_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "q6_abi.c"), "-L", lib, "-lpgf_b200", f"-Wl,-rpath,{lib}", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_compiles_as_c_and_links():
    import torch
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        if not torch.cuda.is_available():
            # no device: the library refuses to create a context (PGF_ERR_NO_DEVICE = 3) -- there is no CPU fallback
            out = subprocess.run([exe, "1000"], capture_output=True, text=True)
            assert out.returncode == 1 and "status 3" in out.stderr


@pytest.mark.gpu
def test_q6_through_the_c_abi_only():
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.run([build(tmp), "2000000"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "rows_in=2000000" in out.stdout and out.stdout.strip().endswith("ok")
