"""include/pgf_b200.h is usable from plain C: examples/q6_abi.c compiles with gcc -std=c11 against the
header, links against libpgf_b200.so and (on a GPU) runs the Q6 shape through the C ABI only."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp, name="q6_abi"):
    exe = os.path.join(tmp, name)
    lib = os.path.join(ROOT, "pg_fusion_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), "-L", lib, "-lpgf_b200", f"-Wl,-rpath,{lib}", "-o", exe]
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
        build(tmp, "q6_sharded_abi")      # the multi-GPU consumer (pgf_comm_*, pgf_pipeline_run_sharded) compiles and links too


@pytest.mark.gpu
def test_q6_through_the_c_abi_only():
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.run([build(tmp), "2000000"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "rows_in=2000000" in out.stdout and out.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_sharded_q6_through_the_c_abi_only():
    """examples/q6_sharded_abi.c: one process per GPU, collectives inside the library.  On a 1-GPU box the single-rank
    form runs (pgf_pipeline_run_sharded degenerates to pgf_pipeline_run); with >= 2 GPUs two ranks exchange a
    communicator id through a file and must print the same merged result as the single rank."""
    import torch
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp, "q6_sharded_abi")
        one = subprocess.run([exe, "0", "1", os.path.join(tmp, "id1"), "3000000"], capture_output=True, text=True, timeout=300)
        assert one.returncode == 0 and one.stdout.strip().endswith("ok"), one.stdout + one.stderr
        merged = one.stdout.split("merged ")[1].split("\n")[0]
        if torch.cuda.device_count() >= 2:
            procs = [subprocess.Popen([exe, str(r), "2", os.path.join(tmp, "id2"), "3000000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                     for r in range(2)]
            for p in procs:
                out, err = p.communicate(timeout=300)
                assert p.returncode == 0 and out.strip().endswith("ok"), out + err
                got = out.split("merged ")[1].split("\n")[0]
                assert got.split("count=")[1] == merged.split("count=")[1]           # counts: exact
                a, b = float(got.split("revenue=")[1].split()[0]), float(merged.split("revenue=")[1].split()[0])
                assert abs(a - b) <= 1e-12 * abs(b)
