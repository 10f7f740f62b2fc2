"""What the compiled library contains, checked without a GPU (nvcc cross-compiles here):
every device image is sm_100a, the fused pipeline kernels move their tiles with bulk asynchronous
copies completed on mbarriers (the SASS the TMA path compiles to, /opt/skills/guides/B200_PROFILING.md
"SASS mnemonics"), and the registered Float64 shapes -- the kernels the headline numbers are measured
on -- do not spill registers."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pg_fusion_b200", "libpgf_b200.so")
BUILD = os.path.join(ROOT, "pg_fusion_b200", "csrc", "build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


def test_every_device_image_is_sm_100a():
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    images = re.findall(r"ELF file\s+\d+:\s+(\S+)", out.stdout)
    assert len(images) >= 8 and all(name.endswith(".sm_100a.cubin") for name in images), images
    ptx = subprocess.run([CUOBJDUMP, "-lptx", LIB], capture_output=True, text=True, timeout=120)
    assert "sm_" not in ptx.stdout.replace("sm_100a", ""), ptx.stdout     # no PTX for other targets (no JIT fallback path)


@pytest.fixture(scope="module")
def shapes_sass(tmp_path_factory):
    cubin_dir = tmp_path_factory.mktemp("cubin")
    out = subprocess.run([CUOBJDUMP, "-xelf", "pipeline_inst_shapes", LIB], capture_output=True, text=True, timeout=120, cwd=cubin_dir)
    assert out.returncode == 0, out.stderr
    cubins = [f for f in os.listdir(cubin_dir) if f.endswith(".cubin")]
    assert len(cubins) == 1, cubins
    sass = subprocess.run([CUOBJDUMP, "-sass", os.path.join(cubin_dir, cubins[0])], capture_output=True, text=True, timeout=600)
    assert sass.returncode == 0, sass.stderr
    return sass.stdout


def test_pipeline_kernels_stage_tiles_with_bulk_copies_on_mbarriers(shapes_sass):
    kernels = re.split(r"\n\s*Function : ", shapes_sass)[1:]
    pipelines = [k for k in kernels if k.startswith("_ZN3pgf15pipeline_kernel")]
    assert len(pipelines) >= 5, [k.split("\n")[0][:60] for k in kernels]
    for k in pipelines:
        name = k.split("\n")[0]
        assert "UBLKCP" in k, f"{name}: no bulk asynchronous copy (cp.async.bulk) in the SASS"
        assert "SYNCS" in k, f"{name}: no mbarrier instructions in the SASS"


def test_registered_float64_shapes_do_not_spill():
    log = open(os.path.join(BUILD, "pipeline_inst_shapes.ptxas.log")).read()
    entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
    assert len(entries) >= 5
    # template arguments <SINK, ACC, GROUPED, NJ, MAXE, Shape>: ACC 0 = Float64 accumulators (Lj1ELj0E...)
    f64 = [e for e in entries if re.match(r"_ZN3pgf15pipeline_kernelILj1ELj0E", e[0])]
    assert len(f64) >= 3          # Q6, Q1 with 8 and 7 aggregates
    for name, stack, st, ld, regs in f64:
        assert int(st) == 0 and int(ld) == 0, f"{name[:80]} spills {st}/{ld} bytes"
        assert int(regs) <= 128
    # the budget the launch bounds allow: 16 consumer warps + producer on one SM need <= 96 registers ... 128 for 14 warps
    worst = max(int(e[4]) for e in entries)
    assert worst <= 128, worst
    # the Decimal128 GROUP BY shape (Q1 "D": <SINK_AGG, CLS_I128, grouped>) runs 15 + 1 warps so that it may use 128
    # registers: under the 96 of a 576-thread block its two rows of five 128-bit products spilled 200 bytes
    q1d = [e for e in entries if re.match(r"_ZN3pgf15pipeline_kernelILj1ELj2ELb1E", e[0])]
    assert len(q1d) == 1 and (int(q1d[0][2]), int(q1d[0][3])) == (0, 0), q1d


def test_topk_selection_kernel_keeps_its_entries_in_registers():
    """ORDER BY ... LIMIT k: one selection kernel per level (topk_select_kernel), eight order summaries per thread
    in registers, shuffle reductions -- no per-round pass over global memory."""
    log = open(os.path.join(BUILD, "pipeline.ptxas.log")).read()
    m = re.search(r"Compiling entry function '(\S*topk_select_kernel\S*)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                  r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
    assert m, "topk_select_kernel not found in the ptxas log"
    assert int(m.group(3)) <= 64 and int(m.group(5)) <= 64, m.groups()
    assert "topk_local_kernel" not in log and "topk_final_kernel" not in log


def test_compaction_pipeline_kernels_use_bulk_copies_mbarriers_and_l2_hints(tmp_path):
    """probe_kernel.cuh: the join / build-side pipelines stage their predicate columns with cp.async.bulk on
    mbarriers too, and their hot loop (stages A and B) keeps its state in registers: what spills is bounded."""
    out = subprocess.run([CUOBJDUMP, "-xelf", "pipeline_inst_probe", LIB], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    assert out.returncode == 0, out.stderr
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 1, cubins
    sass = subprocess.run([CUOBJDUMP, "-sass", os.path.join(tmp_path, cubins[0])], capture_output=True, text=True, timeout=600).stdout
    kernels = [k for k in re.split(r"\n\s*Function : ", sass)[1:] if k.startswith("_ZN3pgf21probe_pipeline_kernel")]
    # <ACC, T0, SPLIT>: three accumulator classes x {generic, one-string-term} fused, plus the two stage-A/B-only
    # instantiations of the split execution (no accumulators: one class serves all)
    assert len(kernels) == 8, [k.split("\n")[0][:60] for k in kernels]
    for k in kernels:
        name = k.split("\n")[0]
        assert "UBLKCP" in k and "SYNCS" in k, name
        assert "VOTE" in k and "POPC" in k, f"{name}: no ballot / popcount compaction"
    log = open(os.path.join(BUILD, "pipeline_inst_probe.ptxas.log")).read()
    entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
    stage_c = [e for e in entries if e[0].startswith("_ZN3pgf23entries_pipeline_kernel")]
    entries = [e for e in entries if e[0].startswith("_ZN3pgf21probe_pipeline_kernel")]
    assert len(entries) == 8 and len(stage_c) == 3
    for name, stack, st, ld, regs in stage_c:   # stage C as a kernel of its own: 64 registers (32 warps per SM), no spills
        assert int(regs) <= 64 and int(st) == 0 and int(ld) == 0, f"{name[:80]}: {regs} registers, spills {st}/{ld}"
    for name, stack, st, ld, regs in entries:
        if name.endswith("Lb1EEEvNS_7DevPlanE"):   # stages A + B alone need far fewer registers than the fused kernel
            assert int(regs) <= 80 and int(st) == 0 and int(ld) == 0, f"{name[:80]}: {regs} registers, spills {st}/{ld}"
    for name, stack, st, ld, regs in entries:
        # <ACC, T0>: ACC 2 = Decimal128 sums (four-word accumulators in stage C, which runs for joined rows only)
        limit = 256 if "kernelILj2E" in name else 16
        assert int(st) <= limit and int(ld) <= limit, f"{name[:80]} spills {st}/{ld} bytes"
    for k in kernels:
        # the tag windows of stage B travel by asynchronous copies into shared memory (no register scoreboard is held
        # across chunks), and the next page descriptor by a bulk copy of its own
        assert "LDGSTS" in k and "DEPBAR" in k, k.split("\n")[0]
