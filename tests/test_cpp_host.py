"""Host side above the C ABI in C++ (include/pgf_b200_plan.hpp): the plan-node surface, the
install_runtime_filters / install_b200_operators rewrites and the lowering to pgf_pipeline.

CPU part: tests/cpp/plan_dump.cpp is compiled with g++ -Werror, linked against libpgf_b200.so and
run without a device; every pgf_pipeline it lowers from a DataFusion-shaped plan tree must be
byte-identical to the one the ctypes PipelineBuilder produces for the same query (the builder is
what the GPU parity tests run, so the two host layers are interchangeable)."""
import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import pytest

from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, TypeTag, _lib
from pg_fusion_b200.worker import PipelineBuilder

from . import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp(name: str, tmp: str) -> str:
    exe = os.path.join(tmp, name)
    lib = os.path.join(ROOT, "pg_fusion_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-L", lib, "-lpgf_b200", f"-Wl,-rpath,{lib}", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    exe = build_cpp("plan_dump", str(tmp_path_factory.mktemp("cpp")))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    trees, pods, checks, cur = {}, {}, {}, None
    for line in out.stdout.splitlines():
        if line.startswith("TREE "):
            cur = line[5:]
            trees[cur] = []
        elif line == "END":
            cur = None
        elif cur is not None:
            trees[cur].append(line)
        elif line.startswith("POD "):
            _, name, idx, hexbytes = line.split()
            pods.setdefault(name, {})[int(idx)] = bytes.fromhex(hexbytes)
        elif line.startswith("CHECK "):
            _, name, verdict = line.split()
            checks[name] = verdict
    return SimpleNamespace(rc=out.returncode, trees=trees, pods=pods, checks=checks, stderr=out.stderr)


def fake_scan(scan_id, schema):
    scan = SimpleNamespace(scan_id=scan_id, schema=list(schema))
    scan.pipeline = lambda: PipelineBuilder(None, scan)
    return scan


def pod_bytes(builder) -> bytes:
    return C.string_at(C.addressof(builder.p), C.sizeof(builder.p))


def diff(a: bytes, b: bytes) -> str:
    bad = [i for i in range(min(len(a), len(b))) if a[i] != b[i]]
    return f"sizes {len(a)} / {len(b)}, first differing offsets {bad[:8]}"


def test_host_checks_pass(dump):
    assert dump.rc == 0, dump.stderr
    assert dump.checks and all(v == "ok" for v in dump.checks.values()), dump.checks
    for name in ("partition_1_is_a_plan_error", "unabsorbed_node_has_no_cpu_operator", "pool_exhaustion_is_soft",
                 "or_predicate_not_absorbed", "runtime_filter_targets", "two_probes_on_one_stream_are_fused", "three_probes_on_one_stream_stay_datafusion",
                 "result_pages_one_per_step", "nothing_after_the_close_step", "limits_record_their_reasons"):
        assert name in dump.checks


def test_struct_size_matches_the_ctypes_mirror(dump):
    assert len(dump.pods["q6"][0]) == C.sizeof(_lib.Pipeline)


def test_q6_plan_lowers_to_the_builder_pod(dump):
    want = pod_bytes(U.gpu_q6(fake_scan(1, U.Q6_SCHEMA)))
    assert dump.pods["q6"] == {0: want}, diff(dump.pods["q6"][0], want)
    assert dump.trees["q6"][0].startswith("B200PipelineExec: scan_id=1") and "WorkerPgScanExec" in dump.trees["q6"][1]


@pytest.mark.parametrize("name", ["q1", "q1_partial_final"])
def test_q1_plan_with_cse_projection_and_sort_lowers_to_the_builder_pod(dump, name):
    # the aggregate arguments go through DataFusion's common-subexpression projection; the SortExec
    # above the aggregate is absorbed; Partial -> Final over one partition collapses to Single
    want = pod_bytes(U.gpu_q1(fake_scan(2, U.Q1_SCHEMA)).order_by([("key", 0, False), ("key", 1, False)]))
    assert dump.pods[name] == {0: want}, diff(dump.pods[name][0], want)


def test_boolean_predicates_lower_to_the_builder_pod(dump):
    # WHERE active AND deleted = false AND k < 2500: a bare Boolean column is the term `active = true`
    schema = [ColumnSpec(TypeTag.Boolean), ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Boolean), ColumnSpec(TypeTag.Float64)]
    want = pod_bytes(fake_scan(8, schema).pipeline().filter(0, Cmp.EQ, True).filter(2, Cmp.EQ, False).filter(1, Cmp.LT, 2500)
                     .aggregate([1], [(AggFunc.SUM, [Factor.of(3)]), (AggFunc.COUNT_STAR, None)]))
    assert dump.pods["flags_filter"] == {0: want}, diff(dump.pods["flags_filter"][0], want)


def test_two_join_probes_on_one_stream_lower_to_the_builder_pods(dump):
    # right-deep chain customer |><| (orders |><| lineitem): lineitem probes the orders table with l_orderkey and, for
    # the rows that matched, the customer table with the matched order's o_custkey (payload 0 of the first join)
    customer, orders, lineitem = fake_scan(3, U.CUSTOMER_SCHEMA), fake_scan(4, U.ORDERS_SCHEMA), fake_scan(5, U.LINEITEM_Q3_SCHEMA)
    want = [pod_bytes(orders.pipeline().build_join(0, [1])),          # inner join's build side first: o_orderkey; o_custkey
            pod_bytes(customer.pipeline().build_join(0, [])),
            pod_bytes(lineitem.pipeline().join(0, 0).join(0, (1, 0)).aggregate([], [(AggFunc.COUNT_STAR, None)]))]   # (table handles are filled in at execute time)
    got = dump.pods["two_probes"]
    assert sorted(got) == [0, 1, 2]
    for i in range(3):
        assert got[i] == want[i], f"pipeline {i}: " + diff(got[i], want[i])


def q3_builders(rf1=None, rf2=None, limit=10):
    customer, orders, lineitem = fake_scan(3, U.CUSTOMER_SCHEMA), fake_scan(4, U.ORDERS_SCHEMA), fake_scan(5, U.LINEITEM_Q3_SCHEMA)
    b1 = customer.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, [], rf1)
    b2 = orders.pipeline()
    if rf1 is not None:
        b2.bloom_probe(rf1, 1)
    b2 = b2.filter(2, Cmp.LT, U.Q3_DATE).join(0, 1).build_join(0, [2, 3], rf2)
    b3 = lineitem.pipeline()
    if rf2 is not None:
        b3.bloom_probe(rf2, 0)
    b3 = (b3.filter(3, Cmp.GT, U.Q3_DATE).join(0, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])])
          .order_by(U.Q3_ORDER, limit=limit))
    return [b1, b2, b3]     # build sides before the pipeline that probes them


def test_q3_plan_lowers_to_three_dependent_pipelines(dump):
    want = [pod_bytes(b) for b in q3_builders()]
    got = dump.pods["q3"]
    assert sorted(got) == [0, 1, 2]
    for i in range(3):
        assert got[i] == want[i], f"pipeline {i}: " + diff(got[i], want[i])
    tree = dump.trees["q3"]
    assert [ln.strip().split(":")[0] for ln in tree] == ["B200PipelineExec", "B200PipelineExec", "B200PipelineExec",
                                                         "WorkerPgScanExec", "WorkerPgScanExec", "WorkerPgScanExec"]


def test_q3_runtime_filters_follow_the_reference_rewrite(dump):
    # install_runtime_filters visits the inner join first (runtime_filter_plan.rs:27-48): filter 101 is built
    # from c_custkey and probed on o_custkey, filter 102 from o_orderkey and probed on l_orderkey
    rf1, rf2 = SimpleNamespace(handle=101, generation=1), SimpleNamespace(handle=102, generation=1)
    want = [pod_bytes(b) for b in q3_builders(rf1, rf2)]
    got = dump.pods["q3_filters"]
    for i in range(3):
        assert got[i] == want[i], f"pipeline {i}: " + diff(got[i], want[i])
    before = "\n".join(dump.trees["q3_filters_before"])
    assert before.count("RuntimeFilterBuildExec: key_index=0") == 2
    assert "WorkerPgScanExec: scan_id=4, runtime_filter(col=1)" in before
    assert "WorkerPgScanExec: scan_id=5, runtime_filter(col=0)" in before
    # pool exhausted after one slot: only the customer -> orders filter exists
    want = [pod_bytes(b) for b in q3_builders(rf1, None)]
    got = dump.pods["q3_pool_exhausted"]
    for i in range(3):
        assert got[i] == want[i], f"pipeline {i}: " + diff(got[i], want[i])


def test_decimal_plans_lower_to_the_builder_pods(dump):
    # Decimal128 literals at the column's scale, Date32 day numbers, a flipped comparison
    # (literal >= column) and the (x + c) spelling of a factor
    want = pod_bytes(U.gpu_q6_d(fake_scan(6, U.Q6_D_SCHEMA)))
    assert dump.pods["q6_decimal"] == {0: want}, diff(dump.pods["q6_decimal"][0], want)
    want = pod_bytes(U.gpu_q1_d(fake_scan(7, U.Q1_D_SCHEMA)))
    assert dump.pods["q1_decimal"] == {0: want}, diff(dump.pods["q1_decimal"][0], want)


# ---- randomly generated plans: the C++ lowering and the Python builder must agree byte for byte ----
T_I16, T_I32, T_I64, T_F64, T_VIEW, T_DEC = 2, 3, 4, 6, 8, 10
OPS = {"lt": Cmp.LT, "le": Cmp.LE, "gt": Cmp.GT, "ge": Cmp.GE, "eq": Cmp.EQ, "ne": Cmp.NE}


def random_plan(rng, scan_id):
    """Returns (spec lines for tests/cpp/plan_from_spec.cpp, the equivalent PipelineBuilder)."""
    from pg_fusion_b200 import ColumnSpec, TypeTag
    cls = rng.choice([T_F64, T_DEC, T_I64])                       # arithmetic class of the aggregates
    ncols = rng.randint(2, 8)
    types = [rng.choice([cls, cls, T_VIEW, T_I32, T_I16, T_I64]) for _ in range(ncols)]
    types[0] = cls
    schema = [ColumnSpec(TypeTag(t)) for t in types]
    lines = [f"scan {scan_id} {ncols} " + " ".join(str(t) for t in types)]
    b = fake_scan(scan_id, schema).pipeline()

    def lit_for(t):
        if t == T_F64:
            v = rng.choice([0.05, 24.0, -1.5, 1e9, 0.07, rng.uniform(-1e6, 1e6)])
            return "f", repr(v), v
        if t == T_VIEW:
            s = "".join(rng.choice("ABCXYZ0123456789-") for _ in range(rng.randint(1, 12)))
            return "s", s, s.encode()
        if t == T_DEC:
            v = rng.choice([5, 7, 2400, -100, rng.randint(-2**62, 2**62)])
            return "d", str(v), v
        v = rng.choice([0, 1, -1, 10471, rng.randint(-2**31, 2**31 - 1)]) if t != T_I64 else rng.randint(-2**62, 2**62)
        return "i", str(v), v

    for _ in range(rng.randint(0, 6)):
        c = rng.randrange(ncols)
        op = rng.choice(list(OPS))
        kind, text, value = lit_for(types[c])
        flip = rng.random() < 0.3
        # written flipped in the plan (<literal> <op'> <column>): the lowering must turn it back
        lines.append(f"term {c} {op} {kind} {text}" + (" flip" if flip else ""))
        b.filter(c, OPS[op], value)
    keyable = [c for c in range(ncols) if types[c] != T_F64]
    keys, words = [], 0
    for c in rng.sample(keyable, min(len(keyable), rng.randint(0, 3))):
        w = 1 if types[c] in (T_I16, T_I32, T_I64) else 2
        if words + w <= 4:
            keys.append(c)
            words += w
            lines.append(f"key {c}")
    cls_cols = [c for c in range(ncols) if types[c] == cls]
    aggs = []
    for _ in range(rng.randint(1, 6)):
        func = rng.choice(["sum", "avg", "count", "countstar"] if cls != T_I64 else ["sum", "count", "countstar"])
        if func == "countstar":
            lines.append("agg countstar 0")
            aggs.append((AggFunc.COUNT_STAR, None))
            continue
        nf = rng.randint(1, 3)
        factors, parts = [], []
        for f in range(nf):
            c = rng.choice(cls_cols)
            kind = rng.choice([0, 0, 1, 2])
            if kind == 0:
                parts.append(f"0 {c} i 0")
                factors.append(Factor.of(c))
            else:
                k, text, value = lit_for(cls)
                parts.append(f"{kind} {c} {k} {text}")
                factors.append(Factor.const_minus(value, c) if kind == 1 else Factor.const_plus(value, c))
        lines.append(f"agg {func} {nf} " + " ".join(parts))
        aggs.append(({"sum": AggFunc.SUM, "avg": AggFunc.AVG, "count": AggFunc.COUNT}[func], factors))
    b.aggregate(keys, aggs)
    order = []
    for _ in range(rng.randint(0, 2)):
        is_agg = rng.random() < 0.5 or not keys
        index = rng.randrange(len(aggs)) if is_agg else rng.randrange(len(keys))
        desc, nulls_first = rng.random() < 0.5, rng.random() < 0.5
        lines.append(f"sort {int(is_agg)} {index} {int(desc)} {int(nulls_first)}")
        order.append(("agg" if is_agg else "key", index, desc, nulls_first))
    limit = rng.choice([0, 10, 64, 1000]) if order else 0
    if order:
        if limit:
            lines.append(f"limit {limit}")
        b.order_by(order, limit=limit)
    lines.append("end")
    return lines, b


def test_random_plans_lower_identically_in_both_host_layers(tmp_path):
    import random
    exe = build_cpp("plan_from_spec", str(tmp_path))
    rng = random.Random(20240611)
    specs, builders = [], []
    for i in range(400):
        lines, b = random_plan(rng, 100 + i)
        specs.extend(lines)
        builders.append((lines, b))
    out = subprocess.run([exe], input="\n".join(specs) + "\n", capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    got = out.stdout.splitlines()
    assert len(got) == len(builders)
    for line, (spec, b) in zip(got, builders):
        assert line.startswith("POD "), (line, spec)
        want = pod_bytes(b)
        have = bytes.fromhex(line[4:])
        assert have == want, (diff(have, want), spec)


def test_planner_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """plan_dump and 100 random plans again, built with -fsanitize=address,undefined (leak check on):
    the lowering juggles short-lived expression trees and must not read freed nodes."""
    import random
    lib = os.path.join(ROOT, "pg_fusion_b200")
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:protect_shadow_gap=0", UBSAN_OPTIONS="halt_on_error=1")
    rng = random.Random(7)
    spec = "\n".join("\n".join(random_plan(rng, 500 + i)[0]) for i in range(100)) + "\n"
    for name, stdin in (("plan_dump", ""), ("plan_from_spec", spec)):
        exe = os.path.join(str(tmp_path), name + "_asan")
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-pthread",
               "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
               "-L", lib, "-lpgf_b200", f"-Wl,-rpath,{lib}", "-o", exe]
        built = subprocess.run(cmd, capture_output=True, text=True)
        if built.returncode != 0 and "asan" in built.stderr.lower():
            pytest.skip("sanitizer runtime not available")
        assert built.returncode == 0, built.stderr
        out = subprocess.run([exe], input=stdin, capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0 and "ERROR" not in out.stderr and "runtime error" not in out.stderr, out.stderr[-3000:]
