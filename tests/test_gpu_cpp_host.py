"""The C++ host layer end to end on the GPU: tests/cpp/driver.cpp plans the Q6 / Q1 / Q3 physical
plans (DataFusion node types), rewrites them with install_runtime_filters + install_b200_operators
and executes the resulting B200PipelineExec trees.  Every value it prints is compared with the CPU
oracle run over the same generated pages (this process regenerates them with the same seeds)."""
import json
import subprocess

import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import GenTable

from . import util as U
from .test_cpp_host import build_cpp

pytestmark = pytest.mark.gpu

Q6_ROWS, Q1_ROWS, NCUST, NORD, NLI = 300_000, 200_000, 1500, 15_000, 60_000


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("cpp_gpu")
    exe = build_cpp("driver", str(tmp))
    pages_file = str(tmp / "q1_pages.bin")
    out = subprocess.run([exe, str(Q6_ROWS), str(Q1_ROWS), str(NCUST), str(NORD), str(NLI), pages_file],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = {}
    for line in out.stdout.splitlines():
        d = json.loads(line)
        res[d["query"]] = d
    assert "done" in res
    res["_pages_file"] = pages_file
    return res


@pytest.fixture(scope="module")
def ctx():
    with pg.Context(0) as c:
        yield c


def table(ctx, gen, rows, schema, **kw):
    scan = ctx.gen_scan(gen, rows, seed=42, **kw)
    t = O.OTable.from_pages(scan.read_pages(), 65536, U.orc_cols(schema))
    scan.release()
    return t


def test_q6_plan(runs, ctx):
    want = U.oracle_q6(table(ctx, GenTable.LINEITEM_Q6, Q6_ROWS, U.Q6_SCHEMA))
    got = runs["q6"]
    assert got["root"] == "B200PipelineExec" and got["columns"] == ["revenue", "count(*)"]
    (revenue, count), = got["rows"]
    assert count == want.aggs[0][1] == want.rows_filtered
    U.assert_close(revenue, want.aggs[0][0], 1e-12, "revenue")
    (m,) = got["pipelines"]
    assert m["rows_in"] == Q6_ROWS and m["rows_filtered"] == want.rows_filtered and m["kernel_launches"] >= 1
    assert m["variant"] == "q6_f64"            # the plan tree reaches the registered Q6 instantiation


@pytest.mark.parametrize("name", ["q1", "q1_partial_final"])
def test_q1_plan(runs, ctx, name):
    want = U.oracle_q1(table(ctx, GenTable.LINEITEM_Q1, Q1_ROWS, U.Q1_SCHEMA))
    got = runs[name]
    assert got["columns"][:2] == ["l_returnflag", "l_linestatus"] and len(got["columns"]) == 10
    keys = [(r[0].encode(), r[1].encode()) for r in got["rows"]]
    assert keys == sorted(keys) and len(keys) == 4          # SortExec absorbed: ORDER BY l_returnflag, l_linestatus
    by_key = dict(zip(want.keys, want.aggs))
    assert set(keys) == set(by_key)
    for k, row in zip(keys, got["rows"]):
        w = by_key[k]
        for a in range(7):
            U.assert_close(row[2 + a], w[a], 1e-12, f"{name} agg {a}")
        assert row[9] == w[7]                                # count(*): exact
    assert got["pipelines"][0]["rows_filtered"] == want.rows_filtered


def q3_oracle(ctx):
    ct = table(ctx, GenTable.CUSTOMER_Q3, NCUST, U.CUSTOMER_SCHEMA)
    ot = table(ctx, GenTable.ORDERS_Q3, NORD, U.ORDERS_SCHEMA, scale_rows=NCUST)
    lt = table(ctx, GenTable.LINEITEM_Q3, NLI, U.LINEITEM_Q3_SCHEMA, scale_rows=NORD)
    return U.oracle_q3(ct, ot, lt)


@pytest.mark.parametrize("name", ["q3", "q3_runtime_filters", "q3_streamed"])
def test_q3_plan_top10(runs, ctx, name):
    want, wstats = q3_oracle(ctx)
    got = runs[name]
    assert got["columns"] == ["l_orderkey", "revenue", "o_orderdate", "o_shippriority"]   # the final projection's order
    want10 = U.top10(want)
    assert len(got["rows"]) == 10
    assert [(r[0], r[2].encode(), r[3]) for r in got["rows"]] == [(w[0], w[2], w[3]) for w in want10]
    for r, w in zip(got["rows"], want10):
        U.assert_close(r[1], w[1], 1e-12, "revenue")
    cust, orders, lineitem = got["pipelines"]               # build sides run first, in dependency order
    assert cust["rows_out"] == wstats["customers"] and orders["rows_out"] == wstats["orders"]
    assert lineitem["rows_out"] == want.rows_joined
    if name == "q3_runtime_filters":
        assert cust["bloom_rows"] == wstats["customers"] and orders["bloom_rows"] == wstats["orders"]
        assert orders["rows_bloom"] < orders["rows_in"] and lineitem["rows_bloom"] < lineitem["rows_in"]
    else:
        assert orders["rows_bloom"] == orders["rows_in"] and cust["bloom_rows"] == 0


def test_two_join_probes_fused_by_the_cpp_planner(runs, ctx):
    """The right-deep plan of the same join -- customer |><| (orders |><| lineitem) -- fuses two probes into the lineitem
    stream (the second with a payload column of the first as its key): the same joined rows as the left-deep Q3."""
    want, _ = q3_oracle(ctx)
    got = runs["q3_two_probes"]
    assert got["rows"] == [[want.rows_joined]]
    orders, cust, lineitem = got["pipelines"]               # build sides in probe order: the inner join's first
    assert lineitem["rows_out"] == want.rows_joined and lineitem["variant"] == "compact_1_string_term"


def test_q3_plan_all_groups(runs, ctx):
    want, _ = q3_oracle(ctx)
    got = runs["q3_all_groups"]
    assert len(got["rows"]) == len(want.keys)
    rows = {(r[0], r[2].encode(), r[3]): r[1] for r in got["rows"]}
    for k, a in zip(want.keys, want.aggs):
        U.assert_close(rows[k], a[0], 1e-12, "revenue")
    rev = [r[1] for r in got["rows"]]
    assert rev == sorted(rev, reverse=True)                  # SortExec without fetch: the whole output is ordered


def test_q1_result_pages_from_the_cpp_layer(runs, ctx):
    got = runs["q1"]
    npages = runs["q1_result_pages"]["pages"]
    pages = np.fromfile(runs["_pages_file"], dtype=np.uint8).reshape(npages, 65536)
    # the receiving schema is the PLAN's: l_returnflag / l_linestatus are NOT NULL columns, so DataFusion's
    # AggregateExec output fields for them are non-nullable; SUM / AVG nullable, COUNT(*) not
    cols = [(O.T_UTF8VIEW, False)] * 2 + [(O.T_FLOAT64, True)] * 7 + [(O.T_INT64, False)]
    t = O.OTable.from_pages(pages, 65536, cols)               # runs the reference's import checks on every page
    assert O.import_check(0x4152, 0, np.ascontiguousarray(pages[0, 20:]), [(O.T_UTF8VIEW, True)] * 2 + cols[2:]) != 0
    assert t.rows == 4
    # result pages carry the pod's column order (keys, then aggregates), which is also Q1's output order
    flags = t.column(0)                                       # view columns decode to a list of bytes
    assert [bytes(f) for f in flags] == [r[0].encode() for r in got["rows"]]
    counts, _ = t.column(9)
    assert list(counts) == [r[9] for r in got["rows"]]
    sums, _ = t.column(2)
    assert [float(s) for s in sums] == [r[2] for r in got["rows"]]


def test_ineligible_plan_and_library_errors(runs):
    assert runs["ineligible"] == {"query": "ineligible", "kept": "AggregateExec", "reasons": 1, "not_implemented": True}
    assert runs["unknown_scan"]["status"] == 5               # PGF_ERR_UNKNOWN_HANDLE as DataFusionError::Execution


def i128(v):
    hi, lo = v
    return (hi << 64) | lo


def test_decimal_plans_are_exact(runs, ctx):
    want = U.oracle_q6_d(table(ctx, GenTable.LINEITEM_Q6_D, Q6_ROWS, U.Q6_D_SCHEMA))
    (revenue, count), = runs["q6_decimal"]["rows"]
    assert i128(revenue) == want.aggs[0][0] and count == want.aggs[0][1]           # bit-exact i128 sum
    assert runs["q6_decimal"]["pipelines"][0]["variant"] == "q6_decimal"
    want, _ = U.oracle_q1_d(table(ctx, GenTable.LINEITEM_Q1_D, Q1_ROWS, U.Q1_D_SCHEMA))
    got = runs["q1_decimal"]["rows"]
    assert len(got) == len(want) == 4
    for row in got:
        w = want[(row[0], row[1])]
        assert tuple(i128(v) for v in row[2:9]) == tuple(w[:7]) and row[9] == w[7]
