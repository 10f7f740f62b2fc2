"""Committed known-answer vectors (tests/golden/vectors.json, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (guards the checker).  GPU: the product reproduces them
through the C ABI -- Bloom words bit-exact, counts exact, Float64 sums within 1e-12 relative."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pyorc as O

from . import util as U

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def unhex(v):
    return float.fromhex(v) if isinstance(v, str) else v


def bloom_keys(spec):
    return np.arange(1, 1001, dtype=np.int64) if spec["keys"] == "1..1000" else np.array(spec["keys"], dtype=np.int64)


def check_bloom_words(words, spec):
    if "words" in spec:
        assert [int(w) for w in words] == spec["words"]
    else:
        assert hashlib.sha256(np.ascontiguousarray(words).tobytes()).hexdigest() == spec["sha256_of_le_words"]
        assert int(sum(bin(int(w)).count("1") for w in words)) == spec["popcount"]


def test_oracle_reproduces_golden_vectors():
    for spec in GOLD["bloom"]:
        b = O.Bloom(O.bloom_params(spec["bit_count"], spec["hash_count"], spec["seed"]))
        b.insert_keys(bloom_keys(spec))
        check_bloom_words(b.words, spec)
        if "bit_positions_of_key_1" in spec:
            assert [b.bit_index(1, i) for i in range(4)] == spec["bit_positions_of_key_1"]
            assert [b.bit_index(2**64 - 1, i) for i in range(4)] == spec["bit_positions_of_key_minus_1"]
    schemas = {"q6_F": U.Q6_SCHEMA, "q1_F": U.Q1_SCHEMA, "q3_lineitem_F": U.LINEITEM_Q3_SCHEMA,
               "q3_orders_F": U.ORDERS_SCHEMA, "q3_customer_F": U.CUSTOMER_SCHEMA}
    for rc in GOLD["row_caps"]:
        assert O.fixed_row_cap(U.orc_cols(schemas[rc["shape"]]), 65516) == rc["rows_per_page"]
    for name, ops in GOLD["operators"].items():
        n, seed = int(name.split("_n")[1].split("_")[0]), int(name.split("seed")[1])
        li = U.lineitem(n, seed)
        q6 = U.oracle_q6(O.OTable.from_pages(U.q6_pages(li), 65536, U.orc_cols(U.Q6_SCHEMA)))
        assert q6.rows_filtered == ops["q6"]["rows_filtered"]
        assert [v for v in q6.aggs[0]] == [unhex(v) for v in ops["q6"]["aggs"]]   # same code, same order: bit-exact


@pytest.mark.gpu
def test_gpu_reproduces_golden_vectors():
    import pg_fusion_b200 as pg
    with pg.Context(0) as ctx:
        for spec in GOLD["bloom"]:
            rf = ctx.runtime_filter(pg.BloomParams.new(spec["bit_count"], spec["hash_count"], spec["seed"]))
            rf.try_acquire_builder()
            rf.insert_keys(bloom_keys(spec))
            rf.publish_ready()
            check_bloom_words(rf.words(), spec)
        for name, ops in GOLD["operators"].items():
            n, seed = int(name.split("_n")[1].split("_")[0]), int(name.split("seed")[1])
            li = U.lineitem(n, seed)
            for shape, schema, pages, plan in (("q6", U.Q6_SCHEMA, U.q6_pages(li), U.gpu_q6), ("q1", U.Q1_SCHEMA, U.q1_pages(li), U.gpu_q1)):
                scan = ctx.declare_scan(schema)
                scan.push_pages(pages)
                scan.finish()
                res = plan(scan).run()
                assert res.rows_filtered == ops[shape]["rows_filtered"]
                if shape == "q6":
                    want = {(): [unhex(v) for v in ops["q6"]["aggs"]]}
                else:
                    want = {tuple(k.encode().split(b"|")): [unhex(v) for v in a] for k, a in ops["q1"]["groups"].items()}
                got = res.by_key()
                assert set(got) == set(want)
                for k in want:
                    for j, (x, y) in enumerate(zip(got[k], want[k])):
                        U.assert_close(x, y, 1e-12, f"{name} {shape} group {k} agg {j}")
                scan.release()
