#!/usr/bin/env python
"""Generates tests/golden/vectors.json: small, committed known-answer vectors for the hot path.

Sources (nothing here is produced by the product under test):
  * Bloom: the reference's own known-answer values (runtime_filter/src/tests.rs:48-83 and the
    SURVEY.md 8c vectors) -- transcribed constants, re-checked here against the oracle;
  * layout: fixed row caps of the scan shapes (page/row_estimator/src/lib.rs:353-371 rule);
  * operators: results of the oracle (oracle/orc_ops.c, the CPU restatement of the DataFusion 44
    semantics) on seeded synthetic TPC-H-shaped tables.  The reference itself cannot run in this
    image (Rust), so these are oracle outputs, not reference outputs ("parity unpinned" at the
    DataFusion boundary).
Run from the repository root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyorc as O  # noqa: E402
from tests import util as U  # noqa: E402


def hexf(x):
    return None if x is None else (float(x).hex() if isinstance(x, float) else int(x))


def main():
    out = {"bloom": [], "row_caps": [], "operators": {}}
    # ---- Bloom known answers
    for bits, k, seed, key, words in ((512, 4, 42, 10, [0x2, 0x800000000, 0, 0, 0x4000000000000, 0, 0x100000, 0]),):
        b = O.Bloom(O.bloom_params(bits, k, seed))
        b.insert_u64(key)
        assert [int(w) for w in b.words] == words
        out["bloom"].append({"bit_count": bits, "hash_count": k, "seed": seed, "keys": [key], "words": words})
    guc = O.Bloom(O.bloom_params(1 << 20, 4, 0x7067667573696f6e))
    assert [guc.bit_index(1, i) for i in range(4)] == [179616, 13537, 896034, 729955]
    guc.insert_keys(np.arange(1, 1001, dtype=np.int64))
    digest = hashlib.sha256(np.ascontiguousarray(guc.words).tobytes()).hexdigest()
    assert digest == "748e72253a2859a687a1452a75ee832856704e4ea717e4cd7305e78e177f8a3c"
    out["bloom"].append({"bit_count": 1 << 20, "hash_count": 4, "seed": 0x7067667573696f6e, "keys": "1..1000",
                         "popcount": int(sum(bin(int(w)).count("1") for w in guc.words)), "sha256_of_le_words": digest,
                         "bit_positions_of_key_1": [179616, 13537, 896034, 729955],
                         "bit_positions_of_key_minus_1": [guc.bit_index((-1) & (2**64 - 1), i) for i in range(4)]})
    # ---- layout: rows per 64 KiB page of the scan shapes
    for name, schema in (("q6_F", U.Q6_SCHEMA), ("q1_F", U.Q1_SCHEMA), ("q3_lineitem_F", U.LINEITEM_Q3_SCHEMA),
                         ("q3_orders_F", U.ORDERS_SCHEMA), ("q3_customer_F", U.CUSTOMER_SCHEMA)):
        out["row_caps"].append({"shape": name, "rows_per_page": O.fixed_row_cap(U.orc_cols(schema), 65516)})
    # ---- operators on seeded tables
    for n, seed in ((20_000, 7), (3_000, 8)):
        li = U.lineitem(n, seed)
        q6 = U.oracle_q6(O.OTable.from_pages(U.q6_pages(li), 65536, U.orc_cols(U.Q6_SCHEMA)))
        q1 = U.oracle_q1(O.OTable.from_pages(U.q1_pages(li), 65536, U.orc_cols(U.Q1_SCHEMA)))
        out["operators"][f"lineitem_n{n}_seed{seed}"] = {
            "q6": {"rows_filtered": int(q6.rows_filtered), "aggs": [hexf(v) for v in q6.aggs[0]]},
            "q1": {"rows_filtered": int(q1.rows_filtered),
                   "groups": {(k[0] + b"|" + k[1]).decode(): [hexf(v) for v in a] for k, a in zip(q1.keys, q1.aggs)}}}
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote tests/golden/vectors.json")


if __name__ == "__main__":
    main()
