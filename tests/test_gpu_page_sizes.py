"""pg_fusion.page_size is a GUC (pg/extension/src/guc.rs:31-32): the library takes it per context.
The same shapes must give the oracle's results with 8 KiB, 16 KiB and 256 KiB pages (different rows
per page, tiles per page and ring shapes), for generated scans too."""
import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import AggFunc, BloomParams, Cmp, ColumnSpec, Factor, GenTable, TypeTag

from . import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("page_size", [8192, 16384, 262144])
def test_shapes_with_other_page_sizes(page_size):
    with pg.Context(0, page_size=page_size) as ctx:
        li = U.lineitem(40_000, 3)
        for schema, pages, plan, orc in ((U.Q6_SCHEMA, U.q6_pages(li, page_size), U.gpu_q6, U.oracle_q6),
                                         (U.Q1_SCHEMA, U.q1_pages(li, page_size), U.gpu_q1, U.oracle_q1)):
            assert pages.shape[1] == page_size
            scan = ctx.declare_scan(schema)
            scan.push_pages(pages)
            scan.finish()
            res = plan(scan).run()
            want = orc(O.OTable.from_pages(pages, page_size, U.orc_cols(schema)))
            assert res.rows_in == 40_000 and res.rows_filtered == want.rows_filtered
            U.assert_agg_equal(res, want)
            scan.release()
        # generated pages of this size: layout valid, Q3 shape (joins + Bloom) equals the oracle
        cust = ctx.gen_scan(GenTable.CUSTOMER_Q3, 1500, seed=42)
        orders = ctx.gen_scan(GenTable.ORDERS_Q3, 15_000, seed=42, scale_rows=1500)
        lit = ctx.gen_scan(GenTable.LINEITEM_Q3, 60_000, seed=42, scale_rows=15_000)
        tabs = [O.OTable.from_pages(s.read_pages(), page_size, U.orc_cols(sc)) for s, sc in
                ((cust, U.CUSTOMER_SCHEMA), (orders, U.ORDERS_SCHEMA), (lit, U.LINEITEM_Q3_SCHEMA))]
        assert tabs[2].rows == 60_000
        want, _ = U.oracle_q3(*tabs)
        res, st = U.gpu_q3(ctx, cust, orders, lit, (BloomParams.new(1 << 12, 4, 7), BloomParams.new(1 << 16, 4, 7)))
        assert res.rows_out == want.rows_joined
        U.assert_agg_equal(res, want)
        top = U.gpu_q3(ctx, cust, orders, lit, None, limit=10)[0]
        assert [k[0] for k in top.keys] == [r[0] for r in U.top10(want)]
