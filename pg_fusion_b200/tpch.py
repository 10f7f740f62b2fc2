"""The TPC-H shaped plans of the reference's benchmark (benches/tpch/queries/q06.sql, q01.sql, q03.sql over
benches/tpch/schema.sql:73-89: money Float64, dates ISO text, keys integer) expressed with the pipeline builder,
i.e. the operator chains DataFusion plans for them.  Used by bench.py, the examples and the
parity tests; the "D" variants are the SURVEY 8d extension schema (Decimal128 money, Date32 dates)."""
from __future__ import annotations

Q3_DATE = b"1995-03-15"
Q3_ORDER = [("agg", 0, True), ("key", 1, False)]   # ORDER BY revenue DESC, o_orderdate (q03.sql)
D_1994, D_1995, D_1998_09_02 = 8766, 9131, 10471   # days since 1970-01-01


def gpu_q6(scan, cols=(0, 1, 2, 3)):
    from . import AggFunc, Cmp, Factor
    q, p, d, s = cols
    return (scan.pipeline()
            .filter(s, Cmp.GE, b"1994-01-01").filter(s, Cmp.LT, b"1995-01-01")
            .filter(d, Cmp.GE, 0.05).filter(d, Cmp.LE, 0.07).filter(q, Cmp.LT, 24.0)
            .aggregate([], [(AggFunc.SUM, [Factor.of(p), Factor.of(d)]), (AggFunc.COUNT_STAR, None)]))


def gpu_q1(scan):
    from . import AggFunc, Cmp, Factor
    q, p, d, t, rf, ls, s = range(7)
    disc_price = [Factor.of(p), Factor.const_minus(1.0, d)]
    charge = disc_price + [Factor.const_plus(1.0, t)]
    aggs = [(AggFunc.SUM, [Factor.of(q)]), (AggFunc.SUM, [Factor.of(p)]), (AggFunc.SUM, disc_price),
            (AggFunc.SUM, charge), (AggFunc.AVG, [Factor.of(q)]), (AggFunc.AVG, [Factor.of(p)]),
            (AggFunc.AVG, [Factor.of(d)]), (AggFunc.COUNT_STAR, None)]
    return scan.pipeline().filter(s, Cmp.LE, b"1998-09-02").aggregate([rf, ls], aggs)


def gpu_q3(ctx, customer, orders, lineitem, bloom_params=None, segment=b"BUILDING", limit=0):
    """Runs the three fused pipelines of the Q3 shape; returns (result, stats dict).
    limit > 0 adds ORDER BY revenue DESC, o_orderdate LIMIT n (device top-k)."""
    from . import AggFunc, Cmp, Factor
    stats = {}
    rf1 = rf2 = None
    if bloom_params is not None:
        rf1 = ctx.runtime_filter(bloom_params[0])
        rf1.try_acquire_builder()
    # customer(BUILDING) -> join table T1 keyed by c_custkey (+ Bloom for the orders scan)
    hint = lambda scan, frac: int(scan.info().rows * frac) + 4096   # build-row buffer sizing (too small = one more pass)
    r1 = customer.pipeline().filter(1, Cmp.EQ, segment).build_join(0, [], rf1, expected_rows=hint(customer, 0.25)).run()
    if rf1 is not None:
        rf1.publish_ready()
    # orders: [Bloom probe] -> o_orderdate < date -> probe T1 -> T2 keyed by o_orderkey with payload
    p2 = orders.pipeline()
    if rf1 is not None:
        p2.bloom_probe(rf1, 1)
        rf2 = ctx.runtime_filter(bloom_params[1])
        rf2.try_acquire_builder()
    r2 = p2.filter(2, Cmp.LT, Q3_DATE).join(r1.join_table, 1).build_join(0, [2, 3], rf2, expected_rows=hint(orders, 0.125)).run()
    if rf2 is not None:
        rf2.publish_ready()
    # lineitem: [Bloom probe] -> l_shipdate > date -> probe T2 -> GROUP BY l_orderkey, o_orderdate, o_shippriority
    p3 = lineitem.pipeline()
    if rf2 is not None:
        p3.bloom_probe(rf2, 0)
    p3 = (p3.filter(3, Cmp.GT, Q3_DATE).join(r2.join_table, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                     expected_groups=max(1024, r2.rows_out)))
    if limit:
        p3.order_by(Q3_ORDER, limit=limit)
    r3 = p3.run()
    ctx.destroy_join_table(r1.join_table)
    ctx.destroy_join_table(r2.join_table)
    stats.update(customer=r1, orders=r2, lineitem=r3, rf1=rf1, rf2=rf2)
    return r3, stats


def gpu_q3_sharded(ctx, customer, orders, lineitem, world, device, bloom_params=None, segment=b"BUILDING", limit=0):
    """The Q3 shape with every scan sharded by pages over `world` ranks (one process per GPU):
    broadcast joins, OR-merged runtime filters, Partial -> Final aggregate."""
    from . import AggFunc, Cmp, Factor
    from . import multi_gpu as MG
    rf1 = rf2 = None
    if bloom_params is not None:
        rf1 = ctx.runtime_filter(bloom_params[0])
        rf1.try_acquire_builder()
    r1 = customer.pipeline().filter(1, Cmp.EQ, segment).build_join(0, [], rf1).run()
    t1 = MG.broadcast_join_table(ctx, r1.join_table, world, device)
    if rf1 is not None:
        MG.or_merge_filter(rf1, world, device)
        rf1.publish_ready()
    p2 = orders.pipeline()
    if rf1 is not None:
        p2.bloom_probe(rf1, 1)
        rf2 = ctx.runtime_filter(bloom_params[1])
        rf2.try_acquire_builder()
    r2 = p2.filter(2, Cmp.LT, Q3_DATE).join(t1, 1).build_join(0, [2, 3], rf2).run()
    t2 = MG.broadcast_join_table(ctx, r2.join_table, world, device)
    if rf2 is not None:
        MG.or_merge_filter(rf2, world, device)
        rf2.publish_ready()
    p3 = lineitem.pipeline()
    if rf2 is not None:
        p3.bloom_probe(rf2, 0)
    total_orders = ctx.join_table_info(t2).rows
    p3 = (p3.filter(3, Cmp.GT, Q3_DATE).join(t2, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                     expected_groups=max(1024, total_orders)))
    if limit:
        p3.order_by(Q3_ORDER, limit=limit)   # applied to the merged (final) groups
    res, stats = MG.merge_partial_aggregate(p3, world, device, max_groups=max(1024, total_orders))
    ctx.destroy_join_table(t1)
    ctx.destroy_join_table(t2)
    return res, dict(customer=r1, orders=r2, lineitem=stats)


def gpu_q6_d(scan):
    from . import AggFunc, Cmp, Factor
    return (scan.pipeline().filter(3, Cmp.GE, D_1994).filter(3, Cmp.LT, D_1995)
            .filter(2, Cmp.GE, 5).filter(2, Cmp.LE, 7).filter(0, Cmp.LT, 2400)
            .aggregate([], [(AggFunc.SUM, [Factor.of(1), Factor.of(2)]), (AggFunc.COUNT_STAR, None)]))


def gpu_q1_d(scan):
    from . import AggFunc, Cmp, Factor
    disc_price = [Factor.of(1), Factor.const_minus(100, 2)]
    charge = disc_price + [Factor.const_plus(100, 3)]
    aggs = [(AggFunc.SUM, [Factor.of(0)]), (AggFunc.SUM, [Factor.of(1)]), (AggFunc.SUM, disc_price), (AggFunc.SUM, charge),
            (AggFunc.AVG, [Factor.of(0)]), (AggFunc.AVG, [Factor.of(1)]), (AggFunc.AVG, [Factor.of(2)]), (AggFunc.COUNT_STAR, None)]
    return scan.pipeline().filter(6, Cmp.LE, D_1998_09_02).aggregate([4, 5], aggs)


def q3_bloom_params(ncust: int, nord: int, bits_per_key: int = 16):
    """Runtime filters sized bits_per_key per expected build key (a fifth of the customers, a tenth of the orders)."""
    from . import BloomParams

    def pow2(n):
        b = 1
        while b < n:
            b <<= 1
        return b
    return (BloomParams.new(pow2(bits_per_key * max(1, ncust // 5)), 4, 7), BloomParams.new(pow2(bits_per_key * max(1, nord // 10)), 4, 7))


def gpu_q3_partitioned(ctx, customer, orders, lineitem, nord_total: int, segment=b"BUILDING", limit=10, rf=None):
    """The Q3 shape over page-sharded scans with HASH-PARTITIONED joins, every collective inside
    the library (pgf_comm_*): the context must carry a communicator.

      customer  -> build rows            -> BROADCAST (small)                     -> T1 on every rank
      orders    -> probe T1 -> build rows (+ runtime filter over o_orderkey)      -> PARTITION by hash(o_orderkey)
                                                                                   -> T2 = this rank's share of the orders
                   runtime filter: OR all-reduce                                  -> the filter knows every build key
      lineitem  -> date filter -> runtime filter (dense lanes) -> rows {l_orderkey; extendedprice, discount}
                                                                                   -> PARTITION by hash(l_orderkey)
      received rows -> probe T2 -> GROUP BY -> top `limit`    (every group lives on the rank that owns its key: no merge)
      the ranks' top rows -> all-gather (a few hundred bytes) -> the global top `limit`

    Only ~1 % of the lineitem rows cross NVLink: the ones the runtime filter cannot rule out.
    Returns (rows [(l_orderkey, revenue, o_orderdate, o_shippriority)], stats)."""
    import struct

    from . import AggFunc, Cmp, ColumnSpec, Factor, RuntimeFilterState, TypeTag
    rank, world = ctx.comm_info()
    stats = {"nvlink_bytes": 0}
    # build-row buffers are sized from TPC-H's selectivities (a hint that is too small costs one more pass, never a
    # wrong result): a fifth of the customers, an eighth of the orders, a thirtieth of the lineitems survive
    hint = lambda scan, frac: int(scan.info().rows * frac) + 4096
    r1 = customer.pipeline().filter(1, Cmp.EQ, segment).build_join(0, [], rows_only=True, expected_rows=hint(customer, 0.25)).run()
    t1, sent = ctx.exchange(r1.join_table, partition=False)
    stats["nvlink_bytes"] += sent
    ctx.destroy_join_table(r1.join_table)
    own_rf = rf is None
    if own_rf:
        rf = ctx.runtime_filter(q3_bloom_params(1, nord_total)[1])
    elif rf.snapshot()[1] == RuntimeFilterState.Ready:   # a filter slot recycled from the previous query
        rf.retire_ready_after_quiescence()
    rf.try_acquire_builder()
    r2 = orders.pipeline().filter(2, Cmp.LT, Q3_DATE).join(t1, 1).build_join(0, [2, 3], rf, rows_only=True, expected_rows=hint(orders, 0.125)).run()
    t2, sent = ctx.exchange(r2.join_table, partition=True)
    stats["nvlink_bytes"] += sent
    ctx.destroy_join_table(r2.join_table)
    rf.or_all_reduce()
    rf.publish_ready()
    r3 = lineitem.pipeline().bloom_probe(rf, 0).filter(3, Cmp.GT, Q3_DATE).build_join(0, [1, 2], rows_only=True, expected_rows=hint(lineitem, 1 / 30)).run()
    rs3, sent = ctx.exchange(r3.join_table, partition=True, rows_only=True)
    stats["nvlink_bytes"] += sent
    ctx.destroy_join_table(r3.join_table)
    schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Float64), ColumnSpec(TypeTag.Float64)]
    # every group is an order of this rank's share AND the key of at least one routed row: the smaller count bounds
    # the groups (a group table sized by the orders alone is 4-8 x larger than it need be and falls out of L2)
    p4 = (ctx.row_set_pipeline(rs3, schema).join(t2, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                     expected_groups=max(1024, min(ctx.join_table_info(t2).rows, ctx.join_table_info(rs3).rows))))
    if limit:
        p4.order_by(Q3_ORDER, limit=limit)
    r4 = p4.run()
    for h in (t1, t2, rs3):
        ctx.destroy_join_table(h)
    if own_rf:
        rf.destroy()
    stats.update(customer=r1, orders=r2, lineitem=r3, final=r4, rf=None if own_rf else rf)
    rows = [(int(k[0]), float(a[0]), bytes(k[1]), int(k[2])) for k, a in zip(r4.keys, r4.aggs)]
    if not limit:
        return rows, stats
    # the ranks' top rows -> every rank -> the global top `limit`
    return merge_topk(ctx.comm_all_gather_host(pack_topk(rows, limit)), limit), stats


_TOPK_REC = __import__("struct").Struct("<qd12sq")   # l_orderkey, revenue, o_orderdate (<= 12 bytes), o_shippriority


def pack_topk(rows, limit: int) -> bytes:
    """Fixed-size buffer of one rank's top rows: row count, then `limit` records (zero padded)."""
    import struct
    body = b"".join(_TOPK_REC.pack(k, v, d, p) for k, v, d, p in rows[:limit])
    return struct.pack("<q", min(len(rows), limit)) + body.ljust(_TOPK_REC.size * limit, b"\0")


def merge_topk(buffers, limit: int):
    """ORDER BY revenue DESC, o_orderdate LIMIT `limit` over the ranks' buffers (every rank computes the same list)."""
    import struct
    merged = []
    for buf in buffers:
        n = struct.unpack_from("<q", buf)[0]
        for i in range(n):
            k, v, d, p = _TOPK_REC.unpack_from(buf, 8 + i * _TOPK_REC.size)
            merged.append((k, v, d.rstrip(b"\0"), p))
    merged.sort(key=lambda r: (-r[1], r[2]))
    return merged[:limit]
