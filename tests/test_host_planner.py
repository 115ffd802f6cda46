"""Host-side logic of the C++ mirror that needs no GPU: SQL -> plan -> optimizer -> pipeline shape, checked
against the reference's golden EXPLAIN strings (tests/golden/ref_strings.json, README.md:100-117), and the
scalar state plumbing (DataValue JSON wire format, scalar merges) against the reference's vectors."""
import json

import pytest

from conftest import golden
from fuse_query_b200 import _fuse_host as h

DT = h.DataType
TAG = {"Null": 0, "Boolean": 1, "Int8": 2, "Int16": 3, "Int32": 4, "Int64": 5, "UInt8": 6, "UInt16": 7, "UInt32": 8, "UInt64": 9,
       "Float32": 10, "Float64": 11, "Utf8": 12}
OPS = {"Add": h.ops.Add, "Sub": h.ops.Sub, "Mul": h.ops.Mul, "Div": h.ops.Div, "Min": h.ops.Min, "Max": h.ops.Max, "Sum": h.ops.Sum,
       "Count": h.ops.Count}


@pytest.fixture()
def ctx():
    # FuseQueryContext::create_ctx(0, ...) as in the reference tests: 8 source pipes; no device bound
    c = h.FuseQueryContext.create_ctx(0)
    c.options.fuse = False   # reference-shaped pipeline (one processor per plan node)
    return c


STRINGS = golden("ref_strings.json")


@pytest.mark.parametrize("case", [c for c in STRINGS if c["sql"]], ids=lambda c: c["source"])
def test_reference_golden_plan_and_pipeline_strings(ctx, case):
    plan = h.Planner().build_from_sql(ctx, case["sql"])
    if case["optimized"]:
        plan = h.FilterPushDownOptimizer.create().optimize(plan)
    if case["pipeline"]:
        assert str(h.PipelineBuilder.create(ctx, plan).build()) == case["expect"]
    else:
        assert str(plan) == case["expect"]


def test_plan_filter_golden(ctx):
    """plan_filter_test.rs:18-21 builds the plan programmatically: project(field number) over filter(number = 1)."""
    case = [c for c in STRINGS if "plan_filter_test" in c["source"]][0]
    src = h.Planner().build_from_sql(ctx, "select number from system.numbers_mt").children_to_plans()[0]
    E = h.ExpressionPlan
    plan = (h.PlanBuilder.from_plan(src).filter(E.BinaryExpression(E.Field("number"), "=", E.Constant(h.DataValue(DT.Int64, 1))))
            .project([E.Field("number")]).build())
    assert str(plan) == case["expect"]


README_SQL = "select (number+1) as c1, number/2 as c2 from system.numbers_mt(10000000) where (c1+c2+1) < 100 limit 3"


def test_readme_explain(ctx):
    """README.md:100-117."""
    plan = h.Optimizer.create().optimize(h.Planner().build_from_sql(ctx, "explain " + README_SQL))
    assert plan.name() == "ExplainPlan"
    assert str(plan) == ("└─ Limit: 3\n  └─ Projection: (number + 1) as c1, (number / 2) as c2\n"
                         "    └─ Filter: ((((number + 1) + (number / 2)) + 1) < 100)\n"
                         "      └─ ReadDataSource: scan parts [8](Read from system.numbers_mt table)")
    pipeline = h.PipelineBuilder.create(ctx, plan.input).build()
    assert str(pipeline) == ("\n  └─ LimitTransform × 1 processor\n    └─ Merge (LimitTransform × 8 processors) to (MergeProcessor × 1)\n"
                             "      └─ LimitTransform × 8 processors\n        └─ ProjectionTransform × 8 processors\n"
                             "          └─ FilterTransform × 8 processors\n            └─ SourceTransform × 8 processors")


def test_fused_pipeline_shape():
    c = h.FuseQueryContext.create_ctx(1)   # one worker: all 8 partitions in one source pipe (pipeline_builder.rs:75-84)
    plan = h.Optimizer.create().optimize(h.Planner().build_from_sql(c, README_SQL))
    assert str(h.PipelineBuilder.create(c, plan).build()) == "\n  └─ GpuPipeTransform × 1 processor"
    agg = h.Planner().build_from_sql(c, "select sum(number)/count(number), max(number), min(number) from system.numbers_mt(10000000000)")
    assert str(h.PipelineBuilder.create(c, agg).build()) == ("\n  └─ AggregateFinalTransform × 1 processor\n    └─ GpuPipeTransform × 1 processor")
    c.worker_threads = 0
    assert str(h.PipelineBuilder.create(c, agg).build()) == (
        "\n  └─ AggregateFinalTransform × 1 processor\n    └─ Merge (GpuPipeTransform × 8 processors) to (MergeProcessor × 1)\n"
        "      └─ GpuPipeTransform × 8 processors")


def test_generate_parts_and_read_plan(ctx):
    names = [p.name for p in h.NumbersTable.generate_parts(10**10)]
    assert names[0] == "10000000000-0-1249999999" and names[7] == "10000000000-8750000000-9999999999"
    assert [p.name for p in h.NumbersTable.generate_parts(5)] == ["5-0-4"]
    assert [p.name for p in h.NumbersTable.generate_parts(19)][-1] == "19-14-18"   # last part takes total % 8
    plan = h.Planner().build_from_sql(ctx, "select number from system.numbers_mt")   # default 10 000 rows (numbers_table.rs:69)
    assert plan.children_to_plans()[0].partitions[0].name == "10000-0-1249"


def test_literal_typing_and_column_names(ctx):
    """plan_parser.rs:223-238: non-negative integer -> UInt64, non-integer -> Float64; names are Debug of the function."""
    plan = h.Planner().build_from_sql(ctx, "select sum(number)/count(number), max(number+1), number*2.5 x from system.numbers_mt(8)".replace(", number*2.5 x", ""))
    assert plan.schema().names() == ["Sum(number) / Count(number)", "Max(number + 1)"]
    assert plan.schema().types() == [DT.UInt64, DT.UInt64]
    p2 = h.Planner().build_from_sql(ctx, "select number*2.5 as x, number+1 from system.numbers_mt(8)")
    assert p2.schema().names() == ["x", "number + 1"] and p2.schema().types() == [DT.Float64, DT.UInt64]


@pytest.mark.parametrize("sql,err", [
    ("select number from system.nope", "Internal Error: Cannot find the table: nope"),
    ("select number from nodb.numbers_mt", "Internal Error: Cannot find the database: nodb"),
    ("select avg(number) from system.numbers_mt", "Internal Error: Unsupported Function: avg"),
    ("select number from system.numbers_mt limit number", "Error during plan: Unexpected expression for LIMIT clause"),
    ("select sum(number), number from system.numbers_mt", "Error during plan: Projection references non-aggregate values"),
    ("select number from system.numbers_mt having number > 1", "Internal Error: HAVING is not implemented yet"),
    ("select number from system.numbers_mt where number > -1", "Error during plan: Unsupported ExpressionPlan: - 1"),
    # UnaryOp { Not } parses in sqlparser 0.6 and falls into sql_to_rex's catch-all (plan_parser.rs:262-265), printed as SQL
    ("select number from system.numbers_mt where not number > 3", "Error during plan: Unsupported ExpressionPlan: NOT number > 3"),
    ("select number from system.numbers_mt where number > 1 and not max(number) = 'x'",
     "Error during plan: Unsupported ExpressionPlan: NOT max(number) = 'x'"),
])
def test_planner_errors(ctx, sql, err):
    with pytest.raises(h.FuseQueryError) as e:
        h.Planner().build_from_sql(ctx, sql)
    assert str(e.value) == err


def test_where_aggregate_rejected(ctx):
    plan = h.Planner().build_from_sql(ctx, "select number from system.numbers_mt where sum(number) > 1")
    with pytest.raises(h.FuseQueryError) as e:
        h.PipelineBuilder.create(ctx, plan).build()
    assert str(e.value) == "Internal Error: Aggregate function (sum([number]) > 1) is found in WHERE in query"


# ---- scalar state plumbing against the reference's vectors ----
def to_value(spec):
    return h.DataValue(TAG[spec["value"]], spec["v"])


@pytest.mark.parametrize("case", [c for c in golden("ref_datavalues.json") if c["kind"].startswith("value_")],
                         ids=lambda c: f"{c['source'].split('/')[-1]}-{c['name']}")
def test_data_value_ops(case):
    fn = h.data_value_aggregate_op if case["kind"] == "value_aggregate" else h.data_value_arithmetic_op
    try:
        got = fn(OPS[case["op"]], to_value(case["left"]), to_value(case["right"]))
    except h.FuseQueryError as e:
        assert case["error"] and str(e) == case["error"]
        return
    assert got == to_value(case["expect"])


def test_partial_state_wire_format():
    """transform_aggregate_partial.rs:61-66: serde_json of DataValue::Struct."""
    st = h.DataValue(DT.Struct, [h.DataValue(DT.UInt64, 4999999950000000), h.DataValue(DT.UInt64, 1250000000)])
    js = st.to_json()
    assert js == '{"Struct":[{"UInt64":4999999950000000},{"UInt64":1250000000}]}'
    assert h.DataValue.from_json(js) == st
    assert h.DataValue.Null().to_json() == '"Null"' and h.DataValue(DT.UInt64).to_json() == '{"UInt64":null}'
    assert h.DataValue.from_json('{"Struct":["Null",{"Int64":-3},{"Float64":2.5}]}').value[1] == h.DataValue(DT.Int64, -3)
    # byte-for-byte the oracle's format too
    from oracle import binding as o
    assert o.value_to_json(o.Value(o.STRUCT, (o.Value(o.U64, 4999999950000000), o.Value(o.U64, 1250000000)))) == js


def test_merge_protocol_without_device():
    """AggregateFinalTransform's merge is host logic (function_aggregator.rs:106-143): sum/count with depth indexing."""
    E = h.ExpressionPlan
    f = E.BinaryExpression(E.Function("sum", [E.Field("number")]), "/", E.Function("count", [E.Field("number")])).to_function()
    assert str(f) == "Sum(number) / Count(number)"
    for js in ('{"Struct":[{"UInt64":10},{"UInt64":4}]}', '{"Struct":[{"UInt64":60},{"UInt64":24}]}'):
        f.merge_state(h.DataValue.from_json(js).value)
    assert f.merge_result() == h.DataValue(DT.UInt64, 2)   # 70 / 28, integer division (function_aggregator_test.rs:118-140)


def test_derived_table_plans_flatten_like_the_reference(ctx):
    """plan_parser.rs:206-208 + plan_node.rs:60-120: a subquery in FROM is planned on its own and its SelectPlan is skipped
    when the plan is flattened, so the pipeline is one chain."""
    plan = h.Planner().build_from_sql(ctx, "select sum(x) from (select number * 2 as x from system.numbers_mt(1000) where number < 10) t")
    assert [p.name() for p in plan.children_to_plans()] == ["ReadSourcePlan", "FilterPlan", "ProjectionPlan", "AggregatePlan"]
    pipeline = h.PipelineBuilder.create(ctx, h.Optimizer.create().optimize(plan)).build()
    text = pipeline.display() if hasattr(pipeline, "display") else str(pipeline)
    # this module's context builds the reference-shaped pipeline: one processor per plan node, in plan order
    order = [line.strip().lstrip("└─ ") for line in text.splitlines() if line.strip()]
    assert order == ["AggregateFinalTransform × 1 processor",
                     "Merge (AggregatePartialTransform × 8 processors) to (MergeProcessor × 1)",
                     "AggregatePartialTransform × 8 processors", "ProjectionTransform × 8 processors",
                     "FilterTransform × 8 processors", "SourceTransform × 8 processors"]
    with pytest.raises(h.FuseQueryError) as e:
        h.Planner().build_from_sql(ctx, "select number from (select number from system.numbers_mt(10)) a, system.numbers_mt(5)")
    assert str(e.value) == "Internal Error: Cannot support JOIN clause"


def test_planner_survives_token_soup(ctx):
    """Robustness of the hand-written SQL front end: random token sequences only ever produce a plan or a FuseQueryError
    (SQLParser / Plan / Internal, error.rs:10-20) — never another exception type, never a crash."""
    import random
    toks = ["select", "from", "where", "limit", "group", "by", "as", "and", "or", "not", "explain", "system", ".", "numbers_mt", "number",
            "(", ")", ",", "*", "+", "-", "/", "%", "=", "<", ">", "<=", ">=", "<>", "!=", "1", "0", "10000", "1.5", "'a'", "sum", "count",
            "max", "min", "avg", "x", "t", ";", "having", "order", "asc", "desc", "join", "union", "-1", "1e10", "99999999999999999999", '"q"', "`b`"]
    rng = random.Random(20201)
    planned = 0
    for _ in range(4000):
        k = rng.randint(1, 14)
        if rng.random() < 0.5:
            arg = rng.choice(["10", "0", "", "1,2", "number", "'a'", "-1", "1.5"])
            q = "select " + " ".join(rng.choice(toks) for _ in range(k)) + f" from system.numbers_mt({arg}) " + \
                " ".join(rng.choice(toks) for _ in range(rng.randint(0, 6)))
        else:
            q = " ".join(rng.choice(toks) for _ in range(k))
        try:
            plan = h.Optimizer.create().optimize(h.Planner().build_from_sql(ctx, q))
            h.PipelineBuilder.create(ctx, plan).build()
            planned += 1
        except h.FuseQueryError as e:
            assert str(e).split(":")[0] in ("Internal Error", "SQLParser Error", "Error during plan"), (q, str(e))
    assert planned > 0


def test_arrow_columns_hand_over_their_own_null_bitmap():
    """tables._arrow_column: values with NULL slots zero-filled, validity as arrow's LSB-first bitmap + bit offset
    (expanded on the device by fq_column_upload_bits); no validity at all for a column without NULLs."""
    import numpy as np
    import pyarrow as pa
    from fuse_query_b200.tables import _arrow_column
    vals, valid = _arrow_column(pa.array([1, 2, 3], type=pa.int32()))
    assert valid is None and vals.dtype == np.int32 and vals.tolist() == [1, 2, 3]
    arr = pa.array([5, None, 7, None, None, 11, 13, None, 17, 19], type=pa.uint64()).slice(3, 6)   # None None 11 13 None 17
    vals, valid = _arrow_column(arr)
    assert vals.dtype == np.uint64 and vals.tolist() == [0, 0, 11, 13, 0, 17]
    kind, buf, off = valid
    assert kind == "bitmap" and off == arr.offset
    bits = np.unpackbits(np.frombuffer(buf, dtype=np.uint8), bitorder="little")[off:off + len(arr)]
    assert bits.tolist() == [0, 0, 1, 1, 0, 1]
    chunked = pa.chunked_array([pa.array([1.5, None]), pa.array([None, 4.0])])
    vals, valid = _arrow_column(chunked)
    assert vals.tolist() == [1.5, 0.0, 0.0, 4.0] and valid[0] == "bitmap"


def test_partial_state_json_parser_survives_mutations():
    """The merge point parses what the partial pipes serialise (transform_aggregate_final.rs:56-66 -> serde_json of
    DataValue::Struct): mutated JSON either parses to a value that round-trips or raises FuseQueryError."""
    import random
    rng = random.Random(3)
    seeds = ['{"Struct":[{"UInt64":4999},{"UInt64":null}]}', '"Null"', '{"Int64":-5}', '{"Float64":1.5e3}', '{"Utf8":"a\\\\"b"}',
             '{"Boolean":true}', '{"Struct":[{"Struct":[{"UInt8":1}]},"Null"]}', '{"UInt64":18446744073709551615}', '{"Float32":null}']
    alphabet = list('{}[]":,\\\\ntruefalsnul0123456789.-+eE ') + ['"UInt64"', '"Struct"', '"Null"', 'null', '{', '}', '[', ']']
    parsed = 0
    for _ in range(30000):
        s = rng.choice(seeds)
        for _ in range(rng.randint(0, 4)):
            k, pos = rng.randint(0, 3), rng.randint(0, len(s))
            if k == 0:
                s = s[:pos] + rng.choice(alphabet) + s[pos:]
            elif k == 1:
                s = s[:pos] + s[pos + 1:]
            elif k == 2:
                s = s[:pos] + s[pos:][::-1][:rng.randint(0, 5)] + s[pos:]
            else:
                s = s[:pos]
        try:
            v = h.DataValue.from_json(s)
        except h.FuseQueryError:
            continue
        assert h.DataValue.from_json(v.to_json()).to_json() == v.to_json(), s
        parsed += 1
    assert parsed > 1000


def test_distributed_execution_refuses_plan_shapes_it_cannot_split():
    """distributed.execute_sql_distributed runs ReadSource [Filter] (Projection | Aggregate) [Limit] per rank and merges;
    a derived table would need its outer stages after the merge and must be refused, not silently truncated."""
    from fuse_query_b200.distributed import execute_sql_distributed
    c = h.FuseQueryContext.create_ctx(8)
    gather = lambda obj: [obj]
    with pytest.raises(h.FuseQueryError) as e:
        execute_sql_distributed(c, "select number + 1 from (select number from system.numbers_mt(100) where number > 2) where number < 8",
                                0, 1, gather)
    assert "distributed execution supports" in str(e.value) and "FilterPlan -> ProjectionPlan -> FilterPlan -> ProjectionPlan" in str(e.value)
    # a plain select passes the shape check and only then misses the device this context does not have
    for sql in ("select sum(number) from system.numbers_mt(1000)", "select number from system.numbers_mt(1000) where number > 3 limit 2"):
        with pytest.raises(h.FuseQueryError) as e:
            execute_sql_distributed(c, sql, 0, 1, gather)
        assert "distributed execution supports" not in str(e.value)


def test_deeply_nested_sql_is_refused_not_crashed(ctx):
    """A 40 KB query of nested parentheses, a 20 000-term sum or 5 000 nested derived tables must end in a
    FuseQueryError — the recursive-descent parser counts its nesting and ExpressionPlan bounds its height at 128
    (the same bound the code generator applies), instead of running the host stack out (exit 139)."""
    deep = "(" * 20000 + "number" + ")" * 20000
    for sql in (f"select {deep} from system.numbers_mt(10)",
                "select " + "+".join(["number"] * 20000) + " from system.numbers_mt(10)",
                "select " + "sum(" * 3000 + "number" + ")" * 3000 + " from system.numbers_mt(10)",
                "select number from " + "(select number from " * 5000 + "system.numbers_mt(10)" + ")" * 5000,
                "select " + "not " * 20000 + "number from system.numbers_mt(10)",
                "select " + "- " * 20000 + "number from system.numbers_mt(10)"):
        with pytest.raises(h.FuseQueryError) as e:
            h.Planner().build_from_sql(ctx, sql)
        assert "depth more than 128" in str(e.value), str(e.value)[:200]
    # 100 levels are fine
    ok = "(" * 100 + "number" + ")" * 100
    h.Planner().build_from_sql(ctx, f"select {ok} from system.numbers_mt(10)")
    h.Planner().build_from_sql(ctx, "select " + "+".join(["number"] * 100) + " from system.numbers_mt(10)")


def test_more_than_eight_select_items_plan_without_a_device(ctx):
    """fq_pipe_desc holds 8 roots; a 9+ item select list must never overrun it (stack-buffer-overflow in
    Lowering::desc).  Without a device the query plans and builds its pipeline; executing throws a FuseQueryError."""
    sql = "select " + ", ".join(f"number + {i} as c{i}" for i in range(12)) + " from system.numbers_mt(100)"
    plan = h.Optimizer.create().optimize(h.Planner().build_from_sql(ctx, sql))
    assert len(plan.children_to_plans()[1].schema().names()) == 12
    with pytest.raises(h.FuseQueryError):
        h.execute_sql(ctx, sql)


def test_order_by_plans_a_sort_over_the_output_columns(ctx):
    """The reference accepts ORDER BY and drops it (sqlparser parses it, plan_parser.rs never reads query.order_by; README.md:28
    lists sorting as open).  Here it becomes a SortPlan between the projection / aggregate and the LIMIT, its keys resolved
    against the query's OUTPUT columns; EXPLAIN shows it, the optimizer keeps it."""
    sql = "explain select number + 1 as c1, number from system.numbers_mt(100) where c1 > 3 order by number desc, c1 + 1 limit 5"
    text = h.execute_sql(ctx, sql)[0].column(0).to_list()[0]
    assert text.splitlines()[:3] == ["└─ Limit: 5", "  └─ Sort: number desc, (c1 + 1)", "    └─ Projection: (number + 1) as c1, number"]
    plan = h.Optimizer.create().optimize(h.Planner().build_from_sql(ctx, sql.replace("explain ", "")))
    assert [p.name() for p in plan.children_to_plans()] == ["ReadSourcePlan", "FilterPlan", "ProjectionPlan", "SortPlan", "LimitPlan"]
    pipe = h.PipelineBuilder.create(ctx, plan).build()
    assert "GpuSortTransform × 1 processor" in str(pipe) and "LimitTransform × 1 processor" in str(pipe)
    for bad, err in [("select number from system.numbers_mt(10) order by nope", "nope"),
                     ("select sum(number) from system.numbers_mt(10) order by sum(number)", "ORDER BY sorts the query's output columns"),
                     ("select number from system.numbers_mt(10) order number", "BY")]:
        with pytest.raises(h.FuseQueryError) as e:
            h.Planner().build_from_sql(ctx, bad)
        assert err in str(e.value), str(e.value)
