"""The C++ host mirror (Function / DataBlock / IProcessor / Pipeline / executors) driving the CUDA path,
written the way the reference's own tests are (src/functions/*_test.rs, src/transforms/*_test.rs,
src/executors/*_test.rs) and checked against the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest

from conftest import golden
from fuse_query_b200 import _fuse_host as h
from oracle import binding as o

pytestmark = pytest.mark.gpu
DT = h.DataType
E = h.ExpressionPlan
TAG = {"Null": 0, "Boolean": 1, "Int8": 2, "Int16": 3, "Int32": 4, "Int64": 5, "UInt8": 6, "UInt16": 7, "UInt32": 8, "UInt64": 9,
       "Float32": 10, "Float64": 11, "Utf8": 12}
NP = {"Boolean": np.bool_, "Int8": np.int8, "Int16": np.int16, "Int32": np.int32, "Int64": np.int64, "UInt8": np.uint8,
      "UInt16": np.uint16, "UInt32": np.uint32, "UInt64": np.uint64, "Float32": np.float32, "Float64": np.float64}
OPS = {"Add": h.ops.Add, "Sub": h.ops.Sub, "Mul": h.ops.Mul, "Div": h.ops.Div, "Eq": h.ops.Eq, "Lt": h.ops.Lt, "LtEq": h.ops.LtEq,
       "Gt": h.ops.Gt, "GtEq": h.ops.GtEq, "And": h.ops.And, "Or": h.ops.Or, "Min": h.ops.Min, "Max": h.ops.Max, "Sum": h.ops.Sum,
       "Count": h.ops.Count}


@pytest.fixture(scope="module")
def gpu():
    return h.GpuContext.create(0)


def make_ctx(gpu, workers=0, **opts):
    c = h.FuseQueryContext.create_ctx(workers, gpu)
    for k, v in opts.items():
        setattr(c.options, k, v)
    return c


def to_array(gpu, spec):
    return h.DataArray.from_numpy(gpu, np.asarray(spec["values"], dtype=NP[spec["array"]]))


def to_value(spec):
    return h.DataValue(TAG[spec["value"]], spec["v"])


def operand(gpu, spec):
    return h.DataColumnarValue.Array(to_array(gpu, spec)) if "array" in spec else h.DataColumnarValue.Scalar(to_value(spec))


def ident(c):
    return f"{c['source'].split('/')[-1]}-{c['name']}"


def has_utf8(case):
    return any(isinstance(case.get(k), dict) and (case[k].get("array") == "Utf8" or case[k].get("value") == "Utf8")
               for k in ("left", "right", "array"))


# ---------------------------------------------------------------------------------------------
# datavalues: the reference's table-driven vectors for array (op) array / scalar (data_array_*_test.rs)
# ---------------------------------------------------------------------------------------------
DV = golden("ref_datavalues.json")


@pytest.mark.parametrize("case", [c for c in DV if c["kind"] in ("array_arithmetic", "array_comparison", "array_logic")], ids=ident)
def test_data_array_ops(gpu, case):
    fn = {"array_arithmetic": h.data_array_arithmetic_op, "array_comparison": h.data_array_comparison_op,
          "array_logic": h.data_array_logic_op}[case["kind"]]
    if has_utf8(case) and not case["error"]:
        # the *_utf8 comparison kernels (datavalues/macros.rs:29-36): strings go to the device as Arrow offsets + bytes
        def utf8_operand(spec):
            return h.DataColumnarValue.Array(h.DataArray.utf8(spec["values"])) if "array" in spec else h.DataColumnarValue.Scalar(h.DataValue(TAG["Utf8"], spec["v"]))
        got = fn(gpu, OPS[case["op"]], utf8_operand(case["left"]), utf8_operand(case["right"]))
        assert h.data_type_name(got.data_type()) == "Boolean" and got.to_list() == case["expect"]["values"]
        return
    if has_utf8(case):
        # the reference's error text is still reproduced by the typing rules before anything reaches the device
        with pytest.raises(h.FuseQueryError) as e:
            h.numerical_coercion({"Add": "+", "Sub": "-", "Mul": "*", "Div": "/"}[case["op"]], TAG[case["left"]["array"]], TAG[case["right"]["array"]])
        assert str(e.value) == case["error"]
        return
    got = fn(gpu, OPS[case["op"]], operand(gpu, case["left"]), operand(gpu, case["right"]))
    exp = case["expect"]
    assert h.data_type_name(got.data_type()) == exp["array"]
    assert got.to_list() == np.asarray(exp["values"], dtype=NP[exp["array"]]).tolist()


@pytest.mark.parametrize("case", [c for c in DV if c["kind"] == "array_aggregate"], ids=ident)
def test_data_array_aggregate(gpu, case):
    if has_utf8(case):   # min_string / max_string on the device (data_array_aggregate.rs:139-154); Sum keeps the reference's error
        arr = h.DataArray.utf8(case["array"]["values"])
        if case["error"]:
            with pytest.raises(h.FuseQueryError) as e:
                h.data_array_aggregate_op(gpu, OPS[case["op"]], arr)
            assert str(e.value) == case["error"]
        else:
            assert h.data_array_aggregate_op(gpu, OPS[case["op"]], arr) == h.DataValue(TAG["Utf8"], case["expect"]["v"])
        return
    got = h.data_array_aggregate_op(gpu, OPS[case["op"]], to_array(gpu, case["array"]))
    assert got == to_value(case["expect"])


# ---------------------------------------------------------------------------------------------
# functions: function_{arithmetic,comparison,logic,aggregator}_test.rs
# ---------------------------------------------------------------------------------------------
FN = golden("ref_functions.json")


def fn_from_sexpr(s):
    """s-expression (as recorded by the extractor) -> Function via the reference's try_create constructors."""
    toks = s.replace("(", " ( ").replace(")", " ) ").split()

    def parse(i):
        assert toks[i] == "("
        head = toks[i + 1]
        i += 2
        if head == "col":
            return h.FieldFunction.try_create(toks[i]), i + 2
        ty = {"i8": DT.Int8, "i16": DT.Int16, "i32": DT.Int32, "i64": DT.Int64, "u8": DT.UInt8, "u16": DT.UInt16, "u32": DT.UInt32,
              "u64": DT.UInt64, "f32": DT.Float32, "f64": DT.Float64}
        if head in ty:
            v = float(toks[i]) if head[0] == "f" else int(toks[i])
            return h.ConstantFunction.try_create(h.DataValue(ty[head], v)), i + 2
        args = []
        while toks[i] != ")":
            a, i = parse(i)
            args.append(a)
        return h.ScalarFunctionFactory.get(head, args), i + 1

    return parse(0)[0]


def block_of(gpu, spec):
    cols = [to_array(gpu, c) for c in spec["columns"]]
    schema = h.DataSchema([h.DataField(n, TAG[c["array"]], False) for n, c in zip(spec["names"], spec["columns"])])
    return h.DataBlock.create(schema, cols)


@pytest.mark.parametrize("case", [c for c in FN if c["kind"] == "function_eval"], ids=ident)
def test_function_eval(gpu, case):
    func = fn_from_sexpr(case["sexpr"])
    block = block_of(gpu, case["block"])
    assert str(func) == case["display"]                       # Display check
    assert func.nullable(block.schema()) == case["nullable"]  # Nullable check
    v = func.eval(gpu, block)
    assert func.return_type(block.schema()) == v.data_type()  # Type check
    assert h.data_type_name(v.data_type()) == case["expect"]["array"]
    assert v.to_array(gpu, 0).to_list() == case["expect"]["values"]


@pytest.mark.parametrize("case", [c for c in FN if c["kind"] == "function_aggregate"], ids=ident)
def test_aggregator_function(gpu, case):
    """function_aggregator_test.rs:168-188: accumulate x evals, accumulate x (evals-1), merge both states."""
    proto = fn_from_sexpr(case["sexpr"])
    block = block_of(gpu, case["block"])
    func1 = proto.clone()
    for _ in range(case["evals"]):
        func1.accumulate(gpu, block)
    state1 = func1.accumulate_result()
    func2 = proto.clone()
    for _ in range(1, case["evals"]):
        func2.accumulate(gpu, block)
    state2 = func2.accumulate_result()
    final_func = proto.clone()
    final_func.set_depth(0)
    final_func.merge_state(state1)
    final_func.merge_state(state2)
    assert final_func.merge_result() == to_value(case["expect"])


def test_function_errors_carry_reference_text(gpu):
    block = block_of(gpu, {"names": ["a"], "columns": [{"array": "UInt64", "values": [4, 0, 2]}]})
    div = h.ArithmeticFunction.try_create(h.ops.Div, [h.ConstantFunction.try_create(h.DataValue(DT.UInt64, 8)), h.FieldFunction.try_create("a")])
    with pytest.raises(h.FuseQueryError) as e:
        div.eval(gpu, block)
    assert str(e.value) == "Internal Error: Divide by zero error"
    with pytest.raises(h.FuseQueryError) as e:
        h.FieldFunction.try_create("a").accumulate_result()
    assert str(e.value) == "Internal Error: Unsupported aggregate operation for function field"
    with pytest.raises(h.FuseQueryError) as e:
        h.ScalarFunctionFactory.get("avg", [h.FieldFunction.try_create("a")])
    assert str(e.value) == "Internal Error: Unsupported Function: avg"
    with pytest.raises(h.FuseQueryError) as e:
        h.FieldFunction.try_create("zz").eval(gpu, block)
    assert "Unable to get field named \"zz\"" in str(e.value)


# ---------------------------------------------------------------------------------------------
# transforms / pipeline: transform_*_test.rs, processor_merge_test.rs (testdata/number.rs fixtures)
# ---------------------------------------------------------------------------------------------
def number_source_transform_for_test(ctx, numbers):
    """testdata/number.rs:53-70."""
    table = ctx.get_table("system", "numbers_mt")
    plan = h.Planner().build_from_sql(ctx, f"select number from system.numbers_mt({numbers})").children_to_plans()[0]
    return h.SourceTransform.try_create(ctx, "system", "numbers_mt", plan.partitions), table.schema()


def rows_of(blocks):
    out = []
    for b in blocks:
        cols = [b.column(i).to_list() for i in range(b.num_columns())]
        out += list(zip(*cols)) if cols else []
    return out


@pytest.mark.parametrize("fuse_blocks", [0, 10000], ids=["whole-partition-blocks", "reference-10000-row-blocks"])
def test_transform_aggregate(gpu, fuse_blocks):
    """transform_aggregate_test.rs:19-57: sum(number)+2 over numbers_mt(16) through Partial -> Merge -> Final = 122."""
    ctx = make_ctx(gpu, 0, fuse=False, block_rows=fuse_blocks)
    pipeline = h.Pipeline.create()
    a, schema = number_source_transform_for_test(ctx, 16)
    pipeline.add_source(a)
    plan = h.PlanBuilder.create(schema).aggregate([], [E.BinaryExpression(E.Function("sum", [E.Field("number")]), "+", E.Constant(h.DataValue(DT.UInt64, 2)))]).build()
    pipeline.add_simple_transform(lambda: h.AggregatePartialTransform.try_create(ctx, plan.schema(), plan.expr))
    pipeline.merge_processor()
    pipeline.add_simple_transform(lambda: h.AggregateFinalTransform.try_create(ctx, plan.schema(), plan.expr))
    blocks = pipeline.execute().collect()
    assert [b.column(0).to_list() for b in blocks if b.num_rows() > 0] == [[122]]
    assert blocks[0].column(0).data_type() == DT.UInt64


def test_transform_filter(gpu):
    """transform_filter_test.rs:19-42: number = 1 over numbers_mt(8) -> [1]."""
    ctx = make_ctx(gpu, 0, fuse=False)
    pipeline = h.Pipeline.create()
    a, schema = number_source_transform_for_test(ctx, 8)
    pipeline.add_source(a)
    pred = E.BinaryExpression(E.Field("number"), "=", E.Constant(h.DataValue(DT.Int64, 1)))   # constant(1): i64 literal in the Rust test
    pipeline.add_simple_transform(lambda: h.FilterTransform.try_create(ctx, pred))
    pipeline.merge_processor()
    blocks = pipeline.execute().collect()
    assert [b.column(0).to_list() for b in blocks if b.num_rows() > 0] == [[1]]


def test_transform_limit_and_source(gpu):
    """transform_limit_test.rs: limit 2 -> 2 rows; transform_source_test.rs: two numbers_mt(8) sources -> 16 rows;
    processor_merge_test.rs:17-26: first block of numbers_mt(16) is [0, 1]."""
    ctx = make_ctx(gpu, 0, fuse=False)
    pipeline = h.Pipeline.create()
    pipeline.add_source(number_source_transform_for_test(ctx, 8)[0])
    pipeline.merge_processor()
    pipeline.add_simple_transform(lambda: h.LimitTransform.try_create(2))
    assert sum(b.num_rows() for b in pipeline.execute().collect()) == 2
    p2 = h.Pipeline.create()
    p2.add_source(number_source_transform_for_test(ctx, 8)[0])
    p2.add_source(number_source_transform_for_test(ctx, 8)[0])
    p2.merge_processor()
    assert sum(b.num_rows() for b in p2.execute().collect()) == 16
    ctx16 = make_ctx(gpu, 0, fuse=False, block_rows=10000)   # reference block split: one block per 2-row partition
    plan = h.Planner().build_from_sql(ctx16, "select number from system.numbers_mt(16)").children_to_plans()[0]
    p3 = h.Pipeline.create()
    for part in plan.partitions:
        p3.add_source(h.SourceTransform.try_create(ctx16, "system", "numbers_mt", [part]))
    p3.merge_processor()
    assert p3.execute().next().column(0).to_list() == [0, 1]


# ---------------------------------------------------------------------------------------------
# executors: SQL in, blocks out — fused and reference-shaped pipelines against the oracle
# ---------------------------------------------------------------------------------------------
NUM = "(col number)"
QUERIES = [
    # (sql, oracle exprs, predicate, is_aggregate, limit)
    ("select sum(number) from system.numbers_mt({n})", [f"(sum {NUM})"], None, True, None),
    ("select max(number+1), min(number), count(number) from system.numbers_mt({n})",
     [f"(max (+ {NUM} (u64 1)))", f"(min {NUM})", f"(count {NUM})"], None, True, None),
    ("select sum(number)/count(number), max(number), min(number) from system.numbers_mt({n})",
     [f"(/ (sum {NUM}) (count {NUM}))", f"(max {NUM})", f"(min {NUM})"], None, True, None),
    ("select sum(number+1)+2 as sumx from system.numbers_mt({n})", [f"(alias sumx (+ (sum (+ {NUM} (u64 1))) (u64 2)))"], None, True, None),
    ("select (number+1) as c1, number/2 as c2 from system.numbers_mt({n}) where (c1+c2+1) < 100 limit 3",
     [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"],
     f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))", False, 3),
    ("select number*3 as t, number from system.numbers_mt({n}) where number*3 >= 30 and number < 1000 limit 50",
     [f"(alias t (* {NUM} (u64 3)))", NUM], f"(and (>= (* {NUM} (u64 3)) (u64 30)) (< {NUM} (u64 1000)))", False, 50),
    ("select min(number), max(number), count(number) from system.numbers_mt({n}) where number/7*7 = number",
     [f"(min {NUM})", f"(max {NUM})", f"(count {NUM})"], f"(= (* (/ {NUM} (u64 7)) (u64 7)) {NUM})", True, None),
]


@pytest.mark.parametrize("mode", ["fused", "fused-generated", "reference-shaped"])
@pytest.mark.parametrize("n", [16, 80000, 10_000_000])
@pytest.mark.parametrize("q", QUERIES, ids=[q[0][:48] for q in QUERIES])
def test_select_executor_matches_oracle(gpu, q, n, mode):
    sql, exprs, pred, is_agg, limit = q
    if mode == "reference-shaped" and n > 80000:
        pytest.skip("one kernel per 10 000-row block per node: covered at the smaller sizes")
    workers = {"fused": 1, "fused-generated": 0, "reference-shaped": 0}[mode]
    ctx = make_ctx(gpu, workers, fuse=mode != "reference-shaped", generated=mode == "fused-generated",
                   block_rows=10000 if mode == "reference-shaped" else 0)
    blocks = h.execute_sql(ctx, sql.format(n=n))
    want = o.run_query(exprs, total=n, predicate=pred, is_aggregate=is_agg, limit=limit, worker_threads=workers, use_threads=True)
    assert rows_of(blocks) == want.rows()
    assert blocks[0].schema().names() == want.names
    assert [h.data_type_name(t) for t in blocks[0].schema().types()] == [o.DTYPE_NAMES[c.dtype] for c in want.columns]


@pytest.mark.parametrize("n", [80008, 1_000_000, 12345, 10001])
def test_numbers_stream_tail_quirk_is_reproduced(gpu, n):
    """SURVEY F7 (numbers_stream.rs:44-46): numbers_mt(80008) emits 16 rows per... the reference drops rows when a
    partition is >= 10 000 rows and not a multiple of 10 000.  The source mirror reproduces it by default."""
    sql = f"select count(number), sum(number), max(number) from system.numbers_mt({n})"
    exprs = [f"(count {NUM})", f"(sum {NUM})", f"(max {NUM})"]
    want = o.run_query(exprs, total=n, is_aggregate=True, worker_threads=0)
    for opts in (dict(fuse=True), dict(fuse=True, generated=True), dict(fuse=False, block_rows=10000)):
        assert rows_of(h.execute_sql(make_ctx(gpu, 0, **opts), sql)) == want.rows()
    if n == 80008:
        assert want.rows()[0][0] == 16   # 8 partitions x 2 rows survive
    fixed = o.run_query(exprs, total=n, is_aggregate=True, worker_threads=0, tail_quirk=False)
    assert rows_of(h.execute_sql(make_ctx(gpu, 1, tail_quirk=False), sql)) == fixed.rows() == [(n, n * (n - 1) // 2, n - 1)]


def test_explain_executor(gpu):
    ctx = make_ctx(gpu, 0, fuse=False)
    blocks = h.execute_sql(ctx, "explain select (number+1) as c1, number/2 as c2 from system.numbers_mt(10000000) where (c1+c2+1) < 100 limit 3")
    text = blocks[0].column(0).to_list()
    assert blocks[0].schema().names() == ["explain"] and len(text) == 2
    assert text[0].startswith("└─ Limit: 3\n  └─ Projection: (number + 1) as c1, (number / 2) as c2\n    └─ Filter: ((((number + 1) + (number / 2)) + 1) < 100)")
    assert "SourceTransform × 8 processors" in text[1]


def test_readme_headline_query_10b_through_sql(gpu):
    """BASELINE configs[3] end to end through plan_select/executor_select, generated and materialised shards."""
    sql = "select sum(number)/count(number), max(number), min(number) from system.numbers_mt(10000000000)"
    for opts in (dict(generated=True), dict(generated=False)):
        blocks = h.execute_sql(make_ctx(gpu, 1, **opts), sql)
        assert rows_of(blocks) == [(1310651184, 9999999999, 0)]
        assert blocks[0].schema().names() == ["Sum(number) / Count(number)", "Max(number)", "Min(number)"]
    h.numbers_cache_clear()


# ---------------------------------------------------------------------------------------------
# multi-process execution: two ranks (gloo plumbing, both on cuda:0) run the fused pipes of their own partitions
# ---------------------------------------------------------------------------------------------
def _dist_worker(rank, world, port, sqls, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fuse_query_b200 import _fuse_host as hh
    from fuse_query_b200.distributed import execute_sql_distributed
    gpu = hh.GpuContext.create(0)
    ctx = hh.FuseQueryContext.create_ctx(1, gpu)

    def gather(obj):
        res = [None] * world
        dist.all_gather_object(res, obj)
        return res

    # with `gpu` the rows of LIMIT queries meet in the ranks' exchange windows on the device (two processes, one GPU here: CUDA IPC)
    results = [execute_sql_distributed(ctx, s, rank, world, gather, gpu=gpu) for s in sqls]
    results += [execute_sql_distributed(ctx, s, rank, world, gather) for s in sqls[1:2]]      # the host path, for comparison
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_distributed_sql(gpu):
    import os
    import torch.multiprocessing as mp
    n = 16_000_000
    sqls = [f"select sum(number)/count(number), max(number), min(number) from system.numbers_mt({n})",
            f"select (number+1) as c1, number/2 as c2 from system.numbers_mt({n}) where (c1+c2+1) < 100 limit 3",
            f"select number from system.numbers_mt({n}) where number/1000000*1000000 = number",
            f"select number from system.numbers_mt({n}) where number/1000000*1000000 = number limit 5",
            f"select number, number - number/7*7 as m from system.numbers_mt({n}) where number/1000000*1000000 = number order by m desc, number limit 6"]
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 29800 + (os.getpid() % 150)
    procs = [mpctx.Process(target=_dist_worker, args=(r, 2, port, sqls, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    s = n * (n - 1) // 2
    assert got[0] == (["Sum(number) / Count(number)", "Max(number)", "Min(number)"], [(s // n, n - 1, 0)])
    assert got[1] == (["c1", "c2"], [(1, 0), (2, 0), (3, 1)])
    assert got[2] == (["number"], [(k * 1000000,) for k in range(16)])
    assert got[3] == (["number"], [(k * 1000000,) for k in range(5)]) and got[5] == got[1]
    # ORDER BY across ranks: every rank's rows meet in partition order, then one device sort (NULLs first, stable) and the LIMIT
    want = sorted([(k * 1000000, (k * 1000000) % 7) for k in range(16)], key=lambda r: (-r[1], r[0]))[:6]
    assert got[4] == (["number", "m"], want)


# ---------------------------------------------------------------------------------------------
# a real (non-generated) table resident in HBM: same pipeline, SQL over user columns
# ---------------------------------------------------------------------------------------------
def test_memory_table_sql_matches_oracle(gpu):
    from fuse_query_b200.tables import register_table
    rng = np.random.default_rng(99)
    n = 123_457
    data = {"k": rng.integers(0, 1 << 50, n, dtype=np.uint64), "v": rng.integers(-10**9, 10**9, n, dtype=np.int64),
            "w": rng.integers(1, 1000, n, dtype=np.uint16)}
    for workers, fuse in ((1, True), (0, True), (0, False)):
        ctx = make_ctx(gpu, workers, fuse=fuse, block_rows=0 if fuse else 10000)
        register_table(ctx, gpu, "default", "t", data)
        table = {k: o.from_numpy(v) for k, v in data.items()}
        blocks = h.execute_sql(ctx, "select sum(v), min(v / w), max(k + w), count(k) from t where v < 500000000")
        want = o.run_query(["(sum (col v))", "(min (/ (col v) (col w)))", "(max (+ (col k) (col w)))", "(count (col k))"], table=table,
                           predicate="(< (col v) (i64 500000000))" if False else "(< (col v) (u64 500000000))", is_aggregate=True,
                           worker_threads=workers, tail_quirk=False)
        assert rows_of(blocks) == want.rows()
        assert blocks[0].schema().names() == ["Sum(v)", "Min(v / w)", "Max(k + w)", "Count(k)"] == want.names
        blocks = h.execute_sql(ctx, "select k, v * w as vw from t where w = 7 limit 20")
        want = o.run_query(["(col k)", "(alias vw (* (col v) (col w)))"], table=table, predicate="(= (col w) (u64 7))", limit=20,
                           worker_threads=workers, tail_quirk=False)
        assert rows_of(blocks) == want.rows() and len(want.rows()) == 20


def test_nullable_memory_table_sql_matches_oracle(gpu):
    """NULLs in user columns (numpy masked arrays, (values, valid) pairs, pyarrow arrays) ride through the same pipes:
    arrow's null propagation for arithmetic / comparison / logic, NULL predicate slots keep nothing, Sum/Min/Max skip
    NULL slots and Count is the block length (data_array_aggregate.rs:29)."""
    import pyarrow as pa
    from fuse_query_b200.tables import register_table
    rng = np.random.default_rng(7)
    n = 96_000
    k = rng.integers(0, 1 << 50, n, dtype=np.uint64)
    v = rng.integers(-10**9, 10**9, n, dtype=np.int64)
    w = rng.integers(1, 1000, n, dtype=np.uint16)
    kv = (rng.random(n) > 0.3).astype(np.uint8)
    vv = (rng.random(n) > 0.1).astype(np.uint8)
    data = {"k": (k, kv), "v": np.ma.MaskedArray(v, mask=~vv.astype(bool)), "w": pa.array(w)}
    table = {"k": o.array(o.U64, k, kv), "v": o.array(o.I64, v, vv), "w": o.array(o.U16, w)}
    for workers, fuse in ((1, True), (0, True), (0, False)):
        ctx = make_ctx(gpu, workers, fuse=fuse, block_rows=0 if fuse else 10000)
        t = register_table(ctx, gpu, "default", "t", data)
        assert [f.nullable for f in t.schema().fields] == [True, True, False]
        blocks = h.execute_sql(ctx, "select sum(v), min(v / w), max(k + w), count(k), sum(k + v) from t where v < 500000000")
        want = o.run_query(["(sum (col v))", "(min (/ (col v) (col w)))", "(max (+ (col k) (col w)))", "(count (col k))",
                            "(sum (+ (col k) (col v)))"], table=table, predicate="(< (col v) (u64 500000000))", is_aggregate=True,
                           worker_threads=workers, tail_quirk=False)
        assert rows_of(blocks) == want.rows()
        blocks = h.execute_sql(ctx, "select k, v * w as vw, k + v from t where w = 7 limit 40")
        want = o.run_query(["(col k)", "(alias vw (* (col v) (col w)))", "(+ (col k) (col v))"], table=table,
                           predicate="(= (col w) (u64 7))", limit=40, worker_threads=workers, tail_quirk=False)
        got = rows_of(blocks)
        assert got == want.rows() and len(got) == 40
        assert any(x is None for r in got for x in r)
        blocks = h.execute_sql(ctx, "select w from t where k > 1000 and v > 0 limit 15")
        want = o.run_query(["(col w)"], table=table, predicate="(and (> (col k) (u64 1000)) (> (col v) (u64 0)))", limit=15,
                           worker_threads=workers, tail_quirk=False)
        assert rows_of(blocks) == want.rows()


def test_parquet_and_arrow_ipc_files_become_resident_tables(gpu, tmp_path):
    """a real ITable fed from column files.  Same queries, same oracle."""
    import pyarrow as pa
    import pyarrow.ipc as ipc
    import pyarrow.parquet as pq
    from fuse_query_b200.tables import register_arrow_ipc, register_parquet
    rng = np.random.default_rng(11)
    n = 50_000
    price = rng.integers(1, 10_000, n, dtype=np.int64)
    qty = rng.integers(1, 100, n, dtype=np.uint32)
    disc = np.round(rng.random(n), 3)
    dv = rng.random(n) > 0.2
    tbl = pa.table({"price": pa.array(price), "qty": pa.array(qty), "disc": pa.array(disc, mask=~dv), "unused": pa.array(price)})
    pq.write_table(tbl, tmp_path / "t.parquet", row_group_size=7_000)
    with ipc.new_file(str(tmp_path / "t.arrow"), tbl.schema) as w:
        w.write_table(tbl, max_chunksize=9_999)
    table = {"price": o.array(o.I64, price), "qty": o.array(o.U32, qty), "disc": o.array(o.F64, disc, dv.astype(np.uint8))}
    for reg, fname, tname in ((register_parquet, "t.parquet", "pq"), (register_arrow_ipc, "t.arrow", "ipc")):
        ctx = make_ctx(gpu, 1, fuse=True)
        t = reg(ctx, gpu, "default", tname, tmp_path / fname, columns=["price", "qty", "disc"])
        assert t.schema().names() == ["price", "qty", "disc"] and t.num_rows() == n
        assert [f.nullable for f in t.schema().fields] == [False, False, True]
        got = rows_of(h.execute_sql(ctx, f"select sum(price * qty), max(disc), min(price / qty), count(disc) from {tname} where qty > 10"))
        want = o.run_query(["(sum (* (col price) (col qty)))", "(max (col disc))", "(min (/ (col price) (col qty)))", "(count (col disc))"],
                           table=table, predicate="(> (col qty) (u64 10))", is_aggregate=True, worker_threads=1, tail_quirk=False)
        assert got == want.rows()
        got = rows_of(h.execute_sql(ctx, f"select price, disc * qty from {tname} where disc < 0.5 limit 25"))
        want = o.run_query(["(col price)", "(* (col disc) (col qty))"], table=table, predicate="(< (col disc) (f64 0.5))", limit=25,
                           worker_threads=1, tail_quirk=False)
        assert got == want.rows() and len(got) == 25


def test_derived_tables_chain_transforms_after_the_fused_pipe(gpu):
    """FROM (subquery): plan_parser.rs:206-208 plans the subquery and PlanNode::children_to_plans flattens the nested
    SelectPlan into one chain, so the outer Filter / Projection / Aggregate / Limit run as ordinary transforms after the
    fused device pipe of the inner query.  Expected rows by hand (the oracle has no nested plans)."""
    for workers, fuse in ((1, True), (0, True), (0, False)):
        ctx = make_ctx(gpu, workers, fuse=fuse, block_rows=0 if fuse else 10000)
        got = rows_of(h.execute_sql(ctx, "select number + 1 from (select number from system.numbers_mt(100000) where number > 2) where number < 8"))
        assert got == [(4,), (5,), (6,), (7,), (8,)]
        got = rows_of(h.execute_sql(ctx, "select max(x), count(x), min(x) from (select number * 2 as x from system.numbers_mt(100000) where number < 10) t"))
        assert got == [(18, 10, 0)]
        got = rows_of(h.execute_sql(ctx, "select sum(x) / count(x) from (select number * 2 as x from system.numbers_mt(160000))"))
        assert got == [(159999,)]     # 160000 rows: partitions of whole 10 000-row blocks (no NumbersStream tail quirk)
        # Sum over a filtered subquery meets the emptied blocks of partitions 1..7: the reference's poisoned state
        with pytest.raises(h.FuseQueryError) as e:
            h.execute_sql(ctx, "select sum(x) from (select number * 2 as x from system.numbers_mt(100000) where number < 10) t")
        assert str(e.value) == "Internal Error: DataValue to array cannot be NONE NULL"
        got = rows_of(h.execute_sql(ctx, "select number from (select number from system.numbers_mt(100000) limit 5) as t where number > 1"))
        assert got == [(2,), (3,), (4,)]
        got = rows_of(h.execute_sql(ctx, "select count(number) from (select number from (select number from system.numbers_mt(100000) where number > 9) where number < 20)"))
        assert got == [(10,)]
        # the reference's alias rewrite (optimizer_filter_push_down.rs:66-78) turns the outer `x < 8` into `number < 8`, which the
        # block below it (columns: x) cannot satisfy: the same arrow error
        with pytest.raises(h.FuseQueryError) as e:
            h.execute_sql(ctx, "select x from (select number as x from system.numbers_mt(100)) where x < 8")
        assert 'Unable to get field named "number"' in str(e.value)
    ctx = make_ctx(gpu, 8, fuse=True)
    plan = h.Planner().build_from_sql(ctx, "select number + 1 from (select number from system.numbers_mt(100) where number > 2) where number < 8")
    assert [p.name() for p in plan.children_to_plans()] == ["ReadSourcePlan", "FilterPlan", "ProjectionPlan", "FilterPlan", "ProjectionPlan"]


def test_mysql_result_set_matches_the_reference_writer(gpu):
    """servers/mysql/mysql_stream.rs:22-86: column types by data type, cells by arrow's array_value_to_string."""
    import pyarrow as pa
    from fuse_query_b200.tables import register_table
    ctx = make_ctx(gpu, 1, fuse=True)
    cols, rows = h.mysql_result_set(h.execute_sql(ctx, "select number + 1 as c1, number / 2 as c2 from system.numbers_mt(1000000) "
                                                       "where (c1 + c2 + 1) < 100 limit 3"))
    assert cols == [("c1", "MYSQL_TYPE_LONG"), ("c2", "MYSQL_TYPE_LONG")]
    assert rows == [["1", "0"], ["2", "0"], ["3", "1"]]          # README.md:120-126
    cols, rows = h.mysql_result_set(h.execute_sql(ctx, "select sum(number)/count(number), max(number) from system.numbers_mt(160000)"))
    assert cols == [("Sum(number) / Count(number)", "MYSQL_TYPE_LONG"), ("Max(number)", "MYSQL_TYPE_LONG")]
    assert rows == [["79999", "159999"]]
    tbl = pa.table({"f": pa.array([1.0, 0.1, None, 1e21, -2.5e-7, 3.0000000000000004], type=pa.float64()),
                    "i": pa.array([-3, 0, 7, None, 9, 10], type=pa.int16())})
    register_table(ctx, gpu, "default", "m", tbl)
    cols, rows = h.mysql_result_set(h.execute_sql(ctx, "select f, i, f * i from m"))
    assert cols == [("f", "MYSQL_TYPE_FLOAT"), ("i", "MYSQL_TYPE_LONG"), ("f * i", "MYSQL_TYPE_FLOAT")]
    assert rows == [["1", "-3", "-3"], ["0.1", "0", "0"], ["", "7", ""], ["1000000000000000000000", "", ""],
                    ["-0.00000025", "9", "-0.00000225"], ["3.0000000000000004", "10", "30.000000000000004"]]
    assert h.mysql_result_set([]) == ([], [])
    with pytest.raises(h.FuseQueryError) as e:                       # a Boolean column has no MySQL type in the reference
        h.mysql_result_set(h.execute_sql(ctx, "select number < 3 from system.numbers_mt(10)"))
    assert str(e.value) == "Internal Error: Unsupported column type:Boolean"


def test_pyarrow_table_with_nulls_registers_nullable_fields(gpu):
    import pyarrow as pa
    from fuse_query_b200.tables import register_table
    ctx = make_ctx(gpu, 1, fuse=True)
    tbl = pa.table({"x": pa.array([1, None, 3, None, 5], type=pa.uint64()), "y": pa.array([10, 20, 30, 40, 50], type=pa.int32())})
    t = register_table(ctx, gpu, "default", "pt", tbl)
    assert [f.nullable for f in t.schema().fields] == [True, False]
    assert rows_of(h.execute_sql(ctx, "select x + y, y from pt")) == [(11, 10), (None, 20), (33, 30), (None, 40), (55, 50)]
    assert rows_of(h.execute_sql(ctx, "select sum(x), count(x), max(x + y) from pt")) == [(9, 5, 55)]
    assert rows_of(h.execute_sql(ctx, "select y from pt where x > 1")) == [(30,), (50,)]


# ---------------------------------------------------------------------------------------------
# with a WHERE clause the reference folds Sum per 10 000-row block and an emptied block poisons it
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sql_where,pred,expect_error", [
    ("number >= 60000", f"(>= {NUM} (u64 60000))", True),                 # partitions 0..2 have only empty blocks
    ("number/10000*10000 = number", f"(= (* (/ {NUM} (u64 10000)) (u64 10000)) {NUM})", False),   # one row kept in every block
    ("number/20000*20000 = number", f"(= (* (/ {NUM} (u64 20000)) (u64 20000)) {NUM})", True),    # every other block empty
    ("number < 5", f"(< {NUM} (u64 5))", True),
])
def test_sum_with_where_reproduces_reference_block_poisoning(gpu, sql_where, pred, expect_error):
    n = 160_000   # 8 partitions of 2 blocks
    sql = f"select sum(number), count(number) from system.numbers_mt({n}) where {sql_where}"
    exprs = [f"(sum {NUM})", f"(count {NUM})"]
    msg = "Internal Error: DataValue to array cannot be NONE NULL"
    for workers in (0, 1, 2):
        try:
            want = o.run_query(exprs, total=n, predicate=pred, is_aggregate=True, worker_threads=workers).rows()
            oerr = None
        except o.OracleError as e:
            want, oerr = None, str(e)
        assert (oerr == msg) == expect_error
        for opts in (dict(fuse=True), dict(fuse=True, generated=True), dict(fuse=False, block_rows=10000)):
            ctx = make_ctx(gpu, workers, **opts)
            if expect_error:
                with pytest.raises(h.FuseQueryError) as e:
                    h.execute_sql(ctx, sql)
                assert str(e.value) == msg
            else:
                assert rows_of(h.execute_sql(ctx, sql)) == want
    # block_quirks = False: the mathematically merged answer instead of the reference's failure
    ctx = make_ctx(gpu, 1, block_quirks=False)
    x = np.arange(n, dtype=np.uint64)
    keep = {"number >= 60000": x >= 60000, "number/10000*10000 = number": x % 10000 == 0, "number/20000*20000 = number": x % 20000 == 0,
            "number < 5": x < 5}[sql_where]
    assert rows_of(h.execute_sql(ctx, sql)) == [(int(x[keep].sum()), int(keep.sum()))]


def test_more_than_eight_select_expressions_split_into_several_pipes(gpu):
    """A pipe holds FQ_MAX_EXPRS = 8 select expressions: wider projections, wider aggregate lists and a WHERE over a table of
    more than 8 columns (FilterTransform gathers every column) run as several launches and still match the oracle."""
    from fuse_query_b200.tables import register_table
    n = 80_000
    items = [f"number + {i} as c{i}" for i in range(11)]
    want = o.run_query([f"(alias c{i} (+ (col number) (u64 {i})))" for i in range(11)], total=n, predicate="(< (col number) (u64 7))",
                       worker_threads=1)
    for fuse in (True, False):
        ctx = make_ctx(gpu, 1, fuse=fuse)
        blocks = h.execute_sql(ctx, "select " + ", ".join(items) + f" from system.numbers_mt({n}) where number < 7")
        assert rows_of(blocks) == want.rows() and blocks[0].schema().names() == want.names
    aggs = [f"sum(number + {i})" for i in range(10)]
    ctx = make_ctx(gpu, 1)
    blocks = h.execute_sql(ctx, "select " + ", ".join(aggs) + f" from system.numbers_mt({n})")
    want = o.run_query([f"(sum (+ (col number) (u64 {i})))" for i in range(10)], total=n, is_aggregate=True, worker_threads=1)
    assert rows_of(blocks) == want.rows() and blocks[0].schema().names() == want.names
    rng = np.random.default_rng(5)
    data = {f"k{i}": rng.integers(0, 1000, 5000, dtype=np.uint64) for i in range(10)}
    ctx = make_ctx(gpu, 1, fuse=False, block_rows=0)
    register_table(ctx, gpu, "default", "wide", data)
    blocks = h.execute_sql(ctx, "select k0, k9 from wide where k3 < 10")
    keep = data["k3"] < 10
    assert rows_of(blocks) == list(zip(data["k0"][keep].tolist(), data["k9"][keep].tolist()))


def test_validity_masks_are_canonicalised_on_upload(gpu):
    """register_table's (values, valid) pairs may hold any non-zero byte for "valid"; the device sees 0/1 only."""
    from fuse_query_b200.tables import register_table
    vals = np.arange(1, 9, dtype=np.int64)
    valid = np.array([255, 0, 2, 1, 0, 128, 7, 0], dtype=np.uint8)
    ctx = make_ctx(gpu, 1)
    register_table(ctx, gpu, "default", "m", {"v": (vals, valid)})
    blocks = h.execute_sql(ctx, "select sum(v), count(v), min(v), max(v) from m")
    assert rows_of(blocks) == [(int(vals[valid != 0].sum()), 8, 1, 7)]
    blocks = h.execute_sql(ctx, "select v + 1 as w from m where v > 0")
    assert rows_of(blocks) == [(int(x) + 1,) for x in vals[valid != 0]]


@pytest.mark.parametrize("generated", [False, True])
def test_projection_errors_beyond_the_limit_follow_the_reference_blocks(gpu, generated):
    """The reference projects every kept row of a 10 000-row block before LimitStream cuts it, and LimitStream polls its
    input once more before it notices that the limit is complete (transform_projection.rs:45-56, stream_limit.rs:28-62): a
    zero divisor in a kept row AFTER the LIMIT-th row is an error when it sits in the same block or in the next one; two
    blocks on it is never evaluated.  The fused kernel only projects the rows it writes, so GpuPipeTransform settles the
    errors over exactly the reference's rows (block_quirks, default on); with the option off the launch's own outcome
    stands."""
    from fuse_query_b200.tables import register_table
    n = 80_000                                    # one source pipe (workers = 1): eight 10 000-row blocks
    cases = [(50, True), (9_999, True), (10_000, True), (19_999, True), (20_000, False), (30_050, False), (3, True)]   # zero divisor at row -> error?
    for row, want_error in cases:
        d = np.ones(n, dtype=np.uint64)
        d[row] = 0
        x = np.arange(n, dtype=np.uint64)
        table = {"x": o.from_numpy(x), "d": o.from_numpy(d)}
        for pred_sql, pred in ((" where x >= 0", "(>= (col x) (u64 0))"), ("", None)):
            def oracle():
                return o.run_query(["(alias q (/ (col x) (col d)))"], table=table, predicate=pred, limit=5, worker_threads=1, tail_quirk=False)
            ctx = make_ctx(gpu, 1)
            register_table(ctx, gpu, "default", "t", {"x": x, "d": d})
            sql = f"select x / d as q from t{pred_sql} limit 5"
            if want_error:
                with pytest.raises(o.OracleError) as oe:
                    oracle()
                with pytest.raises(h.FuseQueryError) as e:
                    h.execute_sql(ctx, sql)
                assert str(e.value) == str(oe.value) == "Internal Error: Divide by zero error"
            else:
                assert rows_of(h.execute_sql(ctx, sql)) == oracle().rows()
            # quirk off: only rows that are written are evaluated -> an error only if the zero divisor is among the first 5 rows
            ctx.options.block_quirks = False
            if row < 5:
                with pytest.raises(h.FuseQueryError):
                    h.execute_sql(ctx, sql)
            else:
                assert rows_of(h.execute_sql(ctx, sql)) == [(i,) for i in range(5)]
    # the predicate itself: evaluated over whole blocks up to the one completing the limit, never beyond
    ctx = make_ctx(gpu, 1)
    ctx.options.generated = generated
    sql = "select number from system.numbers_mt(80000) where 100 / (number - 20000) < 1000 limit 5"     # zero divisor at row 20 000: block 2
    want = o.run_query(["(col number)"], total=80_000, predicate="(< (/ (u64 100) (- (col number) (u64 20000))) (u64 1000))", limit=5,
                       worker_threads=1)
    assert rows_of(h.execute_sql(ctx, sql)) == want.rows() and len(want.rows()) == 5
    ctx.options.limit_early_exit = False         # full scan: the kernel sees the zero divisor, the reference never pulls that block
    assert rows_of(h.execute_sql(ctx, sql)) == want.rows()


def test_group_by_through_sql_matches_the_group_by_oracle(gpu):
    """AggregatePlan{group_expr, aggr_expr} executed as planned (plan_parser.rs:279-308; the reference's pipeline builder
    drops group_expr, pipeline_builder.rs:50-65): schema = group fields then aggregate fields (plan_builder.rs:63-83), one
    row per group, arithmetic over aggregates applied per group.  Oracle-anchored (oracle/groupby.py); order unspecified."""
    from fuse_query_b200.tables import register_table
    from oracle.groupby import run_group_by
    n = 160_000
    key = "(- (col number) (* (/ (col number) (u64 7)) (u64 7)))"
    aggs = ["(sum (col number))", "(count (col number))", "(/ (sum (col number)) (count (col number)))", "(max (+ (col number) (u64 1)))"]
    sql = ("select number - number / 7 * 7, sum(number), count(number), sum(number) / count(number), max(number + 1) "
           f"from system.numbers_mt({n}) where number >= 10 group by number - number / 7 * 7")
    names, want = run_group_by([key], aggs, total=n, predicate="(>= (col number) (u64 10))")
    for generated in (False, True):
        for workers in (0, 1):
            ctx = make_ctx(gpu, workers)
            ctx.options.generated = generated
            blocks = h.execute_sql(ctx, sql)
            assert blocks[0].schema().names() == names == ["number - number / 7 * 7", "Sum(number)", "Count(number)", "Sum(number) / Count(number)",
                                                           "Max(number + 1)"]
            assert sorted(rows_of(blocks)) == want
    # select list order does not matter: the plan's schema is group fields first (the reference's own rule)
    ctx = make_ctx(gpu, 1)
    blocks = h.execute_sql(ctx, f"select count(number), number / 40000 from system.numbers_mt({n}) group by number / 40000")
    assert blocks[0].schema().names() == ["number / 40000", "Count(number)"] and sorted(rows_of(blocks)) == [(i, 40000) for i in range(4)]
    # keys only (DISTINCT), and a table with NULL keys / NULL values
    blocks = h.execute_sql(ctx, f"select number / 50000 from system.numbers_mt({n}) group by number / 50000")
    assert sorted(rows_of(blocks)) == [(0,), (1,), (2,), (3,)]
    rng = np.random.default_rng(3)
    m = 50_000
    k = rng.integers(0, 20, m).astype(np.int16)
    k_ok = rng.random(m) > 0.1
    v = rng.integers(-500, 500, m).astype(np.int64)
    v_ok = rng.random(m) > 0.3
    register_table(ctx, gpu, "default", "g", {"k": (k, k_ok), "v": (v, v_ok)})
    blocks = h.execute_sql(ctx, "select k, sum(v), min(v), count(v), sum(v) / count(v) from g group by k")
    table = {"k": o.Array(o.I16, k, k_ok.astype(np.uint8)), "v": o.Array(o.I64, v, v_ok.astype(np.uint8))}
    _, want = run_group_by(["(col k)"], ["(sum (col v))", "(min (col v))", "(count (col v))", "(/ (sum (col v)) (count (col v)))"], table=table)
    got = rows_of(blocks)
    assert sorted(got, key=lambda r: (r[0] is not None, r[0])) == want and len(want) == 21
    # the reference's own behaviour (GROUP BY planned, then ignored) stays available
    ctx.options.group_by = False
    blocks = h.execute_sql(ctx, f"select number / 40000, count(number) from system.numbers_mt({n}) group by number / 40000")
    assert rows_of(blocks) == [(n,)]
    # errors: an aggregate as a key, more key bits than the table key holds
    ctx.options.group_by = True
    with pytest.raises(h.FuseQueryError):
        h.execute_sql(ctx, f"select sum(number), count(number) from system.numbers_mt({n}) group by sum(number)")


def test_order_by_through_sql_matches_the_sort_oracle(gpu):
    """ORDER BY executes as a device sort (GpuSortTransform -> fq_sort_indices / fq_column_take) between the projection or
    aggregation and the LIMIT.  The reference does not sort (README.md:28), so the oracle is the stated semantics
    (oracle/sort.py: ASC unless DESC, NULLs first, stable in partition order)."""
    from fuse_query_b200.tables import register_table
    from oracle.sort import sort_indices
    n = 160_000                      # 8 partitions of whole 10 000-row blocks (SURVEY F7 drops the tail of others)
    for workers in (0, 1):          # 8 pipes merged in partition order / one pipe
        ctx = make_ctx(gpu, workers)
        blocks = h.execute_sql(ctx, f"select number, number / 3 as t from system.numbers_mt({n}) where number - number / 7 * 7 = 3 "
                                    "order by number - number / 5 * 5 desc, number limit 7")
        x = np.arange(n, dtype=np.uint64)
        x = x[x % 7 == 3]
        perm = sort_indices([x % 5, x], descending=[True, False])[:7]
        assert rows_of(blocks) == [(int(v), int(v) // 3) for v in x[perm]]
    # GROUP BY ... ORDER BY an aggregate's output column (by its display name or alias), descending
    ctx = make_ctx(gpu, 1)
    blocks = h.execute_sql(ctx, f"select number - number / 7 * 7 as k, sum(number) as s from system.numbers_mt({n}) group by number - number / 7 * 7 order by s desc")
    x = np.arange(n, dtype=np.uint64)
    sums = [(int(k), int(x[x % 7 == k].sum())) for k in range(7)]
    assert rows_of(blocks) == sorted(sums, key=lambda r: -r[1])
    # a table with NULLs: NULL keys first, ties keep table order; the payload's validity travels with the rows
    rng = np.random.default_rng(5)
    m = 30_000
    a = rng.integers(-20, 20, m).astype(np.int32)
    a_ok = rng.random(m) > 0.2
    b = rng.normal(size=m)
    b_ok = rng.random(m) > 0.5
    register_table(ctx, gpu, "default", "srt", {"a": (a, a_ok), "b": (b, b_ok)})
    got = rows_of(h.execute_sql(ctx, "select a, b from srt order by a desc, b"))
    perm = sort_indices([a, b], [a_ok, b_ok], [True, False])
    want = [(int(a[i]) if a_ok[i] else None, float(b[i]) if b_ok[i] else None) for i in perm]
    assert got == want
    assert rows_of(h.execute_sql(ctx, "select a from srt where a > 100 order by a")) == []


def test_utf8_arrays_on_the_device(gpu):
    """Arrow string arrays on the device (fq_utf8): every comparison operator array/array, array/scalar and scalar/array
    (flipped like data_array_comparison.rs:75-85), NULL slots, min / max with ties and empties — against python's own
    bytewise string order (Rust str ordering, which arrow's *_utf8 kernels use)."""
    from fuse_query_b200 import cabi
    import random
    rng = random.Random(7)
    alphabet = ["", "a", "ab", "abc", "b", "ba", "x1", "x2", "é", "zz", "Zebra", "a\u00e9", "日本"]
    n = 5000
    left = [rng.choice(alphabet) + rng.choice(alphabet) for _ in range(n)]
    right = [rng.choice(alphabet) + rng.choice(alphabet) for _ in range(n)]
    left_n = [None if rng.random() < 0.1 else v for v in left]
    ctx = cabi.Context(0)
    L, R, LN = ctx.utf8(left), ctx.utf8(right), ctx.utf8(left_n)
    key = lambda s: s.encode()
    ops = {"=": lambda a, b: key(a) == key(b), "<": lambda a, b: key(a) < key(b), "<=": lambda a, b: key(a) <= key(b),
           ">": lambda a, b: key(a) > key(b), ">=": lambda a, b: key(a) >= key(b)}
    for op, f in ops.items():
        assert L.compare(op, R) == [f(a, b) for a, b in zip(left, right)]
        assert L.compare(op, "ab") == [f(a, "ab") for a in left]
        assert LN.compare(op, R) == [None if a is None else f(a, b) for a, b in zip(left_n, right)]
    valid = [v for v in left_n if v is not None]
    assert left_n[LN.minmax("min")] == min(valid, key=key) and LN.minmax("min") == left_n.index(min(valid, key=key))
    assert left_n[LN.minmax("max")] == max(valid, key=key) and LN.minmax("max") == left_n.index(max(valid, key=key))
    assert ctx.utf8([]).minmax("min") == -1 and ctx.utf8([None, None]).minmax("max") == -1
    # through the host mirror: scalar (op) array flips the operator; a WHERE over a Utf8 column filters on the device result
    arr = h.DataColumnarValue.Array(h.DataArray.utf8(left))
    sc = h.DataColumnarValue.Scalar(h.DataValue(TAG["Utf8"], "b"))
    assert h.data_array_comparison_op(gpu, OPS["Lt"], sc, arr).to_list() == [key("b") < key(a) for a in left]
    with pytest.raises(h.FuseQueryError) as e:
        h.data_array_comparison_op(gpu, OPS["Eq"], sc, sc)
    assert str(e.value) == "Internal Error: Cannot do data_array =, left:Utf8, right:Utf8"
    with pytest.raises(h.FuseQueryError) as e:
        h.data_array_comparison_op(gpu, OPS["Eq"], arr, h.DataColumnarValue.Scalar(h.DataValue(TAG["Int8"], 1)))
    assert str(e.value) == "Internal Error: Unsupported (Utf8) = (Int8)"
    ctx.close()
