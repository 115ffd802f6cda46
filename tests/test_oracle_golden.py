"""Pins the CPU oracle against every known-answer vector of the reference's own tests for the path
(SURVEY.md §8c).  Fixtures: tests/golden/ref_*.json, produced by tests/golden/extract_reference_vectors.py
from /root/reference/src/**/*_test.rs (file:line recorded per case)."""
import math

import pytest

from conftest import golden
from oracle import binding as o

TAG = {n: i for i, n in enumerate(o.DTYPE_NAMES)}
SYM = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "Eq": "=", "Lt": "<", "LtEq": "<=", "Gt": ">", "GtEq": ">=",
       "And": "and", "Or": "or", "Min": "min", "Max": "max", "Sum": "sum", "Count": "count"}


def to_array(spec):
    return o.array(TAG[spec["array"]], spec["values"])


def to_value(spec):
    return o.Value(TAG[spec["value"]], spec["v"])


def operand(spec):
    return to_array(spec) if "array" in spec else to_value(spec)


def same(a, b):
    if isinstance(a, float) or isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    return a == b


DV = golden("ref_datavalues.json")
FN = golden("ref_functions.json")
PIPE = golden("ref_pipeline.json")


def ident(c):
    return f"{c['source'].split('/')[-1]}-{c['name']}"


@pytest.mark.parametrize("case", [c for c in DV if c["kind"].startswith("array_") and c["kind"] != "array_aggregate"],
                         ids=ident)
def test_array_binary_ops(case):
    fn = {"array_arithmetic": o.array_arithmetic, "array_comparison": o.array_comparison, "array_logic": o.array_logic}[case["kind"]]
    try:
        got = fn(SYM[case["op"]], operand(case["left"]), operand(case["right"]))
    except o.OracleError as e:
        assert case["error"], f"unexpected error {e}"
        assert str(e) == case["error"]
        return
    exp = case["expect"]
    assert o.DTYPE_NAMES[got.dtype] == exp["array"]
    assert got.to_list() == to_array(exp).to_list()  # expected literals rounded to the lane type (f32)


@pytest.mark.parametrize("case", [c for c in DV if c["kind"] == "array_aggregate"], ids=ident)
def test_array_aggregate(case):
    try:
        got = o.array_aggregate(SYM[case["op"]], to_array(case["array"]))
    except o.OracleError as e:
        assert str(e) == case["error"]
        return
    assert got == to_value(case["expect"])


@pytest.mark.parametrize("case", [c for c in DV if c["kind"].startswith("value_")], ids=ident)
def test_value_ops(case):
    fn = o.value_aggregate if case["kind"] == "value_aggregate" else o.value_arithmetic
    try:
        got = fn(SYM[case["op"]], to_value(case["left"]), to_value(case["right"]))
    except o.OracleError as e:
        assert case["error"], f"unexpected error {e}"
        assert str(e) == case["error"]
        return
    exp = to_value(case["expect"])
    assert got.tag == exp.tag and same(got.value, exp.value)


def block(spec):
    return {n: to_array(c) for n, c in zip(spec["names"], spec["columns"])}


@pytest.mark.parametrize("case", [c for c in FN if c["kind"] == "function_eval"], ids=ident)
def test_function_eval(case):
    f = o.Function(case["sexpr"])
    assert f.display() == case["display"]
    blk = block(case["block"])
    got = f.eval(blk)
    assert o.DTYPE_NAMES[f.return_type(blk)] == o.DTYPE_NAMES[got.dtype] == case["expect"]["array"]
    assert got.to_list() == case["expect"]["values"]


@pytest.mark.parametrize("case", [c for c in FN if c["kind"] == "function_aggregate"], ids=ident)
def test_function_aggregate_protocol(case):
    """accumulate x evals -> state1; accumulate x (evals-1) -> state2; merge both; merge_result
    (function_aggregator_test.rs:168-188)."""
    blk = block(case["block"])
    proto = o.Function(case["sexpr"])
    f1 = proto.clone()
    for _ in range(case["evals"]):
        f1.accumulate(blk)
    f2 = proto.clone()
    for _ in range(1, case["evals"]):
        f2.accumulate(blk)
    final = proto.clone()
    final.set_depth(0)
    final.merge_state(f1.accumulate_result())
    final.merge_state(f2.accumulate_result())
    assert final.merge_result() == to_value(case["expect"])


@pytest.mark.parametrize("case", PIPE, ids=ident)
def test_pipeline_known_answers(case):
    r = o.run_query(case["exprs"], total=case["total"], predicate=case.get("predicate"), is_aggregate=case["is_aggregate"],
                    limit=case.get("limit"), worker_threads=case["worker_threads"])
    if "expect_rows" in case:
        assert [list(t) for t in r.rows()] == case["expect_rows"]
    if "expect_n_rows" in case:
        assert r.n_rows == case["expect_n_rows"]
    if "expect_first_rows" in case:
        assert [list(t) for t in r.rows()[:len(case["expect_first_rows"])]] == case["expect_first_rows"]
    if "expect_dtypes" in case:
        assert [o.DTYPE_NAMES[c.dtype] for c in r.columns] == case["expect_dtypes"]
    if "expect_names" in case:
        assert r.names == case["expect_names"]


# ---------------------------------------------------------------------------------------------
# size-independent properties of the oracle itself (so that it can be trusted beyond the golden sizes)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("total", [80_000, 240_000, 1_600_000])
def test_aggregates_do_not_depend_on_how_partitions_are_chunked(total):
    """pipeline_builder.rs:75-84 chunks the 8 partitions over `worker_threads` pipes; partial states merge by wrapping add /
    min / max (function_aggregator.rs:106-139), so every chunking gives the closed form."""
    exprs = ["(/ (sum (col number)) (count (col number)))", "(max (+ (col number) (u64 1)))", "(min (col number))", "(count (col number))",
             "(sum (* (col number) (col number)))"]
    n = total
    want = ((n * (n - 1) // 2) // n, n, 0, n, sum(i * i for i in range(n)) % (1 << 64))
    for workers in (0, 1, 2, 3, 4, 8, 16):
        for threads in (False, True):
            r = o.run_query(exprs, total=total, is_aggregate=True, worker_threads=workers, use_threads=threads)
            assert r.rows() == [want], (workers, threads)


def test_tail_quirk_rows_match_the_block_formula():
    """numbers_stream.rs:37-54: a partition of c >= 10000 rows, c % 10000 = r > 0, emits 10000 * (c / 10000 - 1) + r + 1 rows."""
    for total in (80_008, 100_000, 123_456, 799_999):
        parts = o.generate_parts(total)
        expect = 0
        for b, e in parts:
            c = e - b + 1
            nblk, rem = divmod(c, 10_000)
            expect += c if (nblk == 0 or rem == 0) else 10_000 * (nblk - 1) + rem + 1
        r = o.run_query(["(count (col number))"], total=total, is_aggregate=True, worker_threads=0)
        assert r.rows() == [(expect,)]
        fixed = o.run_query(["(count (col number))"], total=total, is_aggregate=True, worker_threads=0, tail_quirk=False)
        assert fixed.rows() == [(total,)]
