#!/usr/bin/env python3
"""Extract the known-answer vectors of the reference's own unit tests into JSON fixtures.

Run in the build container (the only place /root/reference exists):

    python tests/golden/extract_reference_vectors.py

It parses the table-driven `let tests = vec![ ... ]` literals of the reference's `*_test.rs` files
with a small Rust-literal parser (no Rust toolchain exists in this image, so the tests cannot be
executed; their *expected values* are what pins the oracle) and writes tests/golden/ref_*.json.
Every case records the file:line it came from.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("FQ_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# tokenizer / parser for the subset of Rust used in the test tables
# ---------------------------------------------------------------------------------------------
TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<num>-?\d[\d_]*(?:\.\d+)?(?:[iuf](?:8|16|32|64))?)
  | (?P<path>[A-Za-z_][A-Za-z0-9_]*(?:::[A-Za-z_][A-Za-z0-9_]*)*!?)
  | (?P<punct>[\[\]\(\)\{\},:;&\?\.\*=<>\|])
""", re.X)


class Tok:
    def __init__(self, kind, text, pos):
        self.kind, self.text, self.pos = kind, text, pos

    def __repr__(self):
        return f"{self.kind}:{self.text!r}"


def tokenize(src: str, start: int, end: int):
    toks, i = [], start
    while i < end:
        m = TOKEN.match(src, i)
        if not m:
            raise SyntaxError(f"cannot tokenize at {i}: {src[i:i+40]!r}")
        i = m.end()
        if m.lastgroup == "ws":
            continue
        toks.append(Tok(m.lastgroup, m.group(), m.start()))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else Tok("eof", "", -1)

    def eat(self, text=None):
        tok = self.peek()
        if text is not None and tok.text != text:
            raise SyntaxError(f"expected {text!r} got {tok!r} at {tok.pos}")
        self.i += 1
        return tok

    def list_until(self, close):
        items = []
        while self.peek().text != close:
            items.append(self.expr())
            if self.peek().text == ",":
                self.eat()
        self.eat(close)
        return items

    def postfix(self, node):
        while True:
            tok = self.peek()
            if tok.text == "?":
                self.eat()
            elif tok.text == "." and self.peek(1).kind == "path" and self.peek(2).text == "(":
                method = self.peek(1).text
                self.eat(); self.eat(); self.eat("(")
                args = self.list_until(")")
                if method in ("clone", "to_string", "to_owned", "as_str", "unwrap"):
                    pass
                else:
                    node = {"method": method, "recv": node, "args": args}
            else:
                return node

    def expr(self):
        tok = self.eat()
        if tok.text == "&":
            return self.expr()
        if tok.kind == "str":
            return self.postfix(json.loads(tok.text))
        if tok.kind == "num":
            txt = re.sub(r"[iuf](8|16|32|64)$", "", tok.text).replace("_", "")
            return float(txt) if "." in txt else int(txt)
        if tok.text == "[":
            return self.postfix(self.list_until("]"))
        if tok.text == "(":
            inner = self.list_until(")")
            return self.postfix(inner[0] if len(inner) == 1 else inner)
        if tok.kind == "path":
            name = tok.text
            if name in ("true", "false"):
                return name == "true"
            if name == "vec!":
                self.eat("[")
                return self.postfix(self.list_until("]"))
            nxt = self.peek()
            if nxt.text == "(":
                self.eat()
                args = self.list_until(")")
                return self.postfix(self.call(name, args, tok.pos))
            if nxt.text == "{" and name[0].isupper():
                self.eat()
                fields = {"__struct__": name, "__pos__": tok.pos}
                while self.peek().text != "}":
                    key = self.eat().text
                    if self.peek().text in (",", "}"):  # field-init shorthand
                        fields[key] = {"ident": key}
                    else:
                        self.eat(":")
                        fields[key] = self.expr()
                    if self.peek().text == ",":
                        self.eat()
                self.eat("}")
                return fields
            if "::" in name:
                return self.postfix({"enum": name})
            return self.postfix({"ident": name})
        raise SyntaxError(f"unexpected {tok!r} at {tok.pos}")

    @staticmethod
    def call(name, args, pos):
        if name in ("Arc::new", "Some", "Box::new"):
            return args[0]
        m = re.match(r"(\w+)Array::from$", name)
        if m:
            kind = {"String": "Utf8"}.get(m.group(1), m.group(1))
            return {"array": kind, "values": args[0]}
        m = re.match(r"DataValue::(\w+)$", name)
        if m:
            kind = {"String": "Utf8"}.get(m.group(1), m.group(1))
            v = args[0]
            if isinstance(v, dict) and v.get("enum") == "None" or v == {"ident": "None"}:
                v = None
            return {"value": kind, "v": v}
        return {"call": name, "args": args, "__pos__": pos}


def expr_end(src: str, start: int, limit: int) -> int:
    """Index of the `;` that ends the statement starting at `start` (brackets and strings respected)."""
    depth, i = 0, start
    while i < limit:
        c = src[i]
        if c == '"':
            i += 1
            while src[i] != '"':
                i += 2 if src[i] == "\\" else 1
        elif c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == ";" and depth == 0:
            return i
        i += 1
    return limit


def line_of(src: str, pos: int) -> int:
    return src.count("\n", 0, pos) + 1


def find_tables(src: str):
    """Yield (fn_name, parsed `tests` vec, bindings) for every #[test] fn holding `let tests = vec![`."""
    for m in re.finditer(r"fn (test_\w+)\(\)", src):
        body_start = m.end()
        nxt = re.search(r"\n#\[(?:tokio::)?test", src[body_start:])
        body_end = body_start + nxt.start() if nxt else len(src)
        t = re.search(r"let tests = ", src[body_start:body_end])
        if not t:
            continue
        # simple `let name = <expr>;` bindings before the table (field_a, schema, block ...)
        bindings = {}
        for b in re.finditer(r"let (\w+) = ", src[body_start:body_start + t.start()]):
            try:
                b0 = body_start + b.end()
                p = Parser(tokenize(src, b0, expr_end(src, b0, body_end)))
                bindings[b.group(1)] = p.expr()
            except SyntaxError:
                pass
        t0 = body_start + t.end()
        p = Parser(tokenize(src, t0, expr_end(src, t0, body_end)))
        yield m.group(1), p.expr(), bindings


def resolve(node, bindings):
    if isinstance(node, dict):
        if "ident" in node and node["ident"] in bindings:
            return resolve(bindings[node["ident"]], bindings)
        return {k: resolve(v, bindings) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, bindings) for v in node]
    return node


def opname(node):
    return node["enum"].split("::")[-1]


# ---------------------------------------------------------------------------------------------
# per-file converters
# ---------------------------------------------------------------------------------------------
def rel(path):
    return os.path.relpath(path, REF)


def expected_or_error(t, i, src_errors_are_per_case=True):
    return t["expect"][i], (t["error"][i] if i < len(t["error"]) else "")


def extract_datavalues():
    cases = []
    base = os.path.join(REF, "src/datavalues")

    def add(path, src, t, **kw):
        kw["source"] = f"{rel(path)}:{line_of(src, t['__pos__'])}"
        kw["name"] = t["name"]
        cases.append(kw)

    for fname, kind in (("data_array_arithmetic_test.rs", "arithmetic"), ("data_array_comparison_test.rs", "comparison"),
                        ("data_array_logic_test.rs", "logic")):
        path = os.path.join(base, fname)
        src = open(path).read()
        for fn, tests, _ in find_tables(src):
            for t in tests:
                op = opname(t["op"])
                if "args" in t:  # array (op) array, one result or error per argument pair
                    for i, pair in enumerate(t["args"]):
                        add(path, src, t, kind=f"array_{kind}", form="array-array", op=op, left=pair[0], right=pair[1],
                            expect=t["expect"][i], error=t["error"][i] if i < len(t["error"]) else "")
                else:
                    form = "scalar-array" if fn.startswith("test_scalar_array") else "array-scalar"
                    l, r = (t["scalar"], t["array"]) if form == "scalar-array" else (t["array"], t["scalar"])
                    add(path, src, t, kind=f"array_{kind}", form=form, op=op, left=l, right=r, expect=t["expect"],
                        error=t["error"])
    path = os.path.join(base, "data_array_aggregate_test.rs")
    src = open(path).read()
    for fn, tests, _ in find_tables(src):
        for t in tests:
            for i, arr in enumerate(t["args"]):
                add(path, src, t, kind="array_aggregate", op=opname(t["op"]), array=arr, expect=t["expect"][i],
                    error=t["error"][i] if i < len(t["error"]) else "")
    for fname, kind in (("data_value_aggregate_test.rs", "value_aggregate"), ("data_value_arithmetic_test.rs", "value_arithmetic")):
        path = os.path.join(base, fname)
        src = open(path).read()
        for fn, tests, _ in find_tables(src):
            for t in tests:
                for i, pair in enumerate(t["args"]):
                    add(path, src, t, kind=kind, op=opname(t["op"]), left=pair[0], right=pair[1], expect=t["expect"][i],
                        error=t["error"][i] if i < len(t["error"]) else "")
    return cases


def fn_to_sexpr(node):
    """Function constructor calls of the reference tests -> the oracle's s-expression."""
    if isinstance(node, dict) and "call" in node:
        name, args = node["call"], node["args"]
        if name == "FieldFunction::try_create":
            return f"(col {args[0]})"
        if name == "ConstantFunction::try_create":
            v = args[0]
            ty = {"Int8": "i8", "Int16": "i16", "Int32": "i32", "Int64": "i64", "UInt8": "u8", "UInt16": "u16",
                  "UInt32": "u32", "UInt64": "u64", "Float32": "f32", "Float64": "f64", "Utf8": "str",
                  "Boolean": "bool"}[v["value"]]
            return f"({ty} {v['v']})"
        sym = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "Eq": "=", "Lt": "<", "LtEq": "<=", "Gt": ">",
               "GtEq": ">=", "And": "and", "Or": "or", "Count": "count", "Min": "min", "Max": "max", "Sum": "sum"}
        if name.endswith("Function::try_create"):
            op = sym[opname(args[0])]
            return "(" + " ".join([op] + [fn_to_sexpr(a) for a in args[1]]) + ")"
    raise ValueError(f"cannot convert {node!r}")


def block_of(node):
    """DataBlock::create(schema, vec![arrays]) with schema = DataSchema::new(vec![DataField::new(name, ..)])"""
    schema, arrays = node["args"]
    fields = schema["args"][0]
    names = [f["args"][0] for f in fields]
    return {"names": names[:len(arrays)], "columns": arrays}


def extract_functions():
    cases = []
    base = os.path.join(REF, "src/functions")
    sym = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "Eq": "=", "Lt": "<", "LtEq": "<=", "Gt": ">", "GtEq": ">=",
           "And": "and", "Or": "or"}
    for fname in ("function_arithmetic_test.rs", "function_comparison_test.rs", "function_logic_test.rs"):
        path = os.path.join(base, fname)
        src = open(path).read()
        for fn, tests, bindings in find_tables(src):
            for t in tests:
                t = resolve(t, bindings)
                args = [fn_to_sexpr(a) for a in t["args"]]
                cases.append({"kind": "function_eval", "source": f"{rel(path)}:{line_of(src, t['__pos__'])}",
                              "name": t["name"], "sexpr": f"({sym[opname(t['op'])]} {args[0]} {args[1]})",
                              "display": t["display"], "nullable": t["nullable"], "block": block_of(t["block"]),
                              "expect": t["expect"], "error": t["error"]})
    path = os.path.join(base, "function_aggregator_test.rs")
    src = open(path).read()
    for fn, tests, bindings in find_tables(src):
        for t in tests:
            t = resolve(t, bindings)
            cases.append({"kind": "function_aggregate", "source": f"{rel(path)}:{line_of(src, t['__pos__'])}",
                          "name": t["name"], "sexpr": fn_to_sexpr(t["func"]), "evals": t["evals"],
                          "block": block_of(t["block"]), "expect": t["expect"], "error": t["error"]})
    path = os.path.join(base, "function_factory_test.rs")
    if os.path.exists(path):
        src = open(path).read()
        cases.append({"kind": "factory_source", "source": rel(path), "text_sha": str(hash(src) & 0xFFFFFFFF)})
    return [c for c in cases if c["kind"] != "factory_source"]


def rust_str_literals(src: str, var: str):
    """`let expect = "\\\n ...";` multi-line literals (backslash-newline continuations)."""
    out = []
    for m in re.finditer(r"let " + var + r"\s*=\s*(\"(?:[^\"\\]|\\.|\\\n)*\")", src):
        lit = m.group(1)
        lit = re.sub(r"\\\n\s*", "", lit)  # Rust line continuation swallows leading whitespace
        out.append((json.loads(lit), line_of(src, m.start())))
    return out


def extract_strings():
    """SQL -> plan / pipeline EXPLAIN golden strings."""
    cases = []
    for relpath in ("src/planners/plan_select_test.rs", "src/planners/plan_filter_test.rs", "src/planners/plan_explain_test.rs",
                    "src/optimizers/optimizer_filter_push_down_test.rs", "src/processors/pipeline_builder_test.rs",
                    "src/planners/plan_expression_test.rs", "src/planners/plan_builder_test.rs"):
        path = os.path.join(REF, relpath)
        if not os.path.exists(path):
            continue
        src = open(path).read()
        sqls = re.findall(r"build_from_sql\(\s*ctx\.clone\(\),\s*(\"(?:[^\"\\]|\\.)*\")", src)
        expects = rust_str_literals(src, "expect")
        for i, (text, line) in enumerate(expects):
            cases.append({"kind": "golden_string", "source": f"{relpath}:{line}", "sql": json.loads(sqls[i]) if i < len(sqls) else None,
                          "optimized": "optimizer" in relpath, "pipeline": "pipeline_builder" in relpath, "expect": text})
    return cases


def extract_pipeline():
    """Known answers of the reference's pipeline tests.  The query shapes are restated by hand (the
    tests build plans programmatically); each expected literal is checked to still be present in the
    reference file at extraction time."""
    spec = [
        # (file, regex that must match, case)
        ("src/transforms/transform_aggregate_test.rs", r"UInt64Array::from\(vec!\[122\]\)",
         {"name": "sum(number)+2 over numbers_mt(16), partial->merge->final", "total": 16, "worker_threads": 0,
          "exprs": ["(+ (sum (col number)) (u64 2))"], "is_aggregate": True, "expect_rows": [[122]],
          "expect_dtypes": ["UInt64"]}),
        ("src/transforms/transform_filter_test.rs", r"UInt64Array::from\(vec!\[1\]\)",
         {"name": "filter number = 1 over numbers_mt(8)", "total": 8, "worker_threads": 0, "exprs": ["(col number)"],
          "predicate": "(= (col number) (u64 1))", "is_aggregate": False, "expect_rows": [[1]], "expect_dtypes": ["UInt64"]}),
        ("src/transforms/transform_limit_test.rs", r"assert_eq!\(2, rows\)",
         {"name": "limit 2 over numbers_mt(8)", "total": 8, "worker_threads": 0, "exprs": ["(col number)"], "limit": 2,
          "is_aggregate": False, "expect_n_rows": 2}),
        ("src/transforms/transform_source_test.rs", r"assert_eq!\(16, rows\)",
         {"name": "two numbers_mt(8) sources merged -> 16 rows (one source here: 8 rows each)", "total": 8,
          "worker_threads": 0, "exprs": ["(col number)"], "is_aggregate": False, "expect_n_rows": 8}),
        ("src/processors/processor_merge_test.rs", r"UInt64Array::from\(vec!\[0, 1\]\)",
         {"name": "first block of numbers_mt(16) is [0, 1] (8 partitions of 2 rows)", "total": 16, "worker_threads": 0,
          "exprs": ["(col number)"], "is_aggregate": False, "expect_first_rows": [[0], [1]], "expect_n_rows": 16}),
    ]
    cases = []
    for relpath, pattern, case in spec:
        src = open(os.path.join(REF, relpath)).read()
        m = re.search(pattern, src)
        if not m:
            raise SystemExit(f"{relpath}: expected literal /{pattern}/ not found")
        case.update(kind="pipeline", source=f"{relpath}:{line_of(src, m.start())}")
        cases.append(case)
    # README sample output (README.md:120-126)
    readme = open(os.path.join(REF, "README.md")).read()
    m = re.search(r"\|\s+1 \|\s+0 \|\n\|\s+2 \|\s+0 \|\n\|\s+3 \|\s+1 \|", readme)
    if not m:
        raise SystemExit("README sample rows not found")
    cases.append({"kind": "pipeline", "source": f"README.md:{line_of(readme, m.start())}",
                  "name": "README filter/projection/limit sample", "total": 10000000, "worker_threads": 8,
                  "exprs": ["(alias c1 (+ (col number) (u64 1)))", "(alias c2 (/ (col number) (u64 2)))"],
                  "predicate": "(< (+ (+ (+ (col number) (u64 1)) (/ (col number) (u64 2))) (u64 1)) (u64 100))",
                  "limit": 3, "is_aggregate": False, "expect_rows": [[1, 0], [2, 0], [3, 1]],
                  "expect_names": ["c1", "c2"]})
    return cases


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: fixtures can only be regenerated in the build container")
    out = {"ref_datavalues.json": extract_datavalues(), "ref_functions.json": extract_functions(),
           "ref_strings.json": extract_strings(), "ref_pipeline.json": extract_pipeline()}
    for name, cases in out.items():
        for c in cases:
            c.pop("__pos__", None)
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(cases, f, indent=1, sort_keys=True, default=str)
        print(f"{name}: {len(cases)} cases")


if __name__ == "__main__":
    main()
