/*
 * fq_oracle.h — CPU restatement of fuse-query's vectorised
 * Source -> Filter -> Projection -> AggregatePartial -> Merge -> AggregateFinal / Limit
 * hot path.  TEST INFRASTRUCTURE ONLY: this is the checker the CUDA path is compared
 * against (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * leg).  Nothing under fuse_query_b200/ may include, link or call it.
 *
 * The reference (Rust, /root/reference) cannot be compiled in this image (no rustc /
 * cargo) and its arithmetic lives in the un-vendored crate `arrow = "2.0"` (feature
 * "simd", Cargo.toml:19; resolves to 2.0.0).  This file restates the reference's own
 * dispatch / coercion / state plumbing line by line (citations below are relative to
 * /root/reference/src) and the published Arrow 2.0.0 semantics of the kernels it calls
 * (wrapping integer add/sub/mul, truncating integer divide that errors on a zero
 * divisor, `sum`/`min`/`max` returning None on empty or all-null input, num::cast-style
 * numeric `cast` that yields null when out of range).
 *
 * Parity pin: every known-answer vector of the reference's own tests for this path is
 * extracted by tests/golden/extract_reference_vectors.py into the tests/golden JSON files and
 * replayed against this oracle by tests/test_oracle_golden.py.  Behaviours no reference
 * test pins (u64 sum wrap at 10^10 rows, the NumbersStream tail quirk, empty-block sum)
 * are listed as "unpinned" in DESIGN.md.  NULL propagation and wrapping at every
 * numeric width, which those vectors do not exercise, are cross-checked against an
 * independent Arrow implementation (pyarrow.compute) by tests/test_oracle_vs_pyarrow.py.
 */
#ifndef FQ_ORACLE_H
#define FQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DataValue / DataType tags, in the declaration order of datavalues/data_value.rs:19-35 */
enum {
  ORC_NULL = 0, ORC_BOOL = 1, ORC_I8 = 2, ORC_I16 = 3, ORC_I32 = 4, ORC_I64 = 5,
  ORC_U8 = 6, ORC_U16 = 7, ORC_U32 = 8, ORC_U64 = 9, ORC_F32 = 10, ORC_F64 = 11,
  ORC_UTF8 = 12, ORC_STRUCT = 13
};

/* operator enums, datavalues/data_value_operator.rs:5-81 */
enum { ORC_AGG_MIN = 0, ORC_AGG_MAX = 1, ORC_AGG_SUM = 2, ORC_AGG_COUNT = 3 };
enum { ORC_CMP_EQ = 0, ORC_CMP_LT = 1, ORC_CMP_LTEQ = 2, ORC_CMP_GT = 3, ORC_CMP_GTEQ = 4 };
enum { ORC_AR_ADD = 0, ORC_AR_SUB = 1, ORC_AR_MUL = 2, ORC_AR_DIV = 3 };
enum { ORC_LG_AND = 0, ORC_LG_OR = 1 };

/* DataValue (data_value.rs:19-35).  `some` = 0 encodes Type(None). */
typedef struct orc_value {
  int32_t tag;
  int32_t some;
  union { int64_t i; uint64_t u; double f; } v; /* BOOL/Ixx in .i, Uxx in .u, Fxx in .f */
  char *s;                  /* ORC_UTF8 payload (owned, malloc) */
  struct orc_value *items;  /* ORC_STRUCT payload (owned) */
  int32_t n_items;
  int32_t _pad;
} orc_value;

/* Arrow-style primitive array.  BOOL is one byte per element, UTF8 is char*[].
 * valid == NULL means "no nulls"; otherwise one byte per element (1 = valid). */
typedef struct orc_array {
  int32_t dtype;
  int32_t owned;   /* 1: data/valid were malloc'ed by the oracle */
  int64_t len;
  void *data;
  uint8_t *valid;
} orc_array;

/* DataColumnarValue (data_columnar_value.rs:8-13) */
typedef struct orc_columnar {
  int32_t is_scalar;
  int32_t _pad;
  orc_value scalar;
  orc_array array;
} orc_columnar;

/* A block: named columns of equal length (datablocks/data_block.rs:10-62) */
#define ORC_MAX_COLS 16
typedef struct orc_block {
  int32_t n_cols;
  int32_t _pad;
  const char *names[ORC_MAX_COLS];
  orc_array cols[ORC_MAX_COLS];
} orc_block;

#define ORC_ERRLEN 512

/* ---- memory ---- */
void orc_value_free(orc_value *v);
void orc_array_free(orc_array *a);
void orc_columnar_free(orc_columnar *c);
void orc_block_free(orc_block *b);
void orc_free(void *p);

/* ---- source: datasources/system/numbers_table.rs:29-55, numbers_stream.rs:27-62 ---- */
int32_t orc_generate_parts(uint64_t total, uint64_t *begins, uint64_t *ends); /* returns n parts (1 or 8) */
/* returns the number of 10 000-row BlockRanges of [begin,end]; fills up to cap of them.
 * tail_quirk = 1 reproduces numbers_stream.rs:44-46 (drops rows), 0 = no rows dropped. */
int64_t orc_block_ranges(uint64_t begin, uint64_t end, uint64_t block_size, int32_t tail_quirk,
                         uint64_t *b, uint64_t *e, int64_t cap);

/* ---- datavalues ---- */
int32_t orc_numerical_coercion(const char *op, int32_t l, int32_t r, int32_t *out, char *err);
int32_t orc_array_arithmetic(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err);
int32_t orc_array_comparison(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err);
int32_t orc_array_logic(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err);
int32_t orc_array_aggregate(int32_t op, const orc_array *a, orc_value *out, char *err);
int32_t orc_value_arithmetic(int32_t op, const orc_value *l, const orc_value *r, orc_value *out, char *err);
int32_t orc_value_aggregate(int32_t op, const orc_value *l, const orc_value *r, orc_value *out, char *err);
int32_t orc_value_to_array(const orc_value *v, int64_t n, orc_array *out, char *err);
char *orc_value_to_json(const orc_value *v);                 /* serde_json of DataValue */
int32_t orc_value_from_json(const char *json, orc_value *out, char *err);
char *orc_value_display(const orc_value *v);                 /* data_value.rs:200-239 */
const char *orc_dtype_name(int32_t dtype);                   /* arrow DataType Debug */

/* ---- functions (enum Function, functions/function.rs:16-146) ----
 * Built from an s-expression restating planners::ExpressionPlan:
 *   (col name) (u64 1) (i64 -3) (i8 1) (f64 2.5) (str xx) (+ a b) (- a b) (* a b) (/ a b)
 *   (= a b) (< a b) (<= a b) (> a b) (>= a b) (and a b) (or a b)
 *   (sum a) (min a) (max a) (count a) (fn NAME a...) (alias NAME a)
 * orc_fn_parse performs ExpressionPlan::to_function (plan_expression.rs:40-75). */
typedef struct orc_fn orc_fn;
orc_fn *orc_fn_parse(const char *sexpr, char *err);
orc_fn *orc_fn_clone(const orc_fn *f);
void orc_fn_free(orc_fn *f);
char *orc_fn_display(const orc_fn *f);         /* Debug of Function: column names */
char *orc_plan_display(const char *sexpr, char *err); /* Debug of ExpressionPlan */
int32_t orc_fn_is_aggregate(const orc_fn *f);  /* plan_expression.rs:77-89 */
void orc_fn_set_depth(orc_fn *f, uint64_t depth);
int32_t orc_fn_return_type(const orc_fn *f, const orc_block *schema_of, int32_t *out, char *err);
int32_t orc_fn_eval(orc_fn *f, const orc_block *block, orc_columnar *out, char *err);
int32_t orc_fn_accumulate(orc_fn *f, const orc_block *block, char *err);
int32_t orc_fn_accumulate_result(const orc_fn *f, orc_value *out_struct, char *err);
int32_t orc_fn_merge_state(orc_fn *f, const orc_value *states_struct, char *err);
int32_t orc_fn_merge_result(const orc_fn *f, orc_value *out, char *err);

/* ---- pipeline (processors/pipeline_builder.rs:26-106 + transforms) ---- */
typedef struct orc_query {
  /* source: system.numbers_mt(total) when table == NULL, else an in-memory table whose
   * rows are partitioned exactly like numbers_mt's (generate_parts over row indices) */
  uint64_t total;
  const orc_block *table;
  uint64_t block_size;        /* 10000 in the reference (numbers_stream.rs:29) */
  int32_t tail_quirk;         /* 1 = reproduce numbers_stream.rs:44-46 */
  int32_t worker_threads;     /* FuseQueryContext.worker_threads (pipeline_builder.rs:75-84) */
  int32_t use_threads;        /* 1 = one pthread per source pipe (MergeProcessor fan-in) */
  int32_t fused;              /* 0 = reference-shaped passes; 1 = best-case fused CPU scan (numbers_mt u64 only) */
  const char *predicate;      /* s-expr or NULL (already alias-rewritten, optimizer_filter_push_down.rs) */
  int32_t n_exprs;
  int32_t is_aggregate;       /* AggregatePlan vs ProjectionPlan */
  const char *const *exprs;   /* s-exprs */
  int64_t limit;              /* -1 = none */
} orc_query;

typedef struct orc_result {
  orc_block block;            /* concatenated output rows, pipe order */
  int64_t n_rows;
  int64_t n_blocks_out;       /* blocks (incl. empty) reaching the sink */
  int64_t rows_scanned;       /* rows materialised by the sources */
  char *partial_states_json;  /* aggregate queries: '\n'-joined JSON rows of every partial block */
  double seconds;             /* wall time of the execute phase */
} orc_result;

int32_t orc_query_run(const orc_query *q, orc_result *out, char *err);
void orc_result_free(orc_result *r);

/* Best-case CPU figure for bench.py, NOT the reference's structure: the headline query
 * sum(number)/count(number), max(number), min(number) over numbers [0, total) as ONE fused pass per thread, numbers
 * generated in registers, no blocks, no arrays, no per-node passes.  out[0..3] = wrapping sum, count, max, min.
 * Returns the wall seconds of the pass. */
double orc_fused_headline(uint64_t total, int32_t threads, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
