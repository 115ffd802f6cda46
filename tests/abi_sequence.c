/* abi_sequence.c — the call sequence the Rust safe wrapper (ffi/fuse-gpu) makes, in plain C against include/fuse_gpu.h.
 * No Python, no C++: this is what a maintainer binding the library from another language links and runs first.
 *   gcc -std=c11 -O1 -I include tests/abi_sequence.c -L fuse_query_b200 -l:libfuse_gpu.so -Wl,-rpath,$PWD/fuse_query_b200 -o abi_sequence
 * exit 0 = every step checked, 77 = no CUDA device (the library loaded and said so), anything else = failure. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fuse_gpu.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    fq_status st_ = (call);                                                                           \
    if (st_ != FQ_OK) { fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, st_, fq_last_error(ctx)); return 1; } \
  } while (0)
#define EXPECT(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: expectation failed: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

static int push(fq_expr_node *nodes, int *n, int kind, int op, int left, int right, int column, fq_dtype dtype, uint64_t u) {
  fq_expr_node x;
  memset(&x, 0, sizeof x);
  x.kind = kind; x.op = op; x.left = left; x.right = right; x.column = column; x.dtype = dtype; x.value.u = u;
  nodes[*n] = x;
  return (*n)++;
}

int main(void) {
  fq_ctx *ctx = NULL;
  EXPECT(fq_abi_version() == FQ_ABI_VERSION);
  fq_status st = fq_ctx_create(0, &ctx);
  if (st == FQ_ERR_CUDA) { printf("no CUDA device: %s\n", fq_last_error(NULL)); return 77; }
  EXPECT(st == FQ_OK && ctx != NULL);
  EXPECT(fq_ctx_sm_count(ctx) > 0);

  /* Source: system.numbers_mt(1 600 000) as one resident shard (numbers_stream.rs:68-83) */
  const uint64_t n = 1600000;
  fq_column *col = NULL;
  CHECK(fq_column_alloc(ctx, FQ_U64, n, &col));
  CHECK(fq_numbers_fill(ctx, col, 0, 0, n, NULL));
  const fq_column *cols[1] = {col};
  fq_source src = {n, 1, 0, cols, 0};

  /* AggregatePartial: sum(number) / count(number), max(number), min(number) */
  fq_expr_node nodes[32];
  int nn = 0;
  fq_pipe_desc d;
  memset(&d, 0, sizeof d);
  d.n_cols = 1; d.col_dtypes[0] = FQ_U64; d.predicate = -1; d.kind = FQ_PIPE_AGGREGATE;
  int f = push(nodes, &nn, FQ_EXPR_FIELD, 0, -1, -1, 0, FQ_NULL, 0);
  int sum = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_SUM, f, -1, 0, FQ_NULL, 0);
  int cnt = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_COUNT, f, -1, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_DIV, sum, cnt, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_MAX, f, -1, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_MIN, f, -1, 0, FQ_NULL, 0);
  d.nodes = nodes; d.n_nodes = nn;
  fq_pipe *agg = NULL;
  CHECK(fq_pipe_compile(ctx, &d, &agg));
  fq_dtype t;
  CHECK(fq_pipe_expr_dtype(ctx, agg, 0, &t));
  EXPECT(t == FQ_U64);
  CHECK(fq_pipe_launch_aggregate(ctx, agg, &src, 0, NULL));
  fq_value states[8];
  int32_t n_states = 0;
  uint64_t rows = 0;
  CHECK(fq_pipe_fetch_aggregate(ctx, agg, states, 8, &n_states, &rows));
  EXPECT(n_states == 4 && rows == n);
  EXPECT(states[0].some && states[0].v.u == n * (n - 1) / 2 && states[1].v.u == n && states[2].v.u == n - 1 && states[3].v.u == 0);

  /* the merge point for a group of one rank: the launch leaves the merged state on the device */
  fq_group *grp = NULL;
  CHECK(fq_group_create(ctx, 0, 1, 4096, &grp));
  CHECK(fq_pipe_set_group(ctx, agg, grp));
  CHECK(fq_pipe_launch_aggregate(ctx, agg, &src, 0, NULL));
  CHECK(fq_pipe_fetch_merged(ctx, agg, states, 8, &n_states, &rows));
  EXPECT(rows == n && states[0].v.u == n * (n - 1) / 2);
  CHECK(fq_pipe_set_group(ctx, agg, NULL));

  /* Filter -> Projection -> Limit: (number+1) AS c1, number/2 AS c2 WHERE (c1+c2+1) < 100 LIMIT 3 (README.md:120-126) */
  nn = 0;
  memset(&d, 0, sizeof d);
  d.n_cols = 1; d.col_dtypes[0] = FQ_U64; d.kind = FQ_PIPE_PROJECT;
  f = push(nodes, &nn, FQ_EXPR_FIELD, 0, -1, -1, 0, FQ_NULL, 0);
  int one = push(nodes, &nn, FQ_EXPR_CONSTANT, 0, -1, -1, 0, FQ_U64, 1);
  int two = push(nodes, &nn, FQ_EXPR_CONSTANT, 0, -1, -1, 0, FQ_U64, 2);
  int hundred = push(nodes, &nn, FQ_EXPR_CONSTANT, 0, -1, -1, 0, FQ_U64, 100);
  int c1 = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_ADD, f, one, 0, FQ_NULL, 0);
  int c2 = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_DIV, f, two, 0, FQ_NULL, 0);
  int s1 = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_ADD, c1, c2, 0, FQ_NULL, 0);
  int s2 = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_ADD, s1, one, 0, FQ_NULL, 0);
  d.predicate = push(nodes, &nn, FQ_EXPR_COMPARISON, FQ_CMP_LT, s2, hundred, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = c1;
  d.exprs[d.n_exprs++] = c2;
  d.nodes = nodes; d.n_nodes = nn;
  fq_pipe *sel = NULL;
  CHECK(fq_pipe_compile(ctx, &d, &sel));
  fq_column *outs[2] = {NULL, NULL};
  CHECK(fq_column_alloc(ctx, FQ_U64, 3, &outs[0]));
  CHECK(fq_column_alloc(ctx, FQ_U64, 3, &outs[1]));
  CHECK(fq_pipe_launch_project(ctx, sel, &src, outs, NULL, 3, 3, 0, NULL));
  uint64_t selected = 0, written = 0, limit_row = 0;
  CHECK(fq_pipe_fetch_project(ctx, sel, &selected, &written));
  EXPECT(selected == 66 && written == 3);
  CHECK(fq_pipe_fetch_limit_row(ctx, sel, &limit_row));
  EXPECT(limit_row == 2);
  uint64_t a[3], b[3];
  CHECK(fq_column_download(ctx, outs[0], 0, a, 3, NULL));
  CHECK(fq_column_download(ctx, outs[1], 0, b, 3, NULL));
  CHECK(fq_stream_synchronize(ctx, NULL));
  EXPECT(a[0] == 1 && a[1] == 2 && a[2] == 3 && b[0] == 0 && b[1] == 0 && b[2] == 1);

  /* GROUP BY number / 400000: sum(number), count(number) — 4 groups of 400 000 rows */
  nn = 0;
  memset(&d, 0, sizeof d);
  d.n_cols = 1; d.col_dtypes[0] = FQ_U64; d.predicate = -1; d.kind = FQ_PIPE_GROUPBY;
  f = push(nodes, &nn, FQ_EXPR_FIELD, 0, -1, -1, 0, FQ_NULL, 0);
  int k = push(nodes, &nn, FQ_EXPR_CONSTANT, 0, -1, -1, 0, FQ_U64, 400000);
  d.keys[d.n_keys++] = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_DIV, f, k, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_SUM, f, -1, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_COUNT, f, -1, 0, FQ_NULL, 0);
  d.nodes = nodes; d.n_nodes = nn;
  fq_pipe *gb = NULL;
  CHECK(fq_pipe_compile(ctx, &d, &gb));
  CHECK(fq_pipe_groupby_reserve(ctx, gb, 16));
  CHECK(fq_pipe_launch_groupby(ctx, gb, &src, 0, NULL));
  uint64_t groups = 0;
  CHECK(fq_pipe_fetch_groupby(ctx, gb, &groups));
  EXPECT(groups == 4);
  fq_column *kc[1] = {NULL}, *lc[2] = {NULL, NULL};
  CHECK(fq_column_alloc(ctx, FQ_U64, groups, &kc[0]));
  CHECK(fq_column_alloc(ctx, FQ_U64, groups, &lc[0]));
  CHECK(fq_column_alloc(ctx, FQ_U64, groups, &lc[1]));
  CHECK(fq_pipe_export_groups(ctx, gb, kc, NULL, lc, NULL, groups, NULL));
  uint64_t gk[4], gs[4], gc[4];
  CHECK(fq_column_download(ctx, kc[0], 0, gk, 4, NULL));
  CHECK(fq_column_download(ctx, lc[0], 0, gs, 4, NULL));
  CHECK(fq_column_download(ctx, lc[1], 0, gc, 4, NULL));
  CHECK(fq_stream_synchronize(ctx, NULL));
  for (int i = 0; i < 4; i++) {
    const uint64_t lo = gk[i] * 400000, hi = lo + 399999;
    EXPECT(gk[i] < 4 && gc[i] == 400000 && gs[i] == (lo + hi) * 400000 / 2);
  }

  /* an error crosses the ABI as a status + message, never as an abort: zero divisor -> arrow's DivideByZero text */
  nn = 0;
  memset(&d, 0, sizeof d);
  d.n_cols = 1; d.col_dtypes[0] = FQ_U64; d.predicate = -1; d.kind = FQ_PIPE_AGGREGATE;
  f = push(nodes, &nn, FQ_EXPR_FIELD, 0, -1, -1, 0, FQ_NULL, 0);
  int ten = push(nodes, &nn, FQ_EXPR_CONSTANT, 0, -1, -1, 0, FQ_U64, 10);
  int q = push(nodes, &nn, FQ_EXPR_ARITHMETIC, FQ_AR_DIV, ten, f, 0, FQ_NULL, 0);
  d.exprs[d.n_exprs++] = push(nodes, &nn, FQ_EXPR_AGGREGATOR, FQ_AGG_SUM, q, -1, 0, FQ_NULL, 0);
  d.nodes = nodes; d.n_nodes = nn;
  fq_pipe *bad = NULL;
  CHECK(fq_pipe_compile(ctx, &d, &bad));
  CHECK(fq_pipe_launch_aggregate(ctx, bad, &src, 0, NULL));
  st = fq_pipe_fetch_aggregate(ctx, bad, states, 8, &n_states, &rows);
  EXPECT(st == FQ_ERR_DIVIDE_BY_ZERO && strcmp(fq_last_error(ctx), "Internal Error: Divide by zero error") == 0);

  /* a prepared statement: the select and the GROUP BY recorded once on a stream of the library's, replayed twice */
  void *stream = NULL;
  fq_graph *graph = NULL;
  CHECK(fq_stream_create(ctx, &stream));
  CHECK(fq_graph_begin(ctx, stream));
  CHECK(fq_pipe_launch_project(ctx, sel, &src, outs, NULL, 3, 3, FQ_RUN_LIMIT_EARLY_EXIT, stream));
  CHECK(fq_pipe_launch_groupby(ctx, gb, &src, 0, stream));
  CHECK(fq_graph_end(ctx, stream, &graph));
  for (int rep = 0; rep < 2; rep++) {
    CHECK(fq_graph_launch(ctx, graph, stream));
    CHECK(fq_pipe_fetch_project(ctx, sel, &selected, &written));
    CHECK(fq_pipe_fetch_groupby(ctx, gb, &groups));
    EXPECT(written == 3 && groups == 4);
    CHECK(fq_column_download(ctx, outs[0], 0, a, 3, stream));
    CHECK(fq_stream_synchronize(ctx, stream));
    EXPECT(a[0] == 1 && a[1] == 2 && a[2] == 3);
  }
  fq_graph_destroy(ctx, graph);

  /* ORDER BY number DESC LIMIT 3 over the first million rows: permutation, gather, copy (fq_sort_indices / fq_column_take / fq_column_copy) */
  const uint64_t m = 1000000;
  fq_column *head = NULL, *perm = NULL, *sorted = NULL, *top = NULL;
  CHECK(fq_column_slice(ctx, col, 0, m, &head));
  CHECK(fq_column_alloc(ctx, FQ_U32, m, &perm));
  CHECK(fq_column_alloc(ctx, FQ_U64, m, &sorted));
  CHECK(fq_column_alloc(ctx, FQ_U64, 3, &top));
  const fq_column *keys[1] = {head};
  const uint8_t desc[1] = {1};
  CHECK(fq_sort_indices(ctx, keys, desc, 1, m, perm, stream));
  CHECK(fq_column_take(ctx, head, perm, m, sorted, NULL, stream));
  CHECK(fq_column_copy(ctx, top, 0, sorted, 0, 3, stream));
  CHECK(fq_column_download(ctx, top, 0, a, 3, stream));
  CHECK(fq_stream_synchronize(ctx, stream));
  EXPECT(a[0] == m - 1 && a[1] == m - 2 && a[2] == m - 3);
  CHECK(fq_ctx_trim(ctx));
  fq_column_free(ctx, top); fq_column_free(ctx, sorted); fq_column_free(ctx, perm); fq_column_free(ctx, head);
  fq_stream_destroy(ctx, stream);

  fq_pipe_destroy(ctx, bad);
  fq_pipe_destroy(ctx, gb);
  fq_pipe_destroy(ctx, sel);
  fq_pipe_destroy(ctx, agg);
  fq_group_destroy(ctx, grp);
  fq_column_free(ctx, kc[0]); fq_column_free(ctx, lc[0]); fq_column_free(ctx, lc[1]);
  fq_column_free(ctx, outs[0]); fq_column_free(ctx, outs[1]);
  fq_column_free(ctx, col);
  printf("abi sequence ok: %" PRIu64 " launches\n", fq_ctx_launch_count(ctx));
  fq_ctx_destroy(ctx);
  return 0;
}
