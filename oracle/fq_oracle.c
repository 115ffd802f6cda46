/*
 * fq_oracle.c — CPU restatement of the fuse-query hot path.  See fq_oracle.h.
 * TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline); never linked into the product.
 * Citations are file:line under /root/reference/src.
 */
#define _GNU_SOURCE
#include "fq_oracle.h"

#include <ctype.h>
#include <inttypes.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * small utilities
 * ---------------------------------------------------------------------------------------- */
static int fail(char *err, const char *fmt, ...) {
  if (err) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, ORC_ERRLEN, fmt, ap);
    va_end(ap);
  }
  return 1;
}
/* FuseQueryError::Internal Display, error.rs:18-19 */
#define INTERNAL "Internal Error: "
#define PLANERR "Error during plan: "

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "fq_oracle: out of memory\n"); abort(); }
  return p;
}
static char *xstrdup(const char *s) {
  size_t n = strlen(s);
  char *p = xmalloc(n + 1);
  memcpy(p, s, n + 1);
  return p;
}
void orc_free(void *p) { free(p); }

/* growable string */
typedef struct { char *p; size_t n, cap; } sb;
static void sb_put(sb *b, const char *s) {
  size_t l = strlen(s);
  if (b->n + l + 1 > b->cap) {
    b->cap = (b->n + l + 1) * 2 + 32;
    b->p = realloc(b->p, b->cap);
  }
  memcpy(b->p + b->n, s, l + 1);
  b->n += l;
}
static void sb_printf(sb *b, const char *fmt, ...) {
  char tmp[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tmp, sizeof tmp, fmt, ap);
  va_end(ap);
  sb_put(b, tmp);
}
static char *sb_take(sb *b) {
  if (!b->p) return xstrdup("");
  return b->p;
}

/* arrow DataType Debug names */
const char *orc_dtype_name(int32_t t) {
  static const char *names[] = {"Null", "Boolean", "Int8", "Int16", "Int32", "Int64", "UInt8",
                                "UInt16", "UInt32", "UInt64", "Float32", "Float64", "Utf8", "Struct"};
  return (t >= 0 && t <= ORC_STRUCT) ? names[t] : "?";
}
static size_t elem_size(int32_t t) {
  switch (t) {
    case ORC_BOOL: case ORC_I8: case ORC_U8: return 1;
    case ORC_I16: case ORC_U16: return 2;
    case ORC_I32: case ORC_U32: case ORC_F32: return 4;
    case ORC_I64: case ORC_U64: case ORC_F64: return 8;
    case ORC_UTF8: return sizeof(char *);
    default: return 0;
  }
}
static int is_signed_int(int32_t t) { return t >= ORC_I8 && t <= ORC_I64; }
static int is_float(int32_t t) { return t == ORC_F32 || t == ORC_F64; }
/* data_type.rs:9-25 (Float16 has no array type anywhere in the reference; omitted) */
static int is_numeric(int32_t t) { return t >= ORC_I8 && t <= ORC_F64; }

#define NUMERIC_CASES(X)                                                                       \
  X(ORC_I8, int8_t) X(ORC_I16, int16_t) X(ORC_I32, int32_t) X(ORC_I64, int64_t)                \
  X(ORC_U8, uint8_t) X(ORC_U16, uint16_t) X(ORC_U32, uint32_t) X(ORC_U64, uint64_t)            \
  X(ORC_F32, float) X(ORC_F64, double)

/* ------------------------------------------------------------------------------------------
 * values and arrays
 * ---------------------------------------------------------------------------------------- */
static orc_value val_null(void) { orc_value v; memset(&v, 0, sizeof v); return v; }
static orc_value val_none(int32_t tag) { orc_value v = val_null(); v.tag = tag; return v; }
static orc_value val_u64(uint64_t x) { orc_value v = val_none(ORC_U64); v.some = 1; v.v.u = x; return v; }

static orc_value val_clone(const orc_value *s) {
  orc_value v = *s;
  if (s->s) v.s = xstrdup(s->s);
  if (s->items) {
    v.items = xmalloc(sizeof(orc_value) * (size_t)(s->n_items ? s->n_items : 1));
    for (int i = 0; i < s->n_items; i++) v.items[i] = val_clone(&s->items[i]);
  }
  return v;
}
void orc_value_free(orc_value *v) {
  if (!v) return;
  free(v->s);
  if (v->items) {
    for (int i = 0; i < v->n_items; i++) orc_value_free(&v->items[i]);
    free(v->items);
  }
  memset(v, 0, sizeof *v);
}
void orc_array_free(orc_array *a) {
  if (!a) return;
  if (a->owned) {
    if (a->dtype == ORC_UTF8 && a->data) {
      char **s = (char **)a->data;
      for (int64_t i = 0; i < a->len; i++) free(s[i]);
    }
    free(a->data);
    free(a->valid);
  }
  memset(a, 0, sizeof *a);
}
void orc_columnar_free(orc_columnar *c) {
  if (!c) return;
  orc_value_free(&c->scalar);
  orc_array_free(&c->array);
}
void orc_block_free(orc_block *b) {
  if (!b) return;
  for (int i = 0; i < b->n_cols; i++) {
    orc_array_free(&b->cols[i]);
    free((void *)b->names[i]);
  }
  memset(b, 0, sizeof *b);
}
static orc_array arr_alloc(int32_t dtype, int64_t len) {
  orc_array a;
  memset(&a, 0, sizeof a);
  a.dtype = dtype;
  a.len = len;
  a.owned = 1;
  a.data = xmalloc(elem_size(dtype) * (size_t)len);
  if (dtype == ORC_UTF8) memset(a.data, 0, elem_size(dtype) * (size_t)len);
  return a;
}
static orc_array arr_borrow(const orc_array *s) {
  orc_array a = *s;
  a.owned = 0;
  return a;
}
static orc_array arr_clone(const orc_array *s) {
  orc_array a = arr_alloc(s->dtype, s->len);
  if (s->dtype == ORC_UTF8) {
    for (int64_t i = 0; i < s->len; i++) {
      const char *p = ((char **)s->data)[i];
      ((char **)a.data)[i] = p ? xstrdup(p) : NULL;
    }
  } else if (s->len) {
    memcpy(a.data, s->data, elem_size(s->dtype) * (size_t)s->len);
  }
  if (s->valid) {
    a.valid = xmalloc((size_t)s->len);
    memcpy(a.valid, s->valid, (size_t)s->len);
  }
  return a;
}
static int64_t arr_null_count(const orc_array *a) {
  if (!a->valid) return a->dtype == ORC_NULL ? a->len : 0;
  int64_t n = 0;
  for (int64_t i = 0; i < a->len; i++) n += !a->valid[i];
  return n;
}
/* combine_option_bitmap: result slot valid iff valid in both inputs */
static uint8_t *combine_valid(const orc_array *l, const orc_array *r) {
  if (!l->valid && !r->valid) return NULL;
  uint8_t *v = xmalloc((size_t)l->len);
  for (int64_t i = 0; i < l->len; i++)
    v[i] = (uint8_t)((l->valid ? l->valid[i] : 1) & (r->valid ? r->valid[i] : 1));
  return v;
}

/* DataValue Display / Debug, data_value.rs:200-239 (+ macros.rs:201-208): Some(x) -> "{}" of x,
 * None -> "NULL", Null -> "Null".  Rust Display of floats prints 1.0 as "1". */
static void fmt_float(sb *b, double f, int is32) {
  if (isnan(f)) { sb_put(b, "NaN"); return; }
  if (isinf(f)) { sb_put(b, f < 0 ? "-inf" : "inf"); return; }
  char tmp[64];
  /* shortest round-trip representation, then strip to Rust's Display form */
  for (int prec = 1; prec <= 17; prec++) {
    snprintf(tmp, sizeof tmp, "%.*g", prec, f);
    double back = is32 ? (double)strtof(tmp, NULL) : strtod(tmp, NULL);
    if (back == f) break;
  }
  if (strchr(tmp, 'e')) {
    /* Rust never prints exponents in Display; expand */
    snprintf(tmp, sizeof tmp, "%.0f", f);
  }
  sb_put(b, tmp);
}
static void value_display(sb *b, const orc_value *v) {
  if (v->tag == ORC_NULL) { sb_put(b, "Null"); return; }
  if (v->tag == ORC_STRUCT) {
    sb_put(b, "[");
    for (int i = 0; i < v->n_items; i++) {
      if (i) sb_put(b, ", ");
      value_display(b, &v->items[i]);
    }
    sb_put(b, "]");
    return;
  }
  if (!v->some) { sb_put(b, "NULL"); return; }
  switch (v->tag) {
    case ORC_BOOL: sb_put(b, v->v.i ? "true" : "false"); break;
    case ORC_I8: case ORC_I16: case ORC_I32: case ORC_I64: sb_printf(b, "%" PRId64, v->v.i); break;
    case ORC_U8: case ORC_U16: case ORC_U32: case ORC_U64: sb_printf(b, "%" PRIu64, v->v.u); break;
    case ORC_F32: fmt_float(b, v->v.f, 1); break;
    case ORC_F64: fmt_float(b, v->v.f, 0); break;
    case ORC_UTF8: sb_put(b, v->s); break;
    default: sb_put(b, "?");
  }
}
char *orc_value_display(const orc_value *v) {
  sb b = {0};
  value_display(&b, v);
  return sb_take(&b);
}

/* read element i of a numeric/bool array into a DataValue (typed_cast_from_array_to_data_value,
 * macros.rs:219-229) */
static orc_value arr_get(const orc_array *a, int64_t i) {
  orc_value v = val_none(a->dtype);
  if (a->valid && !a->valid[i]) return v;
  v.some = 1;
  switch (a->dtype) {
    case ORC_BOOL: v.v.i = ((uint8_t *)a->data)[i] != 0; break;
    case ORC_I8: v.v.i = ((int8_t *)a->data)[i]; break;
    case ORC_I16: v.v.i = ((int16_t *)a->data)[i]; break;
    case ORC_I32: v.v.i = ((int32_t *)a->data)[i]; break;
    case ORC_I64: v.v.i = ((int64_t *)a->data)[i]; break;
    case ORC_U8: v.v.u = ((uint8_t *)a->data)[i]; break;
    case ORC_U16: v.v.u = ((uint16_t *)a->data)[i]; break;
    case ORC_U32: v.v.u = ((uint32_t *)a->data)[i]; break;
    case ORC_U64: v.v.u = ((uint64_t *)a->data)[i]; break;
    case ORC_F32: v.v.f = ((float *)a->data)[i]; break;
    case ORC_F64: v.v.f = ((double *)a->data)[i]; break;
    case ORC_UTF8: {
      const char *s = ((char **)a->data)[i];
      if (s) v.s = xstrdup(s); else v.some = 0;
      break;
    }
    default: v.some = 0;
  }
  return v;
}
/* DataValue::try_from_array, data_value.rs:115-161 */
static int value_try_from_array(const orc_array *a, int64_t i, orc_value *out, char *err) {
  if (a->dtype == ORC_NULL || a->dtype == ORC_STRUCT)
    return fail(err, INTERNAL "Can't create a scalar of array of type \"%s\"", orc_dtype_name(a->dtype));
  *out = arr_get(a, i);
  return 0;
}

/* DataValue::to_array, data_value.rs:76-112 */
int32_t orc_value_to_array(const orc_value *v, int64_t n, orc_array *out, char *err) {
  if (v->tag == ORC_NULL) { /* NullArray::new(size) */
    memset(out, 0, sizeof *out);
    out->dtype = ORC_NULL;
    out->len = n;
    out->owned = 1;
    return 0;
  }
  if (v->tag == ORC_UTF8) { /* String(v): None allowed -> null entries */
    *out = arr_alloc(ORC_UTF8, n);
    if (!v->some) {
      out->valid = xmalloc((size_t)n);
      memset(out->valid, 0, (size_t)n);
    } else {
      for (int64_t i = 0; i < n; i++) ((char **)out->data)[i] = xstrdup(v->s);
    }
    return 0;
  }
  if (v->tag == ORC_STRUCT || !v->some) {
    char *d = orc_value_display(v);
    fail(err, INTERNAL "DataValue to array cannot be NONE %s", d);
    free(d);
    return 1;
  }
  *out = arr_alloc(v->tag, n);
  switch (v->tag) {
    case ORC_BOOL: memset(out->data, v->v.i ? 1 : 0, (size_t)n); break;
#define FILL_I(TAG, T) case TAG: { T *p = out->data; T x = (T)v->v.i; for (int64_t i = 0; i < n; i++) p[i] = x; break; }
#define FILL_U(TAG, T) case TAG: { T *p = out->data; T x = (T)v->v.u; for (int64_t i = 0; i < n; i++) p[i] = x; break; }
#define FILL_F(TAG, T) case TAG: { T *p = out->data; T x = (T)v->v.f; for (int64_t i = 0; i < n; i++) p[i] = x; break; }
    FILL_I(ORC_I8, int8_t) FILL_I(ORC_I16, int16_t) FILL_I(ORC_I32, int32_t) FILL_I(ORC_I64, int64_t)
    FILL_U(ORC_U8, uint8_t) FILL_U(ORC_U16, uint16_t) FILL_U(ORC_U32, uint32_t) FILL_U(ORC_U64, uint64_t)
    FILL_F(ORC_F32, float) FILL_F(ORC_F64, double)
    default: break;
  }
  return 0;
}

/* DataColumnarValue::to_array, data_columnar_value.rs:24-29 */
static int columnar_to_array(const orc_columnar *c, int64_t n, orc_array *out, char *err) {
  if (!c->is_scalar) { *out = arr_borrow(&c->array); return 0; }
  return orc_value_to_array(&c->scalar, n, out, err);
}
static int32_t columnar_dtype(const orc_columnar *c) { return c->is_scalar ? c->scalar.tag : c->array.dtype; }

/* ------------------------------------------------------------------------------------------
 * coercion — datavalues/data_type.rs:27-98
 * ---------------------------------------------------------------------------------------- */
int32_t orc_numerical_coercion(const char *op, int32_t l, int32_t r, int32_t *out, char *err) {
  if (!is_numeric(l) || !is_numeric(r))
    return fail(err, INTERNAL "Unsupported (%s) %s (%s)", orc_dtype_name(l), op, orc_dtype_name(r));
  if (l == r) { *out = l; return 0; }
  /* most informative first: F64, F32, I64, I32, I16, I8, U64, U32, U16, U8 */
  static const int order[] = {ORC_F64, ORC_F32, ORC_I64, ORC_I32, ORC_I16, ORC_I8, ORC_U64, ORC_U32, ORC_U16, ORC_U8};
  for (size_t k = 0; k < sizeof order / sizeof order[0]; k++)
    if (l == order[k] || r == order[k]) { *out = order[k]; return 0; }
  return fail(err, INTERNAL "Unsupported (%s) %s (%s)", orc_dtype_name(l), op, orc_dtype_name(r));
}
static int equal_coercion(const char *op, int32_t l, int32_t r, int32_t *out, char *err) {
  if (l == r) { *out = l; return 0; }
  return orc_numerical_coercion(op, l, r, out, err);
}

/* arrow::compute::cast, numeric -> numeric: same type is a clone; otherwise num::cast::cast per
 * element, out-of-range -> null (Arrow 2.0.0 cast_numeric_arrays). */
typedef struct { int kind; int64_t i; uint64_t u; double f; } wide; /* kind 0 signed, 1 unsigned, 2 float */
static wide load_wide(const orc_array *a, int64_t k) {
  wide w = {0, 0, 0, 0};
  switch (a->dtype) {
    case ORC_I8: w.i = ((int8_t *)a->data)[k]; break;
    case ORC_I16: w.i = ((int16_t *)a->data)[k]; break;
    case ORC_I32: w.i = ((int32_t *)a->data)[k]; break;
    case ORC_I64: w.i = ((int64_t *)a->data)[k]; break;
    case ORC_U8: w.kind = 1; w.u = ((uint8_t *)a->data)[k]; break;
    case ORC_U16: w.kind = 1; w.u = ((uint16_t *)a->data)[k]; break;
    case ORC_U32: w.kind = 1; w.u = ((uint32_t *)a->data)[k]; break;
    case ORC_U64: w.kind = 1; w.u = ((uint64_t *)a->data)[k]; break;
    case ORC_F32: w.kind = 2; w.f = ((float *)a->data)[k]; break;
    case ORC_F64: w.kind = 2; w.f = ((double *)a->data)[k]; break;
    default: break;
  }
  return w;
}
static int store_cast(orc_array *out, int64_t k, wide w) { /* returns 1 if representable */
  int32_t t = out->dtype;
  if (is_float(t)) {
    double f = w.kind == 2 ? w.f : (w.kind == 1 ? (double)w.u : (double)w.i);
    if (t == ORC_F32) ((float *)out->data)[k] = w.kind == 2 ? (float)w.f : (w.kind == 1 ? (float)w.u : (float)w.i);
    else ((double *)out->data)[k] = f;
    return 1;
  }
  static const int64_t smin[] = {INT8_MIN, INT16_MIN, INT32_MIN, INT64_MIN};
  static const int64_t smax[] = {INT8_MAX, INT16_MAX, INT32_MAX, INT64_MAX};
  static const uint64_t umax[] = {UINT8_MAX, UINT16_MAX, UINT32_MAX, UINT64_MAX};
  if (is_signed_int(t)) {
    int idx = t - ORC_I8;
    int64_t v;
    if (w.kind == 0) { if (w.i < smin[idx] || w.i > smax[idx]) return 0; v = w.i; }
    else if (w.kind == 1) { if (w.u > (uint64_t)smax[idx]) return 0; v = (int64_t)w.u; }
    else {
      if (isnan(w.f)) return 0;
      double tr = trunc(w.f);
      if (idx == 3) { if (!(tr >= -9223372036854775808.0 && tr < 9223372036854775808.0)) return 0; }
      else if (tr < (double)smin[idx] || tr > (double)smax[idx]) return 0;
      v = (int64_t)tr;
    }
    switch (t) {
      case ORC_I8: ((int8_t *)out->data)[k] = (int8_t)v; break;
      case ORC_I16: ((int16_t *)out->data)[k] = (int16_t)v; break;
      case ORC_I32: ((int32_t *)out->data)[k] = (int32_t)v; break;
      default: ((int64_t *)out->data)[k] = v;
    }
    return 1;
  }
  int idx = t - ORC_U8;
  uint64_t v;
  if (w.kind == 0) { if (w.i < 0 || (uint64_t)w.i > umax[idx]) return 0; v = (uint64_t)w.i; }
  else if (w.kind == 1) { if (w.u > umax[idx]) return 0; v = w.u; }
  else {
    if (isnan(w.f)) return 0;
    double tr = trunc(w.f);
    if (idx == 3) { if (!(tr > -1.0 && tr < 18446744073709551616.0)) return 0; }
    else if (tr < 0 || tr > (double)umax[idx]) return 0;
    v = (uint64_t)tr;
  }
  switch (t) {
    case ORC_U8: ((uint8_t *)out->data)[k] = (uint8_t)v; break;
    case ORC_U16: ((uint16_t *)out->data)[k] = (uint16_t)v; break;
    case ORC_U32: ((uint32_t *)out->data)[k] = (uint32_t)v; break;
    default: ((uint64_t *)out->data)[k] = v;
  }
  return 1;
}
static int arr_cast(const orc_array *in, int32_t to, orc_array *out, char *err) {
  if (in->dtype == to) { *out = arr_borrow(in); return 0; } /* same-type cast is a clone */
  if (!is_numeric(in->dtype) || !is_numeric(to))
    return fail(err, INTERNAL "Cast error: Casting from %s to %s not supported", orc_dtype_name(in->dtype), orc_dtype_name(to));
  *out = arr_alloc(to, in->len);
  uint8_t *valid = NULL;
  if (in->valid) { valid = xmalloc((size_t)in->len); memcpy(valid, in->valid, (size_t)in->len); }
  for (int64_t k = 0; k < in->len; k++) {
    if (valid && !valid[k]) { store_cast(out, k, (wide){0, 0, 0, 0}); continue; }
    if (!store_cast(out, k, load_wide(in, k))) {
      if (!valid) { valid = xmalloc((size_t)in->len); memset(valid, 1, (size_t)in->len); }
      valid[k] = 0;
      store_cast(out, k, (wide){0, 0, 0, 0});
    }
  }
  out->valid = valid;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * element-wise arithmetic — datavalues/data_array_arithmetic.rs:14-55, macros.rs:16-53
 * ---------------------------------------------------------------------------------------- */
static const char *arith_sym(int op) { static const char *s[] = {"+", "-", "*", "/"}; return s[op & 3]; }

/* integers wrap (two's complement); evaluated in uint64 then truncated to the lane type */
#define ARITH_INT_LOOP(T, OPSYM)                                                              \
  { const T *a = la.data; const T *b = ra.data; T *o = out->data;                             \
    for (int64_t i = 0; i < n; i++) o[i] = (T)((uint64_t)a[i] OPSYM (uint64_t)b[i]); }
#define ARITH_FLT_LOOP(T, OPSYM)                                                              \
  { const T *a = la.data; const T *b = ra.data; T *o = out->data;                             \
    for (int64_t i = 0; i < n; i++) o[i] = a[i] OPSYM b[i]; }
/* arrow `divide`: DivideByZero on any zero divisor in a valid slot (floats too: Native: Zero) */
#define DIV_LOOP(T, DIVEXPR)                                                                  \
  { const T *a = la.data; const T *b = ra.data; T *o = out->data;                             \
    for (int64_t i = 0; i < n; i++) {                                                         \
      if (valid && !valid[i]) { o[i] = 0; continue; }                                         \
      if (b[i] == 0) { divzero = 1; break; }                                                  \
      o[i] = DIVEXPR; } }

int32_t orc_array_arithmetic(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err) {
  orc_array l0, r0, la, ra;
  memset(&l0, 0, sizeof l0); memset(&r0, 0, sizeof r0); memset(&la, 0, sizeof la); memset(&ra, 0, sizeof ra);
  memset(out, 0, sizeof *out);
  int rc = 1;
  /* scalar broadcast, data_array_arithmetic.rs:19-32 */
  if (!l->is_scalar && !r->is_scalar) { l0 = arr_borrow(&l->array); r0 = arr_borrow(&r->array); }
  else if (!l->is_scalar) { l0 = arr_borrow(&l->array); if (orc_value_to_array(&r->scalar, l->array.len, &r0, err)) goto done; }
  else if (!r->is_scalar) { if (orc_value_to_array(&l->scalar, r->array.len, &l0, err)) goto done; r0 = arr_borrow(&r->array); }
  else { if (orc_value_to_array(&l->scalar, 1, &l0, err)) goto done; if (orc_value_to_array(&r->scalar, 1, &r0, err)) goto done; }

  int32_t t;
  if (orc_numerical_coercion(arith_sym(op), l0.dtype, r0.dtype, &t, err)) goto done;
  if (arr_cast(&l0, t, &la, err)) goto done;
  if (arr_cast(&r0, t, &ra, err)) goto done;
  if (la.len != ra.len) {
    fail(err, INTERNAL "Compute error: Cannot perform math operation on arrays of different length");
    goto done;
  }
  int64_t n = la.len;
  *out = arr_alloc(t, n);
  uint8_t *valid = combine_valid(&la, &ra);
  out->valid = valid;
  int divzero = 0;
  switch (op) {
    case ORC_AR_ADD:
      switch (t) {
        case ORC_I8: ARITH_INT_LOOP(int8_t, +) break; case ORC_I16: ARITH_INT_LOOP(int16_t, +) break;
        case ORC_I32: ARITH_INT_LOOP(int32_t, +) break; case ORC_I64: ARITH_INT_LOOP(int64_t, +) break;
        case ORC_U8: ARITH_INT_LOOP(uint8_t, +) break; case ORC_U16: ARITH_INT_LOOP(uint16_t, +) break;
        case ORC_U32: ARITH_INT_LOOP(uint32_t, +) break; case ORC_U64: ARITH_INT_LOOP(uint64_t, +) break;
        case ORC_F32: ARITH_FLT_LOOP(float, +) break; default: ARITH_FLT_LOOP(double, +) break;
      }
      break;
    case ORC_AR_SUB:
      switch (t) {
        case ORC_I8: ARITH_INT_LOOP(int8_t, -) break; case ORC_I16: ARITH_INT_LOOP(int16_t, -) break;
        case ORC_I32: ARITH_INT_LOOP(int32_t, -) break; case ORC_I64: ARITH_INT_LOOP(int64_t, -) break;
        case ORC_U8: ARITH_INT_LOOP(uint8_t, -) break; case ORC_U16: ARITH_INT_LOOP(uint16_t, -) break;
        case ORC_U32: ARITH_INT_LOOP(uint32_t, -) break; case ORC_U64: ARITH_INT_LOOP(uint64_t, -) break;
        case ORC_F32: ARITH_FLT_LOOP(float, -) break; default: ARITH_FLT_LOOP(double, -) break;
      }
      break;
    case ORC_AR_MUL:
      switch (t) {
        case ORC_I8: ARITH_INT_LOOP(int8_t, *) break; case ORC_I16: ARITH_INT_LOOP(int16_t, *) break;
        case ORC_I32: ARITH_INT_LOOP(int32_t, *) break; case ORC_I64: ARITH_INT_LOOP(int64_t, *) break;
        case ORC_U8: ARITH_INT_LOOP(uint8_t, *) break; case ORC_U16: ARITH_INT_LOOP(uint16_t, *) break;
        case ORC_U32: ARITH_INT_LOOP(uint32_t, *) break; case ORC_U64: ARITH_INT_LOOP(uint64_t, *) break;
        case ORC_F32: ARITH_FLT_LOOP(float, *) break; default: ARITH_FLT_LOOP(double, *) break;
      }
      break;
    default: /* Div; signed MIN / -1 wraps to MIN (Rust would panic; unpinned, off-path) */
      switch (t) {
        case ORC_I8: DIV_LOOP(int8_t, (int8_t)((int32_t)a[i] / (int32_t)b[i])) break;
        case ORC_I16: DIV_LOOP(int16_t, (int16_t)((int32_t)a[i] / (int32_t)b[i])) break;
        case ORC_I32: DIV_LOOP(int32_t, (b[i] == -1 ? (int32_t)(0u - (uint32_t)a[i]) : a[i] / b[i])) break;
        case ORC_I64: DIV_LOOP(int64_t, (b[i] == -1 ? (int64_t)(0ull - (uint64_t)a[i]) : a[i] / b[i])) break;
        case ORC_U8: DIV_LOOP(uint8_t, (uint8_t)(a[i] / b[i])) break;
        case ORC_U16: DIV_LOOP(uint16_t, (uint16_t)(a[i] / b[i])) break;
        case ORC_U32: DIV_LOOP(uint32_t, a[i] / b[i]) break;
        case ORC_U64: DIV_LOOP(uint64_t, a[i] / b[i]) break;
        case ORC_F32: DIV_LOOP(float, a[i] / b[i]) break;
        default: DIV_LOOP(double, a[i] / b[i]) break;
      }
  }
  if (divzero) {
    orc_array_free(out);
    fail(err, INTERNAL "Divide by zero error"); /* ArrowError::DivideByZero via error.rs:24-28 */
    goto done;
  }
  rc = 0;
done:
  orc_array_free(&la); orc_array_free(&ra); orc_array_free(&l0); orc_array_free(&r0);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * comparison — datavalues/data_array_comparison.rs:14-94, macros.rs:55-140
 * ---------------------------------------------------------------------------------------- */
static const char *cmp_sym(int op) { static const char *s[] = {"=", "<", "<=", ">", ">="}; return s[op % 5]; }
static const char *cmp_fn(int op) { static const char *s[] = {"eq", "lt", "lt_eq", "gt", "gt_eq"}; return s[op % 5]; }

#define CMP_APPLY(op, x, y)                                                                   \
  ((op) == ORC_CMP_EQ ? (x) == (y) : (op) == ORC_CMP_LT ? (x) < (y) : (op) == ORC_CMP_LTEQ ? (x) <= (y) \
   : (op) == ORC_CMP_GT ? (x) > (y) : (x) >= (y))

/* array (op) array on equal types */
static int cmp_arrays(int op, const orc_array *la, const orc_array *ra, orc_array *out, char *err) {
  if (la->dtype == ORC_BOOL || la->dtype == ORC_NULL || la->dtype == ORC_STRUCT)
    return fail(err, INTERNAL "Unsupported arithmetic_compute::%s for data type: %s", cmp_fn(op), orc_dtype_name(la->dtype));
  if (la->len != ra->len)
    return fail(err, INTERNAL "Compute error: Cannot perform comparison operation on arrays of different length");
  int64_t n = la->len;
  *out = arr_alloc(ORC_BOOL, n);
  out->valid = combine_valid(la, ra);
  uint8_t *o = out->data;
  switch (la->dtype) {
#define X(TAG, T) case TAG: { const T *a = la->data; const T *b = ra->data; \
    for (int64_t i = 0; i < n; i++) o[i] = (uint8_t)CMP_APPLY(op, a[i], b[i]); break; }
    NUMERIC_CASES(X)
#undef X
    case ORC_UTF8: {
      char **a = la->data; char **b = ra->data;
      for (int64_t i = 0; i < n; i++) {
        int c = strcmp(a[i] ? a[i] : "", b[i] ? b[i] : "");
        o[i] = (uint8_t)CMP_APPLY(op, c, 0);
      }
      break;
    }
    default: break;
  }
  return 0;
}
/* array (op) scalar, arrow `*_scalar` kernels; the scalar must convert to the lane type
 * (macros.rs:85-96 `$RIGHT.try_into()?`, 231-248) */
static int cmp_array_scalar(int op, const orc_array *la, const orc_value *s, orc_array *out, char *err) {
  if (la->dtype == ORC_BOOL || la->dtype == ORC_NULL || la->dtype == ORC_STRUCT)
    return fail(err, INTERNAL "Unsupported data type %s", orc_dtype_name(la->dtype));
  if (la->dtype == ORC_UTF8) {
    if (s->tag != ORC_UTF8 || !s->some) {
      char *d = orc_value_display(s);
      fail(err, INTERNAL "compute_utf8_op_scalar failed to cast literal value %s", d);
      free(d);
      return 1;
    }
  } else if (s->tag != la->dtype || !s->some) {
    char *d = orc_value_display(s);
    fail(err, INTERNAL "Cannot convert %s to %s", d, orc_dtype_name(la->dtype));
    free(d);
    return 1;
  }
  int64_t n = la->len;
  *out = arr_alloc(ORC_BOOL, n);
  if (la->valid) { out->valid = xmalloc((size_t)n); memcpy(out->valid, la->valid, (size_t)n); }
  uint8_t *o = out->data;
  switch (la->dtype) {
#define XI(TAG, T, FIELD) case TAG: { const T *a = la->data; T b = (T)s->v.FIELD; \
    for (int64_t i = 0; i < n; i++) o[i] = (uint8_t)CMP_APPLY(op, a[i], b); break; }
    XI(ORC_I8, int8_t, i) XI(ORC_I16, int16_t, i) XI(ORC_I32, int32_t, i) XI(ORC_I64, int64_t, i)
    XI(ORC_U8, uint8_t, u) XI(ORC_U16, uint16_t, u) XI(ORC_U32, uint32_t, u) XI(ORC_U64, uint64_t, u)
    XI(ORC_F32, float, f) XI(ORC_F64, double, f)
#undef XI
    case ORC_UTF8: {
      char **a = la->data;
      for (int64_t i = 0; i < n; i++) {
        int c = strcmp(a[i] ? a[i] : "", s->s);
        o[i] = (uint8_t)CMP_APPLY(op, c, 0);
      }
      break;
    }
    default: break;
  }
  return 0;
}

int32_t orc_array_comparison(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err) {
  memset(out, 0, sizeof *out);
  if (l->is_scalar && r->is_scalar) /* data_array_comparison.rs:87-92 */
    return fail(err, INTERNAL "Cannot do data_array %s, left:%s, right:%s", cmp_sym(op),
                orc_dtype_name(columnar_dtype(l)), orc_dtype_name(columnar_dtype(r)));
  int32_t t;
  orc_array la, ra;
  memset(&la, 0, sizeof la); memset(&ra, 0, sizeof ra);
  int rc = 1;
  if (!l->is_scalar && !r->is_scalar) { /* :18-40 */
    if (equal_coercion(cmp_sym(op), l->array.dtype, r->array.dtype, &t, err)) return 1;
    if (arr_cast(&l->array, t, &la, err)) goto done;
    if (arr_cast(&r->array, t, &ra, err)) goto done;
    rc = cmp_arrays(op, &la, &ra, out, err);
    goto done;
  }
  /* array-scalar :42-64, scalar-array with the operator flipped :66-85 */
  const orc_array *arr = l->is_scalar ? &r->array : &l->array;
  const orc_value *sc = l->is_scalar ? &l->scalar : &r->scalar;
  if (equal_coercion(cmp_sym(op), arr->dtype, sc->tag, &t, err)) return 1;
  orc_array s1, s1c;
  memset(&s1, 0, sizeof s1); memset(&s1c, 0, sizeof s1c);
  orc_value sv = val_null();
  if (arr_cast(arr, t, &la, err)) goto done2;
  if (orc_value_to_array(sc, 1, &s1, err)) goto done2;
  if (arr_cast(&s1, t, &s1c, err)) goto done2;
  if (value_try_from_array(&s1c, 0, &sv, err)) goto done2;
  int eff = op;
  if (l->is_scalar) {
    static const int flip[] = {ORC_CMP_EQ, ORC_CMP_GT, ORC_CMP_GTEQ, ORC_CMP_LT, ORC_CMP_LTEQ};
    eff = flip[op % 5];
  }
  rc = cmp_array_scalar(eff, &la, &sv, out, err);
done2:
  orc_value_free(&sv);
  orc_array_free(&s1c); orc_array_free(&s1);
done:
  orc_array_free(&la); orc_array_free(&ra);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * logic — datavalues/data_array_logic.rs:10-31
 * ---------------------------------------------------------------------------------------- */
int32_t orc_array_logic(int32_t op, const orc_columnar *l, const orc_columnar *r, orc_array *out, char *err) {
  memset(out, 0, sizeof *out);
  const char *sym = op == ORC_LG_AND ? "and" : "or";
  if (l->is_scalar || r->is_scalar)
    return fail(err, INTERNAL "Cannot do data_array %s, left:%s, right:%s", sym,
                orc_dtype_name(columnar_dtype(l)), orc_dtype_name(columnar_dtype(r)));
  const orc_array *la = &l->array, *ra = &r->array;
  if (la->dtype != ORC_BOOL)
    return fail(err, INTERNAL "Cannot downcast_array from datatype:%s item to:BooleanArray", orc_dtype_name(la->dtype));
  if (ra->dtype != ORC_BOOL)
    return fail(err, INTERNAL "Cannot downcast_array from datatype:%s item to:BooleanArray", orc_dtype_name(ra->dtype));
  if (la->len != ra->len)
    return fail(err, INTERNAL "Compute error: Cannot perform bitwise operation on arrays of different length");
  *out = arr_alloc(ORC_BOOL, la->len);
  out->valid = combine_valid(la, ra);
  const uint8_t *a = la->data, *b = ra->data;
  uint8_t *o = out->data;
  for (int64_t i = 0; i < la->len; i++) o[i] = op == ORC_LG_AND ? (a[i] & b[i]) : (a[i] | b[i]);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * block -> scalar aggregates — datavalues/data_array_aggregate.rs:14-163, macros.rs:142-171
 * ---------------------------------------------------------------------------------------- */
static const char *agg_name(int op) { static const char *s[] = {"min", "max", "sum", "count"}; return s[op & 3]; }

int32_t orc_array_aggregate(int32_t op, const orc_array *a, orc_value *out, char *err) {
  *out = val_null();
  int32_t t = a->dtype;
  if (!(is_numeric(t) || t == ORC_UTF8) || (t == ORC_UTF8 && op == ORC_AGG_SUM))
    return fail(err, INTERNAL "Unsupported data_array_%s for data type: %s", agg_name(op), orc_dtype_name(t));
  if (op == ORC_AGG_COUNT) { *out = val_u64((uint64_t)a->len); return 0; } /* count = len, :113 */
  *out = val_none(t);
  /* arrow sum / min / max: None when every slot is null (incl. len == 0) */
  if (arr_null_count(a) == a->len) return 0;
  int64_t n = a->len;
  const uint8_t *valid = a->valid;
  if (t == ORC_UTF8) { /* min_string / max_string */
    char **s = a->data;
    const char *best = NULL;
    for (int64_t i = 0; i < n; i++) {
      if (valid && !valid[i]) continue;
      if (!best || (op == ORC_AGG_MIN ? strcmp(s[i], best) < 0 : strcmp(s[i], best) > 0)) best = s[i];
    }
    out->some = 1;
    out->s = xstrdup(best);
    return 0;
  }
  out->some = 1;
  switch (t) {
#define AGG_INT(TAG, T, FIELD, WT)                                                            \
    case TAG: { const T *p = a->data;                                                         \
      if (op == ORC_AGG_SUM) { /* wrapping add in the lane type */                            \
        uint64_t acc = 0;                                                                     \
        if (!valid) for (int64_t i = 0; i < n; i++) acc += (uint64_t)p[i];                    \
        else for (int64_t i = 0; i < n; i++) if (valid[i]) acc += (uint64_t)p[i];             \
        out->v.FIELD = (WT)(T)acc;                                                            \
      } else {                                                                                \
        int seen = 0; T best = 0;                                                             \
        for (int64_t i = 0; i < n; i++) {                                                     \
          if (valid && !valid[i]) continue;                                                   \
          if (!seen) { best = p[i]; seen = 1; }                                               \
          else if (op == ORC_AGG_MIN ? p[i] < best : p[i] > best) best = p[i];                \
        }                                                                                     \
        out->v.FIELD = (WT)best;                                                              \
      }                                                                                       \
      break; }
    AGG_INT(ORC_I8, int8_t, i, int64_t) AGG_INT(ORC_I16, int16_t, i, int64_t)
    AGG_INT(ORC_I32, int32_t, i, int64_t) AGG_INT(ORC_I64, int64_t, i, int64_t)
    AGG_INT(ORC_U8, uint8_t, u, uint64_t) AGG_INT(ORC_U16, uint16_t, u, uint64_t)
    AGG_INT(ORC_U32, uint32_t, u, uint64_t) AGG_INT(ORC_U64, uint64_t, u, uint64_t)
#undef AGG_INT
#define AGG_FLT(TAG, T)                                                                       \
    case TAG: { const T *p = a->data;                                                         \
      if (op == ORC_AGG_SUM) { T acc = 0; /* sequential order */                              \
        for (int64_t i = 0; i < n; i++) if (!valid || valid[i]) acc += p[i];                  \
        out->v.f = acc;                                                                       \
      } else { /* min_max_helper: n = first valid; replace when cmp(n, item) */               \
        int seen = 0; T best = 0;                                                             \
        for (int64_t i = 0; i < n; i++) {                                                     \
          if (valid && !valid[i]) continue;                                                   \
          if (!seen) { best = p[i]; seen = 1; }                                               \
          else if (op == ORC_AGG_MIN ? best > p[i] : best < p[i]) best = p[i];                \
        }                                                                                     \
        out->v.f = best;                                                                      \
      }                                                                                       \
      break; }
    AGG_FLT(ORC_F32, float) AGG_FLT(ORC_F64, double)
#undef AGG_FLT
    default: break;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * scalar (+) scalar — datavalues/data_value_arithmetic.rs:10-27
 * ---------------------------------------------------------------------------------------- */
int32_t orc_value_arithmetic(int32_t op, const orc_value *l, const orc_value *r, orc_value *out, char *err) {
  if (l->tag == ORC_NULL) { *out = val_clone(r); return 0; }
  if (r->tag == ORC_NULL) { *out = val_clone(l); return 0; }
  orc_columnar cl, cr;
  memset(&cl, 0, sizeof cl); memset(&cr, 0, sizeof cr);
  cl.is_scalar = cr.is_scalar = 1;
  cl.scalar = *l; cr.scalar = *r; /* borrowed */
  orc_array res;
  if (orc_array_arithmetic(op, &cl, &cr, &res, err)) return 1;
  int rc = value_try_from_array(&res, 0, out, err);
  orc_array_free(&res);
  return rc;
}

/* scalar min/max/sum — datavalues/data_value_aggregate.rs:8-101, macros.rs:173-199 */
int32_t orc_value_aggregate(int32_t op, const orc_value *l, const orc_value *r, orc_value *out, char *err) {
  if (l->tag == ORC_NULL) { *out = val_clone(r); return 0; }
  if (r->tag == ORC_NULL) { *out = val_clone(l); return 0; }
  int ok = l->tag == r->tag && (is_numeric(l->tag) || l->tag == ORC_UTF8);
  if (ok && l->tag == ORC_UTF8 && (op == ORC_AGG_SUM || op == ORC_AGG_COUNT)) ok = 0;
  if (!ok)
    return fail(err, INTERNAL "Unsupported data_value_%s for data type: left:%s, right:%s", agg_name(op),
                orc_dtype_name(l->tag), orc_dtype_name(r->tag));
  if (op == ORC_AGG_COUNT) { *out = val_u64(1); return 0; } /* sic: data_value_aggregate.rs:20 */
  if (!l->some && !r->some) { *out = val_none(l->tag); return 0; }
  if (!r->some) { *out = val_clone(l); return 0; }
  if (!l->some) { *out = val_clone(r); return 0; }
  *out = val_none(l->tag);
  out->some = 1;
  int32_t t = l->tag;
  if (t == ORC_UTF8) {
    int c = strcmp(l->s, r->s);
    out->s = xstrdup((op == ORC_AGG_MIN ? c <= 0 : c >= 0) ? l->s : r->s);
    return 0;
  }
  if (is_float(t)) {
    double a = l->v.f, b = r->v.f;
    /* f64::min/max (NaN-ignoring); sum in the lane type */
    if (op == ORC_AGG_SUM) out->v.f = t == ORC_F32 ? (double)((float)a + (float)b) : a + b;
    else out->v.f = op == ORC_AGG_MIN ? fmin(a, b) : fmax(a, b);
    return 0;
  }
  if (is_signed_int(t)) {
    int64_t a = l->v.i, b = r->v.i;
    if (op == ORC_AGG_SUM) {
      uint64_t s = (uint64_t)a + (uint64_t)b;
      switch (t) { case ORC_I8: out->v.i = (int8_t)s; break; case ORC_I16: out->v.i = (int16_t)s; break;
                   case ORC_I32: out->v.i = (int32_t)s; break; default: out->v.i = (int64_t)s; }
    } else out->v.i = op == ORC_AGG_MIN ? (a < b ? a : b) : (a > b ? a : b);
    return 0;
  }
  uint64_t a = l->v.u, b = r->v.u;
  if (op == ORC_AGG_SUM) {
    uint64_t s = a + b;
    switch (t) { case ORC_U8: out->v.u = (uint8_t)s; break; case ORC_U16: out->v.u = (uint16_t)s; break;
                 case ORC_U32: out->v.u = (uint32_t)s; break; default: out->v.u = s; }
  } else out->v.u = op == ORC_AGG_MIN ? (a < b ? a : b) : (a > b ? a : b);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * serde_json of DataValue (externally tagged enum, data_value.rs:19-35) — the partial-state
 * wire format of transform_aggregate_partial.rs:61-66 / transform_aggregate_final.rs:56-66
 * ---------------------------------------------------------------------------------------- */
static const char *json_tag(int32_t t) {
  static const char *names[] = {"Null", "Boolean", "Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16",
                                "UInt32", "UInt64", "Float32", "Float64", "String", "Struct"};
  return names[t];
}
static void json_float(sb *b, double f, int is32) {
  if (isnan(f) || isinf(f)) { sb_put(b, "null"); return; }
  char tmp[64];
  for (int prec = 1; prec <= 17; prec++) {
    snprintf(tmp, sizeof tmp, "%.*g", prec, f);
    double back = is32 ? (double)strtof(tmp, NULL) : strtod(tmp, NULL);
    if (back == f) break;
  }
  sb_put(b, tmp);
  if (!strpbrk(tmp, ".en")) sb_put(b, ".0"); /* ryu always prints a fraction */
}
static void value_json(sb *b, const orc_value *v) {
  if (v->tag == ORC_NULL) { sb_put(b, "\"Null\""); return; }
  sb_printf(b, "{\"%s\":", json_tag(v->tag));
  if (v->tag == ORC_STRUCT) {
    sb_put(b, "[");
    for (int i = 0; i < v->n_items; i++) { if (i) sb_put(b, ","); value_json(b, &v->items[i]); }
    sb_put(b, "]");
  } else if (!v->some) sb_put(b, "null");
  else switch (v->tag) {
    case ORC_BOOL: sb_put(b, v->v.i ? "true" : "false"); break;
    case ORC_I8: case ORC_I16: case ORC_I32: case ORC_I64: sb_printf(b, "%" PRId64, v->v.i); break;
    case ORC_U8: case ORC_U16: case ORC_U32: case ORC_U64: sb_printf(b, "%" PRIu64, v->v.u); break;
    case ORC_F32: json_float(b, v->v.f, 1); break;
    case ORC_F64: json_float(b, v->v.f, 0); break;
    case ORC_UTF8: {
      sb_put(b, "\"");
      for (const char *p = v->s; *p; p++) {
        if (*p == '"' || *p == '\\') sb_printf(b, "\\%c", *p);
        else if ((unsigned char)*p < 0x20) sb_printf(b, "\\u%04x", *p);
        else { char c[2] = {*p, 0}; sb_put(b, c); }
      }
      sb_put(b, "\"");
      break;
    }
    default: sb_put(b, "null");
  }
  sb_put(b, "}");
}
char *orc_value_to_json(const orc_value *v) {
  sb b = {0};
  value_json(&b, v);
  return sb_take(&b);
}
static void skip_ws(const char **p) { while (**p && isspace((unsigned char)**p)) (*p)++; }
static int json_string(const char **p, char **out, char *err) {
  if (**p != '"') return fail(err, INTERNAL "expected string at `%.20s`", *p);
  (*p)++;
  sb b = {0};
  while (**p && **p != '"') {
    char c = **p;
    if (c == '\\') {
      (*p)++;
      c = **p;
      if (c == 'n') c = '\n'; else if (c == 't') c = '\t';
      else if (c == 'u') { unsigned x = 0; sscanf(*p + 1, "%4x", &x); c = (char)x; *p += 4; }
    }
    char s[2] = {c, 0};
    sb_put(&b, s);
    (*p)++;
  }
  if (**p != '"') { free(b.p); return fail(err, INTERNAL "EOF while parsing a string"); }
  (*p)++;
  *out = sb_take(&b);
  return 0;
}
static int json_value(const char **p, orc_value *out, char *err) {
  *out = val_null();
  skip_ws(p);
  if (**p == '"') { /* unit variant "Null" */
    char *s;
    if (json_string(p, &s, err)) return 1;
    int ok = strcmp(s, "Null") == 0;
    if (!ok) fail(err, INTERNAL "unknown variant `%s`", s);
    free(s);
    return !ok;
  }
  if (**p != '{') return fail(err, INTERNAL "expected value at `%.20s`", *p);
  (*p)++;
  skip_ws(p);
  char *tag;
  if (json_string(p, &tag, err)) return 1;
  int32_t t = -1;
  for (int k = 1; k <= ORC_STRUCT; k++) if (strcmp(tag, json_tag(k)) == 0) t = k;
  if (t < 0) { fail(err, INTERNAL "unknown variant `%s`", tag); free(tag); return 1; }
  free(tag);
  skip_ws(p);
  if (**p != ':') return fail(err, INTERNAL "expected `:`");
  (*p)++;
  skip_ws(p);
  out->tag = t;
  if (t == ORC_STRUCT) {
    if (**p != '[') return fail(err, INTERNAL "expected `[`");
    (*p)++;
    int cap = 4;
    out->items = xmalloc(sizeof(orc_value) * (size_t)cap);
    skip_ws(p);
    while (**p && **p != ']') {
      if (out->n_items == cap) { cap *= 2; out->items = realloc(out->items, sizeof(orc_value) * (size_t)cap); }
      if (json_value(p, &out->items[out->n_items], err)) return 1;
      out->n_items++;
      skip_ws(p);
      if (**p == ',') { (*p)++; skip_ws(p); }
    }
    if (**p != ']') return fail(err, INTERNAL "EOF while parsing a list");
    (*p)++;
  } else if (strncmp(*p, "null", 4) == 0) {
    *p += 4;
  } else {
    out->some = 1;
    if (t == ORC_BOOL) {
      if (strncmp(*p, "true", 4) == 0) { out->v.i = 1; *p += 4; }
      else if (strncmp(*p, "false", 5) == 0) { out->v.i = 0; *p += 5; }
      else return fail(err, INTERNAL "expected a boolean");
    } else if (t == ORC_UTF8) {
      if (json_string(p, &out->s, err)) return 1;
    } else {
      char *end;
      if (is_float(t)) out->v.f = strtod(*p, &end);
      else if (is_signed_int(t)) out->v.i = strtoll(*p, &end, 10);
      else out->v.u = strtoull(*p, &end, 10);
      if (end == *p) return fail(err, INTERNAL "expected a number at `%.20s`", *p);
      *p = end;
    }
  }
  skip_ws(p);
  if (**p != '}') return fail(err, INTERNAL "expected `}`");
  (*p)++;
  return 0;
}
int32_t orc_value_from_json(const char *json, orc_value *out, char *err) {
  const char *p = json;
  if (json_value(&p, out, err)) { orc_value_free(out); return 1; }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * source — numbers_table.rs:29-55 and numbers_stream.rs:27-62
 * ---------------------------------------------------------------------------------------- */
int32_t orc_generate_parts(uint64_t total, uint64_t *begins, uint64_t *ends) {
  const uint64_t workers = 8;
  uint64_t chunk = total / workers;
  if (chunk == 0) { /* total == 0 underflows in the reference (u64 wrap in release) */
    begins[0] = 0;
    ends[0] = total - 1;
    return 1;
  }
  uint64_t remain = total % workers;
  for (uint64_t p = 0; p < workers; p++) {
    begins[p] = p * chunk;
    ends[p] = (p + 1) * chunk - 1;
    if (p == workers - 1 && remain > 0) ends[p] += remain;
  }
  return (int32_t)workers;
}
int64_t orc_block_ranges(uint64_t begin, uint64_t end, uint64_t block_size, int32_t tail_quirk,
                         uint64_t *b, uint64_t *e, int64_t cap) {
  uint64_t count = end - begin + 1;
  uint64_t nblk = count / block_size, remain = count % block_size;
  int64_t k = 0;
  if (nblk == 0) {
    if (k < cap) { b[k] = begin; e[k] = end; }
    return 1;
  }
  for (uint64_t i = 0; i < nblk; i++) {
    uint64_t bb = begin + block_size * i, be = begin + block_size * (i + 1) - 1;
    if (i == nblk - 1 && remain > 0)
      be = tail_quirk ? bb + remain /* numbers_stream.rs:44-46 */ : be + remain;
    if (k < cap) { b[k] = bb; e[k] = be; }
    k++;
  }
  return k;
}

/* ------------------------------------------------------------------------------------------
 * ExpressionPlan (planners/plan_expression.rs:13-29) parsed from an s-expression
 * ---------------------------------------------------------------------------------------- */
enum { P_ALIAS, P_FIELD, P_CONST, P_BINARY, P_FUNCTION, P_WILDCARD };
typedef struct plan {
  int kind;
  char *s; /* alias / field name / operator */
  orc_value v;
  int nargs;
  struct plan *args[8];
} plan;
static void plan_free(plan *p) {
  if (!p) return;
  free(p->s);
  orc_value_free(&p->v);
  for (int i = 0; i < p->nargs; i++) plan_free(p->args[i]);
  free(p);
}
static char *tok(const char **s) {
  while (**s && isspace((unsigned char)**s)) (*s)++;
  const char *b = *s;
  if (**s == '(' || **s == ')') { (*s)++; }
  else while (**s && !isspace((unsigned char)**s) && **s != '(' && **s != ')') (*s)++;
  size_t n = (size_t)(*s - b);
  char *t = xmalloc(n + 1);
  memcpy(t, b, n);
  t[n] = 0;
  return t;
}
static int type_from_name(const char *h) {
  static const struct { const char *n; int t; } m[] = {
      {"bool", ORC_BOOL}, {"i8", ORC_I8}, {"i16", ORC_I16}, {"i32", ORC_I32}, {"i64", ORC_I64},
      {"u8", ORC_U8}, {"u16", ORC_U16}, {"u32", ORC_U32}, {"u64", ORC_U64}, {"f32", ORC_F32},
      {"f64", ORC_F64}, {"str", ORC_UTF8}};
  for (size_t i = 0; i < sizeof m / sizeof m[0]; i++) if (strcmp(h, m[i].n) == 0) return m[i].t;
  return -1;
}
static int is_binary_head(const char *h) {
  static const char *ops[] = {"+", "-", "*", "/", "=", "<", "<=", ">", ">=", "and", "or", "AND", "OR"};
  for (size_t i = 0; i < sizeof ops / sizeof ops[0]; i++) if (strcmp(h, ops[i]) == 0) return 1;
  return 0;
}
static plan *plan_parse(const char **s, char *err) {
  char *t = tok(s);
  if (strcmp(t, "(") != 0) {
    fail(err, PLANERR "expected `(` at `%s`", t);
    free(t);
    return NULL;
  }
  free(t);
  char *h = tok(s);
  plan *p = xmalloc(sizeof *p);
  memset(p, 0, sizeof *p);
  int ty;
  if (strcmp(h, "col") == 0) { p->kind = P_FIELD; p->s = tok(s); }
  else if (strcmp(h, "wildcard") == 0) { p->kind = P_WILDCARD; }
  else if (strcmp(h, "null") == 0) { p->kind = P_CONST; }
  else if ((ty = type_from_name(h)) >= 0) {
    p->kind = P_CONST;
    p->v = val_none(ty);
    char *lit = tok(s);
    if (strcmp(lit, "none") != 0) {
      p->v.some = 1;
      if (ty == ORC_UTF8) p->v.s = xstrdup(lit);
      else if (ty == ORC_BOOL) p->v.v.i = strcmp(lit, "true") == 0;
      else if (is_float(ty)) p->v.v.f = ty == ORC_F32 ? (double)strtof(lit, NULL) : strtod(lit, NULL);
      else if (is_signed_int(ty)) p->v.v.i = strtoll(lit, NULL, 10);
      else p->v.v.u = strtoull(lit, NULL, 10);
    }
    free(lit);
  } else if (strcmp(h, "alias") == 0) {
    p->kind = P_ALIAS;
    p->s = tok(s);
    p->args[0] = plan_parse(s, err);
    p->nargs = 1;
    if (!p->args[0]) { p->nargs = 0; free(h); plan_free(p); return NULL; }
  } else {
    int fn = strcmp(h, "fn") == 0;
    p->kind = (!fn && is_binary_head(h)) ? P_BINARY : P_FUNCTION;
    p->s = fn ? tok(s) : xstrdup(h);
    for (;;) {
      const char *save = *s;
      char *n = tok(&save);
      int close = strcmp(n, ")") == 0, eof = n[0] == 0;
      free(n);
      if (close || eof) break;
      if (p->nargs == 8) { fail(err, PLANERR "too many arguments"); free(h); plan_free(p); return NULL; }
      plan *a = plan_parse(s, err);
      if (!a) { free(h); plan_free(p); return NULL; }
      p->args[p->nargs++] = a;
    }
    if (p->kind == P_BINARY && p->nargs != 2) {
      fail(err, PLANERR "binary operator %s needs 2 arguments", h);
      free(h); plan_free(p);
      return NULL;
    }
  }
  free(h);
  char *c = tok(s);
  if (strcmp(c, ")") != 0) {
    fail(err, PLANERR "expected `)` got `%s`", c);
    free(c); plan_free(p);
    return NULL;
  }
  free(c);
  return p;
}
/* Debug of ExpressionPlan, plan_expression.rs:92-105 */
static void plan_display(sb *b, const plan *p) {
  switch (p->kind) {
    case P_ALIAS: plan_display(b, p->args[0]); sb_printf(b, " as %s", p->s); break;
    case P_FIELD: sb_put(b, p->s); break;
    case P_CONST: value_display(b, &p->v); break;
    case P_BINARY:
      sb_put(b, "("); plan_display(b, p->args[0]); sb_printf(b, " %s ", p->s);
      plan_display(b, p->args[1]); sb_put(b, ")");
      break;
    case P_FUNCTION:
      sb_printf(b, "%s([", p->s);
      for (int i = 0; i < p->nargs; i++) { if (i) sb_put(b, ", "); plan_display(b, p->args[i]); }
      sb_put(b, "])");
      break;
    default: sb_put(b, "*");
  }
}
char *orc_plan_display(const char *sexpr, char *err) {
  const char *s = sexpr;
  plan *p = plan_parse(&s, err);
  if (!p) return NULL;
  sb b = {0};
  plan_display(&b, p);
  plan_free(p);
  return sb_take(&b);
}
/* plan_expression.rs:77-89 */
static int plan_is_aggregate(const plan *p) {
  switch (p->kind) {
    case P_ALIAS: return plan_is_aggregate(p->args[0]);
    case P_BINARY: return plan_is_aggregate(p->args[0]) || plan_is_aggregate(p->args[1]);
    case P_FUNCTION: {
      static const char *aggs[] = {"max", "min", "avg", "count", "sum"};
      for (int i = 0; i < 5; i++) if (strcasecmp(p->s, aggs[i]) == 0) return 1;
      return 0;
    }
    default: return 0;
  }
}

/* ------------------------------------------------------------------------------------------
 * enum Function — functions/function_{aggregator,arithmetic,comparison,logic,field,constant,alias,factory}.rs
 * ---------------------------------------------------------------------------------------- */
enum { FN_ALIAS, FN_CONST, FN_FIELD, FN_ARITH, FN_CMP, FN_LOGIC, FN_AGG };
struct orc_fn {
  int kind, op;
  uint64_t depth;
  char *name;        /* field / alias */
  orc_value value;   /* constant value or aggregator state (starts Null, function_aggregator.rs:30) */
  orc_fn *l, *r;     /* arithmetic/comparison/logic children; aggregator/alias argument in l */
  int is_agg_plan;   /* ExpressionPlan::is_aggregate of the plan this came from */
};
void orc_fn_free(orc_fn *f) {
  if (!f) return;
  free(f->name);
  orc_value_free(&f->value);
  orc_fn_free(f->l);
  orc_fn_free(f->r);
  free(f);
}
orc_fn *orc_fn_clone(const orc_fn *f) {
  if (!f) return NULL;
  orc_fn *c = xmalloc(sizeof *c);
  *c = *f;
  c->name = f->name ? xstrdup(f->name) : NULL;
  c->value = val_clone(&f->value);
  c->l = orc_fn_clone(f->l);
  c->r = orc_fn_clone(f->r);
  return c;
}
static orc_fn *fn_new(int kind, int op) {
  orc_fn *f = xmalloc(sizeof *f);
  memset(f, 0, sizeof *f);
  f->kind = kind;
  f->op = op;
  return f;
}
/* set_depth: function_arithmetic.rs:48-52 recurses, alias/aggregator/comparison/logic/field only
 * record it (function_alias.rs:36-38 etc.), constants ignore it */
void orc_fn_set_depth(orc_fn *f, uint64_t depth) {
  switch (f->kind) {
    case FN_CONST: break;
    case FN_ARITH:
      orc_fn_set_depth(f->l, depth);
      orc_fn_set_depth(f->r, depth + 1);
      f->depth = depth;
      break;
    default: f->depth = depth;
  }
}
/* ScalarFunctionFactory::get, function_factory.rs:17-39 (takes ownership of args) */
static orc_fn *factory_get(const char *name, orc_fn **args, int nargs, char *err) {
  static const struct { const char *n; int kind, op; } tbl[] = {
      {"+", FN_ARITH, ORC_AR_ADD}, {"-", FN_ARITH, ORC_AR_SUB}, {"*", FN_ARITH, ORC_AR_MUL}, {"/", FN_ARITH, ORC_AR_DIV},
      {"=", FN_CMP, ORC_CMP_EQ}, {"<", FN_CMP, ORC_CMP_LT}, {">", FN_CMP, ORC_CMP_GT}, {"<=", FN_CMP, ORC_CMP_LTEQ},
      {">=", FN_CMP, ORC_CMP_GTEQ}, {"and", FN_LOGIC, ORC_LG_AND}, {"or", FN_LOGIC, ORC_LG_OR},
      {"count", FN_AGG, ORC_AGG_COUNT}, {"min", FN_AGG, ORC_AGG_MIN}, {"max", FN_AGG, ORC_AGG_MAX}, {"sum", FN_AGG, ORC_AGG_SUM}};
  for (size_t i = 0; i < sizeof tbl / sizeof tbl[0]; i++) {
    if (strcasecmp(name, tbl[i].n) != 0) continue;
    int need = tbl[i].kind == FN_AGG ? 1 : 2;
    if (nargs < need) { /* the reference would panic on args[k]; report instead */
      fail(err, INTERNAL "Function %s expects %d argument(s)", name, need);
      for (int k = 0; k < nargs; k++) orc_fn_free(args[k]);
      return NULL;
    }
    orc_fn *f = fn_new(tbl[i].kind, tbl[i].op);
    f->l = args[0];
    if (need == 2) f->r = args[1];
    for (int k = need; k < nargs; k++) orc_fn_free(args[k]);
    return f;
  }
  fail(err, INTERNAL "Unsupported Function: %s", name);
  for (int k = 0; k < nargs; k++) orc_fn_free(args[k]);
  return NULL;
}
/* ExpressionPlan::plan_to_function, plan_expression.rs:40-71 */
static orc_fn *plan_to_function(const plan *p, uint64_t depth, char *err) {
  switch (p->kind) {
    case P_FIELD: { orc_fn *f = fn_new(FN_FIELD, 0); f->name = xstrdup(p->s); return f; }
    case P_CONST: { orc_fn *f = fn_new(FN_CONST, 0); f->value = val_clone(&p->v); return f; }
    case P_BINARY: {
      orc_fn *a[2];
      a[0] = plan_to_function(p->args[0], depth, err);
      if (!a[0]) return NULL;
      a[1] = plan_to_function(p->args[1], depth + 1, err);
      if (!a[1]) { orc_fn_free(a[0]); return NULL; }
      orc_fn *f = factory_get(p->s, a, 2, err);
      if (f) orc_fn_set_depth(f, depth);
      return f;
    }
    case P_FUNCTION: {
      orc_fn *a[8];
      for (int i = 0; i < p->nargs; i++) {
        a[i] = plan_to_function(p->args[i], depth + 1, err);
        if (!a[i]) { for (int k = 0; k < i; k++) orc_fn_free(a[k]); return NULL; }
        orc_fn_set_depth(a[i], depth);
      }
      orc_fn *f = factory_get(p->s, a, p->nargs, err);
      if (f) orc_fn_set_depth(f, depth);
      return f;
    }
    case P_ALIAS: {
      orc_fn *in = plan_to_function(p->args[0], depth, err);
      if (!in) return NULL;
      orc_fn_set_depth(in, depth);
      orc_fn *f = fn_new(FN_ALIAS, 0);
      f->name = xstrdup(p->s);
      f->l = in;
      return f;
    }
    default:
      fail(err, INTERNAL "Cannot transform wildcard to function");
      return NULL;
  }
}
orc_fn *orc_fn_parse(const char *sexpr, char *err) {
  const char *s = sexpr;
  plan *p = plan_parse(&s, err);
  if (!p) return NULL;
  orc_fn *f = plan_to_function(p, 0, err);
  if (f) f->is_agg_plan = plan_is_aggregate(p);
  plan_free(p);
  return f;
}
int32_t orc_fn_is_aggregate(const orc_fn *f) { return f->is_agg_plan; }

/* Debug of Function = Display of the variant (function.rs:134-146) */
static void fn_display(sb *b, const orc_fn *f) {
  switch (f->kind) {
    case FN_ALIAS: case FN_FIELD: sb_put(b, f->name); break;                  /* function_alias.rs:61-65, function_field.rs:75-79 */
    case FN_CONST: value_display(b, &f->value); break;                        /* function_constant.rs:53-57 */
    case FN_ARITH: fn_display(b, f->l); sb_printf(b, " %s ", arith_sym(f->op)); fn_display(b, f->r); break;
    case FN_CMP: fn_display(b, f->l); sb_printf(b, " %s ", cmp_sym(f->op)); fn_display(b, f->r); break;
    case FN_LOGIC: fn_display(b, f->l); sb_printf(b, " %s ", f->op == ORC_LG_AND ? "and" : "or"); fn_display(b, f->r); break;
    case FN_AGG: { /* "{:?}({:?})" with the derived Debug of the operator enum */
      static const char *n[] = {"Min", "Max", "Sum", "Count"};
      sb_printf(b, "%s(", n[f->op & 3]);
      fn_display(b, f->l);
      sb_put(b, ")");
      break;
    }
  }
}
char *orc_fn_display(const orc_fn *f) {
  sb b = {0};
  fn_display(&b, f);
  return sb_take(&b);
}

static int block_index_of(const orc_block *b, const char *name, char *err) {
  for (int i = 0; i < b->n_cols; i++) if (strcmp(b->names[i], name) == 0) return i;
  fail(err, INTERNAL "Invalid argument error: Unable to get field named \"%s\"", name);
  return -1;
}
static int64_t block_rows(const orc_block *b) { return b->n_cols ? b->cols[0].len : 0; }

int32_t orc_fn_return_type(const orc_fn *f, const orc_block *schema_of, int32_t *out, char *err) {
  switch (f->kind) {
    case FN_ALIAS: return orc_fn_return_type(f->l, schema_of, out, err);
    case FN_CONST: *out = f->value.tag; return 0;
    case FN_FIELD: {
      int i = block_index_of(schema_of, f->name, err);
      if (i < 0) return 1;
      *out = schema_of->cols[i].dtype;
      return 0;
    }
    case FN_ARITH: { /* function_arithmetic.rs:36-42 */
      int32_t a, b;
      if (orc_fn_return_type(f->l, schema_of, &a, err) || orc_fn_return_type(f->r, schema_of, &b, err)) return 1;
      return orc_numerical_coercion(arith_sym(f->op), a, b, out, err);
    }
    case FN_CMP: case FN_LOGIC: *out = ORC_BOOL; return 0;
    default: /* function_aggregator.rs:38-43 */
      if (f->op == ORC_AGG_COUNT) { *out = ORC_U64; return 0; }
      return orc_fn_return_type(f->l, schema_of, out, err);
  }
}

int32_t orc_fn_eval(orc_fn *f, const orc_block *block, orc_columnar *out, char *err) {
  memset(out, 0, sizeof *out);
  switch (f->kind) {
    case FN_ALIAS: case FN_AGG: return orc_fn_eval(f->l, block, out, err); /* function_aggregator.rs:53-55 */
    case FN_CONST: out->is_scalar = 1; out->scalar = val_clone(&f->value); return 0;
    case FN_FIELD: {
      int i = block_index_of(block, f->name, err);
      if (i < 0) return 1;
      out->array = arr_borrow(&block->cols[i]);
      return 0;
    }
    default: {
      orc_columnar l, r;
      if (orc_fn_eval(f->l, block, &l, err)) return 1;
      if (orc_fn_eval(f->r, block, &r, err)) { orc_columnar_free(&l); return 1; }
      int rc = f->kind == FN_ARITH ? orc_array_arithmetic(f->op, &l, &r, &out->array, err)
               : f->kind == FN_CMP ? orc_array_comparison(f->op, &l, &r, &out->array, err)
                                   : orc_array_logic(f->op, &l, &r, &out->array, err);
      orc_columnar_free(&l);
      orc_columnar_free(&r);
      return rc;
    }
  }
}

/* function_aggregator.rs:57-100 */
int32_t orc_fn_accumulate(orc_fn *f, const orc_block *block, char *err) {
  switch (f->kind) {
    case FN_ALIAS: return orc_fn_accumulate(f->l, block, err);
    case FN_CONST: case FN_FIELD: return 0;
    case FN_ARITH: case FN_CMP: case FN_LOGIC:
      if (orc_fn_accumulate(f->l, block, err)) return 1;
      return orc_fn_accumulate(f->r, block, err);
    default: break;
  }
  int64_t rows = block_rows(block);
  orc_columnar val;
  if (orc_fn_eval(f->l, block, &val, err)) return 1; /* evaluated even for count, :59 */
  orc_value next = val_null(), part = val_null();
  int rc = 1;
  if (f->op == ORC_AGG_COUNT) {
    orc_value n = val_u64((uint64_t)rows);
    rc = orc_value_arithmetic(ORC_AR_ADD, &f->value, &n, &next, err);
  } else {
    orc_array arr;
    if (columnar_to_array(&val, rows, &arr, err)) goto done;
    int r2 = orc_array_aggregate(f->op, &arr, &part, err);
    orc_array_free(&arr);
    if (r2) goto done;
    rc = f->op == ORC_AGG_SUM ? orc_value_arithmetic(ORC_AR_ADD, &f->value, &part, &next, err)
                              : orc_value_aggregate(f->op, &f->value, &part, &next, err);
  }
  if (!rc) { orc_value_free(&f->value); f->value = next; }
done:
  orc_value_free(&part);
  orc_columnar_free(&val);
  return rc;
}

static int unsupported_agg(const orc_fn *f, char *err) {
  if (f->kind == FN_FIELD) return fail(err, INTERNAL "Unsupported aggregate operation for function field");
  return fail(err, INTERNAL "Unsupported aggregate operation for function %s",
              f->kind == FN_CMP ? cmp_sym(f->op) : (f->op == ORC_LG_AND ? "and" : "or"));
}
/* appends the flattened leaf states (function_arithmetic.rs:69-75) */
static int accumulate_result(const orc_fn *f, orc_value **items, int *n, int *cap, char *err) {
  switch (f->kind) {
    case FN_ALIAS: return accumulate_result(f->l, items, n, cap, err);
    case FN_ARITH:
      if (accumulate_result(f->l, items, n, cap, err)) return 1;
      return accumulate_result(f->r, items, n, cap, err);
    case FN_CONST: case FN_AGG:
      if (*n == *cap) { *cap = *cap ? *cap * 2 : 4; *items = realloc(*items, sizeof(orc_value) * (size_t)*cap); }
      (*items)[(*n)++] = val_clone(&f->value);
      return 0;
    default: return unsupported_agg(f, err);
  }
}
int32_t orc_fn_accumulate_result(const orc_fn *f, orc_value *out_struct, char *err) {
  *out_struct = val_null();
  orc_value *items = NULL;
  int n = 0, cap = 0;
  if (accumulate_result(f, &items, &n, &cap, err)) {
    for (int i = 0; i < n; i++) orc_value_free(&items[i]);
    free(items);
    return 1;
  }
  out_struct->tag = ORC_STRUCT;
  out_struct->items = items ? items : xmalloc(sizeof(orc_value));
  out_struct->n_items = n;
  return 0;
}
/* function_aggregator.rs:106-139; states[self.depth] panics when out of range in the reference */
int32_t orc_fn_merge_state(orc_fn *f, const orc_value *st, char *err) {
  switch (f->kind) {
    case FN_ALIAS: return orc_fn_merge_state(f->l, st, err);
    case FN_CONST: return 0;
    case FN_ARITH:
      if (orc_fn_merge_state(f->l, st, err)) return 1;
      return orc_fn_merge_state(f->r, st, err);
    case FN_AGG: {
      if (f->depth >= (uint64_t)st->n_items)
        return fail(err, INTERNAL "index out of bounds: the len is %d but the index is %" PRIu64, st->n_items, f->depth);
      const orc_value *val = &st->items[f->depth];
      orc_value next;
      int rc = (f->op == ORC_AGG_COUNT || f->op == ORC_AGG_SUM)
                   ? orc_value_arithmetic(ORC_AR_ADD, &f->value, val, &next, err)
                   : orc_value_aggregate(f->op, &f->value, val, &next, err);
      if (rc) return 1;
      orc_value_free(&f->value);
      f->value = next;
      return 0;
    }
    default: return unsupported_agg(f, err);
  }
}
int32_t orc_fn_merge_result(const orc_fn *f, orc_value *out, char *err) {
  switch (f->kind) {
    case FN_ALIAS: return orc_fn_merge_result(f->l, out, err);
    case FN_CONST: case FN_AGG: *out = val_clone(&f->value); return 0;
    case FN_ARITH: { /* function_arithmetic.rs:82-88 */
      orc_value a, b;
      if (orc_fn_merge_result(f->l, &a, err)) return 1;
      if (orc_fn_merge_result(f->r, &b, err)) { orc_value_free(&a); return 1; }
      int rc = orc_value_arithmetic(f->op, &a, &b, out, err);
      orc_value_free(&a);
      orc_value_free(&b);
      return rc;
    }
    default: return unsupported_agg(f, err);
  }
}

/* ------------------------------------------------------------------------------------------
 * transforms and the pipeline
 * ---------------------------------------------------------------------------------------- */
typedef struct { orc_block *blocks; int64_t n, cap; } blocklist;
static void bl_push(blocklist *l, orc_block b) {
  if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->blocks = realloc(l->blocks, sizeof(orc_block) * (size_t)l->cap); }
  l->blocks[l->n++] = b;
}
static void bl_free(blocklist *l) {
  for (int64_t i = 0; i < l->n; i++) orc_block_free(&l->blocks[i]);
  free(l->blocks);
  memset(l, 0, sizeof *l);
}

/* arrow filter: keep rows whose mask bit is set (a null mask slot keeps nothing), order kept.
 * transform_filter.rs:38-55 */
static int filter_block(const orc_block *in, orc_fn *pred, orc_block *out, char *err) {
  memset(out, 0, sizeof *out);
  orc_fn *p = orc_fn_clone(pred); /* funcs[0].clone(), transform_filter.rs:43 */
  orc_columnar cv;
  int rc = orc_fn_eval(p, in, &cv, err);
  orc_fn_free(p);
  if (rc) return 1;
  int64_t rows = block_rows(in);
  orc_array mask;
  if (columnar_to_array(&cv, rows, &mask, err)) { orc_columnar_free(&cv); return 1; }
  if (mask.dtype != ORC_BOOL) {
    orc_array_free(&mask); orc_columnar_free(&cv);
    return fail(err, INTERNAL "cannot downcast to boolean array");
  }
  const uint8_t *m = mask.data;
  int64_t keep = 0;
  for (int64_t i = 0; i < rows; i++) keep += (m[i] && (!mask.valid || mask.valid[i]));
  out->n_cols = in->n_cols;
  for (int c = 0; c < in->n_cols; c++) {
    const orc_array *src = &in->cols[c];
    out->names[c] = xstrdup(in->names[c]);
    out->cols[c] = arr_alloc(src->dtype, keep);
    size_t es = elem_size(src->dtype);
    int64_t k = 0;
    if (src->valid) out->cols[c].valid = xmalloc((size_t)keep);
    for (int64_t i = 0; i < rows; i++) {
      if (!(m[i] && (!mask.valid || mask.valid[i]))) continue;
      if (src->dtype == ORC_UTF8) {
        const char *s = ((char **)src->data)[i];
        ((char **)out->cols[c].data)[k] = s ? xstrdup(s) : NULL;
      } else memcpy((char *)out->cols[c].data + es * (size_t)k, (const char *)src->data + es * (size_t)i, es);
      if (src->valid) out->cols[c].valid[k] = src->valid[i];
      k++;
    }
  }
  orc_array_free(&mask);
  orc_columnar_free(&cv);
  return 0;
}
/* transform_projection.rs:45-56 */
static int project_block(const orc_block *in, orc_fn **funcs, int n, char **names, orc_block *out, char *err) {
  memset(out, 0, sizeof *out);
  int64_t rows = block_rows(in);
  for (int i = 0; i < n; i++) {
    orc_fn *f = orc_fn_clone(funcs[i]);
    orc_columnar cv;
    int rc = orc_fn_eval(f, in, &cv, err);
    orc_fn_free(f);
    if (rc) { orc_block_free(out); return 1; }
    out->names[i] = xstrdup(names[i]);
    if (!cv.is_scalar && cv.array.owned) { /* result of an expression node: moved (Arc in the reference) */
      out->cols[i] = cv.array;
      memset(&cv.array, 0, sizeof cv.array);
    } else if (!cv.is_scalar) {
      out->cols[i] = arr_clone(&cv.array);
    } else if (orc_value_to_array(&cv.scalar, rows, &out->cols[i], err)) {
      free((void *)out->names[i]);
      out->names[i] = NULL;
      orc_columnar_free(&cv);
      orc_block_free(out);
      return 1;
    }
    out->n_cols = i + 1;
    orc_columnar_free(&cv);
  }
  return 0;
}
/* LimitStream::limit, stream_limit.rs:28-48.  returns 0 = pass `out`, 1 = end of stream */
typedef struct { int64_t limit, current; } limit_state;
static int limit_block(limit_state *st, orc_block *blk) {
  int64_t rows = block_rows(blk);
  if (st->current == st->limit) return 1;
  if (st->current + rows < st->limit) { st->current += rows; return 0; }
  int64_t keep = st->limit - st->current;
  st->current = st->limit;
  for (int c = 0; c < blk->n_cols; c++) blk->cols[c].len = keep < blk->cols[c].len ? keep : blk->cols[c].len; /* arrow limit = slice(0, min) */
  return 0;
}

typedef struct {
  const orc_query *q;
  /* partitions of this source pipe */
  int n_parts;
  uint64_t begins[8], ends[8];
  orc_fn *pred;
  orc_fn **funcs; /* clones owned by the way */
  char **names;
  /* outputs */
  blocklist out;          /* projection path */
  char **state_json;      /* aggregate path: one JSON row per aggregate expr */
  int64_t rows_scanned;
  int failed;
  char err[ORC_ERRLEN];
} way;

static void materialise_numbers(uint64_t b, uint64_t e, orc_block *blk) {
  /* NumbersStream::poll_next, numbers_stream.rs:68-83: (begin..=end).collect::<Vec<u64>>() then
   * UInt64Array::from(vec) copies into an Arrow buffer */
  int64_t n = (int64_t)(e - b + 1);
  uint64_t *vec = xmalloc(sizeof(uint64_t) * (size_t)n);
  for (int64_t i = 0; i < n; i++) vec[i] = b + (uint64_t)i;
  memset(blk, 0, sizeof *blk);
  blk->n_cols = 1;
  blk->names[0] = xstrdup("number");
  blk->cols[0] = arr_alloc(ORC_U64, n);
  memcpy(blk->cols[0].data, vec, sizeof(uint64_t) * (size_t)n);
  free(vec);
}
static void slice_table(const orc_block *t, uint64_t b, uint64_t e, orc_block *blk) {
  memset(blk, 0, sizeof *blk);
  blk->n_cols = t->n_cols;
  int64_t n = (int64_t)(e - b + 1);
  for (int c = 0; c < t->n_cols; c++) {
    blk->names[c] = xstrdup(t->names[c]);
    blk->cols[c] = arr_borrow(&t->cols[c]);
    blk->cols[c].data = (char *)t->cols[c].data + elem_size(t->cols[c].dtype) * b;
    if (t->cols[c].valid) blk->cols[c].valid = t->cols[c].valid + b;
    blk->cols[c].len = n;
  }
}

static void *way_run(void *arg) {
  way *w = arg;
  const orc_query *q = w->q;
  limit_state lim = {q->limit, 0};
  int ended = 0;
  for (int p = 0; p < w->n_parts && !ended && !w->failed; p++) {
    int64_t nb = orc_block_ranges(w->begins[p], w->ends[p], q->block_size, q->tail_quirk, NULL, NULL, 0);
    uint64_t *bb = xmalloc(sizeof(uint64_t) * (size_t)nb), *be = xmalloc(sizeof(uint64_t) * (size_t)nb);
    orc_block_ranges(w->begins[p], w->ends[p], q->block_size, q->tail_quirk, bb, be, nb);
    for (int64_t k = 0; k < nb && !ended; k++) {
      orc_block blk, tmp;
      if (q->table) slice_table(q->table, bb[k], be[k], &blk); else materialise_numbers(bb[k], be[k], &blk);
      w->rows_scanned += block_rows(&blk);
      if (w->pred) { /* FilterTransform */
        if (filter_block(&blk, w->pred, &tmp, w->err)) { w->failed = 1; orc_block_free(&blk); break; }
        orc_block_free(&blk);
        blk = tmp;
      }
      if (q->is_aggregate) { /* AggregatePartialTransform hot loop, transform_aggregate_partial.rs:53-59 */
        for (int i = 0; i < q->n_exprs; i++)
          if (orc_fn_accumulate(w->funcs[i], &blk, w->err)) { w->failed = 1; break; }
        orc_block_free(&blk);
        if (w->failed) break;
        continue;
      }
      if (project_block(&blk, w->funcs, q->n_exprs, w->names, &tmp, w->err)) { w->failed = 1; orc_block_free(&blk); break; }
      orc_block_free(&blk);
      if (q->limit >= 0 && limit_block(&lim, &tmp)) { orc_block_free(&tmp); ended = 1; break; }
      bl_push(&w->out, tmp);
    }
    free(bb);
    free(be);
  }
  if (q->is_aggregate && !w->failed) { /* transform_aggregate_partial.rs:61-72 */
    w->state_json = xmalloc(sizeof(char *) * (size_t)q->n_exprs);
    memset(w->state_json, 0, sizeof(char *) * (size_t)q->n_exprs);
    for (int i = 0; i < q->n_exprs; i++) {
      orc_value st;
      if (orc_fn_accumulate_result(w->funcs[i], &st, w->err)) { w->failed = 1; break; }
      w->state_json[i] = orc_value_to_json(&st);
      orc_value_free(&st);
    }
  }
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int32_t orc_query_run(const orc_query *q, orc_result *out, char *err) {
  memset(out, 0, sizeof *out);
  if (q->n_exprs < 1 || q->n_exprs > ORC_MAX_COLS) return fail(err, PLANERR "need 1..%d select expressions", ORC_MAX_COLS);
  int rc = 1;
  orc_fn *pred = NULL, *funcs[ORC_MAX_COLS] = {0};
  char *names[ORC_MAX_COLS] = {0};
  way ways[8];
  uint8_t *kept[8] = {0};
  int n_ways = 0;
  memset(ways, 0, sizeof ways);

  /* schema used for type checks */
  orc_block schema;
  memset(&schema, 0, sizeof schema);
  if (q->table) schema = *q->table;
  else { schema.n_cols = 1; schema.names[0] = "number"; schema.cols[0].dtype = ORC_U64; }

  if (q->predicate) {
    pred = orc_fn_parse(q->predicate, err);
    if (!pred) goto done;
    if (pred->is_agg_plan) { /* transform_filter.rs:23-29 */
      char *d = orc_plan_display(q->predicate, err);
      fail(err, INTERNAL "Aggregate function %s is found in WHERE in query", d ? d : "?");
      free(d);
      goto done;
    }
  }
  for (int i = 0; i < q->n_exprs; i++) {
    funcs[i] = orc_fn_parse(q->exprs[i], err);
    if (!funcs[i]) goto done;
    if (!q->is_aggregate && funcs[i]->is_agg_plan) { /* transform_projection.rs:24-31 */
      char *d = orc_plan_display(q->exprs[i], err);
      fail(err, INTERNAL "Unsupported aggregator function: %s", d ? d : "?");
      free(d);
      goto done;
    }
    names[i] = orc_fn_display(funcs[i]); /* ExpressionPlan::to_field, plan_expression.rs:31-38 */
  }

  /* ReadSource: partitions chunked over workers, pipeline_builder.rs:73-95 */
  uint64_t pb[8], pe[8];
  uint64_t total = q->table ? (uint64_t)block_rows(q->table) : q->total;
  int n_parts = orc_generate_parts(total, pb, pe);
  int workers = q->worker_threads;
  int chunk = (workers == 0 || workers >= n_parts) ? 1 : n_parts / workers;
  for (int s = 0; s < n_parts; s += chunk) {
    way *w = &ways[n_ways++];
    w->q = q;
    for (int k = s; k < s + chunk && k < n_parts; k++) { w->begins[w->n_parts] = pb[k]; w->ends[w->n_parts] = pe[k]; w->n_parts++; }
    w->pred = pred ? orc_fn_clone(pred) : NULL;
    w->funcs = xmalloc(sizeof(orc_fn *) * (size_t)q->n_exprs);
    for (int i = 0; i < q->n_exprs; i++) w->funcs[i] = orc_fn_clone(funcs[i]);
    w->names = names;
  }

  double t0 = now_s();
  if (q->use_threads && n_ways > 1) { /* MergeProcessor: one task per input, processor_merge.rs:46-62 */
    pthread_t th[8];
    for (int i = 0; i < n_ways; i++) pthread_create(&th[i], NULL, way_run, &ways[i]);
    for (int i = 0; i < n_ways; i++) pthread_join(th[i], NULL);
  } else {
    for (int i = 0; i < n_ways; i++) way_run(&ways[i]);
  }
  for (int i = 0; i < n_ways; i++) {
    out->rows_scanned += ways[i].rows_scanned;
    if (ways[i].failed) { fail(err, "%s", ways[i].err); goto done; }
  }

  if (q->is_aggregate) {
    /* AggregateFinalTransform, transform_aggregate_final.rs:50-78 (blocks taken in pipe order; any
     * arrival order gives the same integers) */
    sb js = {0};
    for (int w = 0; w < n_ways; w++) {
      for (int i = 0; i < q->n_exprs; i++) {
        orc_value st;
        if (orc_value_from_json(ways[w].state_json[i], &st, err)) { free(js.p); goto done; }
        if (st.tag == ORC_STRUCT && orc_fn_merge_state(funcs[i], &st, err)) { orc_value_free(&st); free(js.p); goto done; }
        orc_value_free(&st);
        if (js.n) sb_put(&js, "\n");
        sb_put(&js, ways[w].state_json[i]);
      }
    }
    out->partial_states_json = sb_take(&js);
    out->block.n_cols = q->n_exprs;
    for (int i = 0; i < q->n_exprs; i++) {
      orc_value r;
      if (orc_fn_merge_result(funcs[i], &r, err)) goto done;
      int r2 = orc_value_to_array(&r, 1, &out->block.cols[i], err);
      orc_value_free(&r);
      if (r2) goto done;
      out->block.names[i] = xstrdup(names[i]);
    }
    out->n_rows = 1;
    out->n_blocks_out = 1;
    if (q->limit >= 0) { /* LimitTransform x 1 after the final, pipeline_builder.rs:31-41 */
      limit_state lim = {q->limit, 0};
      if (limit_block(&lim, &out->block)) { out->n_rows = 0; out->n_blocks_out = 0; for (int i = 0; i < q->n_exprs; i++) out->block.cols[i].len = 0; }
      else out->n_rows = block_rows(&out->block);
    }
  } else {
    /* Merge then LimitTransform x 1 (only when more than one pipe, pipeline_builder.rs:31-41) */
    limit_state lim = {q->limit, 0};
    int final_limit = q->limit >= 0 && n_ways > 1;
    int32_t dtypes[ORC_MAX_COLS];
    for (int i = 0; i < q->n_exprs; i++)
      if (orc_fn_return_type(funcs[i], &schema, &dtypes[i], err)) goto done;
    int64_t total_rows = 0;
    int ended = 0;
    for (int w = 0; w < n_ways; w++) { kept[w] = xmalloc((size_t)(ways[w].out.n / 8 + 1)); memset(kept[w], 0, (size_t)(ways[w].out.n / 8 + 1)); }
    for (int w = 0; w < n_ways && !ended; w++)
      for (int64_t k = 0; k < ways[w].out.n; k++) {
        orc_block *b = &ways[w].out.blocks[k];
        if (final_limit && limit_block(&lim, b)) { ended = 1; break; }
        out->n_blocks_out++;
        total_rows += block_rows(b);
        kept[w][k >> 3] |= (uint8_t)(1u << (k & 7));
      }
    out->block.n_cols = q->n_exprs;
    for (int i = 0; i < q->n_exprs; i++) {
      out->block.names[i] = xstrdup(names[i]);
      out->block.cols[i] = arr_alloc(dtypes[i], total_rows);
    }
    int64_t at = 0;
    for (int w = 0; w < n_ways; w++)
      for (int64_t k = 0; k < ways[w].out.n; k++) {
        orc_block *b = &ways[w].out.blocks[k];
        if (!(kept[w][k >> 3] & (1u << (k & 7)))) continue;
        int64_t r = block_rows(b);
        for (int i = 0; i < q->n_exprs; i++) {
          size_t es = elem_size(dtypes[i]);
          if (b->cols[i].dtype != dtypes[i]) { fail(err, INTERNAL "projected column type mismatch"); goto done; }
          if (dtypes[i] == ORC_UTF8) {
            for (int64_t j = 0; j < r; j++) {
              const char *s = ((char **)b->cols[i].data)[j];
              ((char **)out->block.cols[i].data)[at + j] = s ? xstrdup(s) : NULL;
            }
          } else if (r) memcpy((char *)out->block.cols[i].data + es * (size_t)at, b->cols[i].data, es * (size_t)r);
          if (b->cols[i].valid) {
            if (!out->block.cols[i].valid) { out->block.cols[i].valid = xmalloc((size_t)total_rows); memset(out->block.cols[i].valid, 1, (size_t)total_rows); }
            memcpy(out->block.cols[i].valid + at, b->cols[i].valid, (size_t)r);
          }
        }
        at += r;
      }
    out->n_rows = total_rows;
  }
  out->seconds = now_s() - t0;
  rc = 0;
done:
  for (int w = 0; w < n_ways; w++) {
    /* a limit may have shortened len of an owned array: restore nothing, free by pointer */
    bl_free(&ways[w].out);
    free(kept[w]);
    orc_fn_free(ways[w].pred);
    if (ways[w].funcs) for (int i = 0; i < q->n_exprs; i++) orc_fn_free(ways[w].funcs[i]);
    free(ways[w].funcs);
    if (ways[w].state_json) for (int i = 0; i < q->n_exprs; i++) free(ways[w].state_json[i]);
    free(ways[w].state_json);
  }
  orc_fn_free(pred);
  for (int i = 0; i < q->n_exprs; i++) { orc_fn_free(funcs[i]); free(names[i]); }
  if (rc) orc_result_free(out);
  return rc;
}

void orc_result_free(orc_result *r) {
  if (!r) return;
  orc_block_free(&r->block);
  free(r->partial_states_json);
  memset(r, 0, sizeof *r);
}

/* ---- best-case CPU variant of the headline query (bench.py only; see the header) ---- */
typedef struct { uint64_t begin, end, sum, mx, mn; } fused_part;
static void *fused_run(void *arg) {
  fused_part *w = (fused_part *)arg;
  uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, mx = 0, mn = UINT64_MAX;
  uint64_t i = w->begin;
  for (; i + 4 <= w->end; i += 4) { /* four independent chains; the compiler vectorises this */
    s0 += i; s1 += i + 1; s2 += i + 2; s3 += i + 3;
    uint64_t hi = i + 3;
    if (hi > mx) mx = hi;
    if (i < mn) mn = i;
  }
  for (; i < w->end; i++) { s0 += i; if (i > mx) mx = i; if (i < mn) mn = i; }
  w->sum = s0 + s1 + s2 + s3; w->mx = mx; w->mn = mn;
  return NULL;
}
double orc_fused_headline(uint64_t total, int32_t threads, uint64_t out[4]) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  fused_part parts[256];
  pthread_t th[256];
  const double t0 = now_s();
  for (int t = 0; t < threads; t++) {
    parts[t].begin = total / (uint64_t)threads * (uint64_t)t;
    parts[t].end = t == threads - 1 ? total : total / (uint64_t)threads * (uint64_t)(t + 1);
    pthread_create(&th[t], NULL, fused_run, &parts[t]);
  }
  uint64_t sum = 0, mx = 0, mn = UINT64_MAX;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    sum += parts[t].sum;
    if (parts[t].begin < parts[t].end) { if (parts[t].mx > mx) mx = parts[t].mx; if (parts[t].mn < mn) mn = parts[t].mn; }
  }
  out[0] = sum; out[1] = total; out[2] = mx; out[3] = mn;
  return now_s() - t0;
}
