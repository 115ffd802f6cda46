// fq_skeleton.cuh — hand-written sm_100a kernel skeletons of the fused fuse-query pipes.
//
// A pipe's expression trees are lowered by codegen.cc into a struct `Q` (typed row loaders, the WHERE
// predicate, per-row accumulate / projection code); the templates below supply everything around it:
// the coalesced 128-bit streaming loads, the register -> warp-shuffle -> shared-memory -> grid
// reductions, the ballot/popc stream compaction with a decoupled look-back across tiles.
//
// The same text is compiled twice: by nvcc into libfuse_gpu.so for the precompiled pipes
// (aot_kernels.cu) and by NVRTC at fq_pipe_compile time for any other expression tree.  It therefore
// includes nothing and relies on compiler builtins only.
//
// What each kernel replaces in the reference (/root/reference/src):
//   fq_agg_kernel     AggregatePartialTransform's loop (transforms/transform_aggregate_partial.rs:53-59) =
//                     per block, per aggregate: Function::eval of the argument (one Arrow pass and one
//                     intermediate array per node, datavalues/data_array_arithmetic.rs:14-55), one Arrow
//                     sum/min/max pass (data_array_aggregate.rs:103-114) and a scalar fold
//                     (functions/function_aggregator.rs:57-100) — here ONE pass, 8 B read per row.
//   fq_select_kernel  FilterTransform + ProjectionTransform (+ LimitStream): predicate -> BooleanArray ->
//                     filter_record_batch -> per-expression arrays (transform_filter.rs:38-55,
//                     transform_projection.rs:45-56, datastreams/stream_limit.rs:28-48).
//   fq_map_kernel     ProjectionTransform without a filter.
//   fq_fill_numbers   NumbersStream::poll_next (datasources/system/numbers_stream.rs:68-83).
#pragma once

typedef unsigned long long fq_u64;
typedef long long fq_i64;
typedef unsigned int fq_u32;
typedef int fq_i32;
typedef unsigned short fq_u16;
typedef short fq_i16;
typedef unsigned char fq_u8;
typedef signed char fq_i8;

#define FQ_E_DIVZERO 1u  // arrow DivideByZero
#define FQ_E_CAST 2u     // arrow cast would have produced a null (out-of-range numeric cast)

// launch shapes (the host reads the same macros through fq_skeleton_config.h)
#ifndef FQ_AGG_THREADS
#define FQ_AGG_THREADS 256
#define FQ_AGG_MIN_BLOCKS 4
#define FQ_AGG_MIN_BLOCKS_U8 2
#define FQ_SEL_THREADS 256
#define FQ_SEL_MIN_BLOCKS 3
#define FQ_SEL_UNROLL 4
#define FQ_MAP_THREADS 256
#define FQ_MAP_MIN_BLOCKS 4
#define FQ_MAP_UNROLL 4
#endif

#define FQ_STATE_HDR 4        // state / partial slots: [0] rows selected, [1] error bits, [2] launches folded, [3] rows scanned
#define FQ_MAX_WARPS 32

// Kernel parameter block (one struct for every kernel so the host launch path is uniform).
struct fq_launch_params {
  fq_u64 n_rows;
  const void *cols[8];
  fq_u64 numbers_begin;  // generated mode: column 0 = numbers_begin + row
  // aggregate
  fq_u64 *partials;      // [gridDim.x][FQ_STATE_HDR + Q::NSLOTS]
  fq_u64 *state;         // [FQ_STATE_HDR + Q::NSLOTS] running state of the pipe
  fq_u32 *ticket;
  fq_u32 accumulate;
  // select / map
  void *outs[8];
  fq_u64 capacity;       // rows written are those with rank < capacity (min(limit, capacity) on the host)
  fq_u64 *tile_status;   // decoupled look-back descriptors, zeroed before the launch
  fq_u32 *tile_counter;  // dynamic tile ids (forward progress for the look-back), zeroed before the launch
  fq_u64 *result;        // [0] rows selected, [1] error bits
  fq_u64 n_tiles;
  fq_u64 stop_after;     // early exit: stop scanning once this many rows were selected (0 = never)
  fq_u32 *done;          // early-exit flag, zeroed before the launch
};

// ---------------------------------------------------------------------------------------------
// memory access helpers
// ---------------------------------------------------------------------------------------------
struct fq_b16 { fq_u32 x, y, z, w; };

// streaming 128-bit load: read-only path, do not allocate in L1 (every byte is touched once)
__device__ __forceinline__ fq_b16 fq_ld16(const void *p) {
  fq_b16 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ fq_u64 fq_ld_volatile(const fq_u64 *p) {
  fq_u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void fq_st_volatile(fq_u64 *p, fq_u64 v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ fq_u32 fq_ld_volatile32(const fq_u32 *p) {
  fq_u32 v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ fq_u64 fq_ld_cg(const fq_u64 *p) {
  fq_u64 v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// one value (scalar tails, ragged last tile)
template <class T> __device__ __forceinline__ T fq_ld1(const void *base, fq_u64 row) { return __ldg((const T *)base + row); }
template <> __device__ __forceinline__ bool fq_ld1<bool>(const void *base, fq_u64 row) { return __ldg((const fq_u8 *)base + row) != 0; }

// Load V consecutive values of type T (V * sizeof(T) bytes, a multiple of 16 or a power of two below)
template <class T, int V>
__device__ __forceinline__ void fq_load_vec(T (&dst)[V], const void *base, fq_u64 group) {
  constexpr int BYTES = V * (int)sizeof(T);
  const char *p = (const char *)base + group * (fq_u64)BYTES;
  if constexpr (BYTES >= 16) {
    union { fq_b16 q[BYTES / 16]; T t[V]; } u;
#pragma unroll
    for (int k = 0; k < BYTES / 16; k++) u.q[k] = fq_ld16(p + 16 * k);
#pragma unroll
    for (int k = 0; k < V; k++) dst[k] = u.t[k];
  } else if constexpr (BYTES == 8) {
    union { fq_u64 q; T t[V]; } u;
    u.q = __ldg((const fq_u64 *)p);
#pragma unroll
    for (int k = 0; k < V; k++) dst[k] = u.t[k];
  } else if constexpr (BYTES == 4) {
    union { fq_u32 q; T t[V]; } u;
    u.q = __ldg((const fq_u32 *)p);
#pragma unroll
    for (int k = 0; k < V; k++) dst[k] = u.t[k];
  } else {
#pragma unroll
    for (int k = 0; k < V; k++) dst[k] = fq_ld1<T>(p, k);
  }
}

// ---------------------------------------------------------------------------------------------
// arithmetic helpers used by generated code (semantics of arrow 2.0 as the reference calls it)
// ---------------------------------------------------------------------------------------------
template <class T> struct fq_traits;
#define FQ_TRAITS(T, UT, ISF, ISS, LO, HI)                                          \
  template <> struct fq_traits<T> {                                                 \
    typedef UT unsigned_t;                                                          \
    enum { is_float = ISF, is_signed = ISS };                                       \
    __device__ static __forceinline__ T lo() { return LO; }                         \
    __device__ static __forceinline__ T hi() { return HI; }                         \
  };
FQ_TRAITS(fq_i8, fq_u8, 0, 1, (fq_i8)-128, (fq_i8)127)
FQ_TRAITS(fq_i16, fq_u16, 0, 1, (fq_i16)-32768, (fq_i16)32767)
FQ_TRAITS(fq_i32, fq_u32, 0, 1, (fq_i32)(-2147483647 - 1), (fq_i32)2147483647)
FQ_TRAITS(fq_i64, fq_u64, 0, 1, (fq_i64)(-9223372036854775807ll - 1), (fq_i64)9223372036854775807ll)
FQ_TRAITS(fq_u8, fq_u8, 0, 0, (fq_u8)0, (fq_u8)255)
FQ_TRAITS(fq_u16, fq_u16, 0, 0, (fq_u16)0, (fq_u16)65535)
FQ_TRAITS(fq_u32, fq_u32, 0, 0, 0u, 4294967295u)
FQ_TRAITS(fq_u64, fq_u64, 0, 0, 0ull, 18446744073709551615ull)
FQ_TRAITS(float, float, 1, 1, -__int_as_float(0x7f800000), __int_as_float(0x7f800000))
FQ_TRAITS(double, double, 1, 1, -__longlong_as_double(0x7ff0000000000000ll), __longlong_as_double(0x7ff0000000000000ll))
#undef FQ_TRAITS

// wrapping integer add / sub / mul (arrow `add` etc. on integer lanes), plain IEEE for floats
template <class T> __device__ __forceinline__ T fq_add(T a, T b) {
  if constexpr (fq_traits<T>::is_float) return a + b;
  else { typedef typename fq_traits<T>::unsigned_t U; return (T)(U)((U)a + (U)b); }
}
template <class T> __device__ __forceinline__ T fq_sub(T a, T b) {
  if constexpr (fq_traits<T>::is_float) return a - b;
  else { typedef typename fq_traits<T>::unsigned_t U; return (T)(U)((U)a - (U)b); }
}
template <class T> __device__ __forceinline__ T fq_mul(T a, T b) {
  if constexpr (fq_traits<T>::is_float) return a * b;
  else { typedef typename fq_traits<T>::unsigned_t U; return (T)(U)((U)a * (U)b); }
}
// arrow `divide`: any zero divisor is an error (integer and float lanes alike); integers truncate
template <class T> __device__ __forceinline__ T fq_div(T a, T b, fq_u32 &err) {
  if (b == (T)0) { err |= FQ_E_DIVZERO; return (T)0; }
  if constexpr (fq_traits<T>::is_float) return a / b;
  else if constexpr (fq_traits<T>::is_signed) {
    typedef typename fq_traits<T>::unsigned_t U;
    if (b == (T)-1) return (T)(U)((U)0 - (U)a);  // MIN / -1 wraps instead of trapping
    return (T)(a / b);
  } else return (T)(a / b);
}
template <class T> __device__ __forceinline__ T fq_min(T a, T b) { return b < a ? b : a; }
template <class T> __device__ __forceinline__ T fq_max(T a, T b) { return b > a ? b : a; }

// arrow numeric `cast` (num::cast): a value that does not fit the target becomes null.  The device
// path carries no validity yet, so such a row raises FQ_E_CAST and the launch reports "unsupported".
template <class T, class S> __device__ __forceinline__ T fq_cast(S x, fq_u32 &err) {
  if constexpr (fq_traits<T>::is_float) return (T)x;
  else if constexpr (fq_traits<S>::is_float) {
    const double d = (double)x;
    const double t = d < 0 ? -floor(-d) : floor(d);
    bool ok;
    if constexpr (sizeof(T) == 8 && fq_traits<T>::is_signed) ok = t >= -9223372036854775808.0 && t < 9223372036854775808.0;
    else if constexpr (sizeof(T) == 8) ok = t > -1.0 && t < 18446744073709551616.0;
    else ok = t >= (double)fq_traits<T>::lo() && t <= (double)fq_traits<T>::hi();
    if (!ok || d != d) { err |= FQ_E_CAST; return (T)0; }
    return (T)t;
  } else if constexpr (fq_traits<S>::is_signed && !fq_traits<T>::is_signed) {
    if (x < 0 || (fq_u64)x > (fq_u64)fq_traits<T>::hi()) { err |= FQ_E_CAST; return (T)0; }
    return (T)x;
  } else if constexpr (!fq_traits<S>::is_signed && fq_traits<T>::is_signed) {
    if ((fq_u64)x > (fq_u64)fq_traits<T>::hi()) { err |= FQ_E_CAST; return (T)0; }
    return (T)x;
  } else {
    if (x < (S)fq_traits<T>::lo() && sizeof(S) > sizeof(T)) { err |= FQ_E_CAST; return (T)0; }
    if (sizeof(S) > sizeof(T) && x > (S)fq_traits<T>::hi()) { err |= FQ_E_CAST; return (T)0; }
    return (T)x;
  }
}

// 64-bit slot packing of accumulator values (what crosses to the host / between launches)
template <class T> __device__ __forceinline__ fq_u64 fq_pack(T x) {
  if constexpr (fq_traits<T>::is_float) return (fq_u64)__double_as_longlong((double)x);
  else if constexpr (fq_traits<T>::is_signed) return (fq_u64)(fq_i64)x;
  else return (fq_u64)x;
}
template <class T> __device__ __forceinline__ T fq_unpack(fq_u64 s) {
  if constexpr (fq_traits<T>::is_float) return (T)__longlong_as_double((fq_i64)s);
  else return (T)s;
}
template <class T> __device__ __forceinline__ T fq_shfl_xor(T x, int m) {
  if constexpr (sizeof(T) == 8) return fq_unpack<T>(__shfl_xor_sync(0xffffffffu, fq_pack<T>(x), m));
  else if constexpr (fq_traits<T>::is_float) return __shfl_xor_sync(0xffffffffu, x, m);
  else return (T)__shfl_xor_sync(0xffffffffu, (fq_i32)x, m);
}

// ---------------------------------------------------------------------------------------------
// block-level reduction of a generated accumulator:  registers -> warp shuffles -> shared -> warp 0
// result valid in thread 0
// ---------------------------------------------------------------------------------------------
template <class Q>
__device__ __forceinline__ void fq_block_reduce(typename Q::Acc &acc, fq_u64 &nsel, fq_u32 &err,
                                                fq_u64 (*sm)[FQ_STATE_HDR + Q::NSLOTS]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    typename Q::Acc o = acc;
    Q::shfl(o, m);
    Q::merge(acc, o);
    nsel += __shfl_xor_sync(0xffffffffu, nsel, m);
    err |= __shfl_xor_sync(0xffffffffu, err, m);
  }
  __syncthreads();  // sm may still be read by a previous use
  if (lane == 0) {
    sm[warp][0] = nsel;
    sm[warp][1] = err;
    Q::store(acc, &sm[warp][FQ_STATE_HDR]);
  }
  __syncthreads();
  if (warp == 0) {
    Q::init(acc);
    nsel = 0;
    err = 0;
    if (lane < nwarps) {
      nsel = sm[lane][0];
      err = (fq_u32)sm[lane][1];
      Q::unpack(acc, &sm[lane][FQ_STATE_HDR]);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      typename Q::Acc o = acc;
      Q::shfl(o, m);
      Q::merge(acc, o);
      nsel += __shfl_xor_sync(0xffffffffu, nsel, m);
      err |= __shfl_xor_sync(0xffffffffu, err, m);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fq_agg_kernel — single-pass multi-aggregate scan.
//
// Grid: persistent, (SM count x resident CTAs) blocks.  Each CTA walks contiguous chunks of
// blockDim.x * UNROLL vector groups (16 B per thread per load, UNROLL independent loads in flight
// per thread, consecutive lanes on consecutive 16-B words -> every warp load is one 512-B run).
// Algorithmic traffic: sizeof(row) bytes read per row, (FQ_STATE_HDR + NSLOTS) * 8 B written per CTA.
// ---------------------------------------------------------------------------------------------
template <class Q, int UNROLL>
__device__ __forceinline__ void fq_agg_kernel(const fq_launch_params &p) {
  constexpr int V = Q::V;
  constexpr int S = FQ_STATE_HDR + Q::NSLOTS;
  __shared__ fq_u64 sm[FQ_MAX_WARPS][S];
  __shared__ fq_u32 s_last;

  typename Q::Acc acc;
  Q::init(acc);
  fq_u32 err = 0;
  fq_u64 nsel = 0;

  const fq_u64 nvec = p.n_rows / V;
  const fq_u64 chunk = (fq_u64)blockDim.x * UNROLL;
  const fq_u64 nfull = nvec / chunk;
  for (fq_u64 c = blockIdx.x; c < nfull; c += gridDim.x) {
    const fq_u64 g0 = c * chunk + threadIdx.x;
    typename Q::Rows rows[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) Q::load(rows[u], p, g0 + (fq_u64)u * blockDim.x);
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int v = 0; v < V; v++) Q::consume(acc, rows[u], v, nsel, err);
  }
  // remainder groups (< one chunk) and the scalar tail (< V rows), spread over the whole grid
  const fq_u64 tid = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x;
  const fq_u64 nthreads = (fq_u64)gridDim.x * blockDim.x;
  for (fq_u64 g = nfull * chunk + tid; g < nvec; g += nthreads) {
    typename Q::Rows r;
    Q::load(r, p, g);
#pragma unroll
    for (int v = 0; v < V; v++) Q::consume(acc, r, v, nsel, err);
  }
  for (fq_u64 row = nvec * V + tid; row < p.n_rows; row += nthreads) {
    typename Q::Rows r;
    Q::load1(r, p, row);
    Q::consume(acc, r, 0, nsel, err);
  }

  fq_block_reduce<Q>(acc, nsel, err, sm);
  if (threadIdx.x == 0) {
    fq_u64 *out = p.partials + (fq_u64)blockIdx.x * S;
    out[0] = nsel;
    out[1] = err;
    Q::store(acc, out + FQ_STATE_HDR);
    __threadfence();
    s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;

  // last CTA: fold every CTA's partial, then fold into (or restart) the running state
  __threadfence();
  Q::init(acc);
  nsel = 0;
  err = 0;
  for (fq_u32 i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    const fq_u64 *in = p.partials + (fq_u64)i * S;
    fq_u64 tmp[Q::NSLOTS > 0 ? Q::NSLOTS : 1];
#pragma unroll
    for (int k = 0; k < Q::NSLOTS; k++) tmp[k] = fq_ld_cg(in + FQ_STATE_HDR + k);
    typename Q::Acc o;
    Q::unpack(o, tmp);
    Q::merge(acc, o);
    nsel += fq_ld_cg(in);
    err |= (fq_u32)fq_ld_cg(in + 1);
  }
  fq_block_reduce<Q>(acc, nsel, err, sm);
  if (threadIdx.x == 0) {
    if (!Q::HAS_PRED) nsel = p.n_rows;
    fq_u64 folded = 1, scanned = p.n_rows;
    if (p.accumulate) {
      typename Q::Acc o;
      Q::unpack(o, p.state + FQ_STATE_HDR);
      Q::merge(acc, o);
      nsel += p.state[0];
      err |= (fq_u32)p.state[1];
      folded += p.state[2];
      scanned += p.state[3];
    }
    p.state[0] = nsel;
    p.state[1] = err;
    p.state[2] = folded;
    p.state[3] = scanned;
    Q::store(acc, p.state + FQ_STATE_HDR);
    *p.ticket = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// fq_select_kernel — fused predicate + order-preserving stream compaction + projection (+ limit).
//
// Tile = blockDim.x * U vector groups; warp w owns the contiguous run of 32 * U groups at
// tile_base + w * 32 * U, so row order = (warp, u, lane, v).  Ranks: one __ballot_sync per (u, v),
// __popc of the lower-lane mask; warp totals through shared memory; tile prefix by a decoupled
// look-back over 64-bit descriptors {flag:2, count:62} (tile ids handed out by atomicAdd so every
// predecessor of a running tile has started).  Selected rows are projected at scatter time.
// Algorithmic traffic: sizeof(row) read per row + sum(sizeof(out_i)) written per selected row.
// ---------------------------------------------------------------------------------------------
#define FQ_TILE_AGG (1ull << 62)
#define FQ_TILE_PREFIX (2ull << 62)
#define FQ_TILE_VALUE(x) ((x) & ((1ull << 62) - 1))
#define FQ_TILE_FLAG(x) ((x) >> 62)

template <class Q, int U>
__device__ __forceinline__ void fq_select_kernel(const fq_launch_params &p) {
  constexpr int V = Q::V;
  __shared__ fq_u32 s_warp_tot[FQ_MAX_WARPS];
  __shared__ fq_u64 s_tile_excl;
  __shared__ fq_u64 s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const fq_u32 lt_mask = (1u << lane) - 1u;
  const fq_u64 tile_groups = (fq_u64)blockDim.x * U;
  const fq_u64 tile_rows = tile_groups * V;
  fq_u32 err = 0;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(p.tile_counter, 1u);
    __syncthreads();
    const fq_u64 tile = s_tile;
    if (tile >= p.n_tiles) break;

    const bool skip = p.stop_after != 0 && fq_ld_volatile32(p.done) != 0;  // uniform per CTA? no: re-read below
    __shared__ fq_u32 s_skip;
    if (threadIdx.x == 0) s_skip = skip ? 1u : 0u;
    __syncthreads();

    typename Q::Rows rows[U];
    fq_u32 keep = 0;  // bit (u * V + v)
    fq_u32 rank[U];
    fq_u32 wtotal = 0;
    if (!s_skip) {
      const fq_u64 g0 = tile * tile_groups + (fq_u64)warp * 32 * U + lane;
      const bool full = (tile + 1) * tile_rows <= p.n_rows;
      if (full) {
#pragma unroll
        for (int u = 0; u < U; u++) Q::load(rows[u], p, g0 + 32ull * u);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int v = 0; v < V; v++) keep |= (Q::pred(rows[u], v, err) ? 1u : 0u) << (u * V + v);
      } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
          const fq_u64 row0 = (g0 + 32ull * u) * V;
#pragma unroll
          for (int v = 0; v < V; v++) {
            if (row0 + v < p.n_rows) {
              typename Q::Rows one;
              Q::load1(one, p, row0 + v);
              Q::copy_row(rows[u], v, one);
              keep |= (Q::pred(rows[u], v, err) ? 1u : 0u) << (u * V + v);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        fq_u32 before = 0, tot = 0;
#pragma unroll
        for (int v = 0; v < V; v++) {
          const fq_u32 b = __ballot_sync(0xffffffffu, (keep >> (u * V + v)) & 1u);
          before += __popc(b & lt_mask);
          tot += __popc(b);
        }
        rank[u] = wtotal + before;
        wtotal += tot;
      }
    }
    if (lane == 0) s_warp_tot[warp] = wtotal;
    __syncthreads();

    fq_u32 warp_off = 0, tile_total = 0;
    for (int w = 0; w < nwarps; w++) {
      const fq_u32 t = s_warp_tot[w];
      if (w < warp) warp_off += t;
      tile_total += t;
    }

    // decoupled look-back, done by warp 0
    if (warp == 0) {
      fq_u64 excl = 0;
      if (s_skip) {
        // the scan already found `stop_after` rows: publish a saturated prefix so successors write nothing
        if (lane == 0) fq_st_volatile(p.tile_status + tile, FQ_TILE_PREFIX | p.stop_after);
        excl = p.stop_after;
      } else if (tile == 0) {
        if (lane == 0) fq_st_volatile(p.tile_status, FQ_TILE_PREFIX | (fq_u64)tile_total);
      } else {
        if (lane == 0) fq_st_volatile(p.tile_status + tile, FQ_TILE_AGG | (fq_u64)tile_total);
        fq_i64 look = (fq_i64)tile - 1;
        for (;;) {
          const fq_i64 idx = look - lane;
          fq_u64 s = FQ_TILE_PREFIX;  // virtual tiles before tile 0: prefix 0
          if (idx >= 0) {
            do { s = fq_ld_volatile(p.tile_status + idx); } while (FQ_TILE_FLAG(s) == 0);
          }
          const fq_u32 pm = __ballot_sync(0xffffffffu, FQ_TILE_FLAG(s) == 2);
          const int first = pm ? (__ffs(pm) - 1) : 32;
          fq_u64 contrib = (lane <= first) ? FQ_TILE_VALUE(s) : 0ull;
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, m);
          excl += contrib;
          if (pm) break;
          look -= 32;
        }
        if (lane == 0) fq_st_volatile(p.tile_status + tile, FQ_TILE_PREFIX | (excl + tile_total));
      }
      if (lane == 0) {
        s_tile_excl = excl;
        const fq_u64 incl = excl + tile_total;
        if (p.stop_after != 0 && incl >= p.stop_after) *(volatile fq_u32 *)p.done = 1u;
        if (tile == p.n_tiles - 1) p.result[0] = incl;
      }
    }
    __syncthreads();

    if (tile_total != 0 && !s_skip) {
      const fq_u64 base = s_tile_excl + warp_off;
#pragma unroll
      for (int u = 0; u < U; u++) {
        fq_u64 pos = base + rank[u];
#pragma unroll
        for (int v = 0; v < V; v++) {
          if ((keep >> (u * V + v)) & 1u) {
            if (pos < p.capacity) Q::emit(rows[u], v, p, pos, err);
            pos++;
          }
        }
      }
    }
  }
  if (err) atomicOr((fq_u32 *)(p.result + 1), err);
}

// ---------------------------------------------------------------------------------------------
// fq_map_kernel — projection of every row (no predicate): out_i[row] = expr_i(row)
// ---------------------------------------------------------------------------------------------
template <class Q, int UNROLL>
__device__ __forceinline__ void fq_map_kernel(const fq_launch_params &p) {
  constexpr int V = Q::V;
  fq_u32 err = 0;
  const fq_u64 nvec = p.n_rows / V;
  const fq_u64 chunk = (fq_u64)blockDim.x * UNROLL;
  const fq_u64 nfull = nvec / chunk;
  for (fq_u64 c = blockIdx.x; c < nfull; c += gridDim.x) {
    const fq_u64 g0 = c * chunk + threadIdx.x;
    typename Q::Rows rows[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) Q::load(rows[u], p, g0 + (fq_u64)u * blockDim.x);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const fq_u64 row0 = (g0 + (fq_u64)u * blockDim.x) * V;
#pragma unroll
      for (int v = 0; v < V; v++)
        if (row0 + v < p.capacity) Q::emit(rows[u], v, p, row0 + v, err);
    }
  }
  const fq_u64 tid = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x;
  const fq_u64 nthreads = (fq_u64)gridDim.x * blockDim.x;
  for (fq_u64 g = nfull * chunk + tid; g < nvec; g += nthreads) {
    typename Q::Rows r;
    Q::load(r, p, g);
#pragma unroll
    for (int v = 0; v < V; v++)
      if (g * V + v < p.capacity) Q::emit(r, v, p, g * V + v, err);
  }
  for (fq_u64 row = nvec * V + tid; row < p.n_rows; row += nthreads) {
    typename Q::Rows r;
    Q::load1(r, p, row);
    if (row < p.capacity) Q::emit(r, 0, p, row, err);
  }
  if (err) atomicOr((fq_u32 *)(p.result + 1), err);
  if (tid == 0) p.result[0] = p.n_rows;
}
