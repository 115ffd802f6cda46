// fq_sort.cuh — ORDER BY on the device: a stable LSD radix sort of row indexes.
//
// The reference lists sorting as not implemented (README.md:28 "[ ] Sorting"; sqlparser accepts ORDER BY and
// plan_parser.rs never reads `query.order_by`), so the semantics are stated here and in oracle/sort.py: keys are
// compared as arrow's sort kernels compare them with the crate's default SortOptions — ascending unless DESC, NULLs first
// — ties keep their input order (stable), floats are ordered by IEEE total order (-NaN < -inf < ... < -0 < +0 < ... < +inf
// < +NaN).  Several keys are sorted last key first, each sort stable, which is the lexicographic order.
//
// What is sorted are (code, row) pairs: `code` is an order-preserving unsigned image of the key (fq_sort_code), `row` a
// 32-bit row index — the payload columns are gathered once at the end (fq_take_kernel), they never ride through the passes.
// One pass handles one 8-bit digit: per-tile digit histograms, one exclusive scan over all (digit, tile) counters in
// digit-major order, and a scatter in which every warp ranks its rows with MATCH.ANY (lanes holding the same digit) against
// per-warp digit counters in shared memory: stable by construction.  Digits whose bits are the same in
// every code (AND == OR, folded during encoding) are skipped: numbers below 2^32 cost four passes, not eight.
// Bound: HBM; per pass 12 B read twice (histogram, scatter; the second read hits L2 for small inputs) + 12 B written per row.
#pragma once

#define FQ_SORT_THREADS 256
#define FQ_SORT_ITEMS 16
#define FQ_SORT_TILE (FQ_SORT_THREADS * FQ_SORT_ITEMS)        // pairs per CTA
#define FQ_SORT_WARPS (FQ_SORT_THREADS / 32)
#define FQ_SORT_WARP_ITEMS (FQ_SORT_TILE / FQ_SORT_WARPS)     // consecutive pairs owned by one warp
#define FQ_SCAN_THREADS 1024
#define FQ_SCAN_ITEMS 8
#define FQ_SCAN_TILE (FQ_SCAN_THREADS * FQ_SCAN_ITEMS)

// order-preserving unsigned code of slot `i` of a column (dtype tags of fuse_gpu.h)
__device__ __forceinline__ fq_u64 fq_sort_code(const void *col, int dtype, fq_u64 i) {
  switch (dtype) {
    case FQ_BOOL:
    case FQ_U8: return ((const fq_u8 *)col)[i];
    case FQ_U16: return ((const unsigned short *)col)[i];
    case FQ_U32: return ((const fq_u32 *)col)[i];
    case FQ_U64: return ((const fq_u64 *)col)[i];
    case FQ_I8: return (fq_u8)(((const signed char *)col)[i]) ^ 0x80u;
    case FQ_I16: return (unsigned short)(((const short *)col)[i]) ^ 0x8000u;
    case FQ_I32: return (fq_u32)(((const int *)col)[i]) ^ 0x80000000u;
    case FQ_I64: return (fq_u64)(((const fq_i64 *)col)[i]) ^ (1ull << 63);
    case FQ_F32: {
      const fq_u32 b = ((const fq_u32 *)col)[i];
      return (b >> 31) ? (fq_u32)~b : (b | 0x80000000u);
    }
    case FQ_F64: {
      const fq_u64 b = ((const fq_u64 *)col)[i];
      return (b >> 63) ? ~b : (b | (1ull << 63));
    }
  }
  return 0;
}

struct fq_sort_encode_params {
  const void *col;          // key column (values); nullptr with flags_only
  const fq_u8 *valid_bytes; // validity, one byte per row, or
  const void *valid_bits;   // an Arrow bitmap with row 0 at bit valid_bit0, or neither
  fq_u64 valid_bit0;
  const fq_u32 *perm;       // rows in their current order (nullptr = identity, and idx_out is written)
  fq_u64 *code_out;
  fq_u32 *idx_out;
  fq_u64 *and_or;           // [0] &= every code, [1] |= every code
  fq_u64 n;
  int dtype, bits;          // bits of the code that carry the key (8 * width)
  int descending;
  int flags_only;           // code = 1 for a valid slot, 0 for NULL (the NULLs-first pass)
};

__global__ void __launch_bounds__(256) fq_sort_encode(const __grid_constant__ fq_sort_encode_params a) {
  fq_u64 m_and = ~0ull, m_or = 0ull;
  const fq_u64 mask = a.bits >= 64 ? ~0ull : ((1ull << a.bits) - 1);
  for (fq_u64 i = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (fq_u64)gridDim.x * blockDim.x) {
    const fq_u64 row = a.perm ? a.perm[i] : i;
    fq_u64 code;
    if (a.flags_only) {
      code = a.valid_bytes ? (a.valid_bytes[row] ? 1u : 0u) : (fq_ld_bit(a.valid_bits, a.valid_bit0 + row) ? 1u : 0u);
    } else {
      code = fq_sort_code(a.col, a.dtype, row);
      if (a.descending) code = ~code & mask;
      // a NULL slot's value bytes are arbitrary: give all NULLs one code so that they keep their input order
      const bool ok = a.valid_bytes ? a.valid_bytes[row] != 0 : (a.valid_bits ? fq_ld_bit(a.valid_bits, a.valid_bit0 + row) : true);
      if (!ok) code = 0;
    }
    a.code_out[i] = code;
    if (!a.perm) a.idx_out[i] = (fq_u32)i;
    m_and &= code;
    m_or |= code;
  }
  for (int o = 16; o; o >>= 1) {
    m_and &= __shfl_xor_sync(0xffffffffu, m_and, o);
    m_or |= __shfl_xor_sync(0xffffffffu, m_or, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAnd((unsigned long long *)a.and_or, (unsigned long long)m_and);
    atomicOr((unsigned long long *)(a.and_or + 1), (unsigned long long)m_or);
  }
}

// The tile is read once, up front, into registers (FQ_SORT_ITEMS independent loads per thread in flight), and ranking is
// split so that no round waits for the one before it.  The first version ranked inside the scatter loop — MATCH, a read of
// the warp's counter, the stores, the leader's write-back, next round — and ncu showed what a chain of shared-memory round
// trips costs: 108 cycles per issued instruction, 58 % of them short-scoreboard stalls, 3.6 % issue-slot use, 10.8 ms for
// 2.5e8 rows.  Now the counting phase leaves, per row, its rank among the warp's earlier rows of the same digit (the value
// the leader's shared-memory atomicAdd returns, plus the row's position among its peers), and the scatter only adds the
// (tile, warp, digit) base: independent work per round.
// Warp w owns the tile's rows [w * WARP_ITEMS, (w + 1) * WARP_ITEMS); round r of a warp = 32 consecutive rows.
#define FQ_SORT_ROUNDS (FQ_SORT_WARP_ITEMS / 32)
__device__ __forceinline__ fq_u64 fq_sort_row(fq_u64 base, int r) {
  return base + (fq_u64)(threadIdx.x >> 5) * FQ_SORT_WARP_ITEMS + r * 32 + (threadIdx.x & 31);
}

// per-warp digit counts of this CTA's tile: cnt[w][d] = rows of warp w's run whose digit is d (digit 256 = past the end).
// RANKS: dg[r] becomes digit | rank << 16, rank = rows of the warp's run before this one with the same digit (rounds are
// issued in order by one warp, and shared-memory atomics of a warp on one address complete in issue order).
template <bool RANKS>
__device__ __forceinline__ void fq_sort_count(fq_u32 (*cnt)[256], fq_u32 (&dg)[FQ_SORT_ROUNDS]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < FQ_SORT_WARPS * 256; i += FQ_SORT_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < FQ_SORT_ROUNDS; r++) {
    const fq_u32 d = dg[r];
    const fq_u32 peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    if constexpr (RANKS) {
      fq_u32 before = 0;
      if (lane == leader && d < 256u) before = atomicAdd(&cnt[warp][d], (fq_u32)__popc(peers));
      before = __shfl_sync(0xffffffffu, before, leader);
      dg[r] = d | ((before + __popc(peers & ((1u << lane) - 1))) << 16);   // lanes in row order: stable
    } else {
      if (lane == leader && d < 256u) atomicAdd(&cnt[warp][d], (fq_u32)__popc(peers));
    }
  }
  __syncthreads();
}

// hist[d * n_tiles + tile] = rows of the tile whose digit is d (digit-major: one scan gives every (digit, tile) its base).
// Plain shared-memory atomics, one per row: no ranks are needed here, and native 32-bit ATOMS keep up even when every row
// of the tile has the same digit (the MATCH-based count of the scatter kernel made this pass 1.84 ms per 2.5e8 rows).
__global__ void __launch_bounds__(FQ_SORT_THREADS) fq_sort_hist(const fq_u64 *code, fq_u64 n, fq_u32 n_tiles, int shift, fq_u32 *hist) {
  __shared__ fq_u32 cnt[256];
  const fq_u64 base = (fq_u64)blockIdx.x * FQ_SORT_TILE;
  cnt[threadIdx.x] = 0;
  fq_u32 dg[FQ_SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < FQ_SORT_ROUNDS; r++) {
    const fq_u64 i = fq_sort_row(base, r);
    dg[r] = i < n ? (fq_u32)(code[i] >> shift) & 255u : 256u;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < FQ_SORT_ROUNDS; r++)
    if (dg[r] < 256u) atomicAdd(&cnt[dg[r]], 1u);
  __syncthreads();
  hist[(fq_u64)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// Dynamic shared memory of the scatter kernel: the tile's pairs in digit-major order, then the counters.
#define FQ_SORT_SMEM (FQ_SORT_TILE * 12 + FQ_SORT_WARPS * 256 * 4 + 256 * 4 + 32 * 4)

__device__ __forceinline__ fq_u32 fq_block_exclusive_scan(fq_u32 x, fq_u32 *warp_sums, fq_u32 &total);

// The tile is first ordered by digit in shared memory and then written out: the rows of one digit form one contiguous run
// in the output, so consecutive threads store consecutive addresses.  (Storing straight from the ranking loop sent every
// lane of a store to a different 32-byte sector: ncu counted 6 of 32 bytes used per sector, 4.8 GB written and 4.7 GB
// read for 3 GB each way, the extra reads being read-modify-write fills of partial sectors.)
__global__ void __launch_bounds__(FQ_SORT_THREADS, 3) fq_sort_scatter(const fq_u64 *code, const fq_u32 *idx, fq_u64 *code_out, fq_u32 *idx_out,
                                                                      const fq_u32 *hist_scanned, fq_u64 n, fq_u32 n_tiles, int shift) {
  extern __shared__ __align__(16) unsigned char fq_sort_smem[];
  fq_u64 *s_code = (fq_u64 *)fq_sort_smem;
  fq_u32 *s_row = (fq_u32 *)(s_code + FQ_SORT_TILE);
  fq_u32(*cnt)[256] = (fq_u32(*)[256])(s_row + FQ_SORT_TILE);
  fq_u32 *g_base = &cnt[0][0] + FQ_SORT_WARPS * 256;   // [256] global position of local slot 0 of the digit's run, minus the run's start
  fq_u32 *ws = g_base + 256;                           // [32]
  const int warp = threadIdx.x >> 5;
  const fq_u64 base = (fq_u64)blockIdx.x * FQ_SORT_TILE;
  fq_u64 c[FQ_SORT_ROUNDS];
  fq_u32 row[FQ_SORT_ROUNDS], dg[FQ_SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < FQ_SORT_ROUNDS; r++) {
    const fq_u64 i = fq_sort_row(base, r);
    const bool in = i < n;
    c[r] = in ? code[i] : 0;
    row[r] = in ? idx[i] : 0;
    dg[r] = in ? (fq_u32)(c[r] >> shift) & 255u : 256u;
  }
  fq_sort_count<true>(cnt, dg);
  {  // thread d: the tile's rows of digit d — each warp's share starts after the earlier warps', the run after the smaller digits
    fq_u32 total = 0;
#pragma unroll
    for (int w = 0; w < FQ_SORT_WARPS; w++) {
      const fq_u32 k = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = total;
      total += k;
    }
    fq_u32 all;
    const fq_u32 start = fq_block_exclusive_scan(total, ws, all);   // local slot of the run's first row
#pragma unroll
    for (int w = 0; w < FQ_SORT_WARPS; w++) cnt[w][threadIdx.x] += start;
    g_base[threadIdx.x] = hist_scanned[(fq_u64)threadIdx.x * n_tiles + blockIdx.x] - start;   // (mod 2^32)
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < FQ_SORT_ROUNDS; r++) {
    const fq_u32 d = dg[r] & 0xffffu;
    if (d < 256u) {
      const fq_u32 at = cnt[warp][d] + (dg[r] >> 16);
      s_code[at] = c[r];
      s_row[at] = row[r];
    }
  }
  __syncthreads();
  const fq_u32 in_tile = (fq_u32)(n - base < (fq_u64)FQ_SORT_TILE ? n - base : (fq_u64)FQ_SORT_TILE);
#pragma unroll
  for (int k = 0; k < FQ_SORT_ITEMS; k++) {
    const fq_u32 l = threadIdx.x + k * FQ_SORT_THREADS;
    if (l < in_tile) {
      const fq_u64 v = s_code[l];
      const fq_u32 at = g_base[(fq_u32)(v >> shift) & 255u] + l;
      code_out[at] = v;
      idx_out[at] = s_row[l];
    }
  }
}

// ---- exclusive scan of a u32 array (the digit-major histogram), in place: tile sums, scan of the sums, tile scans ----
__device__ __forceinline__ fq_u32 fq_block_exclusive_scan(fq_u32 x, fq_u32 *warp_sums /* [32] shared */, fq_u32 &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  fq_u32 inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const fq_u32 y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    fq_u32 s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const fq_u32 y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;   // inclusive over warps
  }
  __syncthreads();
  total = warp_sums[31];
  const fq_u32 before = warp ? warp_sums[warp - 1] : 0;
  __syncthreads();
  return before + inc - x;
}

__global__ void __launch_bounds__(FQ_SCAN_THREADS) fq_scan_tile_sums(const fq_u32 *a, fq_u64 m, fq_u32 *sums) {
  __shared__ fq_u32 ws[32];
  const fq_u64 base = (fq_u64)blockIdx.x * FQ_SCAN_TILE + (fq_u64)threadIdx.x * FQ_SCAN_ITEMS;
  fq_u32 s = 0;
#pragma unroll
  for (int k = 0; k < FQ_SCAN_ITEMS; k++) s += base + k < m ? a[base + k] : 0;
  fq_u32 total;
  fq_block_exclusive_scan(s, ws, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one CTA: sums[] -> exclusive prefix, in place
__global__ void __launch_bounds__(FQ_SCAN_THREADS) fq_scan_sums(fq_u32 *sums, fq_u32 nb) {
  __shared__ fq_u32 ws[32];
  fq_u32 carry = 0;
  for (fq_u32 b0 = 0; b0 < nb; b0 += FQ_SCAN_THREADS) {
    const fq_u32 i = b0 + threadIdx.x;
    const fq_u32 x = i < nb ? sums[i] : 0;
    fq_u32 total;
    const fq_u32 ex = fq_block_exclusive_scan(x, ws, total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(FQ_SCAN_THREADS) fq_scan_tiles(fq_u32 *a, fq_u64 m, const fq_u32 *sums) {
  __shared__ fq_u32 ws[32];
  const fq_u64 base = (fq_u64)blockIdx.x * FQ_SCAN_TILE + (fq_u64)threadIdx.x * FQ_SCAN_ITEMS;
  fq_u32 v[FQ_SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < FQ_SCAN_ITEMS; k++) {
    v[k] = base + k < m ? a[base + k] : 0;
    s += v[k];
  }
  fq_u32 total;
  fq_u32 run = sums[blockIdx.x] + fq_block_exclusive_scan(s, ws, total);
#pragma unroll
  for (int k = 0; k < FQ_SCAN_ITEMS; k++) {
    if (base + k < m) a[base + k] = run;
    run += v[k];
  }
}

// ---- ORDER BY ... LIMIT k: radix select ----
// The k smallest codes are found from the most significant digit down: a 256-bin histogram of the candidates' digit says
// in which bucket the k-th smallest lies; rows of smaller buckets are winners, rows of that bucket are the next round's
// candidates, everything else is dropped.  Two reads of the candidates per round, and the candidates shrink by up to
// 256x per round; what is left (winners + last candidates) is sorted by (code, row) with the passes above.
__global__ void __launch_bounds__(256) fq_topk_hist(const fq_u64 *code, fq_u64 m, int shift, fq_u32 *ghist) {
  __shared__ fq_u32 cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  for (fq_u64 i = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (fq_u64)gridDim.x * blockDim.x)
    atomicAdd(&cnt[(fq_u32)(code[i] >> shift) & 255u], 1u);
  __syncthreads();
  if (cnt[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], cnt[threadIdx.x]);
}

// winners (digit < bucket) -> win[], candidates (digit == bucket) -> (cand_code, cand_row); order inside each is arbitrary
// (one atomicAdd per warp and output; the final sort orders by code and row).  cursors: [0] winners, [1] candidates.
__global__ void __launch_bounds__(256) fq_topk_partition(const fq_u64 *code, const fq_u32 *rows, fq_u64 m, int shift, fq_u32 bucket, fq_u32 *win,
                                                         fq_u64 *cand_code, fq_u32 *cand_row, fq_u32 *cursors) {
  const int lane = threadIdx.x & 31;
  const fq_u64 warps = ((fq_u64)gridDim.x * blockDim.x) >> 5;
  for (fq_u64 i0 = (((fq_u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5; i0 < m; i0 += warps << 5) {
    const fq_u64 i = i0 + lane;
    const bool in = i < m;
    const fq_u64 c = in ? code[i] : 0;
    const fq_u32 d = (fq_u32)(c >> shift) & 255u;
    const bool is_win = in && d < bucket, is_cand = in && d == bucket;
    const fq_u32 mw = __ballot_sync(0xffffffffu, is_win), mc = __ballot_sync(0xffffffffu, is_cand);
    fq_u32 bw = 0, bc = 0;
    if (lane == 0) {
      if (mw) bw = atomicAdd(&cursors[0], (fq_u32)__popc(mw));
      if (mc) bc = atomicAdd(&cursors[1], (fq_u32)__popc(mc));
    }
    bw = __shfl_sync(0xffffffffu, bw, 0);
    bc = __shfl_sync(0xffffffffu, bc, 0);
    const fq_u32 below = (1u << lane) - 1;
    if (is_win) win[bw + __popc(mw & below)] = rows[i];
    if (is_cand) {
      const fq_u32 at = bc + __popc(mc & below);
      cand_code[at] = c;
      cand_row[at] = rows[i];
    }
  }
}

// code = row index: the first sort of the final set, so that the stable key passes that follow break ties by input order
__global__ void __launch_bounds__(256) fq_sort_rows_as_codes(const fq_u32 *rows, fq_u64 *code, fq_u64 m) {
  for (fq_u64 i = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (fq_u64)gridDim.x * blockDim.x) code[i] = rows[i];
}

// ---- gather: out[i] = src[rows[i]] (values of `width` bytes; validity bytes or bits -> validity bytes) ----
struct fq_take_params {
  const void *src;
  void *out;
  const fq_u32 *rows;
  fq_u64 n;
  int width;
  const fq_u8 *valid_bytes;
  const void *valid_bits;
  fq_u64 valid_bit0;
  fq_u8 *out_valid;
};

__global__ void __launch_bounds__(256) fq_take_kernel(const __grid_constant__ fq_take_params a) {
  for (fq_u64 i = (fq_u64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (fq_u64)gridDim.x * blockDim.x) {
    const fq_u64 row = a.rows[i];
    switch (a.width) {
      case 1: ((fq_u8 *)a.out)[i] = ((const fq_u8 *)a.src)[row]; break;
      case 2: ((unsigned short *)a.out)[i] = ((const unsigned short *)a.src)[row]; break;
      case 4: ((fq_u32 *)a.out)[i] = ((const fq_u32 *)a.src)[row]; break;
      default: ((fq_u64 *)a.out)[i] = ((const fq_u64 *)a.src)[row]; break;
    }
    if (a.out_valid)
      a.out_valid[i] = a.valid_bytes ? (a.valid_bytes[row] ? 1 : 0) : (a.valid_bits ? (fq_ld_bit(a.valid_bits, a.valid_bit0 + row) ? 1 : 0) : 1);
  }
}
