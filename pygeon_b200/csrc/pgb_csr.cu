// K2: deterministic COO -> CSR.
//
// symbolic (default, two level): transpose of the cell->dof table by counting (k_inc_count, scan,
//             k_inc_place), then one row kernel (pgb_rowthread.cu / pgb_rowhash.cu / pgb_rowsort.cu)
//             that sorts each row's incidences, finds the sorted distinct columns and the slot of
//             every candidate; scan of the row counts; k_compact.
// symbolic (fallback, mode 1): stable LSD radix sort (8-bit digits) of the packed (row,col) keys
//             with the COO position as payload, then a head-flag compaction that emits indptr /
//             indices / run offsets.  Keys of the first pass are generated on the fly.
// numeric   : k_numeric_t streams the row-sorted local rows (two-level plans); k_numeric reduces
//             vals[perm[p]] over the run of every entry with a warp-shuffle segmented scan (mode 1).
// No floating-point atomics anywhere.  Integer atomics (counters, histogram bins, hash keys, max) are
// used only where the result that reaches the output does not depend on their order, so pattern
// and values are bit-reproducible.
#include "pgb_common.cuh"
#include "pgb_rows.cuh"

#include <vector>

using namespace pgb_rows;

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_BLOCK = 256;
constexpr int SORT_WARPS = SORT_BLOCK / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_BLOCK * SORT_ITEMS;  // 4096 keys per block

constexpr int SCAN_G = PGB_NPART;  // chunks of the two-level scans (<= 2048)

// Key source of a radix pass.  KeyT = uint64_t: packed (row,col) of COO entry i (fallback path);
// KeyT = uint32_t: the dof id of incidence i = cell*L + a (incidence transpose).  In the first
// pass `keys` is null and the key is generated from cell_dofs on the fly.
template <typename KeyT>
struct KeySrc {
  const KeyT* keys;
  const int32_t* dofs;
  uint32_t L, LL;
  int cb;  // column bits
};

__device__ __forceinline__ uint64_t load_key(const KeySrc<uint64_t>& s, uint32_t i) {
  if (s.keys) return s.keys[i];
  uint32_t cell = i / s.LL;
  uint32_t r = i - cell * s.LL;
  uint32_t a = r / s.L, b = r - a * s.L;
  const int32_t* d = s.dofs + (int64_t)cell * s.L;
  return ((uint64_t)(uint32_t)__ldg(d + a) << s.cb) | (uint64_t)(uint32_t)__ldg(d + b);
}

__device__ __forceinline__ uint32_t load_key(const KeySrc<uint32_t>& s, uint32_t i) {
  if (s.keys) return s.keys[i];
  return (uint32_t)__ldg(s.dofs + i);
}

// ---- pass 1/3: per-tile digit histogram, written digit-major ---------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(SORT_BLOCK) k_radix_hist(KeySrc<KeyT> src, uint32_t n, int shift, uint32_t ntiles,
                                                           uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * (uint32_t)SORT_TILE;
#pragma unroll 4
  for (int i = 0; i < SORT_ITEMS; ++i) {
    uint32_t idx = base + i * SORT_BLOCK + threadIdx.x;
    if (idx < n) {
      uint32_t d = (uint32_t)(load_key(src, idx) >> shift) & (RADIX - 1);
      atomicAdd(&h[d], 1u);  // integer count: order independent (measured faster than match.any aggregation)
    }
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// ---- two-level exclusive scan over uint32 (reduce / scan partials / apply) -------------------
// exclusive scan of one value per thread over a 256-thread block; returns exclusive prefix,
// total in `total`.  ws: 8 uint32 of shared memory.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t& total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  __syncthreads();
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i) {
    uint32_t t = ws[i];
    if (i < w) off += t;
    tot += t;
  }
  total = tot;
  return off + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_reduce(const uint32_t* __restrict__ in, uint64_t n, uint64_t chunk,
                                                     uint32_t* __restrict__ parts) {
  __shared__ uint32_t ws[8];
  uint64_t lo = blockIdx.x * chunk, hi = lo + chunk;
  if (hi > n) hi = n;
  uint32_t s = 0;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += 256) s += in[i];
  uint32_t tot;
  block_excl_scan(s, ws, tot);
  if (threadIdx.x == 0) parts[blockIdx.x] = tot;
}

// single block: exclusive scan of parts[0..g) in place (g <= 2048); total -> *total_out (may be null)
__global__ void __launch_bounds__(256) k_scan_parts(uint32_t* parts, int g, uint32_t* total_out) {
  __shared__ uint32_t ws[8];
  uint32_t v[8];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = threadIdx.x * 8 + j;
    v[j] = (i < g) ? parts[i] : 0;
    s += v[j];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, ws, tot);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = threadIdx.x * 8 + j;
    if (i < g) parts[i] = ex;
    ex += v[j];
  }
  if (threadIdx.x == 0 && total_out) *total_out = tot;
}

__global__ void __launch_bounds__(256) k_scan_apply(uint32_t* __restrict__ data, uint64_t n, uint64_t chunk,
                                                    const uint32_t* __restrict__ parts) {
  __shared__ uint32_t ws[8];
  uint64_t lo = blockIdx.x * chunk, hi = lo + chunk;
  if (hi > n) hi = n;
  uint32_t carry = parts[blockIdx.x];
  for (uint64_t t0 = lo; t0 < hi; t0 += 1024) {
    uint64_t i0 = t0 + (uint64_t)threadIdx.x * 4;
    uint32_t v[4];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = (i0 + j < hi) ? data[i0 + j] : 0;
      s += v[j];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, ws, tot) + carry;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < hi) data[i0 + j] = ex;
      ex += v[j];
    }
    carry += tot;
  }
}

// ---- pass 3/3: stable scatter ---------------------------------------------------------------
template <typename KeyT>
struct ScatterSmem {
  uint32_t whist[SORT_WARPS][RADIX];
  uint32_t dstart[RADIX];
  uint32_t goff[RADIX];
  uint32_t ws[8];
  uint32_t svals[SORT_TILE];
  KeyT skeys[SORT_TILE];
};

template <typename KeyT>
__global__ void __launch_bounds__(SORT_BLOCK) k_radix_scatter(KeySrc<KeyT> src, const uint32_t* __restrict__ vals_in, uint32_t n,
                                                              int shift, uint32_t ntiles,
                                                              const uint32_t* __restrict__ hist_scanned,
                                                              KeyT* __restrict__ keys_out,
                                                              uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KeyT>& sm = *reinterpret_cast<ScatterSmem<KeyT>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t tile0 = blockIdx.x * (uint32_t)SORT_TILE;
  const uint32_t count = (n - tile0 < (uint32_t)SORT_TILE) ? (n - tile0) : (uint32_t)SORT_TILE;

#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i) sm.whist[i][tid] = 0;
  sm.goff[tid] = hist_scanned[(uint64_t)tid * ntiles + blockIdx.x];
  __syncthreads();

  KeyT key[SORT_ITEMS];
  uint16_t rank[SORT_ITEMS];
  const uint32_t wbase = tile0 + w * (32 * SORT_ITEMS);
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    uint32_t idx = wbase + i * 32 + lane;
    key[i] = (idx < n) ? load_key(src, idx) : (KeyT)~(KeyT)0;
  }
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    uint32_t idx = wbase + i * 32 + lane;
    bool valid = idx < n;
    uint32_t d = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    uint32_t peers = __match_any_sync(FULL, valid ? d : RADIX);
    uint32_t prev = valid ? sm.whist[w][d] : 0;
    __syncwarp();
    uint32_t r = __popc(peers & lt);
    if (valid && r == 0) sm.whist[w][d] = prev + __popc(peers);
    __syncwarp();
    rank[i] = (uint16_t)(prev + r);
  }
  __syncthreads();
  // per-digit exclusive scan over warps, then over digits
  uint32_t run = 0;
#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i) {
    uint32_t t = sm.whist[i][tid];
    sm.whist[i][tid] = run;
    run += t;
  }
  uint32_t tot;
  uint32_t ds = block_excl_scan(run, sm.ws, tot);
  sm.dstart[tid] = ds;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    uint32_t idx = wbase + i * 32 + lane;
    if (idx < n) {
      uint32_t d = (uint32_t)(key[i] >> shift) & (RADIX - 1);
      uint32_t pos = sm.dstart[d] + sm.whist[w][d] + rank[i];
      sm.skeys[pos] = key[i];
      sm.svals[pos] = vals_in ? vals_in[idx] : idx;
    }
  }
  __syncthreads();
  for (uint32_t j = tid; j < count; j += SORT_BLOCK) {
    KeyT k = sm.skeys[j];
    uint32_t d = (uint32_t)(k >> shift) & (RADIX - 1);
    uint32_t g = sm.goff[d] + (j - sm.dstart[d]);
    keys_out[g] = k;
    vals_out[g] = sm.svals[j];
  }
}

// ---- unique: count heads per chunk, then emit -------------------------------------------------
__device__ __forceinline__ bool is_head(const uint64_t* keys, uint64_t i) {
  return i == 0 || keys[i] != keys[i - 1];
}

__global__ void __launch_bounds__(256) k_count_heads(const uint64_t* __restrict__ keys, uint64_t n, uint64_t chunk,
                                                     uint32_t* __restrict__ parts) {
  __shared__ uint32_t ws[8];
  uint64_t lo = blockIdx.x * chunk, hi = lo + chunk;
  if (hi > n) hi = n;
  uint32_t s = 0;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += 256) s += is_head(keys, i) ? 1u : 0u;
  uint32_t tot;
  block_excl_scan(s, ws, tot);
  if (threadIdx.x == 0) parts[blockIdx.x] = tot;
}

template <typename IP>
__global__ void __launch_bounds__(256) k_emit_unique(const uint64_t* __restrict__ keys, uint64_t n, uint64_t chunk,
                                                     const uint32_t* __restrict__ parts, int cb, int64_t n_rows,
                                                     uint32_t nnz, IP* __restrict__ indptr,
                                                     int32_t* __restrict__ indices, uint32_t* __restrict__ seg) {
  __shared__ uint32_t ws[8];
  uint64_t lo = blockIdx.x * chunk, hi = lo + chunk;
  if (hi > n) hi = n;
  uint32_t carry = parts[blockIdx.x];
  const uint64_t cmask = (1ull << cb) - 1;
  for (uint64_t t0 = lo; t0 < hi; t0 += 1024) {
    uint64_t i0 = t0 + (uint64_t)threadIdx.x * 4;
    uint64_t k[5];
    k[0] = (i0 > 0 && i0 - 1 < n) ? keys[i0 - 1] : ~0ull;
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j + 1] = (i0 + j < hi) ? keys[i0 + j] : 0;
    bool h[4];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = (i0 + j < hi) && (i0 + j == 0 || k[j + 1] != k[j]);
      s += h[j];
    }
    uint32_t tot;
    uint32_t pos = block_excl_scan(s, ws, tot) + carry;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (h[j]) {
        uint64_t i = i0 + j;
        int64_t row = (int64_t)(k[j + 1] >> cb);
        indices[pos] = (int32_t)(k[j + 1] & cmask);
        seg[pos] = (uint32_t)i;
        int64_t prev_row = (i == 0) ? -1 : (int64_t)(k[j] >> cb);
        for (int64_t r = prev_row + 1; r <= row; ++r) indptr[r] = (IP)pos;
        ++pos;
      }
      if (i0 + j == n - 1 && i0 + j < hi) {
        int64_t row = (int64_t)(k[j + 1] >> cb);
        for (int64_t r = row + 1; r <= n_rows; ++r) indptr[r] = (IP)nnz;
        seg[nnz] = (uint32_t)n;
      }
    }
    carry += tot;
  }
}

template <typename IP>
__global__ void k_fill_indptr_empty(int64_t n_rows, IP* indptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) indptr[i] = 0;
}

// ---- two-level symbolic: incidence transpose + per-row register sort ---------------------------
// After the incidences (cell, a) have been sorted by dof id, row r of the matrix owns the
// incidences [row_start[r], row_start[r+1]); each contributes the L candidates
// (col = dofs[cell][b], src = inc*L + b).  In the globally sorted COO order row r therefore
// occupies exactly the positions [L*row_start[r], L*row_start[r+1]), so perm can be written in
// place once the row's candidates are sorted by (col, src).

__global__ void __launch_bounds__(256) k_row_starts(const uint32_t* __restrict__ keys, uint32_t n, int64_t n_rows,
                                                    uint32_t* __restrict__ row_start) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  uint32_t prev = (i == 0) ? 0xffffffffu : keys[i - 1];
  if (k != prev)
    for (uint32_t r = prev + 1u; r <= k; ++r) row_start[r] = i;  // prev + 1 wraps to 0 for i == 0
  if (i == n - 1)
    for (int64_t r = (int64_t)k + 1; r <= n_rows; ++r) row_start[r] = n;
}

// max number of incidences of one dof (integer max: order independent)
__global__ void __launch_bounds__(256) k_max_row(int64_t n_rows, const uint32_t* __restrict__ row_start,
                                                 uint32_t* __restrict__ max_count) {
  __shared__ uint32_t ws[8];
  uint32_t mx = 0;
  for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * 256) {
    uint32_t c = row_start[r + 1] - row_start[r];
    mx = c > mx ? c : mx;
  }
  mx = __reduce_max_sync(FULL, mx);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) ws[w] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) mx = ws[q] > mx ? ws[q] : mx;
    atomicMax(max_count, mx);
  }
}

// Largest incidence count of every chunk of CHUNK_ROWS rows (and of the matrix), and how many rows of
// the chunk fall into each size class (class k: 2^(k-1) < incidences <= 2^k; k = 0: at most one).
// Chunks whose rows are all short go to k_rowthread as a contiguous range; the rows of the other
// chunks are gathered into one row list per size class, so that every row gets the cheapest kernel
// that takes it, wherever the long rows are.
constexpr int CHUNK_ROWS = 4096;  // multiple of ROWTHREAD_GROUP
constexpr int NCLS = 12;          // size classes up to 2^11 = LONG_MAX_INC incidences
__device__ __forceinline__ int size_class(uint32_t cnt) { return cnt <= 1u ? 0 : 32 - __clz(cnt - 1u); }

__global__ void __launch_bounds__(256) k_chunk_max(int64_t n_rows, const uint32_t* __restrict__ row_start,
                                                   uint32_t* __restrict__ chunk_max, uint32_t* __restrict__ chunk_hist,
                                                   uint32_t* __restrict__ chunk_cmax, uint32_t* __restrict__ max_count) {
  __shared__ uint32_t ws[8];
  __shared__ uint32_t hist[NCLS], cmax[NCLS];
  if (threadIdx.x < NCLS) hist[threadIdx.x] = 0, cmax[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * CHUNK_ROWS;
  const int64_t r1 = r0 + CHUNK_ROWS < n_rows ? r0 + CHUNK_ROWS : n_rows;
  uint32_t mx = 0;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
    uint32_t c = row_start[r + 1] - row_start[r];
    mx = c > mx ? c : mx;
    int k = size_class(c);
    k = k < NCLS ? k : NCLS - 1;
    atomicAdd(&hist[k], 1u);  // integer count / max: order independent
    atomicMax(&cmax[k], c);
  }
  mx = __reduce_max_sync(FULL, mx);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) ws[w] = mx;
  __syncthreads();
  if (threadIdx.x < NCLS) {
    chunk_hist[(int64_t)blockIdx.x * NCLS + threadIdx.x] = hist[threadIdx.x];
    chunk_cmax[(int64_t)blockIdx.x * NCLS + threadIdx.x] = cmax[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) mx = ws[q] > mx ? ws[q] : mx;
    chunk_max[blockIdx.x] = mx;
    atomicMax(max_count, mx);
  }
}

// Scatter the rows of the listed chunks into the per-class row lists.  start[chunk*NCLS + k]: where
// the class-k rows of the chunk begin in `list` (0xffffffff in [chunk*NCLS]: chunk not listed).  The
// order inside a (chunk, class) block is arbitrary: every row kernel writes to row-owned locations.
__global__ void __launch_bounds__(256) k_row_lists(int64_t n_rows, const uint32_t* __restrict__ row_start,
                                                   const uint32_t* __restrict__ start, uint32_t* __restrict__ list) {
  __shared__ uint32_t pos[NCLS];
  if (start[(int64_t)blockIdx.x * NCLS] == 0xffffffffu) return;
  if (threadIdx.x < NCLS) pos[threadIdx.x] = start[(int64_t)blockIdx.x * NCLS + threadIdx.x];
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * CHUNK_ROWS;
  const int64_t r1 = r0 + CHUNK_ROWS < n_rows ? r0 + CHUNK_ROWS : n_rows;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
    int k = size_class(row_start[r + 1] - row_start[r]);
    list[atomicAdd(&pos[k < NCLS ? k : NCLS - 1], 1u)] = (uint32_t)r;
  }
}

// ---- incidence transpose by counting (default) ---------------------------------------------------
// (1) k_inc_count: cnt[dof] += 1 for every incidence; the value the integer atomic returns is kept
//     as a provisional rank inside the row (it depends on the execution order, the counts do not).
// (2) exclusive scan of cnt -> row_start.
// (3) k_inc_place: incidence i goes to row_start[dof] + rank  (a permutation of the row's block).
// (4) k_row_canon: every row sorts its (short) incidence list ascending, which is exactly the order
//     a stable sort by dof produces: inc_sorted / inc_pos are bit-identical to the radix path and
//     to themselves from run to run.  The floating-point summation order downstream is therefore
//     fixed; the atomics only ever touch integer counters.
__global__ void __launch_bounds__(256) k_inc_count(uint32_t n_inc, const int32_t* __restrict__ dofs,
                                                   uint32_t* __restrict__ cnt, uint8_t* __restrict__ rank) {
  const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
  if (i0 >= n_inc) return;
  if (i0 + 4 <= n_inc) {
    int4 d = *reinterpret_cast<const int4*>(dofs + i0);
    uint32_t r0 = atomicAdd(cnt + d.x, 1u), r1 = atomicAdd(cnt + d.y, 1u);
    uint32_t r2 = atomicAdd(cnt + d.z, 1u), r3 = atomicAdd(cnt + d.w, 1u);
    uchar4 o;
    o.x = (uint8_t)min(r0, 255u), o.y = (uint8_t)min(r1, 255u);
    o.z = (uint8_t)min(r2, 255u), o.w = (uint8_t)min(r3, 255u);
    *reinterpret_cast<uchar4*>(rank + i0) = o;
  } else {
    for (uint32_t i = i0; i < n_inc; ++i) rank[i] = (uint8_t)min(atomicAdd(cnt + dofs[i], 1u), 255u);
  }
}

// Ranks saturate at 255: the incidences of a dof beyond its first 255 draw their position from a
// second counter (ovf [n_rows], zeroed; may be null when no dof has more than 255 incidences).
__device__ __forceinline__ uint32_t place_pos(const uint32_t* __restrict__ row_start, uint32_t* __restrict__ ovf,
                                              int32_t d, uint32_t rank) {
  uint32_t p = __ldg(row_start + d) + rank;
  if (rank == 255u && ovf) p += atomicAdd(ovf + d, 1u);
  return p;
}

__global__ void __launch_bounds__(256) k_inc_place(uint32_t n_inc, const int32_t* __restrict__ dofs,
                                                   const uint32_t* __restrict__ row_start,
                                                   const uint8_t* __restrict__ rank, uint32_t* __restrict__ ovf,
                                                   uint32_t* __restrict__ inc_tmp) {
  const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
  if (i0 >= n_inc) return;
  if (i0 + 4 <= n_inc) {
    int4 d = *reinterpret_cast<const int4*>(dofs + i0);
    uchar4 k = *reinterpret_cast<const uchar4*>(rank + i0);
    uint32_t p0 = place_pos(row_start, ovf, d.x, k.x), p1 = place_pos(row_start, ovf, d.y, k.y);
    uint32_t p2 = place_pos(row_start, ovf, d.z, k.z), p3 = place_pos(row_start, ovf, d.w, k.w);
    inc_tmp[p0] = i0;
    inc_tmp[p1] = i0 + 1;
    inc_tmp[p2] = i0 + 2;
    inc_tmp[p3] = i0 + 3;
  } else {
    for (uint32_t i = i0; i < n_inc; ++i) inc_tmp[place_pos(row_start, ovf, dofs[i], rank[i])] = i;
  }
}

// G lanes per row, NPL incidences per lane (capacity G*NPL)
template <int G, int NPL>
__global__ void __launch_bounds__(256) k_row_canon(const uint32_t* __restrict__ rows, int64_t row0, int64_t n_rows,
                                                   const uint32_t* __restrict__ row_start,
                                                   const uint32_t* __restrict__ inc_tmp,
                                                   uint32_t* __restrict__ inc_sorted, uint32_t* __restrict__ inc_pos) {
  const int gl = threadIdx.x & (G - 1);
  bool live;
  const int64_t r = pick_row(rows, row0, n_rows, ((int64_t)blockIdx.x * 256 + threadIdx.x) / G, live);
  uint32_t rs = 0, cnt = 0;
  if (live) {
    rs = row_start[r];
    cnt = row_start[r + 1] - rs;
  }
  uint32_t k[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    uint32_t e = (uint32_t)(i * G + gl);  // any initial placement will do: coalesced
    k[i] = (e < cnt) ? inc_tmp[rs + e] : 0xffffffffu;
  }
  warp_bitonic_u32<NPL, G>(k, gl);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    uint32_t e = (uint32_t)(gl * NPL + i);
    if (e < cnt) {
      inc_sorted[rs + e] = k[i];
      inc_pos[k[i]] = rs + e;
    }
  }
}

// row_nnz has been exclusively scanned in place (row_off): copy each row's unique columns from
// their scratch position to the final CSR arrays.
// Rows [row0, row1) whose distinct columns sit at ucol_tmp[row_start[r]*L ...) (k_rowsort / k_rowhash).
// The range that ends the matrix also writes indptr[n_rows].
template <typename IP>
__global__ void __launch_bounds__(256) k_compact(int64_t row0, int64_t row1, int64_t n_rows, int L,
                                                 const uint32_t* __restrict__ row_start,
                                                 const uint32_t* __restrict__ row_off,
                                                 const int32_t* __restrict__ ucol_tmp, IP* __restrict__ indptr,
                                                 int32_t* __restrict__ indices) {
  // 8 lanes per row
  const int gl = threadIdx.x & 7;
  const int64_t r = row0 + (((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3);
  if (r > row1 || (r == row1 && row1 != n_rows)) return;
  uint32_t o0 = row_off[r];
  if (gl == 0) indptr[r] = (IP)o0;
  if (r == n_rows) return;
  uint32_t cnt = row_off[r + 1] - o0;
  const uint64_t base = (uint64_t)row_start[r] * (uint32_t)L;
  for (uint32_t k = gl; k < cnt; k += 8) indices[o0 + k] = ucol_tmp[base + k];
}

// Rows [row0, row1) written by k_rowthread (row0 a multiple of the group size): the distinct columns
// of every group of `grp` consecutive rows are already back to back, so the compaction is one
// contiguous copy per group.
template <typename IP>
__global__ void __launch_bounds__(256) k_compact_groups(int64_t row0, int64_t row1, int64_t n_rows, int L, uint32_t grp,
                                                        const uint32_t* __restrict__ row_start,
                                                        const uint32_t* __restrict__ row_off,
                                                        const int32_t* __restrict__ ucol_tmp, IP* __restrict__ indptr,
                                                        int32_t* __restrict__ indices) {
  const int64_t first = row0 + (int64_t)blockIdx.x * grp;
  const int64_t last = first + grp < row1 ? first + grp : row1;
  const uint64_t src0 = (uint64_t)row_start[first] * (uint32_t)L;
  const uint32_t dst0 = row_off[first];
  const uint32_t len = row_off[last] - dst0;
  for (uint32_t j = threadIdx.x; j < len; j += 256) indices[dst0 + j] = ucol_tmp[src0 + j];
  for (int64_t r = first + threadIdx.x; r < last; r += 256) indptr[r] = (IP)row_off[r];
  if (last == n_rows && threadIdx.x == 0) indptr[n_rows] = (IP)row_off[n_rows];
}

// Rows [row0, row1) written by k_rowins (row0 a multiple of the group size): like k_compact_groups, except that
// a group with flag != 0 keeps one segment per row (a row of it was handed to k_rowhash_list).
template <typename IP>
__global__ void __launch_bounds__(256) k_compact_ins(int64_t row0, int64_t row1, int64_t n_rows, int L, uint32_t grp,
                                                     const uint32_t* __restrict__ row_start,
                                                     const uint32_t* __restrict__ row_off,
                                                     const uint8_t* __restrict__ flag,
                                                     const int32_t* __restrict__ ucol_tmp, IP* __restrict__ indptr,
                                                     int32_t* __restrict__ indices) {
  const int64_t first = row0 + (int64_t)blockIdx.x * grp;
  const int64_t last = first + grp < row1 ? first + grp : row1;
  const uint32_t dst0 = row_off[first];
  if (!flag[first]) {  // the flag is the same for every row of a group
    const uint64_t src0 = (uint64_t)row_start[first] * (uint32_t)L;
    const uint32_t len = row_off[last] - dst0;
    for (uint32_t j = threadIdx.x; j < len; j += 256) indices[dst0 + j] = ucol_tmp[src0 + j];
  } else {
    const int gl = threadIdx.x & 7;
    for (int64_t r = first + (threadIdx.x >> 3); r < last; r += 32) {
      const uint32_t o0 = row_off[r], cnt = row_off[r + 1] - o0;
      const uint64_t base = (uint64_t)row_start[r] * (uint32_t)L;
      for (uint32_t k = gl; k < cnt; k += 8) indices[o0 + k] = ucol_tmp[base + k];
    }
  }
  for (int64_t r = first + threadIdx.x; r < last; r += 256) indptr[r] = (IP)row_off[r];
  if (last == n_rows && threadIdx.x == 0) indptr[n_rows] = (IP)row_off[n_rows];
}

// ---- numeric over row-sorted local rows (plans of the two-level symbolic) ------------------------
// With such a plan K1 writes local row a of cell c straight to position inc_pos[c*L + a] of vals_t,
// i.e. grouped by matrix row in ascending cell order (the order of the incidence transpose).  A CSR
// row r then owns the CONTIGUOUS block vals_t[row_start[r]*L .. row_start[r+1]*L) and the uint8
// column slots at the same positions.  One thread per row streams its block with 256-bit loads
// (every 64-byte HBM atom is consumed, no gathers at all) and adds each value into the row's
// accumulator acc[slot][thread] in shared memory (conflict-free).  Sequential in cell order => the
// oracle's summation order, bit-reproducible, no atomics.  Reads per COO entry: 8 B + 1 B.
__global__ void __launch_bounds__(256) k_inc_pos(uint32_t n_inc, const uint32_t* __restrict__ inc_sorted,
                                                 uint32_t* __restrict__ inc_pos) {
  uint32_t p = blockIdx.x * 256u + threadIdx.x;
  if (p < n_inc) inc_pos[inc_sorted[p]] = p;
}

template <int L>
__device__ __forceinline__ void load_row(const double* __restrict__ p, double (&v)[L]) {
  if constexpr (L % 4 == 0) {
#pragma unroll
    // LDG.E.256, read once: no L1 allocation (P1 128^3: 0.50 -> 0.43 ms, N0 1.97 -> 1.62 ms).  The same
    // hint on the narrow loads (slot words here, SpMV values / indices, K1 ids) is 20-140 % SLOWER.
    for (int b = 0; b < L; b += 4)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(v[b]), "=d"(v[b + 1]), "=d"(v[b + 2]), "=d"(v[b + 3])
                   : "l"(p + b));
  } else if constexpr (L % 2 == 0) {
#pragma unroll
    for (int b = 0; b < L; b += 2) {
      double2 t = __ldg(reinterpret_cast<const double2*>(p + b));
      v[b] = t.x;
      v[b + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int b = 0; b < L; ++b) v[b] = __ldg(p + b);
  }
}

template <int L>
__device__ __forceinline__ void load_slots(const uint8_t* __restrict__ p, uint32_t (&sl)[L]) {
  if constexpr (L % 4 == 0) {
#pragma unroll
    for (int b = 0; b < L; b += 4) {
      uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + b));
      sl[b] = w & 0xffu;
      sl[b + 1] = (w >> 8) & 0xffu;
      sl[b + 2] = (w >> 16) & 0xffu;
      sl[b + 3] = w >> 24;
    }
  } else if constexpr (L % 2 == 0) {
#pragma unroll
    for (int b = 0; b < L; b += 2) {
      uint32_t w = __ldg(reinterpret_cast<const uint16_t*>(p + b));
      sl[b] = w & 0xffu;
      sl[b + 1] = w >> 8;
    }
  } else {
#pragma unroll
    for (int b = 0; b < L; ++b) sl[b] = __ldg(p + b);
  }
}

// STAGED: the row's entries are transposed through a second shared-memory buffer and written with
// coalesced stores (a per-thread store costs one L1 wavefront per lane: rows are 8*cnt bytes apart).
// Pays when the matrix has about as many entries as COO candidates (RT0, BDM1, N0, the Darcy saddle:
// 5-15 % faster); for P1 in 3D (6.3 candidates per entry) the doubled shared memory takes L1 capacity
// away from the loads and the kernel gets 20 % slower, so the caller chooses (flags bit 0).
// (Fetching the slot bytes of 4 local rows with one 16-byte load, rounds aligned to 16 bytes of the
// slot array, was measured slower for every space.)
template <int L, int BLK, typename IP, bool STAGED, int UV = 0>
__global__ void __launch_bounds__(BLK, UV == 2 ? 1536 / BLK : 0) k_numeric_t(int64_t n_rows, int caps, const uint32_t* __restrict__ row_start,
                                                   const uint8_t* __restrict__ slots, const IP* __restrict__ indptr,
                                                   const double* __restrict__ vals_t, double* __restrict__ data) {
  constexpr int U = UV ? UV : ((L <= 4) ? 4 : 2);  // local rows in flight per thread (8 measured slower)
  constexpr int LP = (L + 3) & ~3;      // row pitch (values and slots)
  extern __shared__ double smem[];
  double* acc = smem;  // [slot][thread]
  __shared__ IP s_o0;
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * BLK;
  const int64_t r = r0 + tid;
  const bool live = r < n_rows;
  if constexpr (!STAGED)
    if (!live) return;
  uint32_t rs = 0, re = 0, cnt = 0;
  IP o0 = 0;
  if (live) {
    rs = row_start[r];
    re = row_start[r + 1];
    o0 = indptr[r];
    cnt = (uint32_t)(indptr[r + 1] - o0);
  }
  for (uint32_t s2 = 0; s2 < cnt; ++s2) acc[s2 * BLK + tid] = 0.0;
  for (uint32_t q0 = rs; q0 < re; q0 += U) {
    double v[U][LP];
    uint32_t sl[U][LP];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (q0 + u < re) {
        load_row<LP>(vals_t + (uint64_t)(q0 + u) * LP, v[u]);
        load_slots<LP>(slots + (uint64_t)(q0 + u) * LP, sl[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (q0 + u < re) {
#pragma unroll
        for (int b = 0; b < L; ++b) acc[sl[u][b] * BLK + tid] += v[u][b];
      }
    }
  }
  if constexpr (!STAGED) {
    for (uint32_t s2 = 0; s2 < cnt; ++s2) data[o0 + s2] = acc[s2 * BLK + tid];
  } else {
    double* outs = smem + (size_t)caps * BLK;  // the block's entries in CSR order
    if (tid == 0) s_o0 = o0;
    __syncthreads();
    const IP ob = s_o0;
    const uint32_t off = (uint32_t)(o0 - ob);
    for (uint32_t s2 = 0; s2 < cnt; ++s2) outs[off + s2] = acc[s2 * BLK + tid];
    __syncthreads();
    const int64_t r1 = r0 + BLK < n_rows ? r0 + BLK : n_rows;
    const uint32_t total = (uint32_t)(indptr[r1] - ob);
    for (uint32_t j = tid; j < total; j += BLK) data[ob + j] = outs[j];
  }
}

// ---- numeric: lane-sequential partial sums + warp-shuffle segmented scan ------------------------
// A warp owns NUM_OPW consecutive output entries, i.e. the contiguous sorted positions
// [seg[k0], seg[k0+NUM_OPW)).  It walks them in steps of 32*NUM_C positions: lane l takes the
// NUM_C consecutive positions [B + l*NUM_C, ...) (a warp-wide step reads 1 KB of perm contiguously),
// issues its NUM_C gathers together, sums them sequentially (runs that start and end inside the
// slice are stored directly), and the pieces of runs that cross lane boundaries are combined with
// one segmented inclusive scan over the lanes (5 shuffle steps per step); the open piece of the
// last lane is carried into the next step.  Fixed order => bit-reproducible, no atomics.
constexpr int NUM_BLOCK = 256;
constexpr int NUM_OPW = 128;  // outputs per warp
constexpr int NUM_C = 8;      // positions per lane per step

__global__ void __launch_bounds__(NUM_BLOCK) k_numeric(int64_t nnz, const uint32_t* __restrict__ perm,
                                                       const uint32_t* __restrict__ seg,
                                                       const double* __restrict__ vals, double* __restrict__ data) {
  __shared__ uint32_t sseg[NUM_BLOCK / 32][NUM_OPW + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t k0 = ((int64_t)blockIdx.x * (NUM_BLOCK / 32) + w) * NUM_OPW;
  if (k0 >= nnz) return;
  const int nout = (int)((nnz - k0 < NUM_OPW) ? (nnz - k0) : NUM_OPW);
  for (int i = lane; i <= nout; i += 32) sseg[w][i] = seg[k0 + i];
  __syncwarp();
  const uint32_t* ss = sseg[w];
  const uint32_t S = ss[0], E = ss[nout];
  double carry = 0.0;  // open piece of the run that continues from the previous step
  for (uint32_t B = S; B < E; B += 32 * NUM_C) {
    uint32_t p = B + lane * NUM_C;
    if (p > E) p = E;
    const uint32_t pe = (p + NUM_C < E) ? p + NUM_C : E;
    const bool nonempty = p < pe;
    double v[NUM_C];
#pragma unroll
    for (int u = 0; u < NUM_C; ++u) v[u] = (p + u < pe) ? vals[perm[p + u]] : 0.0;
    // o = run containing p: last o with ss[o] <= p
    int lo = 0, hi = nout;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (ss[mid] <= p) lo = mid; else hi = mid;
    }
    int o = lo;
    const bool starts_at_head = nonempty && (ss[o] == p);
    const int o_first = o;
    double acc = 0.0, first = 0.0;
    bool seen_head = starts_at_head;
    uint32_t next_head = nonempty ? ss[o + 1] : 0u;
#pragma unroll
    for (int u = 0; u < NUM_C; ++u) {
      uint32_t q = p + u;
      if (q < pe) {
        if (q == next_head) {  // a new run starts at q: close the previous piece
          if (seen_head) data[k0 + o] = acc;  // run began inside this slice: complete
          else first = acc;                   // piece of a run that began in an earlier lane / step
          seen_head = true;
          acc = 0.0;
          ++o;
          next_head = ss[o + 1];
        }
        acc += v[u];
      }
    }
    const bool has_head = seen_head;
    if (!has_head) first = acc;  // the whole slice is one piece of an earlier run
    // inclusive segmented scan over lanes of the value each lane passes on
    // (its open piece if a run starts in it, else its single piece), restarted at lanes with a head
    double incl = has_head ? acc : first;
    bool fl = has_head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double t = __shfl_up_sync(FULL, incl, d);
      bool tf = __shfl_up_sync(FULL, fl ? 1 : 0, d) != 0;
      if (lane >= d && !fl) {
        incl += t;
        fl = tf;
      }
    }
    // carry from the previous step reaches every lane up to (and including) the first lane with a head
    const bool reached = !fl;  // no head in lanes 0..lane
    double carry_in = __shfl_up_sync(FULL, incl, 1);
    bool reached_prev = __shfl_up_sync(FULL, reached ? 1 : 0, 1) != 0;
    if (lane == 0) {
      carry_in = carry;
    } else if (reached_prev) {
      carry_in += carry;
    }
    // close the run that continues into this lane (it ends at this lane's first head)
    if (nonempty && !starts_at_head && has_head) data[k0 + o_first] = carry_in + first;
    // value of the open run at the end of this lane
    const double open_tot = has_head ? acc : carry_in + first;
    const uint32_t nxt = __shfl_down_sync(FULL, p, 1);  // first position of the next lane
    const bool last_nonempty = nonempty && (lane == 31 || pe == E || nxt >= E);
    bool ends_here;
    if (lane == 31 || pe == E) ends_here = nonempty && (pe == E || ss[o + 1] == pe);
    else ends_here = nonempty && (ss[o + 1] == pe);
    if (ends_here) data[k0 + o] = open_tot;
    // carry into the next step: the open run of the last non-empty lane, if it continues
    double c_next = (last_nonempty && !ends_here) ? open_tot : 0.0;
    int src_lane = 31;
    unsigned ne = __ballot_sync(FULL, nonempty);
    if (ne) src_lane = 31 - __clz(ne);
    carry = __shfl_sync(FULL, c_next, src_lane);
  }
}

int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

int scan_u32(uint32_t* data, uint64_t n, uint32_t* parts, uint32_t* total_out, cudaStream_t s) {
  uint64_t g = pgb_cdiv((int64_t)n, 1024);
  if (g > SCAN_G) g = SCAN_G;
  if (g < 1) g = 1;
  uint64_t chunk = pgb_cdiv((int64_t)n, (int64_t)g);
  chunk = pgb_cdiv((int64_t)chunk, 1024) * 1024;
  g = pgb_cdiv((int64_t)n, (int64_t)chunk);
  k_scan_reduce<<<(unsigned)g, 256, 0, s>>>(data, n, chunk, parts);
  k_scan_parts<<<1, 256, 0, s>>>(parts, (int)g, total_out);
  k_scan_apply<<<(unsigned)g, 256, 0, s>>>(data, n, chunk, parts);
  PGB_LAUNCH_CHECK();
  return 0;
}

// Stable LSD radix sort of n (key, payload) pairs.  Pass p reads buffer (p-1)&1 ... and the final
// pass writes kbuf[0] / pbuf[0].  Keys of the first pass come from `src` (generated on the fly),
// payloads of the first pass are the positions 0..n-1.
template <typename KeyT>
int radix_sort(KeySrc<KeyT> src, uint32_t n, int total_bits, KeyT* kbuf[2], uint32_t* pbuf[2], uint32_t* hist,
               uint32_t* parts, cudaStream_t s) {
  static bool attr_set[64] = {false};  // function attributes are per device
  int dev = 0;
  PGB_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PGB_CHECK(cudaFuncSetAttribute(k_radix_scatter<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(ScatterSmem<KeyT>)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int passes = (total_bits + RADIX_BITS - 1) / RADIX_BITS;
  const uint32_t ntiles = (uint32_t)pgb_cdiv((int64_t)n, SORT_TILE);
  int dst = (passes % 2 == 1) ? 0 : 1;
  const uint32_t* pin = nullptr;
  for (int p = 0; p < passes; ++p) {
    int shift = p * RADIX_BITS;
    k_radix_hist<KeyT><<<ntiles, SORT_BLOCK, 0, s>>>(src, n, shift, ntiles, hist);
    PGB_LAUNCH_CHECK();
    if (scan_u32(hist, (uint64_t)ntiles * RADIX, parts, nullptr, s)) return 1;
    k_radix_scatter<KeyT><<<ntiles, SORT_BLOCK, sizeof(ScatterSmem<KeyT>), s>>>(src, pin, n, shift, ntiles, hist,
                                                                                 kbuf[dst], pbuf[dst]);
    PGB_LAUNCH_CHECK();
    src.keys = kbuf[dst];
    pin = pbuf[dst];
    dst ^= 1;
  }
  return 0;
}

// optional per-phase timing of the symbolic call (CUDA events on the caller's stream)
bool g_profile = false;
double g_phase_ms[8] = {0};
struct PhaseTimer {
  cudaEvent_t ev[8];
  int n = 0;
  cudaStream_t s;
  explicit PhaseTimer(cudaStream_t st) : s(st) {}
  void mark() {
    if (!g_profile || n >= 8) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], s);
    ++n;
  }
  void finish() {
    if (!g_profile) return;
    for (int i = 0; i < 8; ++i) g_phase_ms[i] = 0.0;
    if (n) cudaEventSynchronize(ev[n - 1]);
    for (int i = 0; i + 1 < n; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      g_phase_ms[i] = ms;
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
    n = 0;
  }
};

// row ranges of the last two-level symbolic_sort and the kernel family each got; symbolic_finish
// (called next, with the same workspace) compacts every range according to its scratch layout
struct RowRange {
  int64_t r0, r1;
  uint32_t max_inc;
  int kind;  // 0: k_rowthread range (grouped scratch); 1: one scratch segment per row (k_rowhash / k_rowsort /
             // k_rowlong ranges, rows handled through the size-class lists); 2: k_rowins range (grouped scratch
             // except where the per-row flag says otherwise)
};
std::vector<RowRange> g_ranges;
const void* g_ranges_ws = nullptr;

constexpr int MODE_TWO_LEVEL = 0, MODE_FULL_SORT = 1, MODE_TWO_LEVEL_KEY64 = 2, MODE_TWO_LEVEL_RADIX = 4;
constexpr uint32_t MAX_ROW_CAND = 512;  // capacity of the largest k_rowsort instance

struct WsLayout {
  // full sort
  int64_t keys_a, keys_b, perm_b;
  // two level
  int64_t ikeys_a, ikeys_b, ipay_a, ipay_b, row_nnz, ucol, chunk, rowlist;
  // shared
  int64_t hist, parts, meta, total;
};

WsLayout ws_layout(int64_t n_coo, int64_t n_inc, int64_t n_rows, int mode) {
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  WsLayout l = {};
  int64_t o = 0;
  int64_t nsort = (mode == MODE_FULL_SORT) ? n_coo : n_inc;
  int64_t ntiles = pgb_cdiv(nsort > 0 ? nsort : 1, SORT_TILE);
  l.meta = o;  o += 256;  // uint32: [0] mode, [1] cb, [2] nnz, [3] max incidences per row, [4] max row nnz,
                          // [5] rows per group of the unique-column scratch (0: one segment per row)
  l.parts = o; o += al(2048 * 4);
  l.hist = o;  o += al(ntiles * RADIX * 4);
  if (mode == MODE_FULL_SORT) {
    l.keys_a = o; o += al(n_coo * 8);
    l.keys_b = o; o += al(n_coo * 8);
    l.perm_b = o; o += al(n_coo * 4);
  } else {
    l.ikeys_a = o;   o += al(n_inc * 4);
    l.ikeys_b = o;   o += al(n_inc * 4);
    l.ipay_a = o;    o += al(n_inc * 4);
    l.ipay_b = o;    o += al(n_inc * 4);
    l.row_nnz = o;   o += al((n_rows + 2) * 4);
    l.ucol = o;      o += al(n_coo * 4);
    l.chunk = o;     o += al((n_rows / CHUNK_ROWS + 2) * (1 + 3 * NCLS) * 4);  // max, class histogram, class maxima, list starts
    l.rowlist = o;   o += al(n_rows * 4);
  }
  l.total = o;
  return l;
}

struct ChunkGrid {
  uint64_t g, chunk;
};
ChunkGrid chunk_grid(int64_t n) {
  uint64_t g = pgb_cdiv(n, 1024);
  if (g > SCAN_G) g = SCAN_G;
  if (g < 1) g = 1;
  uint64_t chunk = pgb_cdiv(pgb_cdiv(n, (int64_t)g), 1024) * 1024;
  g = pgb_cdiv(n, (int64_t)chunk);
  return {g, chunk};
}


bool g_numeric_u2 = false;  // flags bit 6 of pgb_csr_numeric_rows: 2 local rows in flight, 75 % occupancy target

template <int L, int BLK, typename IP, bool STAGED>
int launch_numeric_t_b(int64_t n_rows, int caps, const uint32_t* row_start, const uint8_t* slots, const void* indptr,
                       const double* vals_t, double* data, cudaStream_t s) {
  if constexpr (L == 4 && BLK == 128 && !STAGED) {
    if (g_numeric_u2 && (size_t)caps * BLK * sizeof(double) <= 48 * 1024) {
      k_numeric_t<L, BLK, IP, STAGED, 2><<<(unsigned)pgb_cdiv(n_rows, BLK), BLK, (size_t)caps * BLK * sizeof(double), s>>>(
          n_rows, caps, row_start, slots, (const IP*)indptr, vals_t, data);
      PGB_LAUNCH_CHECK();
      return 0;
    }
  }
  size_t smem = (STAGED ? 2 : 1) * (size_t)caps * BLK * sizeof(double);  // accumulators (+ output staging)
  static size_t attr[64];  // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (smem > 48 * 1024 && smem > attr[dev]) {
    PGB_CHECK(cudaFuncSetAttribute(k_numeric_t<L, BLK, IP, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr[dev] = smem;
  }
  k_numeric_t<L, BLK, IP, STAGED><<<(unsigned)pgb_cdiv(n_rows, BLK), BLK, smem, s>>>(
      n_rows, caps, row_start, slots, (const IP*)indptr, vals_t, data);
  PGB_LAUNCH_CHECK();
  return 0;
}

template <int L, typename IP>
int launch_numeric_t_l(int64_t n_rows, int caps, bool staged, const uint32_t* row_start, const uint8_t* slots,
                       const void* indptr, const double* vals_t, double* data, cudaStream_t s) {
#define PGB_NT(B, S) return launch_numeric_t_b<L, B, IP, S>(n_rows, caps, row_start, slots, indptr, vals_t, data, s)
  if (staged) {
    if (caps <= 48) PGB_NT(128, true);
    if (caps <= 100) PGB_NT(64, true);
    PGB_NT(32, true);
  }
  if (caps <= 200) PGB_NT(128, false);
  PGB_NT(64, false);
#undef PGB_NT
}

template <typename IP>
int launch_numeric_t(int L, int64_t n_rows, int caps, bool staged, const uint32_t* row_start, const uint8_t* slots,
                     const void* indptr, const double* vals_t, double* data, cudaStream_t s) {
  switch (L) {
    case 2: return launch_numeric_t_l<2, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 3: return launch_numeric_t_l<3, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 4: return launch_numeric_t_l<4, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 5: return launch_numeric_t_l<5, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 6: return launch_numeric_t_l<6, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 8: return launch_numeric_t_l<8, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 10: return launch_numeric_t_l<10, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 15: return launch_numeric_t_l<15, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 12: return launch_numeric_t_l<12, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
  }
  pgb_set_error("pgb_csr_numeric_rows: unsupported local size %d", L);
  return 2;
}

}  // namespace

extern "C" void pgb_set_profile(int on) { g_profile = on != 0; }
extern "C" const char* pgb_csr_last_row_kernel(void) { return g_row_kernel; }
/* phases of the last two-level pgb_csr_symbolic_sort call [ms]: 0 radix sort of the incidences,
 * 1 row offsets + max (incl. the host read), 2 per-row sort, 3 scan of the row counts */
extern "C" void pgb_get_phase_ms(double* out, int n) {
  for (int i = 0; i < n && i < 8; ++i) out[i] = g_phase_ms[i];
}

extern "C" int64_t pgb_csr_workspace_bytes(int64_t nc, int L, int64_t n_rows, int mode) {
  return ws_layout(nc * L * L, nc * L, n_rows, mode).total;
}

namespace {
// per-device event and pinned words for the two host reads of a symbolic_sort call: the work queued behind the
// read (the placement kernel, the caller's K1) keeps the GPU busy while the host waits for the event only
struct SymHost {
  cudaEvent_t ev = nullptr;
  uint32_t* pin = nullptr;  // [4]: nnz, max incidences, max row nnz, k_rowlong overflow
  bool pending = false;
  int mode = 0;
};
SymHost g_sym_host[64];
int sym_host(SymHost** out) {
  int dev = 0;
  PGB_CHECK(cudaGetDevice(&dev));
  PGB_REQUIRE(dev >= 0 && dev < 64, "pgb_csr_symbolic_sort: device index out of range");
  SymHost& h = g_sym_host[dev];
  if (!h.ev) {
    PGB_CHECK(cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming));
    PGB_CHECK(cudaMallocHost(&h.pin, 4 * sizeof(uint32_t)));
  }
  *out = &h;
  return 0;
}
int sym_result_check(const uint32_t* out2, int mode, int64_t* nnz_host, int* max_row_nnz_host) {
  *nnz_host = (int64_t)out2[0];
  if (max_row_nnz_host) *max_row_nnz_host = (int)out2[2];
  if (mode != MODE_FULL_SORT && (out2[2] > 255u || out2[3] > 255u)) {
    pgb_set_error("pgb_csr_symbolic_sort: a row has %u entries (> 255, the uint8 slot range); use mode 1",
                  out2[2] > out2[3] ? out2[2] : out2[3]);
    return 3;
  }
  return 0;
}
}  // namespace

extern "C" int pgb_csr_symbolic_result(int64_t* nnz_host, int* max_row_nnz_host) {
  SymHost* h = nullptr;
  if (int rc = sym_host(&h)) return rc;
  PGB_REQUIRE(h->pending, "pgb_csr_symbolic_result: no pgb_csr_symbolic_sort call with nnz_host == NULL is pending");
  h->pending = false;
  PGB_CHECK(cudaEventSynchronize(h->ev));
  return sym_result_check(h->pin, h->mode, nnz_host, max_row_nnz_host);
}

extern "C" int pgb_csr_symbolic_sort(int64_t n_rows, int64_t nc, int L, const int32_t* cell_dofs, int mode,
                                     void* workspace, int64_t workspace_bytes, uint32_t* perm_or_inc,
                                     uint8_t* slots, uint32_t* row_start, int64_t* nnz_host,
                                     int* max_row_nnz_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_coo = nc * L * L, n_inc = nc * L;
  PGB_REQUIRE(mode >= 0 && mode <= 4, "pgb_csr_symbolic_sort: bad mode");
  g_force_key64 = (mode == MODE_TWO_LEVEL_KEY64);
  g_force_bitonic = (mode == 3);
  PGB_REQUIRE(n_coo < ((int64_t)1 << 32) - SORT_TILE, "pgb_csr_symbolic_sort: n_coo must be < 2^32");
  PGB_REQUIRE(n_rows > 0 && n_rows < ((int64_t)1 << 31), "pgb_csr_symbolic_sort: n_rows out of range");
  WsLayout l = ws_layout(n_coo, n_inc, n_rows, mode);
  PGB_REQUIRE(workspace_bytes >= l.total, "pgb_csr_symbolic_sort: workspace too small");
  g_row_kernel[0] = 0;
  SymHost* sh = nullptr;
  if (int rc = sym_host(&sh)) return rc;
  sh->pending = false;
  if (nnz_host) *nnz_host = 0;
  if (max_row_nnz_host) *max_row_nnz_host = 0;
  char* ws = (char*)workspace;
  uint32_t* meta = (uint32_t*)(ws + l.meta);
  uint32_t* parts = (uint32_t*)(ws + l.parts);
  uint32_t* hist = (uint32_t*)(ws + l.hist);
  const int cb = bits_for(n_rows);
  uint32_t m[6] = {(uint32_t)mode, (uint32_t)cb, 0u, 0u, 0u, 0u};
  PGB_CHECK(cudaMemcpyAsync(meta, m, sizeof(m), cudaMemcpyHostToDevice, s));
  if (n_coo == 0) {
    if (mode != MODE_FULL_SORT) PGB_CHECK(cudaMemsetAsync(row_start, 0, sizeof(uint32_t) * (n_rows + 1), s));
    if (!nnz_host) {
      for (int i = 0; i < 4; ++i) sh->pin[i] = 0u;
      PGB_CHECK(cudaEventRecord(sh->ev, s));
      sh->pending = true;
      sh->mode = mode;
    }
    return 0;
  }

  if (mode == MODE_FULL_SORT) {
    uint64_t* kbuf[2] = {(uint64_t*)(ws + l.keys_a), (uint64_t*)(ws + l.keys_b)};
    uint32_t* pbuf[2] = {perm_or_inc, (uint32_t*)(ws + l.perm_b)};
    KeySrc<uint64_t> src{nullptr, cell_dofs, (uint32_t)L, (uint32_t)(L * L), cb};
    if (radix_sort<uint64_t>(src, (uint32_t)n_coo, 2 * cb, kbuf, pbuf, hist, parts, s)) return 1;
    ChunkGrid cg = chunk_grid(n_coo);
    k_count_heads<<<(unsigned)cg.g, 256, 0, s>>>(kbuf[0], (uint64_t)n_coo, cg.chunk, parts);
    k_scan_parts<<<1, 256, 0, s>>>(parts, (int)cg.g, meta + 2);
    PGB_LAUNCH_CHECK();
  } else {
    PGB_REQUIRE(slots && row_start, "pgb_csr_symbolic_sort: slots / row_start required in two-level mode");
    uint32_t* kbuf[2] = {(uint32_t*)(ws + l.ikeys_a), (uint32_t*)(ws + l.ikeys_b)};
    uint32_t* pbuf[2] = {(uint32_t*)(ws + l.ipay_a), (uint32_t*)(ws + l.ipay_b)};  // pbuf[0]: sorted incidences
    uint32_t* row_nnz = (uint32_t*)(ws + l.row_nnz);
    KeySrc<uint32_t> src{nullptr, cell_dofs, (uint32_t)L, (uint32_t)(L * L), cb};
    PhaseTimer pt(s);
    pt.mark();
    uint32_t max_inc = 0;
    uint32_t* chunk_max_d = (uint32_t*)(ws + l.chunk);
    const int64_t n_chunks = pgb_cdiv(n_rows, CHUNK_ROWS);
    // pinned staging for the small device->host / host->device transfers of this call (a pageable
    // buffer costs an extra internal synchronisation): [0, nch) chunk maxima, [nch] the global
    // maximum, then nch*NCLS class counts, nch*NCLS class maxima, nch*NCLS list starts
    static uint32_t* h_pin = nullptr;
    static int64_t h_pin_cap = 0;
    const int64_t h_need = n_chunks * (1 + 3 * NCLS) + 1;
    if (h_need > h_pin_cap) {
      if (h_pin) cudaFreeHost(h_pin);
      h_pin_cap = h_need + 4096;
      PGB_CHECK(cudaMallocHost(&h_pin, sizeof(uint32_t) * h_pin_cap));
    }
    uint32_t* chunk_max = h_pin;
    uint32_t* chunk_hist = h_pin + n_chunks + 1;
    uint32_t* chunk_cmax = chunk_hist + n_chunks * NCLS;
    uint32_t* list_start = chunk_cmax + n_chunks * NCLS;
    uint32_t* chunk_hist_d = chunk_max_d + n_chunks + 1;
    uint32_t* chunk_cmax_d = chunk_hist_d + n_chunks * NCLS;
    uint32_t* list_start_d = chunk_cmax_d + n_chunks * NCLS;
    uint32_t* row_list = (uint32_t*)(ws + l.rowlist);
    const uint32_t* inc_rows = pbuf[0];
    const bool radix = (mode == MODE_TWO_LEVEL_RADIX);
    uint8_t* rank = (uint8_t*)kbuf[0];
    if (radix) {
      if (radix_sort<uint32_t>(src, (uint32_t)n_inc, cb, kbuf, pbuf, hist, parts, s)) return 1;
      pt.mark();
      k_inc_pos<<<pgb_grid(n_inc, 256), 256, 0, s>>>((uint32_t)n_inc, pbuf[0], perm_or_inc);  // inverse permutation
      k_row_starts<<<pgb_grid(n_inc, 256), 256, 0, s>>>(kbuf[0], (uint32_t)n_inc, n_rows, row_start);
    } else {
      // counting transpose: counts -> row_start, provisional ranks in kbuf[0] (bytes)
      PGB_CHECK(cudaMemsetAsync(row_start, 0, sizeof(uint32_t) * (n_rows + 1), s));
      k_inc_count<<<pgb_grid(pgb_cdiv(n_inc, 4), 256), 256, 0, s>>>((uint32_t)n_inc, cell_dofs, row_start, rank);
      PGB_LAUNCH_CHECK();
      if (scan_u32(row_start, (uint64_t)n_rows + 1, parts, nullptr, s)) return 1;
    }
    k_chunk_max<<<(unsigned)n_chunks, 256, 0, s>>>(n_rows, row_start, chunk_max_d, chunk_hist_d, chunk_cmax_d, meta + 3);
    PGB_LAUNCH_CHECK();
    PGB_CHECK(cudaMemcpyAsync(h_pin + n_chunks, meta + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaMemcpyAsync(chunk_max, chunk_max_d, sizeof(uint32_t) * n_chunks, cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaMemcpyAsync(chunk_hist, chunk_hist_d, sizeof(uint32_t) * n_chunks * NCLS * 2, cudaMemcpyDeviceToHost, s));  // counts + maxima
    if (!radix) {
      // the placement does not depend on what the host decides from the chunk maxima, except for dofs with more
      // than 255 incidences (saturated ranks: redone below with the overflow counters): it runs while the host
      // waits for the read and builds the work items
      PGB_CHECK(cudaEventRecord(sh->ev, s));
      k_inc_place<<<pgb_grid(pgb_cdiv(n_inc, 4), 256), 256, 0, s>>>((uint32_t)n_inc, cell_dofs, row_start, rank, nullptr, pbuf[1]);
      PGB_LAUNCH_CHECK();
      PGB_CHECK(cudaEventSynchronize(sh->ev));
    } else {
      PGB_CHECK(cudaStreamSynchronize(s));
    }
    max_inc = h_pin[n_chunks];
    if (!radix) pt.mark();
    // rows beyond the capacity of k_rowsort go to k_rowlong (counting transpose only)
    const bool long_ok = !radix && !g_force_bitonic && !g_force_key64;
    if (max_inc > (uint32_t)LONG_MAX_INC || (!long_ok && (max_inc * (uint32_t)L > MAX_ROW_CAND || max_inc > 256u))) {
      pgb_set_error("pgb_csr_symbolic_sort: a dof is shared by %u cells (more than the row kernels take); "
                    "use mode 1 (full sort)", max_inc);
      return 3;
    }
    // Work items.  Chunks of short rows: contiguous ranges for k_rowthread.  The other chunks: either
    // contiguous ranges, each with the kernel its longest row needs, or - when the long rows are
    // scattered, so that almost every chunk has one - one row list per size class gathered from all of
    // them.  A rough cost model (candidates per row x a factor per kernel family) decides.
    struct Item {
      const uint32_t* rows;  // row list (then r0 = 0) or null
      int64_t r0, r1;
      uint32_t bound;  // largest incidence count
      int kind;        // 0 k_rowthread, 1 k_rowhash, 2 k_rowsort, 3 k_rowlong, 4 k_rowins
    };
    auto bucket = [](uint32_t m) {
      uint32_t b = 1;
      while (b < m) b <<= 1;
      return b;
    };
    auto kind_of = [&](uint32_t m) {  // kernel family for rows of at most m incidences (not k_rowthread)
      if (rowhash_applies(m, m * (uint32_t)L)) return 1;
      if (m * (uint32_t)L <= MAX_ROW_CAND && m <= 256u) return 2;
      return 3;
    };
    auto cost_of = [&](uint32_t m) {
      const double mc = (double)m * L;
      const int k = kind_of(m);
      const double f = k == 3 ? 20.0 : k == 2 ? 6.0 : m > 32u ? 4.0 : mc > 96.0 ? 2.0 : 1.0;
      return mc * f;
    };
    std::vector<char> ckind((size_t)n_chunks);  // 0: k_rowthread range, 1: other
    for (int64_t i = 0; i < n_chunks; ++i) {
      const uint32_t m = chunk_max[i] ? chunk_max[i] : 1u;
      ckind[(size_t)i] = rowthread_applies(m, m * (uint32_t)L) ? 0 : 1;
    }
    {  // many short k_rowthread ranges between other chunks would cost a launch each
      int64_t nrange = 1;
      for (int64_t i = 1; i < n_chunks; ++i) nrange += ckind[(size_t)i] != ckind[(size_t)i - 1];
      if (nrange > 64) {
        for (int64_t i = 0; i < n_chunks;) {
          int64_t j = i;
          while (j < n_chunks && ckind[(size_t)j] == ckind[(size_t)i]) ++j;
          if (ckind[(size_t)i] == 0 && j - i < 8)
            for (int64_t q = i; q < j; ++q) ckind[(size_t)q] = 1;
          i = j;
        }
      }
    }
    uint32_t cls_total[NCLS] = {0}, cls_max[NCLS] = {0};  // rows / largest incidence count per size class
    double cost_ranges = 0.0, cost_lists = 0.0;
    for (int64_t i = 0; i < n_chunks; ++i) {
      if (!ckind[(size_t)i]) continue;
      const int64_t r0 = i * CHUNK_ROWS, r1 = r0 + CHUNK_ROWS < n_rows ? r0 + CHUNK_ROWS : n_rows;
      cost_ranges += (double)(r1 - r0) * cost_of(chunk_max[i] ? chunk_max[i] : 1u);
      for (int k = 0; k < NCLS; ++k) {
        cls_total[k] += chunk_hist[i * NCLS + k];
        cls_max[k] = chunk_cmax[i * NCLS + k] > cls_max[k] ? chunk_cmax[i * NCLS + k] : cls_max[k];
      }
    }
    for (int k = 0; k < NCLS; ++k) cost_lists += (double)cls_total[k] * cost_of(cls_max[k] ? cls_max[k] : 1u);
    const bool use_lists = cost_ranges > 1.25 * cost_lists;
    // ranges (for the launches below and for symbolic_finish)
    std::vector<RowRange>& rg = g_ranges;
    rg.clear();
    for (int64_t i = 0; i < n_chunks; ++i) {
      const uint32_t m = chunk_max[i] ? chunk_max[i] : 1u;
      const int64_t r0 = i * CHUNK_ROWS, r1 = r0 + CHUNK_ROWS < n_rows ? r0 + CHUNK_ROWS : n_rows;
      const int kind = ckind[(size_t)i];
      const bool same_kernel = (kind && use_lists) || bucket(rg.empty() ? 0u : rg.back().max_inc) == bucket(m);
      if (!rg.empty() && rg.back().kind == kind && same_kernel) {
        rg.back().r1 = r1;
        rg.back().max_inc = m > rg.back().max_inc ? m : rg.back().max_inc;
      } else {
        rg.push_back({r0, r1, m, kind});
      }
    }
    if (!use_lists && rg.size() > 64) {  // bound the number of launches: merge neighbouring non-thread ranges
      std::vector<RowRange> c;
      for (const RowRange& x : rg) {
        if (!c.empty() && c.back().kind && x.kind) {
          c.back().r1 = x.r1;
          c.back().max_inc = x.max_inc > c.back().max_inc ? x.max_inc : c.back().max_inc;
        } else {
          c.push_back(x);
        }
      }
      rg.swap(c);
    }
    g_ranges_ws = workspace;
    std::vector<Item> items;
    for (const RowRange& x : rg) {
      if (x.kind == 0) items.push_back({nullptr, x.r0, x.r1, x.max_inc, 0});
      else if (!use_lists) items.push_back({nullptr, x.r0, x.r1, x.max_inc, kind_of(x.max_inc)});
    }
    if (use_lists) {
      uint32_t cls_base[NCLS + 1] = {0}, run[NCLS];
      for (int k = 0; k < NCLS; ++k) cls_base[k + 1] = cls_base[k] + cls_total[k];
      for (int k = 0; k < NCLS; ++k) run[k] = cls_base[k];
      for (int64_t i = 0; i < n_chunks; ++i) {
        if (!ckind[(size_t)i]) {
          list_start[i * NCLS] = 0xffffffffu;
          continue;
        }
        for (int k = 0; k < NCLS; ++k) {
          list_start[i * NCLS + k] = run[k];
          run[k] += chunk_hist[i * NCLS + k];
        }
      }
      PGB_CHECK(cudaMemcpyAsync(list_start_d, list_start, sizeof(uint32_t) * n_chunks * NCLS, cudaMemcpyHostToDevice, s));
      k_row_lists<<<(unsigned)n_chunks, 256, 0, s>>>(n_rows, row_start, list_start_d, row_list);
      PGB_LAUNCH_CHECK();
      for (int k = NCLS - 1; k >= 0; --k) {  // longest rows first
        if (!cls_total[k]) continue;
        const uint32_t bound = cls_max[k] ? cls_max[k] : 1u;
        items.push_back({row_list + cls_base[k], 0, (int64_t)cls_total[k], bound, kind_of(bound)});
      }
    }
    if (!radix) {
      if (max_inc > 255u) {  // saturated ranks: second counter per dof (row_nnz is free until the row kernels run)
        uint32_t* ovf = row_nnz;
        PGB_CHECK(cudaMemsetAsync(ovf, 0, sizeof(uint32_t) * n_rows, s));
        k_inc_place<<<pgb_grid(pgb_cdiv(n_inc, 4), 256), 256, 0, s>>>((uint32_t)n_inc, cell_dofs, row_start, rank, ovf, pbuf[1]);
        PGB_LAUNCH_CHECK();
      }
      inc_rows = pbuf[1];  // row blocks, each in arbitrary order (placed before the host read above)
      for (const Item& x : items) {
        if (x.kind != 2) continue;  // only k_rowsort needs sorted row blocks
        const int64_t nr = x.r1 - x.r0;
#define PGB_RC(G, N) k_row_canon<G, N><<<pgb_grid(nr * G, 256), 256, 0, s>>>(x.rows, x.r0, x.r1, row_start, pbuf[1], pbuf[0], perm_or_inc)
        if (x.bound <= 4) PGB_RC(4, 1);
        else if (x.bound <= 8) PGB_RC(8, 1);
        else if (x.bound <= 16) PGB_RC(16, 1);
        else if (x.bound <= 32) PGB_RC(32, 1);
        else if (x.bound <= 64) PGB_RC(32, 2);
        else if (x.bound <= 128) PGB_RC(32, 4);
        else PGB_RC(32, 8);
#undef PGB_RC
        PGB_LAUNCH_CHECK();
      }
    }
    PGB_CHECK(cudaMemsetAsync(row_nnz + n_rows, 0, 2 * sizeof(uint32_t), s));
    // hash-dedupe items whose rows fit one thread each go to k_rowins; the rows it cannot finish (too many
    // distinct columns) come back in a list and are redone by k_rowhash
    uint32_t* ins_ovf = kbuf[1];  // free in the counting-transpose mode
    uint32_t ins_bound = 0;
    if (!radix && n_rows <= 4 * n_inc)  // (the per-row flags of k_rowins live in the n_inc * 4 bytes of the rank array)
      for (Item& x : items)
        if (x.kind == 1 && rowins_applies(x.bound, L)) {
          x.kind = 4;
          ins_bound = x.bound > ins_bound ? x.bound : ins_bound;
          if (!x.rows)  // a row range: its distinct columns are stored per block of rows (k_compact_ins)
            for (RowRange& g : rg)
              if (g.r0 == x.r0 && g.r1 == x.r1 && g.kind == 1) g.kind = 2;
        }
    if (ins_bound) PGB_CHECK(cudaMemsetAsync(ins_ovf, 0, sizeof(uint32_t), s));
    pt.mark();
    for (const Item& x : items) {
      const uint32_t mc = x.bound * (uint32_t)L;
      int32_t* ucol = (int32_t*)(ws + l.ucol);
      int rc;
      switch (x.kind) {
        case 0:
          rc = launch_rowthread(x.bound, L, x.r0, x.r1, n_rows, row_start, inc_rows, cell_dofs, perm_or_inc, slots, ucol,
                                row_nnz, meta + 4, s);
          break;
        case 1:
          rc = launch_rowhash(x.bound, mc, L, x.rows, x.r0, x.r1, row_start, inc_rows, cell_dofs, perm_or_inc, slots, ucol,
                              row_nnz, meta + 4, s);
          break;
        case 2:
          rc = launch_rowsort(mc, L, x.rows, x.r0, x.r1, n_rows, row_start, pbuf[0], cell_dofs, slots, ucol, row_nnz,
                              meta + 4, s);
          break;
        case 4:
          rc = launch_rowins(x.bound, L, x.rows, x.r0, x.r1, row_start, inc_rows, cell_dofs, perm_or_inc, slots, ucol, row_nnz,
                             meta + 4, ins_ovf, rank /* free after the placement: per-row flags */, s);
          break;
        default:
          rc = launch_rowlong(L, x.rows, x.r0, x.r1, row_start, inc_rows, cell_dofs, perm_or_inc, slots, ucol, row_nnz,
                              meta + 4, meta + 5, s);
      }
      if (rc) return rc;
    }
    if (ins_bound) {  // the rows k_rowins left (count on the device: no host read, a fixed grid loops over the list)
      int rc = launch_rowhash_list(ins_bound * (uint32_t)L, L, ins_ovf, row_start, inc_rows, cell_dofs, perm_or_inc, slots,
                                   (int32_t*)(ws + l.ucol), row_nnz, meta + 4, s);
      if (rc) return rc;
    }
    pt.mark();
    if (scan_u32(row_nnz, (uint64_t)n_rows + 1, parts, nullptr, s)) return 1;
    PGB_CHECK(cudaMemcpyAsync(meta + 2, row_nnz + n_rows, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    pt.mark();
    pt.finish();
  }
  // nnz, max incidences, max row nnz, k_rowlong overflow (distinct columns of a row)
  PGB_CHECK(cudaMemcpyAsync(sh->pin, meta + 2, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (!nnz_host) {  // the caller queues more work (K1 needs inc_pos only) and collects with pgb_csr_symbolic_result
    PGB_CHECK(cudaEventRecord(sh->ev, s));
    sh->pending = true;
    sh->mode = mode;
    return 0;
  }
  PGB_CHECK(cudaStreamSynchronize(s));
  return sym_result_check(sh->pin, mode, nnz_host, max_row_nnz_host);
}

extern "C" int pgb_csr_symbolic_finish(int64_t n_rows, int64_t nc, int L, int mode, int64_t nnz,
                                       const void* workspace, const uint32_t* row_start, void* indptr, int idx64,
                                       int32_t* indices, uint32_t* seg, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_coo = nc * L * L, n_inc = nc * L;
  if (n_coo == 0) {
    if (idx64) k_fill_indptr_empty<int64_t><<<pgb_grid(n_rows + 1, 256), 256, 0, s>>>(n_rows, (int64_t*)indptr);
    else k_fill_indptr_empty<int32_t><<<pgb_grid(n_rows + 1, 256), 256, 0, s>>>(n_rows, (int32_t*)indptr);
    PGB_LAUNCH_CHECK();
    return 0;
  }
  WsLayout l = ws_layout(n_coo, n_inc, n_rows, mode);
  const char* ws = (const char*)workspace;
  const int cb = bits_for(n_rows);
  if (mode == MODE_FULL_SORT) {
    PGB_REQUIRE(seg, "pgb_csr_symbolic_finish: seg required in mode 1");
    const uint64_t* keys = (const uint64_t*)(ws + l.keys_a);
    const uint32_t* parts = (const uint32_t*)(ws + l.parts);
    ChunkGrid cg = chunk_grid(n_coo);
    if (idx64)
      k_emit_unique<int64_t><<<(unsigned)cg.g, 256, 0, s>>>(keys, (uint64_t)n_coo, cg.chunk, parts, cb, n_rows,
                                                            (uint32_t)nnz, (int64_t*)indptr, indices, seg);
    else
      k_emit_unique<int32_t><<<(unsigned)cg.g, 256, 0, s>>>(keys, (uint64_t)n_coo, cg.chunk, parts, cb, n_rows,
                                                            (uint32_t)nnz, (int32_t*)indptr, indices, seg);
  } else {
    PGB_REQUIRE(row_start, "pgb_csr_symbolic_finish: row_start required in two-level mode");
    const uint32_t* row_off = (const uint32_t*)(ws + l.row_nnz);
    const int32_t* ucol = (const int32_t*)(ws + l.ucol);
    PGB_REQUIRE(g_ranges_ws == workspace && !g_ranges.empty(),
                "pgb_csr_symbolic_finish: must follow pgb_csr_symbolic_sort on the same workspace");
    for (const RowRange& x : g_ranges) {
      if (x.kind == 0) {
        unsigned grid = (unsigned)pgb_cdiv(x.r1 - x.r0, (int64_t)ROWTHREAD_GROUP);
        if (idx64)
          k_compact_groups<int64_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, ROWTHREAD_GROUP, row_start, row_off, ucol,
                                                        (int64_t*)indptr, indices);
        else
          k_compact_groups<int32_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, ROWTHREAD_GROUP, row_start, row_off, ucol,
                                                        (int32_t*)indptr, indices);
      } else if (x.kind == 2) {
        const uint8_t* flag = (const uint8_t*)(ws + l.ikeys_a);
        unsigned grid = (unsigned)pgb_cdiv(x.r1 - x.r0, (int64_t)ROWTHREAD_GROUP);
        if (idx64)
          k_compact_ins<int64_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, ROWTHREAD_GROUP, row_start, row_off, flag, ucol,
                                                     (int64_t*)indptr, indices);
        else
          k_compact_ins<int32_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, ROWTHREAD_GROUP, row_start, row_off, flag, ucol,
                                                     (int32_t*)indptr, indices);
      } else {
        unsigned grid = pgb_grid((x.r1 - x.r0 + 1) * 8, 256);
        if (idx64)
          k_compact<int64_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, row_start, row_off, ucol, (int64_t*)indptr, indices);
        else
          k_compact<int32_t><<<grid, 256, 0, s>>>(x.r0, x.r1, n_rows, L, row_start, row_off, ucol, (int32_t*)indptr, indices);
      }
      PGB_LAUNCH_CHECK();
    }
  }
  PGB_LAUNCH_CHECK();
  return 0;
}

int pgb_numeric_bulk_launch(int L, int64_t n_rows, int caps, bool staged, int idx64, const uint32_t* row_start,
                            const uint8_t* slots, const void* indptr, const double* vals_t, double* data,
                            cudaStream_t s);

int pgb_numeric_group_launch(int L, int G, int64_t n_rows, int caps, int idx64, const uint32_t* row_start,
                             const uint8_t* slots, const void* indptr, const double* vals_t, double* data,
                             cudaStream_t s);

extern "C" int pgb_csr_numeric_rows(int64_t n_rows, int L, int max_row_nnz, const uint32_t* row_start,
                                    const uint8_t* slots, const void* indptr, int idx64, const double* vals_t,
                                    double* data, int flags, void* stream) {
  if (n_rows == 0) return 0;
  PGB_REQUIRE(max_row_nnz >= 0 && max_row_nnz <= 255, "pgb_csr_numeric_rows: max_row_nnz out of range");
  cudaStream_t s = (cudaStream_t)stream;
  int caps = max_row_nnz < 1 ? 1 : max_row_nnz;
  const bool staged = (flags & 1) != 0;
  g_numeric_u2 = (flags & 64) != 0;
  if (const int G = (flags >> 2) & 15) {  // G lanes per row, coalesced loads and stores (pgb_numeric_bulk.cu)
    int rc = pgb_numeric_group_launch(L, G, n_rows, caps, idx64, row_start, slots, indptr, vals_t, data, s);
    if (rc >= 0) return rc;
  }
  if (flags & 2) {  // bulk-async staging of the block's contiguous local rows (pgb_numeric_bulk.cu)
    int rc = pgb_numeric_bulk_launch(L, n_rows, caps, staged, idx64, row_start, slots, indptr, vals_t, data, s);
    if (rc >= 0) return rc;
  }
  return idx64 ? launch_numeric_t<int64_t>(L, n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s)
               : launch_numeric_t<int32_t>(L, n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
}

extern "C" int pgb_csr_numeric(int64_t nnz, const uint32_t* perm, const uint32_t* seg,
                               const double* vals, double* data, void* stream) {
  if (nnz == 0) return 0;
  int64_t warps = pgb_cdiv(nnz, NUM_OPW);
  unsigned grid = (unsigned)pgb_cdiv(warps, NUM_BLOCK / 32);
  k_numeric<<<grid, NUM_BLOCK, 0, (cudaStream_t)stream>>>(nnz, perm, seg, vals, data);
  PGB_LAUNCH_CHECK();
  return 0;
}

// ---- grid connectivity: ridges of a 3D simplicial grid (grids/grid.py:149-216) --------------------
// The reference numbers the ridges (edges) as the distinct sorted node pairs of all faces in
// lexicographic order (coo.sum_duplicates() of A[lo, hi], grid.py:187-198) and gives face f the
// ridges (n0,n1), (n1,n2), (n2,n0) of its node order with orientation sign(next - prev)
// (grid.py:172-182).  Here: one packed key (lo << cb | hi) per (face, slot), the stable LSD radix
// sort of this file with the position as payload, head flags + exclusive scan = ridge id.
namespace {

__global__ void __launch_bounds__(256) k_ridge_keys(uint32_t n, const int32_t* __restrict__ fn, int cb,
                                                    uint64_t* __restrict__ keys, int8_t* __restrict__ sign,
                                                    int32_t num_nodes, int32_t* __restrict__ err) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  uint32_t f = i / 3u, j = i - 3u * f;
  int32_t a = fn[i];
  int32_t b = fn[3u * f + (j == 2u ? 0u : j + 1u)];
  if (a == b || a < 0 || b < 0 || a >= num_nodes || b >= num_nodes) *err = 1;  // degenerate face / id out of range
  uint32_t lo = (uint32_t)min(a, b), hi = (uint32_t)max(a, b);
  keys[i] = ((uint64_t)lo << cb) | (uint64_t)hi;
  sign[i] = (int8_t)((b > a) - (b < a));
}

__global__ void __launch_bounds__(256) k_ridge_heads(uint32_t n, const uint64_t* __restrict__ keys,
                                                     uint32_t* __restrict__ head) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) head[i] = is_head(keys, i) ? 1u : 0u;
}

// head_ex[i] = number of heads before position i  =>  ridge id of position i = head_ex[i] - !is_head(i)
__global__ void __launch_bounds__(256) k_ridge_emit(uint32_t n, const uint64_t* __restrict__ keys,
                                                    const uint32_t* __restrict__ pos,
                                                    const uint32_t* __restrict__ head_ex, int cb,
                                                    int32_t* __restrict__ face_ridges,
                                                    int32_t* __restrict__ ridge_peaks) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  const bool h = (i == 0) || keys[i - 1] != k;
  const uint32_t id = head_ex[i] - (h ? 0u : 1u);
  face_ridges[pos[i]] = (int32_t)id;
  if (h) {
    ridge_peaks[2 * (uint64_t)id] = (int32_t)(k >> cb);
    ridge_peaks[2 * (uint64_t)id + 1] = (int32_t)(k & ((1ull << cb) - 1));
  }
}

struct RidgeWs {
  int64_t keys_a, keys_b, pay_a, pay_b, hist, parts, meta, total;
};
RidgeWs ridge_ws(int64_t n) {
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  RidgeWs l = {};
  int64_t o = 0;
  l.meta = o;   o += 256;  // [0] number of ridges, [1] error flag
  l.parts = o;  o += al(2048 * 4);
  l.hist = o;   o += al(pgb_cdiv(n > 0 ? n : 1, SORT_TILE) * RADIX * 4);
  l.keys_a = o; o += al(n * 8);
  l.keys_b = o; o += al(n * 8);
  l.pay_a = o;  o += al(n * 4);
  l.pay_b = o;  o += al(n * 4);
  l.total = o;
  return l;
}

}  // namespace

extern "C" int64_t pgb_ridges3d_workspace_bytes(int64_t nf) { return ridge_ws(3 * nf).total; }

extern "C" int pgb_ridges3d(int64_t nf, int64_t num_nodes, const int32_t* face_nodes, int32_t* face_ridges,
                            int8_t* face_ridge_sign, int32_t* ridge_peaks, int64_t* num_ridges_host,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = 3 * nf;
  *num_ridges_host = 0;
  if (nf == 0) return 0;
  PGB_REQUIRE(n < ((int64_t)1 << 32) - SORT_TILE, "pgb_ridges3d: 3*num_faces must be < 2^32");
  PGB_REQUIRE(num_nodes > 0 && num_nodes < ((int64_t)1 << 31), "pgb_ridges3d: num_nodes out of range");
  RidgeWs l = ridge_ws(n);
  PGB_REQUIRE(workspace_bytes >= l.total, "pgb_ridges3d: workspace too small");
  char* ws = (char*)workspace;
  uint32_t* meta = (uint32_t*)(ws + l.meta);
  uint32_t* parts = (uint32_t*)(ws + l.parts);
  uint32_t* hist = (uint32_t*)(ws + l.hist);
  uint64_t* kbuf[2] = {(uint64_t*)(ws + l.keys_a), (uint64_t*)(ws + l.keys_b)};
  uint32_t* pbuf[2] = {(uint32_t*)(ws + l.pay_a), (uint32_t*)(ws + l.pay_b)};
  const int cb = bits_for(num_nodes);
  const int passes = (2 * cb + RADIX_BITS - 1) / RADIX_BITS;
  PGB_CHECK(cudaMemsetAsync(meta, 0, 256, s));
  // the first pass reads its keys from the buffer the last pass does NOT write into
  uint64_t* keys_in = kbuf[(passes % 2 == 1) ? 1 : 0];
  const unsigned g = pgb_grid(n, 256);
  k_ridge_keys<<<g, 256, 0, s>>>((uint32_t)n, face_nodes, cb, keys_in, face_ridge_sign, (int32_t)num_nodes,
                                 (int32_t*)(meta + 1));
  PGB_LAUNCH_CHECK();
  KeySrc<uint64_t> src{keys_in, nullptr, 0u, 0u, cb};
  if (radix_sort<uint64_t>(src, (uint32_t)n, 2 * cb, kbuf, pbuf, hist, parts, s)) return 1;
  uint32_t* head = pbuf[1];  // free after the sort (sorted keys / positions are in kbuf[0] / pbuf[0])
  k_ridge_heads<<<g, 256, 0, s>>>((uint32_t)n, kbuf[0], head);
  PGB_LAUNCH_CHECK();
  if (scan_u32(head, (uint64_t)n, parts, meta, s)) return 1;
  k_ridge_emit<<<g, 256, 0, s>>>((uint32_t)n, kbuf[0], pbuf[0], head, cb, face_ridges, ridge_peaks);
  PGB_LAUNCH_CHECK();
  uint32_t m[2] = {0, 0};
  PGB_CHECK(cudaMemcpyAsync(m, meta, sizeof(m), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  PGB_REQUIRE(m[1] == 0, "pgb_ridges3d: a face repeats a node or holds a node id outside [0, num_nodes)");
  *num_ridges_host = (int64_t)m[0];
  return 0;
}

// ---- CSR transpose (B -> B^T of the mixed-dimensional saddle blocks; scipy's .T.tocsr()) -------------------
// A stable LSD radix sort of the column indices with the entry position as payload orders the entries by
// (column, row): exactly the CSR arrays of the transpose, deterministic, no atomics.  Then the column
// starts from the sorted keys and one pass that looks the row of every entry up in indptr.
namespace {

struct TrWs {
  int64_t keys_a, keys_b, pay_a, pay_b, hist, parts, total;
};
TrWs transpose_ws(int64_t nnz) {
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  TrWs l = {};
  int64_t o = 0;
  l.parts = o;  o += al(2048 * 4);
  l.hist = o;   o += al(pgb_cdiv(nnz > 0 ? nnz : 1, SORT_TILE) * RADIX * 4);
  l.keys_a = o; o += al(nnz * 4);
  l.keys_b = o; o += al(nnz * 4);
  l.pay_a = o;  o += al(nnz * 4);
  l.pay_b = o;  o += al(nnz * 4);
  l.total = o;
  return l;
}

template <typename IP>
__global__ void __launch_bounds__(256) k_transpose_starts(const uint32_t* __restrict__ keys, uint32_t n, int64_t n_cols,
                                                          IP* __restrict__ out_ip) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  uint32_t prev = (i == 0) ? 0xffffffffu : keys[i - 1];
  if (k != prev)
    for (uint32_t c = prev + 1u; c <= k; ++c) out_ip[c] = (IP)i;  // prev + 1 wraps to 0 for i == 0
  if (i == n - 1)
    for (int64_t c = (int64_t)k + 1; c <= n_cols; ++c) out_ip[c] = (IP)n;
}

template <typename IP>
__global__ void __launch_bounds__(256) k_transpose_emit(uint32_t n, const uint32_t* __restrict__ perm,
                                                        const IP* __restrict__ ip, int64_t n_rows,
                                                        const double* __restrict__ data,
                                                        int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
  uint32_t j = blockIdx.x * 256u + threadIdx.x;
  if (j >= n) return;
  const uint32_t p = perm[j];
  int64_t lo = 0, hi = n_rows;  // the last row r with ip[r] <= p (ip[n_rows] = n > p)
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)ip[mid] <= (int64_t)p) lo = mid;
    else hi = mid;
  }
  out_idx[j] = (int32_t)lo;
  out_val[j] = data[p];
}

}  // namespace

extern "C" int64_t pgb_csr_transpose_workspace_bytes(int64_t nnz) { return transpose_ws(nnz).total; }

extern "C" int pgb_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const void* indptr, int idx64,
                                 const int32_t* indices, const double* data, void* out_indptr, int out_idx64,
                                 int32_t* out_indices, double* out_data, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31) && n_cols >= 0 && n_cols < ((int64_t)1 << 31),
              "pgb_csr_transpose: dimensions must be < 2^31");
  PGB_REQUIRE(nnz >= 0 && nnz < ((int64_t)1 << 32) - SORT_TILE, "pgb_csr_transpose: nnz must be < 2^32");
  if (nnz == 0) {
    PGB_CHECK(cudaMemsetAsync(out_indptr, 0, (size_t)(n_cols + 1) * (out_idx64 ? 8 : 4), s));
    return 0;
  }
  PGB_REQUIRE(out_idx64 || nnz < ((int64_t)1 << 31), "pgb_csr_transpose: int32 indptr cannot hold nnz");
  TrWs l = transpose_ws(nnz);
  PGB_REQUIRE(workspace_bytes >= l.total, "pgb_csr_transpose: workspace too small");
  char* ws = (char*)workspace;
  uint32_t* parts = (uint32_t*)(ws + l.parts);
  uint32_t* hist = (uint32_t*)(ws + l.hist);
  uint32_t* kbuf[2] = {(uint32_t*)(ws + l.keys_a), (uint32_t*)(ws + l.keys_b)};
  uint32_t* pbuf[2] = {(uint32_t*)(ws + l.pay_a), (uint32_t*)(ws + l.pay_b)};
  KeySrc<uint32_t> src{nullptr, indices, 1u, 1u, 0};  // first pass: key i = indices[i], payload i
  if (radix_sort<uint32_t>(src, (uint32_t)nnz, bits_for(n_cols), kbuf, pbuf, hist, parts, s)) return 1;
  const unsigned g = pgb_grid(nnz, 256);
  if (out_idx64) k_transpose_starts<int64_t><<<g, 256, 0, s>>>(kbuf[0], (uint32_t)nnz, n_cols, (int64_t*)out_indptr);
  else k_transpose_starts<int32_t><<<g, 256, 0, s>>>(kbuf[0], (uint32_t)nnz, n_cols, (int32_t*)out_indptr);
  PGB_LAUNCH_CHECK();
  if (idx64)
    k_transpose_emit<int64_t><<<g, 256, 0, s>>>((uint32_t)nnz, pbuf[0], (const int64_t*)indptr, n_rows, data, out_indices,
                                                 out_data);
  else
    k_transpose_emit<int32_t><<<g, 256, 0, s>>>((uint32_t)nnz, pbuf[0], (const int32_t*)indptr, n_rows, data, out_indices,
                                                 out_data);
  PGB_LAUNCH_CHECK();
  return 0;
}
