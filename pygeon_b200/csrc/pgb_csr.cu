// K2: deterministic COO -> CSR.
//
// symbolic  : stable LSD radix sort (8-bit digits) of the packed (row,col) keys with the COO
//             position as payload, then a head-flag compaction that emits indptr / indices /
//             run offsets.  Keys of the first pass are generated on the fly from cell_dofs.
// numeric   : data[k] = sum of vals[perm[p]] over the run of entry k, by a warp-shuffle
//             segmented scan in a fixed order.
// No floating-point atomics anywhere; the only atomics are shared-memory integer counters of the
// digit histogram (their result is order independent), so pattern and values are bit-reproducible.
#include "pgb_common.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_BLOCK = 256;
constexpr int SORT_WARPS = SORT_BLOCK / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_BLOCK * SORT_ITEMS;  // 4096 keys per block
constexpr unsigned FULL = 0xffffffffu;

constexpr int SCAN_G = PGB_NPART;  // chunks of the two-level scans (<= 2048)

struct KeySrc {
  const uint64_t* keys;  // nullptr in the first pass: generate from cell_dofs
  const int32_t* dofs;
  uint32_t L, LL;
  int cb;  // column bits
};

__device__ __forceinline__ uint64_t load_key(const KeySrc& s, uint32_t i) {
  if (s.keys) return s.keys[i];
  uint32_t cell = i / s.LL;
  uint32_t r = i - cell * s.LL;
  uint32_t a = r / s.L, b = r - a * s.L;
  const int32_t* d = s.dofs + (int64_t)cell * s.L;
  return ((uint64_t)(uint32_t)__ldg(d + a) << s.cb) | (uint64_t)(uint32_t)__ldg(d + b);
}

// ---- pass 1/3: per-tile digit histogram, written digit-major ---------------------------------
__global__ void __launch_bounds__(SORT_BLOCK) k_radix_hist(KeySrc src, uint32_t n, int shift, uint32_t ntiles,
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
      atomicAdd(&h[d], 1u);
    }
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// ---- two-level exclusive scan over uint32 (reduce / scan partials / apply) -------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

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
struct ScatterSmem {
  uint32_t whist[SORT_WARPS][RADIX];
  uint32_t dstart[RADIX];
  uint32_t goff[RADIX];
  uint32_t ws[8];
  uint32_t svals[SORT_TILE];
  uint64_t skeys[SORT_TILE];
};

__global__ void __launch_bounds__(SORT_BLOCK) k_radix_scatter(KeySrc src, const uint32_t* __restrict__ vals_in, uint32_t n,
                                                              int shift, uint32_t ntiles,
                                                              const uint32_t* __restrict__ hist_scanned,
                                                              uint64_t* __restrict__ keys_out,
                                                              uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t tile0 = blockIdx.x * (uint32_t)SORT_TILE;
  const uint32_t count = (n - tile0 < (uint32_t)SORT_TILE) ? (n - tile0) : (uint32_t)SORT_TILE;

#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i) sm.whist[i][tid] = 0;
  sm.goff[tid] = hist_scanned[(uint64_t)tid * ntiles + blockIdx.x];
  __syncthreads();

  uint64_t key[SORT_ITEMS];
  uint16_t rank[SORT_ITEMS];
  const uint32_t wbase = tile0 + w * (32 * SORT_ITEMS);
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    uint32_t idx = wbase + i * 32 + lane;
    key[i] = (idx < n) ? load_key(src, idx) : ~0ull;
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
    uint64_t k = sm.skeys[j];
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

// ---- numeric: warp-shuffle segmented reduce --------------------------------------------------
constexpr int NUM_BLOCK = 256;

__global__ void __launch_bounds__(NUM_BLOCK) k_numeric(int64_t nnz, const uint32_t* __restrict__ perm,
                                                       const uint32_t* __restrict__ seg,
                                                       const double* __restrict__ vals, double* __restrict__ data) {
  __shared__ double acc[NUM_BLOCK / 32][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t k0 = ((int64_t)blockIdx.x * (NUM_BLOCK / 32) + w) * 32;
  if (k0 >= nnz) return;
  const int64_t k = k0 + lane;
  const bool own = k < nnz;
  const uint32_t s_l = seg[own ? k : nnz];
  const uint32_t S = __shfl_sync(FULL, s_l, 0);
  const int64_t kend = (k0 + 32 < nnz) ? k0 + 32 : nnz;
  const uint32_t E = seg[kend];
  acc[w][lane] = 0.0;
  __syncwarp();
  int heads_before = 0;
  for (uint32_t chunk = S; chunk < E; chunk += 32) {
    uint32_t p = chunk + lane;
    bool valid = p < E;
    double v = valid ? vals[perm[p]] : 0.0;
    uint32_t rel = s_l - chunk;  // wraps to a huge value when s_l < chunk
    uint32_t mybit = (own && rel < 32u) ? (1u << rel) : 0u;
    uint32_t heads = __reduce_or_sync(FULL, mybit);
    // segmented inclusive scan: add the value d lanes below unless a head lies in (lane-d, lane]
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double t = __shfl_up_sync(FULL, v, d);
      uint32_t upto = (lane == 31) ? FULL : ((2u << lane) - 1u);
      uint32_t below = (lane >= d) ? ((1u << (lane - d + 1)) - 1u) : FULL;
      if (lane >= d && !(heads & upto & ~below)) v += t;
    }
    uint32_t upto = (lane == 31) ? FULL : ((2u << lane) - 1u);
    bool tail = valid && (lane == 31 || ((heads >> (lane + 1)) & 1u) || p == E - 1);
    int out_lane = heads_before + __popc(heads & upto) - 1;
    if (tail) acc[w][out_lane] += v;
    heads_before += __popc(heads);
    __syncwarp();
  }
  if (own) data[k] = acc[w][lane];
}

int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

struct WsLayout {
  int64_t keys_a, keys_b, perm_b, hist, parts, meta, total;
};

WsLayout ws_layout(int64_t n_coo) {
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t ntiles = pgb_cdiv(n_coo > 0 ? n_coo : 1, SORT_TILE);
  WsLayout l;
  int64_t o = 0;
  l.keys_a = o; o += al(n_coo * 8);
  l.keys_b = o; o += al(n_coo * 8);
  l.perm_b = o; o += al(n_coo * 4);
  l.hist = o;   o += al(ntiles * RADIX * 4);
  l.parts = o;  o += al(2048 * 4);
  l.meta = o;   o += 256;  // [0]=which key buffer holds the sorted keys, [1]=cb, [2]=nnz (uint32)
  l.total = o;
  return l;
}

int scan_u32(uint32_t* data, uint64_t n, uint32_t* parts, cudaStream_t s) {
  uint64_t g = pgb_cdiv((int64_t)n, 1024);
  if (g > SCAN_G) g = SCAN_G;
  if (g < 1) g = 1;
  uint64_t chunk = pgb_cdiv((int64_t)n, (int64_t)g);
  chunk = pgb_cdiv((int64_t)chunk, 1024) * 1024;
  g = pgb_cdiv((int64_t)n, (int64_t)chunk);
  k_scan_reduce<<<(unsigned)g, 256, 0, s>>>(data, n, chunk, parts);
  k_scan_parts<<<1, 256, 0, s>>>(parts, (int)g, nullptr);
  k_scan_apply<<<(unsigned)g, 256, 0, s>>>(data, n, chunk, parts);
  PGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int64_t pgb_csr_workspace_bytes(int64_t n_coo, int64_t n_rows) {
  (void)n_rows;
  return ws_layout(n_coo).total;
}

extern "C" int pgb_csr_symbolic_sort(int64_t n_rows, int64_t nc, int L, const int32_t* cell_dofs,
                                     void* workspace, int64_t workspace_bytes, uint32_t* perm,
                                     int64_t* nnz_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_coo = nc * L * L;
  PGB_REQUIRE(n_coo < ((int64_t)1 << 32) - SORT_TILE, "pgb_csr_symbolic_sort: n_coo must be < 2^32");
  PGB_REQUIRE(n_rows > 0 && n_rows < ((int64_t)1 << 31), "pgb_csr_symbolic_sort: n_rows out of range");
  WsLayout l = ws_layout(n_coo);
  PGB_REQUIRE(workspace_bytes >= l.total, "pgb_csr_symbolic_sort: workspace too small");
  if (n_coo == 0) {
    *nnz_host = 0;
    return 0;
  }
  char* ws = (char*)workspace;
  uint64_t* kbuf[2] = {(uint64_t*)(ws + l.keys_a), (uint64_t*)(ws + l.keys_b)};
  uint32_t* pbuf[2] = {perm, (uint32_t*)(ws + l.perm_b)};
  uint32_t* hist = (uint32_t*)(ws + l.hist);
  uint32_t* parts = (uint32_t*)(ws + l.parts);
  uint32_t* meta = (uint32_t*)(ws + l.meta);

  const int cb = bits_for(n_rows);
  const int total_bits = 2 * cb;
  const int passes = (total_bits + RADIX_BITS - 1) / RADIX_BITS;
  const uint32_t n = (uint32_t)n_coo;
  const uint32_t ntiles = (uint32_t)pgb_cdiv(n_coo, SORT_TILE);

  static bool attr_set = false;
  if (!attr_set) {
    PGB_CHECK(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(ScatterSmem)));
    attr_set = true;
  }
  // ping-pong so that the final payload lands in `perm` (pbuf[0]): pass p writes pbuf[dst].
  int dst = (passes % 2 == 1) ? 0 : 1;
  KeySrc src{nullptr, cell_dofs, (uint32_t)L, (uint32_t)(L * L), cb};
  const uint32_t* pin = nullptr;
  for (int p = 0; p < passes; ++p) {
    int shift = p * RADIX_BITS;
    k_radix_hist<<<ntiles, SORT_BLOCK, 0, s>>>(src, n, shift, ntiles, hist);
    PGB_LAUNCH_CHECK();
    if (scan_u32(hist, (uint64_t)ntiles * RADIX, parts, s)) return 1;
    k_radix_scatter<<<ntiles, SORT_BLOCK, sizeof(ScatterSmem), s>>>(src, pin, n, shift, ntiles, hist,
                                                                     kbuf[dst], pbuf[dst]);
    PGB_LAUNCH_CHECK();
    src.keys = kbuf[dst];
    pin = pbuf[dst];
    dst ^= 1;
  }
  const int sorted_buf = dst ^ 1;  // buffer written by the last pass
  // count distinct keys
  uint64_t g = pgb_cdiv(n_coo, 1024);
  if (g > SCAN_G) g = SCAN_G;
  uint64_t chunk = pgb_cdiv(pgb_cdiv(n_coo, (int64_t)g), 1024) * 1024;
  g = pgb_cdiv(n_coo, (int64_t)chunk);
  k_count_heads<<<(unsigned)g, 256, 0, s>>>(kbuf[sorted_buf], (uint64_t)n_coo, chunk, parts);
  k_scan_parts<<<1, 256, 0, s>>>(parts, (int)g, meta + 2);
  PGB_LAUNCH_CHECK();
  uint32_t m[2] = {(uint32_t)sorted_buf, (uint32_t)cb};
  PGB_CHECK(cudaMemcpyAsync(meta, m, sizeof(m), cudaMemcpyHostToDevice, s));
  uint32_t nnz32 = 0;
  PGB_CHECK(cudaMemcpyAsync(&nnz32, meta + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  *nnz_host = (int64_t)nnz32;
  return 0;
}

extern "C" int pgb_csr_symbolic_finish(int64_t n_rows, int64_t n_coo, int64_t nnz,
                                       const void* workspace, void* indptr, int idx64,
                                       int32_t* indices, uint32_t* seg, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_coo == 0) {
    if (idx64) k_fill_indptr_empty<int64_t><<<pgb_grid(n_rows + 1, 256), 256, 0, s>>>(n_rows, (int64_t*)indptr);
    else k_fill_indptr_empty<int32_t><<<pgb_grid(n_rows + 1, 256), 256, 0, s>>>(n_rows, (int32_t*)indptr);
    PGB_LAUNCH_CHECK();
    return 0;
  }
  WsLayout l = ws_layout(n_coo);
  const char* ws = (const char*)workspace;
  uint32_t m[2];
  PGB_CHECK(cudaMemcpyAsync(m, ws + l.meta, sizeof(m), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  const uint64_t* keys = (const uint64_t*)(ws + (m[0] == 0 ? l.keys_a : l.keys_b));
  const uint32_t* parts = (const uint32_t*)(ws + l.parts);
  const int cb = (int)m[1];
  uint64_t g = pgb_cdiv(n_coo, 1024);
  if (g > SCAN_G) g = SCAN_G;
  uint64_t chunk = pgb_cdiv(pgb_cdiv(n_coo, (int64_t)g), 1024) * 1024;
  g = pgb_cdiv(n_coo, (int64_t)chunk);
  if (idx64)
    k_emit_unique<int64_t><<<(unsigned)g, 256, 0, s>>>(keys, (uint64_t)n_coo, chunk, parts, cb, n_rows,
                                                       (uint32_t)nnz, (int64_t*)indptr, indices, seg);
  else
    k_emit_unique<int32_t><<<(unsigned)g, 256, 0, s>>>(keys, (uint64_t)n_coo, chunk, parts, cb, n_rows,
                                                       (uint32_t)nnz, (int32_t*)indptr, indices, seg);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_csr_numeric(int64_t nnz, const uint32_t* perm, const uint32_t* seg,
                               const double* vals, double* data, void* stream) {
  if (nnz == 0) return 0;
  int64_t warps = pgb_cdiv(nnz, 32);
  unsigned grid = (unsigned)pgb_cdiv(warps, NUM_BLOCK / 32);
  k_numeric<<<grid, NUM_BLOCK, 0, (cudaStream_t)stream>>>(nnz, perm, seg, vals, data);
  PGB_LAUNCH_CHECK();
  return 0;
}
