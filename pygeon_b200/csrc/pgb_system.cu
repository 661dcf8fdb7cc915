// System build around the assembled matrix, in HBM: row / column selection of a
// CSR matrix with column renumbering (LinearSystem.reduce_system's R A R^T, numerics/linear_system.py:64-77;
// the owned-row block of the cell-partitioned multi-GPU path), diagonal extraction, the block-Jacobi
// diagonal of the symmetrised Darcy saddle system, and small gather / scatter helpers.
//
// Everything is a count -> scan -> fill pipeline: scratch is O(rows) (+ the outputs); no
// per-entry temporaries, so the 2.3e9-entry config #3 matrix can be reduced beside itself.
#include "pgb_common.cuh"

namespace {

constexpr int SB = 256;
constexpr int SCAN_PARTS = 2048;
constexpr unsigned FULL = 0xffffffffu;

// value of a scan input: masks (1-byte) count as 0 / 1 whatever byte the caller stored
template <typename T>
__device__ __forceinline__ int64_t scan_val(T x) {
  if constexpr (sizeof(T) == 1) return x != 0 ? 1 : 0;
  else return (int64_t)x;
}

__device__ __forceinline__ int64_t warp_incl_scan64(int64_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread over a 256-thread block; total in every thread
__device__ __forceinline__ int64_t block_excl_scan64(int64_t v, int64_t* ws /*[SB/32]*/, int64_t& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t inc = warp_incl_scan64(v);
  __syncthreads();
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  int64_t off = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SB / 32; ++i) {
    int64_t t = ws[i];
    if (i < w) off += t;
    tot += t;
  }
  total = tot;
  return inc - v + off;
}

// ---- two-level exclusive scan of int64 values (in place), n + 1 entries: data[n] receives the total ----
template <typename T>
__global__ void __launch_bounds__(SB) k_scan64_reduce(const T* __restrict__ in, int64_t n, int64_t chunk,
                                                      int64_t* parts) {
  __shared__ int64_t ws[SB / 32];
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, n);
  int64_t s = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += SB) s += scan_val(in[i]);
  int64_t tot;
  block_excl_scan64(s, ws, tot);
  if (threadIdx.x == 0) parts[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SB) k_scan64_parts(int64_t* parts, int g, int64_t* total_out) {
  __shared__ int64_t ws[SB / 32];
  constexpr int PER = SCAN_PARTS / SB;
  int64_t v[PER], s = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int i = threadIdx.x * PER + k;
    v[k] = i < g ? parts[i] : 0;
    s += v[k];
  }
  int64_t tot;
  int64_t ex = block_excl_scan64(s, ws, tot);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int i = threadIdx.x * PER + k;
    if (i < g) parts[i] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0 && total_out) *total_out = tot;
}

// out[i] = exclusive prefix of in[0..i); out may alias in when T == int64_t
template <typename T>
__global__ void __launch_bounds__(SB) k_scan64_apply(const T* in, int64_t* out, int64_t n, int64_t chunk,
                                                     const int64_t* __restrict__ parts) {
  __shared__ int64_t ws[SB / 32];
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, n);
  int64_t carry = parts[blockIdx.x];
  for (int64_t base = lo; base < hi; base += SB) {
    const int64_t i = base + threadIdx.x;
    int64_t v = i < hi ? scan_val(in[i]) : 0;
    int64_t tot;
    int64_t ex = block_excl_scan64(v, ws, tot) + carry;
    if (i < hi) out[i] = ex;
    carry += tot;
    __syncthreads();
  }
}

template <typename T>
int scan64(const T* in, int64_t* out, int64_t n, int64_t* parts /*[SCAN_PARTS + 1]*/, cudaStream_t s) {
  // out[0..n) = exclusive scan, out[n] = total
  int64_t g = pgb_cdiv(n, (int64_t)SB * 8);
  if (g > SCAN_PARTS) g = SCAN_PARTS;
  if (g < 1) g = 1;
  int64_t chunk = pgb_cdiv(pgb_cdiv(n, g), SB) * SB;
  if (chunk < SB) chunk = SB;
  g = pgb_cdiv(n, chunk);
  if (g < 1) g = 1;
  k_scan64_reduce<T><<<(unsigned)g, SB, 0, s>>>(in, n, chunk, parts);
  k_scan64_parts<<<1, SB, 0, s>>>(parts, (int)g, out + n);
  k_scan64_apply<T><<<(unsigned)g, SB, 0, s>>>(in, out, n, chunk, parts);
  PGB_LAUNCH_CHECK();
  return 0;
}

// newid[i] = rank of i among the kept entries or -1; rows[rank] = i.  pre = exclusive scan of the mask.
__global__ void __launch_bounds__(SB) k_mask_emit(int64_t n, const uint8_t* __restrict__ mask,
                                                  const int64_t* __restrict__ pre, int32_t* __restrict__ newid,
                                                  int32_t* __restrict__ rows) {
  const int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x;
  if (i >= n) return;
  if (mask[i]) {
    const int32_t r = (int32_t)pre[i];
    if (newid) newid[i] = r;
    if (rows) rows[r] = (int32_t)i;
  } else if (newid) {
    newid[i] = -1;
  }
}

// ---- extraction: out row i = in row rows[i], columns renumbered through col_map (-1 = dropped) -----
constexpr int XT = 8;  // lanes per row

template <typename IP>
__global__ void __launch_bounds__(SB) k_extract_count(int64_t n_out, const int32_t* __restrict__ rows,
                                                      const IP* __restrict__ ip, const int32_t* __restrict__ indices,
                                                      const int32_t* __restrict__ col_map, int64_t* __restrict__ cnt) {
  const int sub = threadIdx.x % XT;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / XT;
  const int64_t nsub = (int64_t)gridDim.x * (SB / XT);
  for (int64_t i = gid; i < n_out; i += nsub) {
    const int64_t r = rows ? rows[i] : i;
    const IP lo = ip[r], hi = ip[r + 1];
    int c = 0;
    for (IP j = lo + sub; j < hi; j += XT) c += (col_map == nullptr || col_map[indices[j]] >= 0) ? 1 : 0;
#pragma unroll
    for (int d = XT / 2; d > 0; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
    if (sub == 0) cnt[i] = c;
  }
}

template <typename IP>
__global__ void __launch_bounds__(SB) k_extract_fill(int64_t n_out, const int32_t* __restrict__ rows,
                                                     const IP* __restrict__ ip, const int32_t* __restrict__ indices,
                                                     const double* __restrict__ data,
                                                     const int32_t* __restrict__ col_map,
                                                     const int64_t* __restrict__ out_ip, int32_t* __restrict__ out_idx,
                                                     double* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int sub = threadIdx.x % XT;
  const int grp_shift = (lane / XT) * XT;
  const unsigned grp_mask = ((1u << XT) - 1u) << grp_shift;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / XT;
  const int64_t nsub = (int64_t)gridDim.x * (SB / XT);
  const int64_t rounds = (n_out + nsub - 1) / nsub;
  for (int64_t it = 0; it < rounds; ++it) {  // all lanes of a warp run the same number of rounds (ballots)
    const int64_t i = gid + it * nsub;
    const bool live = i < n_out;
    const int64_t r = live ? (rows ? rows[i] : i) : 0;
    const IP lo = live ? ip[r] : (IP)0, hi = live ? ip[r + 1] : (IP)0;
    int64_t w = live ? out_ip[i] : 0;
    IP len = hi - lo;
    // the longest row of the warp's 4 groups decides the trip count
    IP mx = len;
#pragma unroll
    for (int d = 16; d >= XT; d >>= 1) {
      IP o = __shfl_xor_sync(FULL, mx, d);
      mx = o > mx ? o : mx;
    }
    for (IP k = 0; k < mx; k += XT) {
      const IP j = lo + k + sub;
      const bool in = j < hi;
      int32_t c = in ? indices[j] : 0;
      if (in && col_map) c = col_map[c];
      const bool keep = in && c >= 0;
      const unsigned b = __ballot_sync(FULL, keep) & grp_mask;
      if (keep) {
        const int64_t p = w + __popc(b & ((1u << lane) - 1u));
        out_idx[p] = c;
        out_val[p] = data[j];
      }
      w += __popc(b);
    }
  }
}

template <typename IP>
__global__ void __launch_bounds__(SB) k_csr_diag(int64_t n, const IP* __restrict__ ip,
                                                 const int32_t* __restrict__ indices, const double* __restrict__ data,
                                                 double* __restrict__ diag) {
  constexpr int T = 4;
  const int sub = threadIdx.x % T;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / T;
  const int64_t nsub = (int64_t)gridDim.x * (SB / T);
  for (int64_t r = gid; r < n; r += nsub) {
    const IP lo = ip[r], hi = ip[r + 1];
    double v = 0.0;
    for (IP j = lo + sub; j < hi; j += T)
      if (indices[j] == (int32_t)r) v += data[j];  // one entry per (row, col) in the matrices built here
#pragma unroll
    for (int d = T / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    if (sub == 0) diag[r] = v;
  }
}

// rows nf.. of the symmetrised saddle matrix: 1 / sum_{col < nf} a^2 / du[col]; rows < nf: 1 / du
template <typename IP>
__global__ void __launch_bounds__(SB) k_darcy_jacobi(int64_t n, int64_t nf, const IP* __restrict__ ip,
                                                     const int32_t* __restrict__ indices,
                                                     const double* __restrict__ data, const double* __restrict__ du,
                                                     double* __restrict__ dinv) {
  constexpr int T = 4;
  const int sub = threadIdx.x % T;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / T;
  const int64_t nsub = (int64_t)gridDim.x * (SB / T);
  for (int64_t r = gid; r < n; r += nsub) {
    if (r < nf) {
      if (sub == 0) dinv[r] = 1.0 / du[r];
      continue;
    }
    const IP lo = ip[r], hi = ip[r + 1];
    double v = 0.0;
    for (IP j = lo + sub; j < hi; j += T) {
      const int32_t c = indices[j];
      if (c < nf) {
        const double a = data[j];
        v += a * a / du[c];
      }
    }
#pragma unroll
    for (int d = T / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    if (sub == 0) dinv[r] = 1.0 / v;
  }
}

// flags[col] = 1 for every column referenced by the selected rows (rows == nullptr: all rows)
template <typename IP>
__global__ void __launch_bounds__(SB) k_mark_columns(int64_t n_out, const int32_t* __restrict__ rows,
                                                     const IP* __restrict__ ip, const int32_t* __restrict__ indices,
                                                     uint8_t* __restrict__ flags) {
  const int sub = threadIdx.x % XT;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / XT;
  const int64_t nsub = (int64_t)gridDim.x * (SB / XT);
  for (int64_t i = gid; i < n_out; i += nsub) {
    const int64_t r = rows ? rows[i] : i;
    const IP lo = ip[r], hi = ip[r + 1];
    for (IP j = lo + sub; j < hi; j += XT) flags[indices[j]] = 1;  // same value from every writer
  }
}

// touches[r] = 1 when row r has an entry with column >= first
template <typename IP>
__global__ void __launch_bounds__(SB) k_rows_touching(int64_t n, const IP* __restrict__ ip,
                                                      const int32_t* __restrict__ indices, int32_t first,
                                                      uint8_t* __restrict__ touches) {
  const int sub = threadIdx.x % XT;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / XT;
  const int64_t nsub = (int64_t)gridDim.x * (SB / XT);
  for (int64_t r = gid; r < n; r += nsub) {
    const IP lo = ip[r], hi = ip[r + 1];
    int hit = 0;
    for (IP j = lo + sub; j < hi; j += XT) hit |= indices[j] >= first ? 1 : 0;
#pragma unroll
    for (int d = XT / 2; d > 0; d >>= 1) hit |= __shfl_xor_sync(FULL, hit, d);
    if (sub == 0) touches[r] = (uint8_t)hit;
  }
}

__global__ void __launch_bounds__(SB) k_gather(int64_t n, const int32_t* __restrict__ idx,
                                               const double* __restrict__ x, const double* __restrict__ y, double a,
                                               double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x; i < n; i += (int64_t)gridDim.x * SB) {
    const int32_t j = idx[i];
    out[i] = y ? x[j] + a * y[j] : x[j];
  }
}

__global__ void __launch_bounds__(SB) k_scatter_add(int64_t n, const int32_t* __restrict__ idx,
                                                    const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x; i < n; i += (int64_t)gridDim.x * SB)
    dst[idx[i]] += src[i];  // idx holds distinct positions
}

// ---- block stacking (sps.block_array of CSR blocks; utils/bmat.py, numerics/innerproducts.py:162-232) ------
// Row r of a block row is the concatenation, left to right, of row r of its blocks with the columns shifted
// by the block column's offset: sorted as long as every block row is.  Count: lens[r] += length of row r;
// fill: one launch per block in left-to-right order, every row appended at its cursor; the cursors advance by a
// k_block_lengths launch after it (no read / write race inside the fill).
template <typename IP>
__global__ void __launch_bounds__(SB) k_block_lengths(int64_t n, const IP* __restrict__ ip, int64_t* __restrict__ lens) {
  for (int64_t r = (int64_t)blockIdx.x * SB + threadIdx.x; r < n; r += (int64_t)gridDim.x * SB)
    lens[r] += (int64_t)(ip[r + 1] - ip[r]);
}

template <typename IP>
__global__ void __launch_bounds__(SB) k_block_place(int64_t n, const IP* __restrict__ ip,
                                                    const int32_t* __restrict__ indices,
                                                    const double* __restrict__ data, int32_t col_off, double scale,
                                                    const int64_t* __restrict__ cursor, int32_t* __restrict__ out_idx,
                                                    double* __restrict__ out_val) {
  const int sub = threadIdx.x % XT;
  const int64_t gid = ((int64_t)blockIdx.x * SB + threadIdx.x) / XT;
  const int64_t nsub = (int64_t)gridDim.x * (SB / XT);
  for (int64_t r = gid; r < n; r += nsub) {
    const IP lo = ip[r], hi = ip[r + 1];
    const int64_t w = cursor[r];
    for (IP j = lo + sub; j < hi; j += XT) {
      out_idx[w + (int64_t)(j - lo)] = indices[j] + col_off;
      out_val[w + (int64_t)(j - lo)] = scale * data[j];
    }
  }
}

// data[position of (rows[i], rows[i])] += vals[i]; rows distinct.  A row without a stored diagonal raises *missing.
template <typename IP>
__global__ void __launch_bounds__(SB) k_add_diagonal(int64_t m, const int32_t* __restrict__ rows,
                                                     const double* __restrict__ vals, const IP* __restrict__ ip,
                                                     const int32_t* __restrict__ indices, double* __restrict__ data,
                                                     int32_t* missing) {
  for (int64_t i = (int64_t)blockIdx.x * SB + threadIdx.x; i < m; i += (int64_t)gridDim.x * SB) {
    const int32_t r = rows[i];
    IP lo = ip[r], hi = ip[r + 1];
    while (lo < hi) {  // sorted row: lower bound of r
      const IP mid = lo + (hi - lo) / 2;
      if (indices[mid] < r) lo = mid + 1;
      else hi = mid;
    }
    if (lo < ip[r + 1] && indices[lo] == r) data[lo] += vals[i];
    else atomicAdd(missing, 1);
  }
}

unsigned grid_for(int64_t work_items, int per_block) {
  int64_t g = pgb_cdiv(work_items, per_block);
  const int64_t cap = (int64_t)PGB_SMS * 16;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int64_t pgb_system_workspace_bytes(int64_t n) {
  return (int64_t)sizeof(int64_t) * (n + 1 + SCAN_PARTS + 8);
}

extern "C" int pgb_mask_to_map(int64_t n, const uint8_t* mask, int32_t* newid, int32_t* rows,
                               int64_t* count_host, void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(workspace_bytes >= pgb_system_workspace_bytes(n), "pgb_mask_to_map: workspace too small");
  PGB_REQUIRE(n < (int64_t)1 << 31, "pgb_mask_to_map: ids are int32");
  int64_t* pre = (int64_t*)workspace;
  int64_t* parts = pre + n + 1;
  if (n == 0) {
    if (count_host) *count_host = 0;
    return 0;
  }
  if (scan64<uint8_t>(mask, pre, n, parts, s)) return 1;
  if (count_host) {  // the caller sizes `rows` from this: it may pass rows == NULL here and call again
    PGB_CHECK(cudaMemcpyAsync(count_host, pre + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  if (newid || rows) {
    k_mask_emit<<<pgb_grid(n, SB), SB, 0, s>>>(n, mask, pre, newid, rows);
    PGB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int pgb_csr_extract_symbolic(int64_t n_out, const int32_t* rows, const void* indptr, int idx64,
                                        const int32_t* indices, const int32_t* col_map, int64_t* out_indptr,
                                        int64_t* nnz_host, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(workspace_bytes >= pgb_system_workspace_bytes(n_out), "pgb_csr_extract_symbolic: workspace too small");
  int64_t* cnt = (int64_t*)workspace;
  int64_t* parts = cnt + n_out + 1;
  if (n_out == 0) {
    PGB_CHECK(cudaMemsetAsync(out_indptr, 0, sizeof(int64_t), s));
    if (nnz_host) *nnz_host = 0;
    return 0;
  }
  const unsigned g = grid_for(n_out, SB / XT);
  if (idx64) k_extract_count<int64_t><<<g, SB, 0, s>>>(n_out, rows, (const int64_t*)indptr, indices, col_map, cnt);
  else k_extract_count<int32_t><<<g, SB, 0, s>>>(n_out, rows, (const int32_t*)indptr, indices, col_map, cnt);
  PGB_LAUNCH_CHECK();
  if (scan64<int64_t>(cnt, out_indptr, n_out, parts, s)) return 1;
  if (nnz_host) {
    PGB_CHECK(cudaMemcpyAsync(nnz_host, out_indptr + n_out, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

extern "C" int pgb_csr_extract_fill(int64_t n_out, const int32_t* rows, const void* indptr, int idx64,
                                    const int32_t* indices, const double* data, const int32_t* col_map,
                                    const int64_t* out_indptr, int32_t* out_indices, double* out_data,
                                    void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_out == 0) return 0;
  const unsigned g = grid_for(n_out, SB / XT);
  if (idx64)
    k_extract_fill<int64_t><<<g, SB, 0, s>>>(n_out, rows, (const int64_t*)indptr, indices, data, col_map, out_indptr,
                                              out_indices, out_data);
  else
    k_extract_fill<int32_t><<<g, SB, 0, s>>>(n_out, rows, (const int32_t*)indptr, indices, data, col_map, out_indptr,
                                              out_indices, out_data);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_csr_diagonal(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                                const double* data, double* diag, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return 0;
  const unsigned g = grid_for(n, SB / 4);
  if (idx64) k_csr_diag<int64_t><<<g, SB, 0, s>>>(n, (const int64_t*)indptr, indices, data, diag);
  else k_csr_diag<int32_t><<<g, SB, 0, s>>>(n, (const int32_t*)indptr, indices, data, diag);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_darcy_block_jacobi(int64_t n, int64_t num_faces, const void* indptr, int idx64,
                                      const int32_t* indices, const double* data, double* diag_work,
                                      double* dinv, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return 0;
  PGB_REQUIRE(num_faces >= 0 && num_faces <= n, "pgb_darcy_block_jacobi: bad block size");
  if (pgb_csr_diagonal(num_faces, indptr, idx64, indices, data, diag_work, stream)) return 1;
  const unsigned g = grid_for(n, SB / 4);
  if (idx64)
    k_darcy_jacobi<int64_t><<<g, SB, 0, s>>>(n, num_faces, (const int64_t*)indptr, indices, data, diag_work, dinv);
  else
    k_darcy_jacobi<int32_t><<<g, SB, 0, s>>>(n, num_faces, (const int32_t*)indptr, indices, data, diag_work, dinv);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_gather(int64_t n, const int32_t* idx, const double* x, const double* y, double a, double* out,
                          void* stream) {
  if (n == 0) return 0;
  k_gather<<<grid_for(n, SB), SB, 0, (cudaStream_t)stream>>>(n, idx, x, y, a, out);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_scatter_add(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream) {
  if (n == 0) return 0;
  k_scatter_add<<<grid_for(n, SB), SB, 0, (cudaStream_t)stream>>>(n, idx, src, dst);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_csr_mark_columns(int64_t n_out, const int32_t* rows, const void* indptr, int idx64,
                                    const int32_t* indices, uint8_t* flags, void* stream) {
  if (n_out == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = grid_for(n_out, SB / XT);
  if (idx64) k_mark_columns<int64_t><<<g, SB, 0, s>>>(n_out, rows, (const int64_t*)indptr, indices, flags);
  else k_mark_columns<int32_t><<<g, SB, 0, s>>>(n_out, rows, (const int32_t*)indptr, indices, flags);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_csr_rows_touching(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                                     int32_t first_col, uint8_t* touches, void* stream) {
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = grid_for(n, SB / XT);
  if (idx64) k_rows_touching<int64_t><<<g, SB, 0, s>>>(n, (const int64_t*)indptr, indices, first_col, touches);
  else k_rows_touching<int32_t><<<g, SB, 0, s>>>(n, (const int32_t*)indptr, indices, first_col, touches);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_scan_i64(int64_t n, const int64_t* in, int64_t* out, int64_t* total_host, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(workspace_bytes >= (int64_t)sizeof(int64_t) * (SCAN_PARTS + 8), "pgb_scan_i64: workspace too small");
  if (n == 0) {
    PGB_CHECK(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
    if (total_host) *total_host = 0;
    return 0;
  }
  if (scan64<int64_t>(in, out, n, (int64_t*)workspace, s)) return 1;
  if (total_host) {
    PGB_CHECK(cudaMemcpyAsync(total_host, out + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

extern "C" int pgb_csr_block_lengths(int64_t n, const void* indptr, int idx64, int64_t* lens, void* stream) {
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = grid_for(n, SB);
  if (idx64) k_block_lengths<int64_t><<<g, SB, 0, s>>>(n, (const int64_t*)indptr, lens);
  else k_block_lengths<int32_t><<<g, SB, 0, s>>>(n, (const int32_t*)indptr, lens);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_csr_block_place(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                                   const double* data, int32_t col_offset, double scale, int64_t* cursor,
                                   int32_t* out_indices, double* out_data, void* stream) {
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = grid_for(n, SB / XT);
  if (idx64)
    k_block_place<int64_t><<<g, SB, 0, s>>>(n, (const int64_t*)indptr, indices, data, col_offset, scale, cursor,
                                             out_indices, out_data);
  else
    k_block_place<int32_t><<<g, SB, 0, s>>>(n, (const int32_t*)indptr, indices, data, col_offset, scale, cursor,
                                             out_indices, out_data);
  PGB_LAUNCH_CHECK();
  return pgb_csr_block_lengths(n, indptr, idx64, cursor, stream);
}

extern "C" int pgb_csr_add_diagonal(int64_t m, const int32_t* rows, const double* vals, const void* indptr,
                                    int idx64, const int32_t* indices, double* data, int32_t* flag_dev,
                                    void* stream) {
  if (m == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  PGB_CHECK(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), s));
  const unsigned g = grid_for(m, SB);
  if (idx64) k_add_diagonal<int64_t><<<g, SB, 0, s>>>(m, rows, vals, (const int64_t*)indptr, indices, data, flag_dev);
  else k_add_diagonal<int32_t><<<g, SB, 0, s>>>(m, rows, vals, (const int32_t*)indptr, indices, data, flag_dev);
  PGB_LAUNCH_CHECK();
  int32_t missing = 0;
  PGB_CHECK(cudaMemcpyAsync(&missing, flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  PGB_REQUIRE(missing == 0, "pgb_csr_add_diagonal: a listed row has no stored diagonal entry");
  return 0;
}
