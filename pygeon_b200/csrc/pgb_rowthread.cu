// K2 symbolic, short rows: one thread per row (k_rowthread).
#include "pgb_rows.cuh"

namespace {

using namespace pgb_rows;

// ---- short rows: one THREAD per row ---------------------------------------------------------------
// Rows of face / cell / face-node dofs have a handful of incident cells (RT0: 2, P0: 1, BDM1: 2), i.e.
// 5..32 candidate columns.  A sorting network over registers needs no shuffles, no shared memory and
// no predicates (a compare-exchange is two VIMNMX), and neighbouring threads own neighbouring rows,
// so the row offsets, the incidences, the slot bytes and the unique columns are all accessed in
// contiguous runs across the warp.  The thread also puts its (at most MI) incidences into canonical
// order and records inc_pos, like k_rowhash.  Output identical to k_rowsort.
// Block-compacted output: the 128 rows of a block stage their unique columns back to back in shared
// memory (row order) and their slot bytes at the block-relative incidence position; both are then
// written with coalesced stores: the columns to ucol_tmp[row_start[first]*L ...) (k_compact knows
// this layout through meta[5]), the slot bytes to their final place.
template <int N>
struct RowThreadBlock {
  static constexpr int value = 128;
};

template <int MI, int L, int N, typename KeyT>
__global__ void __launch_bounds__(RowThreadBlock<N>::value) k_rowthread(int64_t row0, int64_t n_rows, const uint32_t* __restrict__ row_start,
                                                               const uint32_t* __restrict__ inc_tmp,
                                                               const int32_t* __restrict__ dofs,
                                                               uint32_t* __restrict__ inc_pos, uint8_t* __restrict__ slots,
                                                               int32_t* __restrict__ ucol_tmp,
                                                               uint32_t* __restrict__ row_nnz,
                                                               uint32_t* __restrict__ max_nnz) {
  static_assert(MI * L <= N, "network too small");
  constexpr int LP = (L + 3) & ~3;
  constexpr int B = RowThreadBlock<N>::value;
  constexpr bool K32 = sizeof(KeyT) == 4;
  constexpr int EB = (N <= 8) ? 3 : (N <= 16) ? 4 : (N <= 32) ? 5 : (N <= 64) ? 6 : 7;
  constexpr int CSH = K32 ? EB : 32;
  __shared__ uint32_t s_u[B * MI * L];
  __shared__ __align__(16) uint8_t s_sl[B * MI * LP];
  __shared__ uint32_t s_w[B / 32], s_m[B / 32], s_rs0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t r = row0 + (int64_t)blockIdx.x * B + tid;  // rows [row0, n_rows), row0 a multiple of B
  uint32_t rs = 0, cnt = 0;
  if (r < n_rows) {
    rs = row_start[r];
    cnt = row_start[r + 1] - rs;
  }
  if (tid == 0) s_rs0 = rs;
  for (int j = tid; j < B * MI * LP / 4; j += B) reinterpret_cast<uint32_t*>(s_sl)[j] = 0u;
  uint32_t inc[MI];
#pragma unroll
  for (int q = 0; q < MI; ++q) inc[q] = ((uint32_t)q < cnt) ? inc_tmp[rs + q] : 0xffffffffu;
  if constexpr (MI > 1) reg_bitonic<MI, uint32_t>(inc);
  KeyT key[N];
#pragma unroll
  for (int i = 0; i < N; ++i) key[i] = (KeyT) ~(KeyT)0;
#pragma unroll
  for (int q = 0; q < MI; ++q) {
    if ((uint32_t)q < cnt) {
      inc_pos[inc[q]] = rs + q;
      uint32_t d[L];
      load_dofs<L>(dofs + (int64_t)(inc[q] / L) * L, d);
#pragma unroll
      for (int b = 0; b < L; ++b) key[q * L + b] = ((KeyT)d[b] << CSH) | (KeyT)(q * L + b);
    }
  }
  reg_bitonic<N, KeyT>(key);
  const uint32_t ncand = cnt * L;
  // number of distinct columns, block-exclusive scan
  uint32_t total = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if ((uint32_t)i < ncand) total += (i == 0 || (uint32_t)(key[i] >> CSH) != (uint32_t)(key[i > 0 ? i - 1 : 0] >> CSH)) ? 1u : 0u;
  uint32_t incl = warp_incl_scan(total);
  uint32_t mx = __reduce_max_sync(FULL, total);
  if (lane == 31) s_w[w] = incl;
  if (lane == 0) s_m[w] = mx;
  __syncthreads();  // s_rs0, s_w, s_m, zeroed s_sl
  uint32_t off = incl - total, btotal = 0;
#pragma unroll
  for (int i = 0; i < B / 32; ++i) {
    if (i < w) off += s_w[i];
    btotal += s_w[i];
  }
  const uint32_t rs0 = s_rs0;
  int run = -1;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if ((uint32_t)i < ncand) {
      const uint32_t col = (uint32_t)(key[i] >> CSH);
      if (i == 0 || col != (uint32_t)(key[i > 0 ? i - 1 : 0] >> CSH)) {
        ++run;
        s_u[off + run] = col;
      }
      const uint32_t e0 = (uint32_t)key[i] & ((1u << EB) - 1u);
      const uint32_t q0 = e0 / L;
      s_sl[(rs - rs0 + q0) * LP + (e0 - q0 * L)] = (uint8_t)run;
    }
  }
  if (r < n_rows) row_nnz[r] = total;
  __syncthreads();
  // coalesced write-out
  const int64_t r_end = row0 + ((int64_t)blockIdx.x + 1) * B < n_rows ? row0 + ((int64_t)blockIdx.x + 1) * B : n_rows;
  const uint32_t ninc = row_start[r_end] - rs0;
  for (uint32_t j = tid; j < btotal; j += B) ucol_tmp[(uint64_t)rs0 * L + j] = (int32_t)s_u[j];
  uint32_t* sl_out = reinterpret_cast<uint32_t*>(slots + (uint64_t)rs0 * LP);
  for (uint32_t j = tid; j < ninc * (LP / 4); j += B) sl_out[j] = reinterpret_cast<const uint32_t*>(s_sl)[j];
  if (tid == 0) {
    uint32_t m = s_m[0];
    for (int q = 1; q < B / 32; ++q) m = s_m[q] > m ? s_m[q] : m;
    atomicMax(max_nnz, m);
  }
}

template <int MI, int L, int N>
int launch_rowthread_k(int64_t row0, int64_t row1, int64_t n_rows, const uint32_t* row_start, const uint32_t* inc_tmp, const int32_t* dofs,
                       uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz, uint32_t* max_nnz,
                       cudaStream_t s) {
  constexpr int B = RowThreadBlock<N>::value;
  static_assert(B == ROWTHREAD_GROUP, "k_compact_groups expects groups of ROWTHREAD_GROUP rows");
  unsigned grid = (unsigned)pgb_cdiv(row1 - row0, B);
  const bool k32 = (uint64_t)n_rows * (uint64_t)N <= 0xffffffffull && !g_force_key64;
  char name[64];
  snprintf(name, sizeof(name), "k_rowthread<%d,%d,%d,%s>", MI, L, N, k32 ? "u32" : "u64");
  note_row_kernel(name);
  if (k32)
    k_rowthread<MI, L, N, uint32_t><<<grid, B, 0, s>>>(row0, row1, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz);
  else
    k_rowthread<MI, L, N, uint64_t><<<grid, B, 0, s>>>(row0, row1, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz);
  PGB_LAUNCH_CHECK();
  return 0;
}

// rows with at most 16 incidences and 64 candidates: one thread per row (k_rowthread)

}  // namespace

namespace pgb_rows {

// (a 128-key network for P1 rows in 3D - 24 incidences, 96 candidates - was measured slower than the
// hash dedupe: 1.54 ms vs 0.95 ms on config #2, 168 registers per thread)
bool rowthread_applies(uint32_t max_inc, uint32_t max_cand) {
  if (g_force_bitonic || max_inc == 0 || max_inc > 16) return false;
  const uint32_t mi = max_inc <= 1 ? 1 : max_inc <= 2 ? 2 : max_inc <= 4 ? 4 : max_inc <= 8 ? 8 : 16;
  return mi * (max_cand / max_inc) <= 64;
}

int launch_rowthread(uint32_t max_inc, int L, int64_t row0, int64_t row1, int64_t n_rows, const uint32_t* row_start, const uint32_t* inc_tmp,
                     const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz,
                     uint32_t* max_nnz, cudaStream_t s) {
  const int mi = max_inc <= 1 ? 1 : max_inc <= 2 ? 2 : max_inc <= 4 ? 4 : max_inc <= 8 ? 8 : 16;
#define PGB_RT(MI, LL, N) \
  if (L == LL && mi == MI)  \
    return launch_rowthread_k<MI, LL, N>(row0, row1, n_rows, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz, s)
  PGB_RT(1, 2, 8);  PGB_RT(2, 2, 8);  // d x d cell blocks in 2D
  PGB_RT(1, 3, 8);  PGB_RT(2, 3, 8);  PGB_RT(4, 3, 16); PGB_RT(8, 3, 32); PGB_RT(16, 3, 64);
  PGB_RT(1, 4, 8);  PGB_RT(2, 4, 8);  PGB_RT(4, 4, 16); PGB_RT(8, 4, 32); PGB_RT(16, 4, 64);
  PGB_RT(1, 5, 8);  PGB_RT(2, 5, 16); PGB_RT(4, 5, 32); PGB_RT(8, 5, 64);
  PGB_RT(1, 6, 8);  PGB_RT(2, 6, 16); PGB_RT(4, 6, 32); PGB_RT(8, 6, 64);
  PGB_RT(1, 10, 16); PGB_RT(2, 10, 32); PGB_RT(4, 10, 64);
  PGB_RT(1, 12, 16); PGB_RT(2, 12, 32); PGB_RT(4, 12, 64);
  PGB_RT(1, 8, 8);   PGB_RT(2, 8, 16);  PGB_RT(4, 8, 32);   // RT1 in 2D
  PGB_RT(1, 15, 16); PGB_RT(2, 15, 32); PGB_RT(4, 15, 64);  // RT1 in 3D
#undef PGB_RT
  pgb_set_error("k_rowthread: unsupported local size %d / %u incidences", L, max_inc);
  return 2;
}


}  // namespace pgb_rows
