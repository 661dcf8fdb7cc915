// Row-level kernels of the two-level symbolic phase (one translation unit per kernel family, so
// that the many template instances compile in parallel): shared device helpers and the launchers.
#pragma once
#include "pgb_common.cuh"

namespace pgb_rows {

constexpr unsigned FULL = 0xffffffffu;

// names of the row kernels the last symbolic_sort launched, joined by '+' (pgb_csr_last_row_kernel)
extern char g_row_kernel[256];
void note_row_kernel(const char* name);

// test switches (pgb_csr_symbolic_sort modes 2 and 3)
extern bool g_force_key64;    // 64-bit (col, position) keys in the row sorts
extern bool g_force_bitonic;  // always k_rowsort

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}


// bitonic sort of N = G*NPL keys held by a group of G consecutive lanes (element e = lane*NPL + i,
// `lane` = lane index inside the group)
// one key per lane: the all-ascending form of the network (the first step of every merge pairs
// lane i with its mirror i ^ (k-1)), so that min / max is decided by ONE compile-time bit of the
// lane index: a stage is SHFL + predicate + VIMNMX
template <int G>
__device__ __forceinline__ uint32_t group_bitonic1_u32(uint32_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= G; k <<= 1) {
    {
      uint32_t other = __shfl_xor_sync(FULL, key, k - 1);
      key = (lane & (k >> 1)) ? max(key, other) : min(key, other);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      uint32_t other = __shfl_xor_sync(FULL, key, j);
      key = (lane & j) ? max(key, other) : min(key, other);
    }
  }
  return key;
}

// NPL keys per lane, element e = lane*NPL + i: same all-ascending network; steps with partner
// distance >= NPL go through shuffles (the mirror step exchanges key[i] with key[NPL-1-i] of the
// partner lane), the others are compare-exchanges between registers of one lane
template <int NPL, int G = 32>
__device__ __forceinline__ void warp_bitonic_u32(uint32_t (&key)[NPL], int lane) {
  constexpr int N = G * NPL;
  if constexpr (NPL == 1) {
    key[0] = group_bitonic1_u32<G>(key[0], lane);
    return;
  }
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
    if (k <= NPL) {  // mirror step inside the lane
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int p = i ^ (k - 1);
        if (i < p) {
          uint32_t lo = min(key[i], key[p]), hi = max(key[i], key[p]);
          key[i] = lo, key[p] = hi;
        }
      }
    } else {
      const int lj = k / NPL - 1;
      uint32_t other[NPL];
#pragma unroll
      for (int i = 0; i < NPL; ++i) other[i] = __shfl_xor_sync(FULL, key[NPL - 1 - i], lj);
      const bool up = (lane & (k / NPL / 2)) != 0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) key[i] = up ? max(key[i], other[i]) : min(key[i], other[i]);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      if (j >= NPL) {
        const int lj = j / NPL;
        const bool up = (lane & lj) != 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          uint32_t other = __shfl_xor_sync(FULL, key[i], lj);
          key[i] = up ? max(key[i], other) : min(key[i], other);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          if ((i & j) == 0) {
            uint32_t lo = min(key[i], key[i | j]), hi = max(key[i], key[i | j]);
            key[i] = lo, key[i | j] = hi;
          }
        }
      }
    }
  }
}

// bitonic sorting network over N registers of one thread (no shuffles, no predicates: a compare-exchange is
// two VIMNMX)
template <int N, typename KeyT>
__device__ __forceinline__ void reg_bitonic(KeyT (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int p = i ^ (k - 1);
      if (i < p) {
        KeyT lo = a[i] < a[p] ? a[i] : a[p], hi = a[i] < a[p] ? a[p] : a[i];
        a[i] = lo, a[p] = hi;
      }
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int p = i ^ j;
        if (i < p) {
          KeyT lo = a[i] < a[p] ? a[i] : a[p], hi = a[i] < a[p] ? a[p] : a[i];
          a[i] = lo, a[p] = hi;
        }
      }
    }
  }
}

// dof row of one cell, vector loads where the row pitch allows
template <int L>
__device__ __forceinline__ void load_dofs(const int32_t* __restrict__ p, uint32_t (&d)[L]) {
  if constexpr (L % 4 == 0) {
#pragma unroll
    for (int j = 0; j < L / 4; ++j) {
      int4 t = __ldg(reinterpret_cast<const int4*>(p) + j);
      d[4 * j] = (uint32_t)t.x, d[4 * j + 1] = (uint32_t)t.y, d[4 * j + 2] = (uint32_t)t.z, d[4 * j + 3] = (uint32_t)t.w;
    }
  } else if constexpr (L % 2 == 0) {
#pragma unroll
    for (int j = 0; j < L / 2; ++j) {
      int2 t = __ldg(reinterpret_cast<const int2*>(p) + j);
      d[2 * j] = (uint32_t)t.x, d[2 * j + 1] = (uint32_t)t.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < L; ++j) d[j] = (uint32_t)__ldg(p + j);
  }
}


// i-th row of a launch: row0 + i of the range [row0, row1), or rows[i], i < row1 - row0, when the
// launch works on a row list (rows of one size class picked from anywhere in the matrix)
__device__ __forceinline__ int64_t pick_row(const uint32_t* __restrict__ rows, int64_t row0, int64_t row1, int64_t i,
                                            bool& live) {
  live = row0 + i < row1;
  if (!rows) return row0 + i;
  return live ? (int64_t)rows[i] : 0;
}

// k_rowsort: bitonic sort of all candidates of a row by a sub-warp (pgb_rowsort.cu)
// All launchers work on the row range [row0, row1) (row0 a multiple of 128); n_rows = size of the
// matrix (decides the key width).
// rows (nullable): row list of row1 - row0 entries (then row0 must be 0).
int launch_rowsort(uint32_t max_cand, int L, const uint32_t* rows, int64_t row0, int64_t row1, int64_t n_rows,
                   const uint32_t* row_start, const uint32_t* inc, const int32_t* dofs, uint8_t* slots, int32_t* ucol,
                   uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s);
// k_rowthread: one thread per short row, register sorting network (pgb_rowthread.cu)
bool rowthread_applies(uint32_t max_inc, uint32_t max_cand);
constexpr int ROWTHREAD_GROUP = 128;  // rows whose distinct columns k_rowthread stores back to back
int launch_rowthread(uint32_t max_inc, int L, int64_t row0, int64_t row1, int64_t n_rows, const uint32_t* row_start,
                     const uint32_t* inc_tmp, const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol,
                     uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s);
// k_rowhash: shared-memory hash dedupe by a group of lanes per row (pgb_rowhash.cu)
bool rowhash_applies(uint32_t max_inc, uint32_t max_cand);
int launch_rowhash(uint32_t max_inc, uint32_t max_cand, int L, const uint32_t* rows, int64_t row0, int64_t row1,
                   const uint32_t* row_start, const uint32_t* inc_tmp, const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol,
                   uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s);

int launch_rowhash_list(uint32_t max_cand, int L, const uint32_t* list, const uint32_t* row_start, const uint32_t* inc_tmp,
                        const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz,
                        uint32_t* max_nnz, cudaStream_t s);

// k_rowins: one thread per row of 17..32 incident cells, private hash table (pgb_rowins.cu); rows with too many
// distinct columns are appended to ovf (ovf[0] = count, ovf[1..] = rows) for k_rowhash_list.  Row ranges
// (rows == nullptr): the distinct columns of every block of ROWTHREAD_GROUP rows are stored back to back from
// ucol[row_start[first] * L] (the handed-over rows excluded: ovf_flag[row] = 1, their columns are in the row's own
// segment): k_compact_ins.  Row lists: one segment per row (k_compact).
bool rowins_applies(uint32_t max_inc, int L);
int launch_rowins(uint32_t max_inc, int L, const uint32_t* rows, int64_t row0, int64_t row1, const uint32_t* row_start,
                  const uint32_t* inc_tmp, const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol,
                  uint32_t* row_nnz, uint32_t* max_nnz, uint32_t* ovf, uint8_t* ovf_flag, cudaStream_t s);

// k_rowlong: one block per row, up to LONG_MAX_INC incident cells (pgb_rowlong.cu); inc_tmp: the row
// blocks in any order.  *overflow_flag is raised when a row has more than 255 distinct columns.
constexpr int LONG_MAX_INC = 2048;
int launch_rowlong(int L, const uint32_t* rows, int64_t row0, int64_t row1, const uint32_t* row_start, const uint32_t* inc_tmp,
                   const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz,
                   uint32_t* max_nnz, uint32_t* overflow_flag, cudaStream_t s);

}  // namespace pgb_rows
