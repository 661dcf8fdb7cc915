// K2 symbolic, very long rows (more than 512 candidates or more than 256 incident cells: extreme
// node valences of unstructured meshes, Lagrange2 node rows): one BLOCK per row (k_rowlong).
// Few rows are this long, so the kernel is written for generality, not speed: the incidence list is
// sorted with a shared-memory bitonic network, the candidates are streamed twice through an
// open-addressing table (insert, then look up the rank of the column).  Output identical to the
// other row kernels; a row with more than 255 distinct columns raises the flag that sends the whole
// call to the full-sort fallback (uint8 slots).
#include "pgb_rows.cuh"

namespace {

using namespace pgb_rows;

constexpr int LONG_HS = 1024;  // table entries (at most 255 are ever kept)
constexpr uint32_t LONG_EMPTY = 0xffffffffu;

__device__ __forceinline__ uint32_t long_hash(uint32_t c) { return (c * 2654435761u) >> 22; }

// ascending bitonic sort of n (power of two) keys in shared memory by a 256-thread block
__device__ __forceinline__ void block_bitonic(uint32_t* a, uint32_t n) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n; i += 256) {
        uint32_t p = i ^ j;
        if (p > i) {
          uint32_t x = a[i], y = a[p];
          bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[p] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <int L>
__global__ void __launch_bounds__(256) k_rowlong(const uint32_t* __restrict__ rows, int64_t row0, const uint32_t* __restrict__ row_start,
                                                 const uint32_t* __restrict__ inc_tmp,
                                                 const int32_t* __restrict__ dofs, uint32_t* __restrict__ inc_pos,
                                                 uint8_t* __restrict__ slots, int32_t* __restrict__ ucol_tmp,
                                                 uint32_t* __restrict__ row_nnz, uint32_t* __restrict__ max_nnz,
                                                 uint32_t* __restrict__ overflow_flag) {
  constexpr int LP = (L + 3) & ~3;
  __shared__ uint32_t s_inc[LONG_MAX_INC];
  __shared__ uint32_t s_key[LONG_HS];
  __shared__ uint8_t s_slot[LONG_HS];
  __shared__ uint32_t s_u[256];
  __shared__ uint32_t s_nu;
  const int tid = threadIdx.x;
  const int64_t r = rows ? (int64_t)rows[blockIdx.x] : row0 + blockIdx.x;
  const uint32_t rs = row_start[r], cnt = row_start[r + 1] - rs;
  uint32_t n2 = 32;
  while (n2 < cnt) n2 <<= 1;
  for (uint32_t i = tid; i < n2; i += 256) s_inc[i] = i < cnt ? inc_tmp[rs + i] : LONG_EMPTY;
  for (int i = tid; i < LONG_HS; i += 256) s_key[i] = LONG_EMPTY;
  s_u[tid] = LONG_EMPTY;
  if (tid == 0) s_nu = 0;
  __syncthreads();
  block_bitonic(s_inc, n2);  // canonical order: ascending incidence
  for (uint32_t i = tid; i < cnt; i += 256) inc_pos[s_inc[i]] = rs + i;
  // pass 1: the distinct columns
  const uint32_t ncand = cnt * L;
  for (uint32_t e = tid; e < ncand; e += 256) {
    const uint32_t q = e / L, b = e - q * L;
    const uint32_t col = (uint32_t)__ldg(dofs + (int64_t)(s_inc[q] / L) * L + b);
    uint32_t h = long_hash(col);
    while (true) {
      uint32_t old = atomicCAS(&s_key[h], LONG_EMPTY, col);
      if (old == col) break;
      if (old == LONG_EMPTY) {
        uint32_t k = atomicAdd(&s_nu, 1u);
        if (k < 256) s_u[k] = col;
        break;
      }
      h = (h + 1) & (LONG_HS - 1);
      if (*(volatile uint32_t*)&s_nu > 255u) break;  // hopeless: do not fill the table up
    }
  }
  __syncthreads();
  const uint32_t nu = s_nu;
  if (nu > 255u) {  // uint8 slots cannot address this row: the caller falls back to the full sort
    if (tid == 0) {
      atomicMax(overflow_flag, nu);
      row_nnz[r] = 0;
    }
    return;
  }
  block_bitonic(s_u, 256);  // padding (LONG_EMPTY) sorts to the end
  if ((uint32_t)tid < nu) {
    const uint32_t col = s_u[tid];
    uint32_t h = long_hash(col);
    while (s_key[h] != col) h = (h + 1) & (LONG_HS - 1);
    s_slot[h] = (uint8_t)tid;
  }
  __syncthreads();
  // pass 2: the slot of every candidate
  for (uint32_t e = tid; e < ncand; e += 256) {
    const uint32_t q = e / L, b = e - q * L;
    const uint32_t col = (uint32_t)__ldg(dofs + (int64_t)(s_inc[q] / L) * L + b);
    uint32_t h = long_hash(col);
    while (s_key[h] != col) h = (h + 1) & (LONG_HS - 1);
    slots[(uint64_t)(rs + q) * LP + b] = s_slot[h];
  }
  const uint64_t base = (uint64_t)rs * L;
  if ((uint32_t)tid < nu) ucol_tmp[base + tid] = (int32_t)s_u[tid];
  if (tid == 0) {
    row_nnz[r] = nu;
    atomicMax(max_nnz, nu);
  }
}

}  // namespace

namespace pgb_rows {

int launch_rowlong(int L, const uint32_t* rows, int64_t row0, int64_t row1, const uint32_t* row_start, const uint32_t* inc_tmp,
                   const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz,
                   uint32_t* max_nnz, uint32_t* overflow_flag, cudaStream_t s) {
  char name[64];
  snprintf(name, sizeof(name), "k_rowlong<%d>", L);
  note_row_kernel(name);
  if (row1 <= row0) return 0;
  unsigned grid = (unsigned)(row1 - row0);
#define PGB_RLG(LL) k_rowlong<LL><<<grid, 256, 0, s>>>(rows, row0, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz, overflow_flag)
  switch (L) {
    case 3: PGB_RLG(3); break;
    case 4: PGB_RLG(4); break;
    case 5: PGB_RLG(5); break;
    case 6: PGB_RLG(6); break;
    case 10: PGB_RLG(10); break;
    case 12: PGB_RLG(12); break;
    default: pgb_set_error("k_rowlong: unsupported local size %d", L); return 2;
  }
#undef PGB_RLG
  PGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace pgb_rows
