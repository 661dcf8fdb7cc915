// K2 symbolic, per-row bitonic sort of all candidates (k_rowsort): rows with more than 256 candidates, and the reference implementation the other row kernels are tested against.
#include "pgb_rows.cuh"

namespace {

using namespace pgb_rows;

// Key = (col << CSH) | e0 with e0 the candidate's position in the row before sorting
// (e0 = q*L + b for the q-th incidence of the row and local column b).  KeyT = uint32_t
// (CSH = log2 N) whenever n_rows * N fits in 32 bits, else uint64_t (CSH = 32).
// Output per row: the sorted unique columns (scratch, compacted later), and for every candidate
// the index of its column among them ("slot", uint8): slots[(row_start[r] + q)*L + b].
template <int G, int ITEMS, int L, typename KeyT>
__global__ void __launch_bounds__(256) k_rowsort(const uint32_t* __restrict__ rows, int64_t row0, int64_t n_rows, const uint32_t* __restrict__ row_start,
                                                 const uint32_t* __restrict__ inc_sorted,
                                                 const int32_t* __restrict__ dofs, uint8_t* __restrict__ slots,
                                                 int32_t* __restrict__ ucol_tmp, uint32_t* __restrict__ row_nnz,
                                                 uint32_t* __restrict__ max_nnz) {
  constexpr int N = G * ITEMS;
  constexpr int LP = (L + 3) & ~3;
  constexpr bool K32 = sizeof(KeyT) == 4;
  constexpr int EB = (N <= 8) ? 3 : (N <= 16) ? 4 : (N <= 32) ? 5 : (N <= 64) ? 6 : (N <= 128) ? 7 : (N <= 256) ? 8 : 9;
  constexpr int CSH = K32 ? EB : 32;  // column shift inside the key
  __shared__ uint32_t wmax[8];
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;  // lane within the group
  bool live;
  const int64_t r = pick_row(rows, row0, n_rows, ((int64_t)blockIdx.x * 256 + threadIdx.x) / G, live);
  uint32_t rs = 0, re = 0;
  if (live) {
    rs = row_start[r];
    re = row_start[r + 1];
  }
  const uint32_t ncand = (re - rs) * L;
  const uint32_t base = rs * L;

  // gather candidates, blocked layout: element e = gl*ITEMS + i
  KeyT key[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    uint32_t e = gl * ITEMS + i;
    key[i] = (KeyT) ~(KeyT)0;
    if (e < ncand) {
      uint32_t q = e / L, b = e - q * L;
      uint32_t cell = inc_sorted[rs + q] / L;
      uint32_t col = (uint32_t)__ldg(dofs + (int64_t)cell * L + b);
      key[i] = ((KeyT)col << CSH) | (KeyT)e;
    }
  }
  // bitonic sort, ascending, N elements spread over G lanes
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= ITEMS) {
        const int lj = j / ITEMS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          KeyT other = __shfl_xor_sync(FULL, key[i], lj);
          int e = gl * ITEMS + i;
          bool lower = (e & j) == 0;
          bool asc = (e & k) == 0;
          KeyT mn = key[i] < other ? key[i] : other;
          KeyT mx = key[i] < other ? other : key[i];
          key[i] = (lower == asc) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          if ((i & j) == 0) {
            int e = gl * ITEMS + i;
            bool asc = (e & k) == 0;
            KeyT a = key[i], b = key[i | j];
            bool sw = asc ? (a > b) : (a < b);
            key[i] = sw ? b : a;
            key[i | j] = sw ? a : b;
          }
        }
      }
    }
  }
  // unique columns: head(e) = e == 0 || col(e) != col(e-1)
  uint32_t prev_last = (uint32_t)(__shfl_up_sync(FULL, key[ITEMS - 1], 1) >> CSH);
  uint32_t heads = 0;  // bit i set: element i of this lane starts a run
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    uint32_t e = gl * ITEMS + i;
    uint32_t col = (uint32_t)(key[i] >> CSH);
    uint32_t pc = (i == 0) ? prev_last : (uint32_t)(key[i - 1] >> CSH);
    bool h = (e < ncand) && (e == 0 || col != pc);
    heads |= (h ? 1u : 0u) << i;
    cnt += h;
  }
  // exclusive scan of cnt over the G lanes of the group
  int inc_ = cnt;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    int t = __shfl_up_sync(FULL, inc_, d);
    if (gl >= d) inc_ += t;
  }
  int total = __shfl_sync(FULL, inc_, (lane - gl) + G - 1);
  int run = inc_ - cnt - 1;  // index of the run that is open when this lane starts
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    uint32_t e = gl * ITEMS + i;
    if (e < ncand) {
      if ((heads >> i) & 1u) {
        ++run;
        ucol_tmp[base + run] = (int32_t)(key[i] >> CSH);
      }
      uint32_t e0 = (uint32_t)key[i] & ((1u << EB) - 1u);  // position before sorting
      uint32_t q0 = e0 / L;
      slots[(uint64_t)(rs + q0) * LP + (e0 - q0 * L)] = (uint8_t)run;  // row pitch LP = L rounded up to a multiple of 4
    }
  }
  if (live && gl == 0) row_nnz[r] = (uint32_t)total;
  // largest row (integer max: order independent)
  uint32_t mx = __reduce_max_sync(FULL, (uint32_t)total);
  if (lane == 0) wmax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) mx = wmax[q] > mx ? wmax[q] : mx;
    atomicMax(max_nnz, mx);
  }
}

template <int G, int ITEMS, typename KeyT>
int launch_rowsort_l(int L, const uint32_t* rows, int64_t row0, int64_t n_rows, const uint32_t* row_start, const uint32_t* inc, const int32_t* dofs,
                     uint8_t* slots, int32_t* ucol, uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s) {
  char name[64];
  snprintf(name, sizeof(name), "k_rowsort<%d,%d,%d,%s>", G, ITEMS, L, sizeof(KeyT) == 4 ? "u32" : "u64");
  note_row_kernel(name);
  unsigned grid = pgb_grid((n_rows - row0) * G, 256);
#define PGB_RL(LL) k_rowsort<G, ITEMS, LL, KeyT><<<grid, 256, 0, s>>>(rows, row0, n_rows, row_start, inc, dofs, slots, ucol, row_nnz, max_nnz)
  switch (L) {
    case 3: PGB_RL(3); break;
    case 4: PGB_RL(4); break;
    case 5: PGB_RL(5); break;
    case 6: PGB_RL(6); break;
    case 10: PGB_RL(10); break;
    case 12: PGB_RL(12); break;
    default: pgb_set_error("k_rowsort: unsupported local size %d", L); return 2;
  }
#undef PGB_RL
  PGB_LAUNCH_CHECK();
  return 0;
}

template <int G, int ITEMS>
int launch_rowsort_k(int L, const uint32_t* rows, int64_t row0, int64_t row1, int64_t n_rows, const uint32_t* row_start, const uint32_t* inc, const int32_t* dofs,
                     uint8_t* slots, int32_t* ucol, uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s) {
  // 32-bit keys when (col, position-in-row) fits strictly below the padding value 0xffffffff
  if ((uint64_t)n_rows * (uint64_t)(G * ITEMS) <= 0xffffffffull && !g_force_key64)
    return launch_rowsort_l<G, ITEMS, uint32_t>(L, rows, row0, row1, row_start, inc, dofs, slots, ucol, row_nnz, max_nnz, s);
  return launch_rowsort_l<G, ITEMS, uint64_t>(L, rows, row0, row1, row_start, inc, dofs, slots, ucol, row_nnz, max_nnz, s);
}


}  // namespace

namespace pgb_rows {

bool g_force_key64 = false;
bool g_force_bitonic = false;
char g_row_kernel[256] = "";
void note_row_kernel(const char* name) {
  if (strstr(g_row_kernel, name)) return;
  size_t len = strlen(g_row_kernel);
  snprintf(g_row_kernel + len, sizeof(g_row_kernel) - len, "%s%s", len ? "+" : "", name);
}

int launch_rowsort(uint32_t max_cand, int L, const uint32_t* rows, int64_t row0, int64_t row1, int64_t n_rows,
                   const uint32_t* row_start, const uint32_t* inc,
                   const int32_t* dofs, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz, uint32_t* max_nnz,
                   cudaStream_t s) {
#define PGB_RS(G, I) return launch_rowsort_k<G, I>(L, rows, row0, row1, n_rows, row_start, inc, dofs, slots, ucol, row_nnz, max_nnz, s)
  if (max_cand <= 8) PGB_RS(4, 2);
  if (max_cand <= 16) PGB_RS(8, 2);
  if (max_cand <= 32) PGB_RS(8, 4);
  if (max_cand <= 64) PGB_RS(16, 4);
  if (max_cand <= 128) PGB_RS(32, 4);
  if (max_cand <= 256) PGB_RS(32, 8);
  PGB_RS(32, 16);
#undef PGB_RS
}


}  // namespace pgb_rows
