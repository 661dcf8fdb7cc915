// K2 symbolic, rows with many candidates but few distinct columns: hash dedupe (k_rowhash).
#include "pgb_rows.cuh"

namespace {

using namespace pgb_rows;

// ---- per-row dedupe by hashing (rows with 33..256 candidates) -----------------------------------
// The bitonic network above sorts ALL candidates of a row although only the distinct columns need
// ordering (a P1 node row of the Kuhn mesh has 96 candidates but ~15 distinct columns).  Here one
// warp per row
//   (1) loads the row's incidences (in the arbitrary order k_inc_place left them), sorts them
//       ascending (= the canonical (dof, cell) order) and records inc_pos: the work of k_row_canon;
//   (2) every lane loads the dof row of its incident cell with vector loads and inserts the L
//       columns into an open-addressing table in shared memory (atomicCAS on integers: the
//       resulting SET is order independent);
//   (3) compacts the distinct columns, sorts only those and hands every candidate the rank of its
//       column through the table; a lane writes the LP slot bytes of its incidence as whole words.
// Output identical to k_rowsort (sorted unique columns + uint8 slot per candidate), tested bit for
// bit.  (Merging equal columns with match.any before the insert was measured slower.)
constexpr uint32_t HASH_EMPTY = 0xffffffffu;

template <int HS>
__device__ __forceinline__ uint32_t hash_col(uint32_t c) {
  return (c * 2654435761u) >> (32 - (HS == 64 ? 6 : HS == 128 ? 7 : HS == 256 ? 8 : 9));
}

// sort the nu distinct columns of every group (ucols[0..nu), shared memory; NPU per lane of the
// group, G*NPU >= the largest nu of the warp), write them back sorted and store every column's rank
// in tslot at the column's table position.  Called by all lanes of the warp.
template <int NPU, int HS, int G>
__device__ __forceinline__ void sort_uniques(uint32_t* ucols, int nu, const uint32_t* tkey, uint8_t* tslot, int gl) {
  uint32_t k[NPU];
#pragma unroll
  for (int i = 0; i < NPU; ++i) {
    int e = gl * NPU + i;
    k[i] = (e < nu) ? ucols[e] : HASH_EMPTY;
  }
  __syncwarp();
  warp_bitonic_u32<NPU, G>(k, gl);
#pragma unroll
  for (int i = 0; i < NPU; ++i) {
    int e = gl * NPU + i;
    if (e < nu) {
      ucols[e] = k[i];
      uint32_t h = hash_col<HS>(k[i]);
      while (tkey[h] != k[i]) h = (h + 1) & (HS - 1);
      tslot[h] = (uint8_t)e;
    }
  }
  __syncwarp();
}

// G lanes per row (256/G rows per block), NPL incidences per lane (row capacity G*NPL incidences),
// table of HS entries per row (at most 3/4 full)
template <int G, int NPL, int L, int HS>
__device__ __forceinline__ void rowhash_block(int64_t block_index, const uint32_t* __restrict__ rows, int64_t row0, int64_t n_rows,
                                              const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ inc_tmp,
                                              const int32_t* __restrict__ dofs, uint32_t* __restrict__ inc_pos,
                                              uint8_t* __restrict__ slots, int32_t* __restrict__ ucol_tmp,
                                              uint32_t* __restrict__ row_nnz, uint32_t* __restrict__ max_nnz) {
  constexpr int LP = (L + 3) & ~3;
  constexpr int RPB = 256 / G;  // rows per block
  __shared__ uint32_t s_key[RPB][HS];
  __shared__ uint32_t s_ucol[RPB][HS];
  __shared__ uint8_t s_slot[RPB][HS];
  __shared__ uint32_t wmax[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
  const int g = threadIdx.x / G;
  bool live;
  const int64_t r = pick_row(rows, row0, n_rows, block_index * RPB + g, live);
  uint32_t* tkey = s_key[g];
  uint8_t* tslot = s_slot[g];
  uint32_t* ucols = s_ucol[g];
  uint32_t rs = 0, cnt = 0;
  if (live) {
    rs = row_start[r];
    cnt = row_start[r + 1] - rs;
  }
  uint32_t inc[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    uint32_t e = (uint32_t)(i * G + gl);
    inc[i] = (e < cnt) ? inc_tmp[rs + e] : 0xffffffffu;
  }
#pragma unroll
  for (int k = 0; k < HS / G; ++k) tkey[k * G + gl] = HASH_EMPTY;
  warp_bitonic_u32<NPL, G>(inc, gl);  // canonical order: ascending incidence = ascending (cell, a)
  // Every incident cell contributes the row's own dof once (the worst same-address conflict): it is
  // inserted by the first lane alone.
  const uint32_t hdiag = hash_col<HS>((uint32_t)r);
  __syncwarp();
  if (gl == 0 && cnt) tkey[hdiag] = (uint32_t)r;
  __syncwarp();
  uint16_t hpos[NPL][L];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t e = (uint32_t)(gl * NPL + i);
    if (e < cnt) {
      inc_pos[inc[i]] = rs + e;
      uint32_t d[L];
      load_dofs<L>(dofs + (int64_t)(inc[i] / L) * L, d);
#pragma unroll
      for (int b = 0; b < L; ++b) {
        const uint32_t col = d[b];
        uint32_t h = hdiag;
        if (col != (uint32_t)r) {
          h = hash_col<HS>(col);
          while (true) {
            uint32_t old = atomicCAS(&tkey[h], HASH_EMPTY, col);
            if (old == HASH_EMPTY || old == col) break;
            h = (h + 1) & (HS - 1);
          }
        }
        hpos[i][b] = (uint16_t)h;
      }
    }
  }
  __syncwarp();
  // compact the distinct columns of the group's table
  int nu = 0;
  const uint32_t gmask = (G == 32) ? FULL : ((1u << (G & 31)) - 1u);
#pragma unroll
  for (int k = 0; k < HS / G; ++k) {
    uint32_t v = tkey[k * G + gl];
    unsigned m = (__ballot_sync(FULL, v != HASH_EMPTY) >> gbase) & gmask;
    if (v != HASH_EMPTY) ucols[nu + __popc(m & ((1u << gl) - 1u))] = v;
    nu += __popc(m);
  }
  __syncwarp();
  // sort them; the network size is chosen for the longest row of the warp, so that the warp stays converged
  const int nu_w = (G == 32) ? nu : (int)__reduce_max_sync(FULL, (unsigned)nu);
  if (G >= 16 && nu_w <= 16) sort_uniques<1, HS, 16>(ucols, nu, tkey, tslot, lane & 15);
  else if (nu_w <= G) sort_uniques<1, HS, G>(ucols, nu, tkey, tslot, gl);
  else if (nu_w <= 2 * G) sort_uniques<2, HS, G>(ucols, nu, tkey, tslot, gl);
  else if (nu_w <= 4 * G) sort_uniques<4, HS, G>(ucols, nu, tkey, tslot, gl);
  else if (HS * 3 / 4 <= 8 * G || nu_w <= 8 * G) sort_uniques<8, HS, G>(ucols, nu, tkey, tslot, gl);
  else if (HS * 3 / 4 <= 16 * G || nu_w <= 16 * G) sort_uniques<16, HS, G>(ucols, nu, tkey, tslot, gl);
  else sort_uniques<32, HS, G>(ucols, nu, tkey, tslot, gl);
  // slots of the candidates: the LP bytes of one incidence as whole words
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t e = (uint32_t)(gl * NPL + i);
    if (e < cnt) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(slots + (uint64_t)(rs + e) * LP);
#pragma unroll
      for (int j = 0; j < LP / 4; ++j) {
        uint32_t wv = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (4 * j + t < L) wv |= (uint32_t)tslot[hpos[i][4 * j + t]] << (8 * t);
        dst[j] = wv;
      }
    }
  }
  const uint32_t base = rs * L;
  for (int e = gl; e < nu; e += G) ucol_tmp[base + e] = (int32_t)ucols[e];
  if (live && gl == 0) row_nnz[r] = (uint32_t)nu;
  const uint32_t mxw = __reduce_max_sync(FULL, (unsigned)nu);
  if (lane == 0) wmax[w] = mxw;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t mx = wmax[0];
    for (int q = 1; q < 8; ++q) mx = wmax[q] > mx ? wmax[q] : mx;
    atomicMax(max_nnz, mx);
  }
}

#define PGB_RH_BOUNDS(G, NPL, L, HS) \
  __launch_bounds__(256, (NPL == 1 && L <= 6) ? ((256 / G) * HS * 9 <= 20000 ? 8 : 5) : (NPL == 1 ? 5 : 1))

template <int G, int NPL, int L, int HS>
__global__ void PGB_RH_BOUNDS(G, NPL, L, HS)
k_rowhash(const uint32_t* __restrict__ rows, int64_t row0, int64_t n_rows, const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ inc_tmp,
          const int32_t* __restrict__ dofs, uint32_t* __restrict__ inc_pos, uint8_t* __restrict__ slots,
          int32_t* __restrict__ ucol_tmp, uint32_t* __restrict__ row_nnz, uint32_t* __restrict__ max_nnz) {
  rowhash_block<G, NPL, L, HS>((int64_t)blockIdx.x, rows, row0, n_rows, row_start, inc_tmp, dofs, inc_pos, slots, ucol_tmp, row_nnz,
                               max_nnz);
}

// the rows of a list whose length is only known on the device (list[0] = count, list[1..] = rows: what k_rowins
// could not finish): a fixed grid loops over the list, so that no host read decides the launch
template <int G, int NPL, int L, int HS>
__global__ void PGB_RH_BOUNDS(G, NPL, L, HS)
k_rowhash_list(const uint32_t* __restrict__ list, const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ inc_tmp,
               const int32_t* __restrict__ dofs, uint32_t* __restrict__ inc_pos, uint8_t* __restrict__ slots,
               int32_t* __restrict__ ucol_tmp, uint32_t* __restrict__ row_nnz, uint32_t* __restrict__ max_nnz) {
  const int64_t n = (int64_t)list[0];
  for (int64_t b = blockIdx.x; b * (256 / G) < n; b += gridDim.x) {
    rowhash_block<G, NPL, L, HS>(b, list + 1, 0, n, row_start, inc_tmp, dofs, inc_pos, slots, ucol_tmp, row_nnz, max_nnz);
    __syncthreads();  // the tables and wmax are reused by the next round
  }
}

template <int G, int NPL, int HS>
int launch_rowhash_l(int L, const uint32_t* rows, int64_t row0, int64_t n_rows, const uint32_t* row_start, const uint32_t* inc_tmp, const int32_t* dofs,
                     uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz, uint32_t* max_nnz,
                     cudaStream_t s) {
  char name[64];
  snprintf(name, sizeof(name), "k_rowhash<%d,%d,%d,%d>", G, NPL, L, HS);
  note_row_kernel(name);
  unsigned grid = (unsigned)pgb_cdiv(n_rows - row0, 256 / G);
#define PGB_RH(LL) k_rowhash<G, NPL, LL, HS><<<grid, 256, 0, s>>>(rows, row0, n_rows, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz)
  switch (L) {
    case 3: PGB_RH(3); break;
    case 4: PGB_RH(4); break;
    case 5: PGB_RH(5); break;
    case 6: PGB_RH(6); break;
    case 10: PGB_RH(10); break;
    case 12: PGB_RH(12); break;
    default: pgb_set_error("k_rowhash: unsupported local size %d", L); return 2;
  }
#undef PGB_RH
  PGB_LAUNCH_CHECK();
  return 0;
}


}  // namespace

namespace pgb_rows {

// fused canonical sort of the incidences + hash dedupe; inc_tmp: the row blocks in any order.
// Lanes per row from the largest incidence count, table size from the largest candidate count.
int launch_rowhash(uint32_t max_inc, uint32_t max_cand, int L, const uint32_t* rows, int64_t row0, int64_t n_rows,
                   const uint32_t* row_start,
                   const uint32_t* inc_tmp, const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol,
                   uint32_t* row_nnz, uint32_t* max_nnz, cudaStream_t s) {
#define PGB_RHL(G, N, H) return launch_rowhash_l<G, N, H>(L, rows, row0, n_rows, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz, s)
  if (max_inc <= 8) PGB_RHL(8, 1, 128);  // 65..96 candidates (fewer go to k_rowthread)
  if (max_inc <= 16) {
    if (max_cand <= 96) PGB_RHL(16, 1, 128);
    PGB_RHL(16, 1, 256);  // max_cand <= 192
  }
  if (max_inc <= 32) {
    if (max_cand <= 96) PGB_RHL(32, 1, 128);
    PGB_RHL(32, 1, 512);
  }
  PGB_RHL(32, 4, 512);  // more than 32 incidences: more than 96 candidates
#undef PGB_RHL
}

// rows of a device-side list (see k_rowhash_list); only the shapes k_rowins hands over: at most 32 incidences, L <= 4
int launch_rowhash_list(uint32_t max_cand, int L, const uint32_t* list, const uint32_t* row_start, const uint32_t* inc_tmp,
                        const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol, uint32_t* row_nnz,
                        uint32_t* max_nnz, cudaStream_t s) {
  const unsigned grid = (unsigned)(PGB_SMS * 4);
#define PGB_RHX(LL, H) k_rowhash_list<32, 1, LL, H><<<grid, 256, 0, s>>>(list, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz)
  if (L == 3) {
    if (max_cand <= 96) PGB_RHX(3, 128); else PGB_RHX(3, 512);
  } else if (L == 4) {
    if (max_cand <= 96) PGB_RHX(4, 128); else PGB_RHX(4, 512);
  } else {
    pgb_set_error("k_rowhash_list: unsupported local size %d", L);
    return 2;
  }
#undef PGB_RHX
  PGB_LAUNCH_CHECK();
  return 0;
}

// rows with many candidates but few distinct columns: dedupe by hashing, sort only the distinct ones
bool rowhash_applies(uint32_t max_inc, uint32_t max_cand) {
  return !g_force_bitonic && !g_force_key64 && max_cand > 32 && max_cand <= 256 && max_inc <= 128;
}


}  // namespace pgb_rows
