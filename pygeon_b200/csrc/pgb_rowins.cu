// K2 symbolic, rows of 17..32 incident cells with few distinct columns (P1 node rows in 3D: 24 incident
// tets, 96 candidates, 15 distinct columns): one THREAD per row with a private hash table (k_rowins).
#include "pgb_rows.cuh"

#include <cstdlib>

namespace {

using namespace pgb_rows;

// k_rowhash spends one WARP on such a row: ~290 warp instructions, most lanes idle in the per-row steps
// (table reset, compaction, sort of the 15 distinct columns).  Here a thread does the whole row:
//   (1) its incidences go through a 32-key sorting network in registers (canonical ascending order) and are
//       parked in shared memory, one column of a [24 or 32][128] array per thread;
//   (2) every incident cell's dof row (one vector load, the next one prefetched) is inserted into the thread's
//       own open-addressing table of 32 keys (shared memory, [h][thread]: bank = thread, so no conflicts
//       whatever slot a lane probes, and no atomics); the table position of every candidate replaces the
//       incidence in shared memory (4 x 1 byte);
//   (3) the table goes through the same register network: the distinct columns come out sorted (EMPTY = max
//       key sorts last); their ranks are written back into the table;
//   (4) a candidate's slot is the rank stored at its table position.
// ~165 warp instructions per row on config #2 (345 M in all against 610 M for k_rowhash).  What bounds the
// kernel is the L1 sector rate: a per-thread access costs one 32-byte sector per lane and instruction, so in a
// row RANGE everything that is contiguous per block - the incidences, the slot words, the distinct columns -
// is moved with coalesced loads / stores and redistributed through shared memory; the two accesses that
// stay scattered (the cell's dof row, the inc_pos store) are inherent to the transpose.
// Same outputs as k_rowhash (tested bit for bit).  A row with more than CAP distinct columns is handed to
// k_rowhash_list through the overflow list (ovf[0] = count, ovf[1..] = rows).
constexpr uint32_t INS_EMPTY = 0xffffffffu;
constexpr int INS_B = 128;   // rows per block
static_assert(INS_B == ROWTHREAD_GROUP, "k_compact_ins expects groups of ROWTHREAD_GROUP rows");
constexpr int INS_MI = 32;   // incidences per row
constexpr int INS_HS = 32;   // table entries
constexpr int INS_CAP = 24;  // distinct columns (3/4 of the table)

__device__ __forceinline__ uint32_t ins_hash(uint32_t c) { return (c * 2654435761u) >> 27; }

// RANGE: the block's rows are consecutive, so its incidences and slot words are one contiguous run each: they
// are moved with coalesced loads / stores and redistributed through shared memory (a thread-per-row access
// touches one 32-byte sector per lane and instruction; the kernel is bound by the L1 sector rate).
template <int L, bool RANGE, int MI>
__global__ void __launch_bounds__(INS_B, 7)
k_rowins(const uint32_t* __restrict__ rows, int64_t row0, int64_t n_rows, const uint32_t* __restrict__ row_start,
         const uint32_t* __restrict__ inc_tmp, const int32_t* __restrict__ dofs, uint32_t* __restrict__ inc_pos,
         uint8_t* __restrict__ slots, int32_t* __restrict__ ucol_tmp, uint32_t* __restrict__ row_nnz,
         uint32_t* __restrict__ max_nnz, uint32_t* __restrict__ ovf, uint8_t* __restrict__ ovf_flag) {
  static_assert(L <= 4, "one slot word per incidence");
  constexpr int B = INS_B, HS = INS_HS, NW = 32;  // MI: incidences kept per row (24 or 32); NW: width of the network
  // the runs in global order are stored with one pad word per 32 (sw): thread t reads / writes around word 24 t,
  // which without the padding falls on 4 banks only
  __shared__ __align__(16) uint32_t s_inc[MI * B + MI * B / 32];  // incidences, then the table positions of their candidates
  __shared__ __align__(16) uint32_t s_key[HS * B + HS * B / 32];  // table keys, then the ranks, then the slot words in global order
  auto sw = [](uint32_t j) { return j + (j >> 5); };
  __shared__ uint32_t s_m[B / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t* const my_inc = s_inc + tid;
  uint32_t* const my_key = s_key + tid;
  bool live;
  const int64_t r = pick_row(rows, row0, n_rows, (int64_t)blockIdx.x * B + tid, live);
  uint32_t rs = 0, cnt = 0;
  if (live) {
    rs = row_start[r];
    cnt = row_start[r + 1] - rs;
  }
  uint32_t rs0 = 0, n_blk = 0;  // RANGE: the block's run of incidences
  {
    uint32_t inc[NW];
#pragma unroll
    for (int q = MI; q < NW; ++q) inc[q] = INS_EMPTY;
    if constexpr (RANGE) {
      const int64_t rb = row0 + (int64_t)blockIdx.x * B;
      const int64_t re = rb + B < n_rows ? rb + B : n_rows;
      rs0 = row_start[rb];
      n_blk = row_start[re] - rs0;
      for (uint32_t j = tid; j < n_blk; j += B) s_inc[sw(j)] = inc_tmp[rs0 + j];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < MI; ++q) inc[q] = ((uint32_t)q < cnt) ? s_inc[sw(rs - rs0 + q)] : INS_EMPTY;
      __syncthreads();
    } else {
#pragma unroll
      for (int q = 0; q < MI; ++q) inc[q] = ((uint32_t)q < cnt) ? inc_tmp[rs + q] : INS_EMPTY;
    }
    reg_bitonic<NW, uint32_t>(inc);
#pragma unroll
    for (int q = 0; q < MI; ++q) my_inc[q * B] = inc[q];
  }
#pragma unroll
  for (int h = 0; h < HS; ++h) my_key[h * B] = INS_EMPTY;
  uint32_t nu = 0;
  bool over = false;
  const uint32_t hdiag = ins_hash((uint32_t)r);
  if (cnt) {  // the row's own dof: every incident cell holds it
    my_key[hdiag * B] = (uint32_t)r;
    nu = 1;
  }
  uint32_t in_next = cnt ? my_inc[0] : 0u;
  uint32_t d_next[L];
  if (cnt) load_dofs<L>(dofs + (int64_t)(in_next / L) * L, d_next);
  for (uint32_t q = 0; q < cnt; ++q) {
    const uint32_t in = in_next;
    uint32_t d[L];
#pragma unroll
    for (int b = 0; b < L; ++b) d[b] = d_next[b];
    if (q + 1 < cnt) {  // the next cell's dofs are in flight while this one's are inserted
      in_next = my_inc[(q + 1) * B];
      load_dofs<L>(dofs + (int64_t)(in_next / L) * L, d_next);
    }
    inc_pos[in] = rs + q;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < L; ++b) {
      const uint32_t col = d[b];
      uint32_t h = hdiag;
      if (col != (uint32_t)r) {
        h = ins_hash(col);
        while (true) {
          const uint32_t k = my_key[h * B];
          if (k == col) break;
          if (k == INS_EMPTY) {
            if (nu >= (uint32_t)INS_CAP) over = true;
            else my_key[h * B] = col, ++nu;
            break;
          }
          h = (h + 1) & (HS - 1);
        }
      }
      word |= h << (8 * b);
    }
    my_inc[q * B] = word;
  }
  // sorted distinct columns; their ranks replace the keys
  uint32_t k[HS];
  auto own_segment = [&]() {  // the row's distinct columns at the head of its own scratch segment
    const uint32_t base = rs * L;
    if constexpr (L == 4) {  // 16-byte stores; the tail of the last group stays inside the segment
#pragma unroll
      for (int g = 0; g < INS_CAP / 4; ++g)
        if ((uint32_t)(4 * g) < nu && !over)
          reinterpret_cast<uint4*>(ucol_tmp + base)[g] = make_uint4(k[4 * g], k[4 * g + 1], k[4 * g + 2], k[4 * g + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < INS_CAP; ++e)
        if ((uint32_t)e < nu && !over) ucol_tmp[base + e] = (int32_t)k[e];
    }
  };
  {
#pragma unroll
    for (int h = 0; h < HS; ++h) k[h] = my_key[h * B];
    reg_bitonic<HS, uint32_t>(k);
    if constexpr (!RANGE) own_segment();  // row list: one scratch segment per row (k_compact)
    uint32_t pos[(INS_CAP + 3) / 4];
#pragma unroll
    for (int e = 0; e < (INS_CAP + 3) / 4; ++e) pos[e] = 0;
#pragma unroll
    for (int e = 0; e < INS_CAP; ++e) {
      if ((uint32_t)e < nu && !over) {
        uint32_t h = ins_hash(k[e]);
        while (my_key[h * B] != k[e]) h = (h + 1) & (HS - 1);
        pos[e / 4] |= h << (8 * (e & 3));
      }
    }
#pragma unroll
    for (int e = 0; e < INS_CAP; ++e)
      if ((uint32_t)e < nu && !over) my_key[((pos[e / 4] >> (8 * (e & 3))) & (HS - 1)) * B] = (uint32_t)e;
  }
  uint32_t* const out = reinterpret_cast<uint32_t*>(slots);  // LP = 4: one word per incidence
  if (!over) {
    for (uint32_t q = 0; q < cnt; ++q) {
      const uint32_t word = my_inc[q * B];
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < L; ++b) o |= my_key[((word >> (8 * b)) & (HS - 1)) * B] << (8 * b);
      if constexpr (RANGE) my_inc[q * B] = o;
      else out[rs + q] = o;
    }
    if (live) row_nnz[r] = nu;
  } else {
    ovf[1 + atomicAdd(ovf, 1u)] = (uint32_t)r;  // integer counter: the list's order does not reach any output
  }
  const uint32_t mxw = __reduce_max_sync(FULL, over ? 0u : nu);
  if (lane == 0) s_m[w] = mxw;
  // s_m; RANGE: nobody reads ranks from s_key any more.  A block with a handed-over row keeps one scratch segment
  // per row (the compact run could run into the segment k_rowhash_list fills): k_compact_ins reads the flag.
  const int any_over = __syncthreads_or(over ? 1 : 0);
  if constexpr (RANGE)
    if (live) ovf_flag[r] = any_over ? 1 : 0;
  if (tid == 0) {
    uint32_t mx = s_m[0];
#pragma unroll
    for (int q = 1; q < B / 32; ++q) mx = s_m[q] > mx ? s_m[q] : mx;
    atomicMax(max_nnz, mx);
  }
  if constexpr (RANGE) {  // slot words: [q][thread] -> global order -> one coalesced run (rows handed to k_rowhash are rewritten there)
    for (uint32_t q = 0; q < cnt; ++q) s_key[sw(rs - rs0 + q)] = my_inc[q * B];
    // distinct columns of the block's rows back to back: block scan of the row counts
    const uint32_t incl = warp_incl_scan(nu);
    __syncthreads();  // s_key staged, s_inc and s_m free
    if (lane == 31) s_m[w] = incl;
    __syncthreads();
    uint32_t off = incl - nu, btotal = 0;
#pragma unroll
    for (int i = 0; i < B / 32; ++i) {
      if (i < w) off += s_m[i];
      btotal += s_m[i];
    }
    if (any_over) {
      own_segment();
    } else {
#pragma unroll
      for (int e = 0; e < INS_CAP; ++e)
        if ((uint32_t)e < nu) s_inc[off + e] = k[e];
    }
    for (uint32_t j = tid; j < n_blk; j += B) out[rs0 + j] = s_key[sw(j)];
    __syncthreads();
    if (!any_over)
      for (uint32_t j = tid; j < btotal; j += B) ucol_tmp[(uint64_t)rs0 * L + j] = (int32_t)s_inc[j];
  }
}

}  // namespace

namespace pgb_rows {

bool rowins_applies(uint32_t max_inc, int L) {
  static const bool off = getenv("PGB_NO_ROWINS") != nullptr;  // A/B switch: fall back to k_rowhash
  return !off && !g_force_bitonic && !g_force_key64 && max_inc > 16 && max_inc <= (uint32_t)INS_MI && (L == 3 || L == 4);
}

int launch_rowins(uint32_t max_inc, int L, const uint32_t* rows, int64_t row0, int64_t row1, const uint32_t* row_start,
                  const uint32_t* inc_tmp, const int32_t* dofs, uint32_t* inc_pos, uint8_t* slots, int32_t* ucol,
                  uint32_t* row_nnz, uint32_t* max_nnz, uint32_t* ovf, uint8_t* ovf_flag, cudaStream_t s) {
  char name[64];
  snprintf(name, sizeof(name), "k_rowins<%d,%d>", L, max_inc <= 24 ? 24 : 32);
  note_row_kernel(name);
  const unsigned grid = (unsigned)pgb_cdiv(row1 - row0, INS_B);
#define PGB_RI(LL, RG, M) k_rowins<LL, RG, M><<<grid, INS_B, 0, s>>>(rows, row0, row1, row_start, inc_tmp, dofs, inc_pos, slots, ucol, row_nnz, max_nnz, ovf, ovf_flag)
#define PGB_RI2(LL, M)         \
  do {                         \
    if (rows) PGB_RI(LL, false, M); \
    else PGB_RI(LL, true, M);  \
  } while (0)
  // 24 incidences (P1 on Kuhn tets) leave room for a seventh block per SM
  if (L == 3) {
    if (max_inc <= 24) PGB_RI2(3, 24); else PGB_RI2(3, 32);
  } else {
    if (max_inc <= 24) PGB_RI2(4, 24); else PGB_RI2(4, 32);
  }
#undef PGB_RI2
#undef PGB_RI
  PGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace pgb_rows
