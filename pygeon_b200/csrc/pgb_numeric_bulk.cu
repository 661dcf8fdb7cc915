// K2 numeric, bulk-async variant (flags bit 1 of pgb_csr_numeric_rows).
//
// k_numeric_t (pgb_csr.cu) gives every CSR row to one thread which streams its block of row-sorted local
// rows with 256-bit loads.  DRAM traffic equals the algorithmic bytes, but the loads of a warp go to 32
// different 128-byte lines (the rows of neighbouring threads are ~24 local rows apart), so every 32-byte
// value load and every 4-byte slot load costs one L1 wavefront per lane: ~1e8 wavefronts on BASELINE
// config #2, which is what the kernel's 0.43 ms amount to.
//
// The local rows of the BLK consecutive CSR rows of a thread block are CONTIGUOUS in vals_t / slots
// (that is what the row-sorted layout is for), so here they are brought into shared memory by the
// bulk-async copy engine (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) in chunks of CH local
// rows through an NS-stage ring, issued by one thread and completed on an mbarrier: no LSU wavefronts,
// full-line DRAM requests, the next chunks in flight while the block works on the current one.  Each
// thread then walks the part of the chunk that belongs to its row, in ascending order, and adds into
// its column accumulators in shared memory exactly as k_numeric_t does: same summation order, bit-identical
// results.
#include "pgb_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int NS = 3;  // stages

// dynamic shared memory: [acc: caps * BLK doubles][NS x (CH * LP doubles)][NS x (CH * LP bytes)][NS mbarriers]
// (+ [out: caps * BLK doubles] when STAGED)
template <int L, int BLK, typename IP, bool STAGED>
__global__ void __launch_bounds__(BLK) k_numeric_bulk(int64_t n_rows, int caps, int ch,
                                                      const uint32_t* __restrict__ row_start,
                                                      const uint8_t* __restrict__ slots, const IP* __restrict__ indptr,
                                                      const double* __restrict__ vals_t, double* __restrict__ data) {
  constexpr int LP = (L + 3) & ~3;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  double* sv = acc + (size_t)caps * BLK;
  uint8_t* ss = reinterpret_cast<uint8_t*>(sv + (size_t)NS * ch * LP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ss + (size_t)NS * ch * LP);
  double* outs = reinterpret_cast<double*>(bars + NS);  // STAGED only
  __shared__ uint32_t s_range[3];
  __shared__ IP s_o0;

  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * BLK;
  const int64_t r = r0 + tid;
  const bool live = r < n_rows;
  uint32_t rs = 0, re = 0, cnt = 0;
  IP o0 = 0;
  if (live) {
    rs = row_start[r];
    re = row_start[r + 1];
    o0 = indptr[r];
    cnt = (uint32_t)(indptr[r + 1] - o0);
  }
  if (tid == 0) {
    const int64_t r1 = r0 + BLK < n_rows ? r0 + BLK : n_rows;
    s_range[0] = rs & ~3u;              // chunk origin: a multiple of 4 local rows (16-byte aligned slot bytes)
    s_range[1] = row_start[r1];         // end of the block's local rows
    s_range[2] = row_start[n_rows];     // end of the arrays
    s_o0 = o0;
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t s2 = 0; s2 < cnt; ++s2) acc[s2 * BLK + tid] = 0.0;
  __syncthreads();
  const uint32_t a0 = s_range[0], s1 = s_range[1], n_inc = s_range[2];
  const uint32_t nch = (s1 > a0) ? (s1 - a0 + ch - 1) / ch : 0;

  auto issue = [&](uint32_t k) {  // thread 0 only
    const int st = k % NS;
    const uint32_t c0 = a0 + k * ch;
    uint32_t rows = n_inc - c0 < (uint32_t)ch ? n_inc - c0 : (uint32_t)ch;  // never past the arrays
    const uint32_t bulk = rows & ~3u;  // bulk copies move multiples of 16 bytes
    double* dv = sv + (size_t)st * ch * LP;
    uint8_t* ds = ss + (size_t)st * ch * LP;
    for (uint32_t q = bulk; q < rows; ++q) {  // at most 3 tail rows of the whole array: plain copies
      for (int b = 0; b < LP; ++b) {
        dv[q * LP + b] = vals_t[(uint64_t)(c0 + q) * LP + b];
        ds[q * LP + b] = slots[(uint64_t)(c0 + q) * LP + b];
      }
    }
    if (bulk) {
      mbar_expect_tx(&bars[st], bulk * LP * 9u);
      bulk_g2s(dv, vals_t + (uint64_t)c0 * LP, bulk * LP * 8u, &bars[st]);
      bulk_g2s(ds, slots + (uint64_t)c0 * LP, bulk * LP, &bars[st]);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[st])) : "memory");
    }
  };

  if (tid == 0)
    for (uint32_t k = 0; k < nch && k < (uint32_t)NS; ++k) issue(k);

  for (uint32_t k = 0; k < nch; ++k) {
    const int st = k % NS;
    mbar_wait(&bars[st], (k / NS) & 1u);
    const uint32_t c0 = a0 + k * ch;
    const uint32_t c1 = c0 + ch < s1 ? c0 + ch : s1;
    const uint32_t qa = rs > c0 ? rs : c0, qb = re < c1 ? re : c1;
    const double* dv = sv + (size_t)st * ch * LP;
    const uint8_t* ds = ss + (size_t)st * ch * LP;
    for (uint32_t q = qa; q < qb; ++q) {
      const uint32_t o = (q - c0) * LP;
      double v[LP];
      uint32_t sl[LP];
#pragma unroll
      for (int b = 0; b < LP; b += 2) {
        const double2 t = *reinterpret_cast<const double2*>(dv + o + b);
        v[b] = t.x;
        v[b + 1] = t.y;
      }
#pragma unroll
      for (int b = 0; b < LP; b += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(ds + o + b);
        sl[b] = w & 0xffu;
        sl[b + 1] = (w >> 8) & 0xffu;
        sl[b + 2] = (w >> 16) & 0xffu;
        sl[b + 3] = w >> 24;
      }
#pragma unroll
      for (int b = 0; b < L; ++b) acc[sl[b] * BLK + tid] += v[b];
    }
    __syncthreads();  // every thread is done with this stage
    if (tid == 0 && k + NS < nch) issue(k + NS);
  }

  if constexpr (!STAGED) {
    for (uint32_t s2 = 0; s2 < cnt; ++s2) data[o0 + s2] = acc[s2 * BLK + tid];
  } else {
    const IP ob = s_o0;
    const uint32_t off = (uint32_t)(o0 - ob);
    for (uint32_t s2 = 0; s2 < cnt; ++s2) outs[off + s2] = acc[s2 * BLK + tid];
    __syncthreads();
    const int64_t r1 = r0 + BLK < n_rows ? r0 + BLK : n_rows;
    const uint32_t total = (uint32_t)(indptr[r1] - ob);
    for (uint32_t j = tid; j < total; j += BLK) data[ob + j] = outs[j];
  }
}

// ---- G lanes per CSR row ------------------------------------------------------------------------------
// Lane j of a row's group takes the local rows rs + j, rs + j + G, ... (ascending), so one warp-wide load
// covers 32 / G rows x G CONSECUTIVE local rows: 32 / G x (G * 32 B) contiguous pieces instead of 32
// separate sectors (G = 8: 8 L1 wavefronts per load instruction instead of 32; the slot words of the G
// lanes share one sector).  Every lane keeps its own column accumulators acc[slot][thread]; at the end
// the G partial sums of a column are added in lane order 0..G-1 by one lane per column and written out:
// consecutive lanes write consecutive entries.  The summation order is fixed (bit-reproducible): local
// rows j, j + G, ... within a lane, then the lanes in order.
template <int L, int G, typename IP>
__global__ void __launch_bounds__(256) k_numeric_g(int64_t n_rows, int caps, const uint32_t* __restrict__ row_start,
                                                   const uint8_t* __restrict__ slots, const IP* __restrict__ indptr,
                                                   const double* __restrict__ vals_t, double* __restrict__ data) {
  constexpr int LP = (L + 3) & ~3;
  constexpr int RPB = 256 / G;
  extern __shared__ __align__(16) double acc[];  // [slot][thread]
  const int tid = threadIdx.x;
  const int j = tid % G;
  const int64_t r = (int64_t)blockIdx.x * RPB + tid / G;
  const bool live = r < n_rows;
  uint32_t rs = 0, re = 0, cnt = 0;
  IP o0 = 0;
  if (live) {
    rs = row_start[r];
    re = row_start[r + 1];
    o0 = indptr[r];
    cnt = (uint32_t)(indptr[r + 1] - o0);
  }
  for (uint32_t s2 = 0; s2 < cnt; ++s2) acc[s2 * 256 + tid] = 0.0;
  constexpr int U = 2;  // local rows in flight per lane
  for (uint32_t q0 = rs + j; q0 < re; q0 += U * G) {
    double v[U][LP];
    uint32_t sw[U][LP / 4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = q0 + u * G;
      if (q < re) {
        const double* p = vals_t + (uint64_t)q * LP;
#pragma unroll
        for (int b = 0; b < LP; b += 4) {
          asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(v[u][b]), "=d"(v[u][b + 1]), "=d"(v[u][b + 2]), "=d"(v[u][b + 3])
                       : "l"(p + b));
        }
#pragma unroll
        for (int b = 0; b < LP / 4; ++b) sw[u][b] = __ldg(reinterpret_cast<const uint32_t*>(slots + (uint64_t)q * LP) + b);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (q0 + u * G < re) {
#pragma unroll
        for (int b = 0; b < L; ++b) acc[((sw[u][b / 4] >> (8 * (b % 4))) & 0xffu) * 256 + tid] += v[u][b];
      }
    }
  }
  __syncwarp();
  // column s of the row: lanes 0..G-1 in order, by lane s % G; consecutive lanes write consecutive entries
  const int base = tid - j;
  for (uint32_t s2 = j; s2 < cnt; s2 += G) {
    double sum = acc[s2 * 256 + base];
#pragma unroll
    for (int t = 1; t < G; ++t) sum += acc[s2 * 256 + base + t];
    data[o0 + s2] = sum;
  }
}

template <int L, int G, typename IP>
int launch_g(int64_t n_rows, int caps, const uint32_t* row_start, const uint8_t* slots, const void* indptr,
             const double* vals_t, double* data, cudaStream_t s) {
  size_t smem = (size_t)caps * 256 * sizeof(double);
  if (smem > 160 * 1024) return -1;
  int dev = 0;
  cudaGetDevice(&dev);
  static size_t attr[64] = {};
  if (dev >= 0 && dev < 64 && smem > 48 * 1024 && smem > attr[dev]) {
    PGB_CHECK(cudaFuncSetAttribute(k_numeric_g<L, G, IP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024)));
    attr[dev] = 160 * 1024;
  }
  k_numeric_g<L, G, IP><<<(unsigned)pgb_cdiv(n_rows, 256 / G), 256, smem, s>>>(n_rows, caps, row_start, slots,
                                                                              (const IP*)indptr, vals_t, data);
  PGB_LAUNCH_CHECK();
  return 0;
}

template <int L, typename IP>
int launch_gl(int G, int64_t n_rows, int caps, const uint32_t* row_start, const uint8_t* slots, const void* indptr,
              const double* vals_t, double* data, cudaStream_t s) {
  switch (G) {
    case 2: return launch_g<L, 2, IP>(n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 4: return launch_g<L, 4, IP>(n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 8: return launch_g<L, 8, IP>(n_rows, caps, row_start, slots, indptr, vals_t, data, s);
  }
  return -1;
}

template <typename IP>
int launch_gip(int L, int G, int64_t n_rows, int caps, const uint32_t* row_start, const uint8_t* slots,
               const void* indptr, const double* vals_t, double* data, cudaStream_t s) {
  switch (L) {
    case 3: return launch_gl<3, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 4: return launch_gl<4, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 5: return launch_gl<5, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 6: return launch_gl<6, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 8: return launch_gl<8, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 10: return launch_gl<10, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 12: return launch_gl<12, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
    case 15: return launch_gl<15, IP>(G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
  }
  return -1;
}

template <int L, int BLK, typename IP, bool STAGED>
int launch_b(int64_t n_rows, int caps, const uint32_t* row_start, const uint8_t* slots, const void* indptr,
             const double* vals_t, double* data, cudaStream_t s) {
  constexpr int LP = (L + 3) & ~3;
  // chunk: about 18 KB of values + slots per stage, a multiple of 4 local rows
  int ch = (18 * 1024 / (LP * 9)) & ~3;
  if (ch < 16) ch = 16;
  size_t smem = (size_t)caps * BLK * 8 * (STAGED ? 2 : 1) + (size_t)NS * ch * LP * 9 + NS * 8 + 128;
  if (smem > 200 * 1024) return -1;  // caller falls back to k_numeric_t
  int dev = 0;
  cudaGetDevice(&dev);
  static size_t attr[64] = {};
  if (dev >= 0 && dev < 64 && smem > attr[dev]) {
    PGB_CHECK(cudaFuncSetAttribute(k_numeric_bulk<L, BLK, IP, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(200 * 1024)));
    attr[dev] = 200 * 1024;
  }
  k_numeric_bulk<L, BLK, IP, STAGED><<<(unsigned)pgb_cdiv(n_rows, BLK), BLK, smem, s>>>(
      n_rows, caps, ch, row_start, slots, (const IP*)indptr, vals_t, data);
  PGB_LAUNCH_CHECK();
  return 0;
}

template <int L, typename IP>
int launch_l(int64_t n_rows, int caps, bool staged, const uint32_t* row_start, const uint8_t* slots,
             const void* indptr, const double* vals_t, double* data, cudaStream_t s) {
  if (staged) return launch_b<L, 128, IP, true>(n_rows, caps, row_start, slots, indptr, vals_t, data, s);
  return launch_b<L, 128, IP, false>(n_rows, caps, row_start, slots, indptr, vals_t, data, s);
}

template <typename IP>
int launch_ip(int L, int64_t n_rows, int caps, bool staged, const uint32_t* row_start, const uint8_t* slots,
              const void* indptr, const double* vals_t, double* data, cudaStream_t s) {
  switch (L) {
    case 3: return launch_l<3, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 4: return launch_l<4, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 5: return launch_l<5, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 6: return launch_l<6, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 8: return launch_l<8, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 10: return launch_l<10, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 12: return launch_l<12, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
    case 15: return launch_l<15, IP>(n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
  }
  return -1;
}

}  // namespace

// returns 0 = launched, -1 = not applicable (caller uses k_numeric_t), > 0 = error
int pgb_numeric_bulk_launch(int L, int64_t n_rows, int caps, bool staged, int idx64, const uint32_t* row_start,
                            const uint8_t* slots, const void* indptr, const double* vals_t, double* data,
                            cudaStream_t s) {
  return idx64 ? launch_ip<int64_t>(L, n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s)
               : launch_ip<int32_t>(L, n_rows, caps, staged, row_start, slots, indptr, vals_t, data, s);
}

// G lanes per CSR row (2, 4 or 8): 0 = launched, -1 = not applicable, > 0 = error
int pgb_numeric_group_launch(int L, int G, int64_t n_rows, int caps, int idx64, const uint32_t* row_start,
                             const uint8_t* slots, const void* indptr, const double* vals_t, double* data,
                             cudaStream_t s) {
  return idx64 ? launch_gip<int64_t>(L, G, n_rows, caps, row_start, slots, indptr, vals_t, data, s)
               : launch_gip<int32_t>(L, G, n_rows, caps, row_start, slots, indptr, vals_t, data, s);
}
