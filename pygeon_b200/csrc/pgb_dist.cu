// K4: row-distributed CG / MINRES over NVLink peer memory (one process per GPU, one node).
//
// The solve of the cell-partitioned systems needs, per iteration, one halo exchange of
// ghost vector entries with the neighbouring ranks and two or three global sums of scalars.  Here both
// are done by the Krylov kernels themselves through peer-mapped memory (cudaIpc handles over
// NVLink / NVSwitch) instead of separate collective calls:
//
//  * halo: the kernel that produces the next search / Lanczos vector first recomputes the entries its
//    neighbours need and stores them straight into the neighbours' ghost blocks (P2P stores), then
//    raises a sequence flag at each neighbour.  The SpMV that follows runs the rows without ghost
//    columns first and only then waits for the flags: the transfer overlaps the interior rows.
//  * global sums: every kernel that produces a dot product leaves one partial per block (as on one
//    GPU); the last block to finish (integer ticket) sums them in a fixed order, stores the value into
//    slot [rank] at EVERY rank and raises that slot's flag.  A consumer kernel waits for the `world`
//    flags and adds the slots in rank order, so all ranks obtain bit-identical scalars, take the same
//    branches and stop at the same iteration.  No floating-point atomics; results are reproducible
//    for a fixed partition.
//
// Iteration = 3 kernel launches for CG and for MINRES (the one-thread scalar kernel of the single-GPU
// MINRES driver is folded into the vector update: every block advances a private copy of the state).
// Spin loops are bounded (TIMEOUT_NS): a rank that never shows up turns into an error code,
// not a hung GPU.
#include "pgb_krylov.cuh"

namespace {

constexpr int MAXW = 16;    // ranks per node
constexpr int NRED = 8;     // reduction slots
constexpr unsigned long long TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// lives at the start of every rank's arena; written by peers
struct Header {
  unsigned long long halo_flag[MAXW];          // [src rank]: latest halo sequence number pushed by src
  unsigned long long red_flag[2][NRED][MAXW];  // [parity][slot][src rank]
  double red_val[2][NRED][MAXW];
  unsigned long long bar_flag[MAXW];
  int err;                                     // 1 = a wait timed out
  int pad[15];
};
constexpr size_t HEADER_BYTES = (sizeof(Header) + 4095) / 4096 * 4096;

struct Comm {
  int rank = 0, world = 1, device = 0;
  size_t arena_bytes = 0;
  char* base = nullptr;
  char* peer[MAXW] = {};
  bool opened[MAXW] = {};
  cudaIpcMemHandle_t handle;
  unsigned long long epoch = 16;  // sequence numbers handed out so far (monotonic over the comm's life)
  unsigned long long bar_epoch = 0;
  unsigned* tickets = nullptr;  // device, 16 counters
  int* flag_host = nullptr;     // mapped pinned stop flag
  int* flag_dev = nullptr;
  void* scal = nullptr;         // device scalars (CgScal / 2 x MinresScal)
  double* parts = nullptr;      // device, NRED * NPART partial sums
  double* gsum = nullptr;       // device, 2 * NRED completed global sums
  double* resid = nullptr;
  int resid_cap = 0;
};

struct CommDev {
  int rank, world;
  Header* hdr[MAXW];
  double* gsum;  // [2][NRED] completed global sums (device local)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one thread: spin until *flag >= want (bounded)
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long want, Header* mine) {
  if (ld_acquire_sys(flag) >= want) return;
  if (*(volatile int*)&mine->err) return;  // an earlier wait timed out: do not stack up timeouts
  const unsigned long long t0 = now_ns();
  int spins = 0;
  while (ld_acquire_sys(flag) < want) {
    if (++spins == 1024) {
      spins = 0;
      if (now_ns() - t0 > TIMEOUT_NS) {
        mine->err = 1;
        __threadfence_system();
        return;
      }
    }
  }
}

// The global value of slot `k` for sequence number `seq`: completed by the tail of the kernel that produced
// the partial sums (all_reduce_tail below), i.e. by an EARLIER kernel of this stream - a plain load.
__device__ __forceinline__ double global_sum(const CommDev& c, int k, unsigned long long seq, double* /*sh*/) {
  return c.gsum[(int)(seq & 1ull) * NRED + k];
}

// After write_partial(parts, ...) by every block: the last block to arrive (integer ticket) sums the partials
// in index order, stores the value into slot [rank] of EVERY rank and raises the slot's flag (lane q talks to
// rank q), then waits for the `world` flags at home and adds the slots in rank order: every rank computes
// the same bits.  The all-reduce is complete when the kernel ends; consumers are later kernels and never
// spin (one poller per rank instead of one per block).  NV = 1 or 2 values reduced by one pass.
template <int NV>
__device__ __forceinline__ void all_reduce_tail(const CommDev& c, const int (&k)[NV], unsigned long long seq,
                                                const double* const (&parts)[NV], unsigned* ticket,
                                                double* ws /*[NV * KB/32]*/, int* sh_last) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *sh_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!*sh_last) return;
  __threadfence();
  double v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = 0.0;
  for (int i = threadIdx.x; i < NPART; i += KB) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] += ld_volatile_f64(parts[j] + i);
  }
  block_sum_n<NV>(v, ws);
  if (threadIdx.x >= 32) return;
  const int q = threadIdx.x;  // lane q <-> rank q
  const int par = (int)(seq & 1ull);
  if (q == 0) *ticket = 0;
  if (q < c.world) {
#pragma unroll
    for (int j = 0; j < NV; ++j) c.hdr[q]->red_val[par][k[j]][c.rank] = v[j];
    __threadfence_system();
#pragma unroll
    for (int j = 0; j < NV; ++j) st_release_sys(&c.hdr[q]->red_flag[par][k[j]][c.rank], seq);
  }
  Header* mine = c.hdr[c.rank];
  double got[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    got[j] = 0.0;
    if (q < c.world) {
      wait_flag(&mine->red_flag[par][k[j]][q], seq, mine);
      got[j] = ld_volatile_f64(&mine->red_val[par][k[j]][q]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double tot = 0.0;
    for (int r = 0; r < c.world; ++r) tot += __shfl_sync(FULL, got[j], r);  // rank order: identical on all ranks
    if (q == 0) c.gsum[par * NRED + k[j]] = tot;
  }
}

__device__ __forceinline__ void publish(const CommDev& c, int k, unsigned long long seq, const double* parts,
                                        unsigned* ticket, double* ws, int* sh_last) {
  const int ks[1] = {k};
  const double* const ps[1] = {parts};
  all_reduce_tail<1>(c, ks, seq, ps, ticket, ws, sh_last);
}

__device__ __forceinline__ void publish2(const CommDev& c, int k0, int k1, unsigned long long seq, const double* parts0,
                                         const double* parts1, unsigned* ticket, double* ws, int* sh_last) {
  const int ks[2] = {k0, k1};
  const double* const ps[2] = {parts0, parts1};
  all_reduce_tail<2>(c, ks, seq, ps, ticket, ws, sh_last);
}

// ---- halo description (device side) ------------------------------------------------------------------
struct Halo {
  int nsend, nrecv;             // peers
  int send_peer[MAXW], recv_peer[MAXW];
  long long send_off[MAXW + 1];  // ranges of send_idx per send peer
  long long send_dst[MAXW];      // element offset of this rank's block inside the peer's vector
  const int32_t* send_idx;       // owned entries the peers need
  size_t vec_off[2];             // byte offset of the exchanged vector(s) inside every rank's arena
  unsigned* ticket;
};

// All blocks: store the send entries produced by f(i) into the peers' ghost blocks; the last block raises
// the flags.  `which` selects the arena vector (double-buffered p of CG), seq the halo sequence number.
template <typename F>
__device__ __forceinline__ void halo_push(const CommDev& c, const Halo& h, int which, unsigned long long seq, F f,
                                          int* sh_last) {
  if (h.nsend == 0) return;  // block-uniform: nothing to send (one rank, or an isolated partition)
  for (int s = 0; s < h.nsend; ++s) {
    const int q = h.send_peer[s];
    double* dst = (double*)((char*)c.hdr[q] + HEADER_BYTES + h.vec_off[which]) + h.send_dst[s];
    const long long lo = h.send_off[s], cnt = h.send_off[s + 1] - lo;
    for (long long j = (long long)blockIdx.x * KB + threadIdx.x; j < cnt; j += (long long)gridDim.x * KB)
      dst[j] = f(h.send_idx[lo + j]);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *sh_last = (atomicAdd(h.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (*sh_last && threadIdx.x == 0) {
    *h.ticket = 0;
    __threadfence_system();
    for (int s = 0; s < h.nsend; ++s) st_release_sys(&c.hdr[h.send_peer[s]]->halo_flag[c.rank], seq);
  }
}

__device__ __forceinline__ void halo_wait(const CommDev& c, const Halo& h, unsigned long long seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    Header* mine = c.hdr[c.rank];
    for (int s = 0; s < h.nrecv; ++s) wait_flag(&mine->halo_flag[h.recv_peer[s]], seq, mine);
  }
  __syncthreads();
}

// ---- SpMV over the owned rows: rows without ghost columns [ia, ib) first, the interface rows after the halo ----
struct DSpmv {
  int64_t n, ia, ib;
  int32_t gs;             // first ghost column
  const void* indptr;
  const int32_t* indices;
  const double* data;
  const double* x;
  double* y;
  double* parts;
  const double* r1;       // MINRES
  const MinresScal* mr;   // MINRES: state after the previous iteration
  const int* done;
  int itn;
};

// x is read through the non-coherent path like the single-GPU SpMV (plain loads cost the kernel 10-20 %, a
// per-element choice between the two paths 30 %: measured).  Owned entries were written by an earlier kernel.
// Ghost entries are stored by the peers while this kernel runs, but always before the flag the block waits
// for (halo_wait ends with a block barrier, which loads are not moved across), they are read only after
// that wait, and no 128-byte line mixes owned and ghost entries (ghost_start is a multiple of 16), so a
// line holding ghosts is first fetched when its data is final.
__device__ __forceinline__ double load_x(const double* x, int32_t c) { return __ldg(x + c); }

template <int TPR, typename IP, int MODE>
__global__ void __launch_bounds__(KB, (sizeof(IP) == 4 && MODE == SPMV_MINRES) ? 6 : 0) kd_spmv(DSpmv a, CommDev c, Halo h, unsigned long long halo_seq, int red_slot,
                                              unsigned long long red_seq, unsigned* ticket) {
  __shared__ double ws[KB / 32];
  __shared__ int sh_last;
  if (*a.done) return;
  constexpr int RPB = KB / TPR;
  constexpr int UR = SPMV_UR;
  const int sub = threadIdx.x % TPR;
  const int grp = threadIdx.x / TPR;
  const IP* __restrict__ ip = reinterpret_cast<const IP*>(a.indptr);
  const double* __restrict__ data = a.data;
  const int32_t* __restrict__ indices = a.indices;
  double shift = 0.0, coef = 0.0;
  if (MODE == SPMV_MINRES) {
    shift = a.mr->shift;
    coef = (a.itn >= 2) ? a.mr->beta / a.mr->oldb : 0.0;
  }
  // rows are int32 ids; all row arithmetic in 32 bits (the kernel is close to issue bound at 2 lanes per row)
  const uint32_t n32 = (uint32_t)a.n, ia = (uint32_t)a.ia, ib = (uint32_t)a.ib, nint = ib - ia;
  double dot = 0.0;
  bool waited = (h.nrecv == 0);
  const uint32_t ntiles = (n32 + RPB - 1) / RPB;
  for (uint32_t t = blockIdx.x * UR; t < ntiles; t += gridDim.x * UR) {
    const bool inner = (t + UR) * RPB <= nint;  // block-uniform: the whole tile group lies in [ia, ib)
    if (!waited && !inner) {                    // this block reaches the interface rows
      halo_wait(c, h, halo_seq);
      waited = true;
    }
    uint32_t row[UR];
    IP lo[UR], hi[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const uint32_t v = (t + u) * RPB + grp;  // virtual row: interior first
      const bool ok = v < n32;
      uint32_t r = ia + v;
      if (!inner) r = v < nint ? ia + v : (v - nint < ia ? v - nint : ib + (v - nint - ia));
      row[u] = ok ? r : n32;
      lo[u] = ok ? ip[r] : (IP)0;
      hi[u] = ok ? ip[r + 1] : (IP)0;
    }
    double dv[UR];
    int32_t ci[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      IP j = lo[u] + sub;
      bool ok = j < hi[u];
      dv[u] = ok ? data[j] : 0.0;
      ci[u] = ok ? indices[j] : -1;
    }
    double s[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) s[u] = (ci[u] >= 0) ? dv[u] * load_x(a.x, ci[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < UR; ++u)
      for (IP j = lo[u] + sub + TPR; j < hi[u]; j += TPR) s[u] += data[j] * load_x(a.x, indices[j]);
#pragma unroll
    for (int u = 0; u < UR; ++u) {
#pragma unroll
      for (int d = TPR / 2; d > 0; d >>= 1) s[u] += __shfl_xor_sync(FULL, s[u], d);
    }
    if (sub == 0) {
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        if (row[u] < n32) {
          double sv = s[u];
          const double xv = __ldg(a.x + row[u]);  // owned entry
          if (MODE == SPMV_MINRES) {
            if (shift != 0.0) sv -= shift * xv;
            if (a.itn >= 2) sv -= coef * __ldg(a.r1 + row[u]);
          }
          dot += xv * sv;
          a.y[row[u]] = sv;
        }
      }
    }
  }
  double tot = block_sum(dot, ws);
  write_partial(a.parts, tot);
  publish(c, red_slot, red_seq, a.parts, ticket, ws, &sh_last);
}

template <typename IP, int MODE>
int launch_dspmv(const DSpmv& a, const CommDev& c, const Halo& h, unsigned long long hs, int slot,
                 unsigned long long rs, unsigned* ticket, int tpr, cudaStream_t s) {
#define PGB_DSPMV(T)                                                                                   \
  {                                                                                                    \
    int64_t ntiles = ((a.n + (KB / T) - 1) / (KB / T) + SPMV_UR - 1) / SPMV_UR;                         \
    int g = resident_grid(kd_spmv<T, IP, MODE>, ntiles);                                               \
    kd_spmv<T, IP, MODE><<<g, KB, 0, s>>>(a, c, h, hs, slot, rs, ticket);                              \
  }
  switch (tpr) {
    case 2: PGB_DSPMV(2) break;
    case 4: PGB_DSPMV(4) break;
    case 8: PGB_DSPMV(8) break;
    case 16: PGB_DSPMV(16) break;
    default: PGB_DSPMV(32) break;
  }
#undef PGB_DSPMV
  PGB_LAUNCH_CHECK();
  return 0;
}

// ---- CG ------------------------------------------------------------------------------------------------
enum { S_RR = 0, S_RHO = 1, S_BB = 2, S_PQ = 3 };                  // CG slots
enum { M_ALFA = 0, M_BETA = 1, M_XX = 2, M_B1 = 3, M_BB = 4 };      // MINRES slots
enum { T_HALO = 0, T_A = 1, T_B = 2, T_C = 3 };                     // ticket counters

// r = b (x0 = 0); partials rr = rho' = b.b, rho = b.z
__global__ void __launch_bounds__(KB) kd_cg_init(int64_t n, const double* __restrict__ b,
                                                 const double* __restrict__ dinv, double* __restrict__ r,
                                                 double* __restrict__ x, double* parts, CommDev c,
                                                 unsigned long long seq, unsigned* tickets) {
  __shared__ double ws[2 * (KB / 32)];
  __shared__ int sh_last;
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
    const double bi = b[i];
    r[i] = bi;
    x[i] = 0.0;
    v[0] += bi * bi;
    v[1] += bi * (dinv ? dinv[i] * bi : bi);
  }
  block_sum_n<2>(v, ws);
  write_partial(parts + S_RR * NPART, v[0]);
  write_partial(parts + S_RHO * NPART, v[1]);
  publish2(c, S_RR, S_RHO, seq, parts + S_RR * NPART, parts + S_RHO * NPART, tickets + T_A, ws, &sh_last);
}

// iteration k: convergence test on the global ||r||, beta, halo push of the new p, p_new = z + beta p_old
__global__ void __launch_bounds__(KB) kd_cg_p(int64_t n, int k, const double* __restrict__ r,
                                              const double* __restrict__ dinv, const double* __restrict__ p_old,
                                              double* __restrict__ p_new, CgScal* sc, int* done, double* resid,
                                              volatile int* host_flag, CommDev c, Halo h, unsigned long long red_seq,
                                              unsigned long long halo_seq) {
  __shared__ double sh[1];
  __shared__ int sh_last;
  if (*done) return;
  const double rr = global_sum(c, S_RR, red_seq, sh);
  const double rho = global_sum(c, S_RHO, red_seq, sh);
  const double rnorm = sqrt(rr);
  const double bb = (k == 0) ? rr : sc->bnorm2;  // x0 = 0: r_0 = b
  const double atol = fmax(sc->atol_user, sc->rtol * sqrt(bb));
  const bool stop = rnorm < atol || (k == 0 && rr == 0.0);  // b == 0: scipy returns x = 0 at once
  const double rho_prev = sc->rho_prev;  // written by the previous kd_cg_xr
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (k < sc->resid_cap) resid[k] = rnorm;
    if (k == 0) sc->bnorm2 = rr;
    if (stop) {
      *done = 1;
      sc->done = 1;
      *host_flag = 1;
      __threadfence_system();
    } else {
      sc->rho = rho;
      sc->iters = k + 1;
    }
  }
  if (stop) return;
  const double beta = (k > 0) ? rho / rho_prev : 0.0;
  halo_push(c, h, k & 1, halo_seq, [&](int32_t i) {
    const double z = dinv ? dinv[i] * r[i] : r[i];
    return (k > 0) ? z + beta * p_old[i] : z;
  }, &sh_last);
  const int64_t st = (int64_t)gridDim.x * KB;
  for (int64_t i0 = (int64_t)blockIdx.x * KB + threadIdx.x; i0 < n; i0 += VU * st) {
    double rv[VU], pv[VU], dv[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      const bool ok = i < n;
      rv[u] = ok ? r[i] : 0.0;
      pv[u] = (ok && k > 0) ? p_old[i] : 0.0;
      dv[u] = (ok && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      if (i < n) {
        const double z = dinv ? dv[u] * rv[u] : rv[u];
        p_new[i] = (k > 0) ? z + beta * pv[u] : z;
      }
    }
  }
}

// alpha = rho / (p.q); x += alpha p; r -= alpha q; global rr, rho
__global__ void __launch_bounds__(KB) kd_cg_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ q,
                                               const double* __restrict__ dinv, double* __restrict__ x,
                                               double* __restrict__ r, double* parts, CgScal* sc, const int* done,
                                               CommDev c, unsigned long long pq_seq, unsigned long long out_seq,
                                               unsigned* tickets) {
  __shared__ double sh[1];
  __shared__ double ws[2 * (KB / 32)];
  __shared__ int sh_last;
  if (*done) return;
  const double pq = global_sum(c, S_PQ, pq_seq, sh);
  const double rho_cur = sc->rho;
  const double alpha = rho_cur / pq;
  double v[2] = {0.0, 0.0};
  const int64_t st = (int64_t)gridDim.x * KB;
  for (int64_t i0 = (int64_t)blockIdx.x * KB + threadIdx.x; i0 < n; i0 += VU * st) {
    double pv[VU], qv[VU], xv[VU], rv[VU], dv[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      const bool ok = i < n;
      pv[u] = ok ? p[i] : 0.0;
      qv[u] = ok ? q[i] : 0.0;
      xv[u] = ok ? x[i] : 0.0;
      rv[u] = ok ? r[i] : 0.0;
      dv[u] = (ok && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      if (i < n) {
        x[i] = xv[u] + alpha * pv[u];
        const double ri = rv[u] - alpha * qv[u];
        r[i] = ri;
        v[0] += ri * ri;
        v[1] += ri * (dinv ? dv[u] * ri : ri);
      }
    }
  }
  block_sum_n<2>(v, ws);
  write_partial(parts + S_RR * NPART, v[0]);
  write_partial(parts + S_RHO * NPART, v[1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->rho_prev = rho_cur;
  publish2(c, S_RR, S_RHO, out_seq, parts + S_RR * NPART, parts + S_RHO * NPART, tickets + T_A, ws, &sh_last);
}

__global__ void kd_cg_final(const CgScal* sc, const int* done, double* resid, CommDev c, unsigned long long seq) {
  __shared__ double sh[1];
  if (*done) return;
  const double rr = global_sum(c, S_RR, seq, sh);
  if (threadIdx.x == 0 && sc->iters < sc->resid_cap) resid[sc->iters] = sqrt(rr);
}

// ---- MINRES ------------------------------------------------------------------------------------------
// r1 = b (x0 = 0), x = w = w2 = 0; global beta1^2 = b.psolve(b) and b.b
__global__ void __launch_bounds__(KB) kd_mr_init(int64_t n, const double* __restrict__ b,
                                                 const double* __restrict__ dinv, double* __restrict__ r1,
                                                 double* __restrict__ x, double* __restrict__ w,
                                                 double* __restrict__ w2, double* parts, CommDev c,
                                                 unsigned long long seq, unsigned* tickets) {
  __shared__ double ws[2 * (KB / 32)];
  __shared__ int sh_last;
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
    const double bi = b[i];
    r1[i] = bi;
    x[i] = 0.0;
    w[i] = 0.0;
    w2[i] = 0.0;
    v[0] += bi * (dinv ? dinv[i] * bi : bi);
    v[1] += bi * bi;
  }
  block_sum_n<2>(v, ws);
  write_partial(parts + M_B1 * NPART, v[0]);
  write_partial(parts + M_BB * NPART, v[1]);
  publish2(c, M_B1, M_BB, seq, parts + M_B1 * NPART, parts + M_BB * NPART, tickets + T_A, ws, &sh_last);
}

// state 0 from the global beta1^2; v = psolve(r1)/beta1 with its halo push; out2 = (beta1^2, b.b)
__global__ void __launch_bounds__(KB) kd_mr_start(int64_t n, const double* __restrict__ r1,
                                                  const double* __restrict__ dinv, double* __restrict__ v,
                                                  MinresScal* sc0, int* done, double* out2, double rtol, double shift,
                                                  int maxiter, int resid_cap, CommDev c, Halo h,
                                                  unsigned long long red_seq, unsigned long long halo_seq) {
  __shared__ double sh[1];
  __shared__ int sh_last;
  const double b1 = global_sum(c, M_B1, red_seq, sh);
  const double bb = global_sum(c, M_BB, red_seq, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    MinresScal s = {};
    s.beta1 = sqrt(fmax(b1, 0.0));
    s.beta = s.beta1;
    s.phibar = s.beta1;
    s.rhs1 = s.beta1;
    s.gmin = 1.7976931348623157e308;
    s.cs = -1.0;
    s.rtol = rtol;
    s.shift = shift;
    s.maxiter = maxiter;
    s.resid_cap = resid_cap;
    *sc0 = s;
    out2[0] = b1;
    out2[1] = bb;
    if (b1 == 0.0 || bb == 0.0) *done = 1;
  }
  if (b1 == 0.0 || bb == 0.0) return;
  const double ib = 1.0 / sqrt(fmax(b1, 0.0));
  halo_push(c, h, 0, halo_seq, [&](int32_t i) { return (dinv ? dinv[i] * r1[i] : r1[i]) * ib; }, &sh_last);
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB)
    v[i] = (dinv ? dinv[i] * r1[i] : r1[i]) * ib;
}

// y -= (alfa / beta) r2; global beta^2 = y.psolve(y)
__global__ void __launch_bounds__(KB) kd_mr_c(int64_t n, double* __restrict__ y, const double* __restrict__ r2,
                                              const double* __restrict__ dinv, const MinresScal* sc, const int* done,
                                              double* parts, CommDev c, unsigned long long alfa_seq,
                                              unsigned long long out_seq, unsigned* tickets) {
  __shared__ double sh[1];
  __shared__ double ws[KB / 32];
  __shared__ int sh_last;
  if (*done) return;
  const double alfa = global_sum(c, M_ALFA, alfa_seq, sh);
  const double cf = alfa / sc->beta;
  double bt = 0;
  const int64_t st = (int64_t)gridDim.x * KB;
  for (int64_t i0 = (int64_t)blockIdx.x * KB + threadIdx.x; i0 < n; i0 += VU * st) {
    double yv[VU], r2v[VU], dv[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      const bool ok = i < n;
      yv[u] = ok ? y[i] : 0.0;
      r2v[u] = ok ? r2[i] : 0.0;
      dv[u] = (ok && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      if (i < n) {
        const double yi = yv[u] - cf * r2v[u];
        y[i] = yi;
        bt += yi * (dinv ? dv[u] * yi : yi);
      }
    }
  }
  write_partial(parts + M_BETA * NPART, block_sum(bt, ws));
  publish(c, M_BETA, out_seq, parts + M_BETA * NPART, tickets + T_B, ws, &sh_last);
}

__device__ __noinline__ bool mr_scalar_step_call(MinresScal* s, int itn, double bt, double xx) {
  return mr_scalar_step(s, itn, bt, xx, 0);  // out of line: keeps the registers of the streaming loop low
}

// Every block advances a private copy of the scalar state (identical on all blocks and ranks), block 0
// stores it; then halo push of the new v, w = (v - oldeps w1 - delta w2) denom over w1, x += phi w,
// v = psolve(r2) / beta; global x.x.
__global__ void __launch_bounds__(KB) kd_mr_d(int64_t n, int itn, double* __restrict__ v, double* __restrict__ w1,
                                              const double* __restrict__ w2, const double* __restrict__ r2,
                                              const double* __restrict__ dinv, double* __restrict__ x,
                                              const MinresScal* sc_old, MinresScal* sc_new, int* done, double* resid,
                                              volatile int* host_flag, double* parts, CommDev c, Halo h,
                                              unsigned long long alfa_seq, unsigned long long beta_seq,
                                              unsigned long long xx_seq_in, unsigned long long xx_seq_out,
                                              unsigned long long halo_seq, unsigned* tickets) {
  __shared__ double sh[1];
  __shared__ double coef[5];
  __shared__ int sh_stop;
  __shared__ double ws[KB / 32];
  __shared__ int sh_last;
  if (*done) return;
  const double alfa = global_sum(c, M_ALFA, alfa_seq, sh);
  const double bt = global_sum(c, M_BETA, beta_seq, sh);
  const double xx = (itn >= 2) ? global_sum(c, M_XX, xx_seq_in, sh) : 0.0;
  if (threadIdx.x == 0) {
    MinresScal s = *sc_old;
    s.alfa = alfa;
    const bool stop = mr_scalar_step_call(&s, itn, bt, xx);
    sh_stop = stop ? 1 : 0;
    coef[0] = s.oldeps; coef[1] = s.delta; coef[2] = s.denom; coef[3] = s.phi; coef[4] = s.inv_beta;
    if (blockIdx.x == 0) {
      *sc_new = s;
      if (stop) {
        *done = 1;
        *host_flag = 1;
        __threadfence_system();
      } else if (itn < s.resid_cap) {
        resid[itn] = s.phibar;
      }
    }
  }
  __syncthreads();
  if (sh_stop) return;
  const double oldeps = coef[0], delta = coef[1], denom = coef[2], phi = coef[3], ib = coef[4];
  halo_push(c, h, 0, halo_seq, [&](int32_t i) { return (dinv ? dinv[i] * r2[i] : r2[i]) * ib; }, &sh_last);
  double acc = 0;
  const int64_t st = (int64_t)gridDim.x * KB;
  for (int64_t i0 = (int64_t)blockIdx.x * KB + threadIdx.x; i0 < n; i0 += VU * st) {
    double vv[VU], w1v[VU], w2v[VU], r2v[VU], xv[VU], dv[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      const bool ok = i < n;
      vv[u] = ok ? v[i] : 0.0;
      w1v[u] = ok ? w1[i] : 0.0;
      w2v[u] = ok ? w2[i] : 0.0;
      r2v[u] = ok ? r2[i] : 0.0;
      xv[u] = ok ? x[i] : 0.0;
      dv[u] = (ok && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      if (i < n) {
        const double wn = (vv[u] - oldeps * w1v[u] - delta * w2v[u]) * denom;
        w1[i] = wn;
        const double xi = xv[u] + phi * wn;
        x[i] = xi;
        acc += xi * xi;
        v[i] = (dinv ? dv[u] * r2v[u] : r2v[u]) * ib;
      }
    }
  }
  write_partial(parts + M_XX * NPART, block_sum(acc, ws));
  publish(c, M_XX, xx_seq_out, parts + M_XX * NPART, tickets + T_C, ws, &sh_last);
}

// final stopping test after the last launched iteration (maxiter reached or the host stopped launching)
__global__ void kd_mr_final(int itn, const MinresScal* sc_old, MinresScal* sc_out, const int* done, CommDev c,
                            unsigned long long xx_seq) {
  __shared__ double sh[1];
  if (*done) {
    if (threadIdx.x == 0) *sc_out = *sc_old;
    return;
  }
  const double xx = (itn >= 2) ? global_sum(c, M_XX, xx_seq, sh) : 0.0;
  if (threadIdx.x == 0) {
    MinresScal s = *sc_old;
    mr_scalar_step(&s, itn, 0.0, xx, 1);
    *sc_out = s;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
int comm_dev(const Comm* cm, CommDev* out) {
  out->rank = cm->rank;
  out->world = cm->world;
  out->gsum = cm->gsum;
  for (int q = 0; q < MAXW; ++q) out->hdr[q] = q < cm->world ? (Header*)cm->peer[q] : nullptr;
  return 0;
}

struct HaloHost {
  int nsend, nrecv;
  const int* send_peer;
  const int64_t* send_off;
  const int64_t* send_dst;
  const int32_t* send_idx;
  const int* recv_peer;
};

int make_halo(const Comm* cm, int nsend, const int* send_peer, const int64_t* send_off, const int64_t* send_dst,
              const int32_t* send_idx, int nrecv, const int* recv_peer, size_t off0, size_t off1, Halo* h) {
  PGB_REQUIRE(nsend >= 0 && nsend <= MAXW && nrecv >= 0 && nrecv <= MAXW, "too many halo peers");
  memset(h, 0, sizeof(*h));
  h->nsend = nsend;
  h->nrecv = nrecv;
  for (int s = 0; s < nsend; ++s) {
    PGB_REQUIRE(send_peer[s] >= 0 && send_peer[s] < cm->world && send_peer[s] != cm->rank, "bad send peer");
    h->send_peer[s] = send_peer[s];
    h->send_off[s] = send_off[s];
    h->send_dst[s] = send_dst[s];
  }
  h->send_off[nsend] = nsend ? send_off[nsend] : 0;
  for (int s = 0; s < nrecv; ++s) {
    PGB_REQUIRE(recv_peer[s] >= 0 && recv_peer[s] < cm->world && recv_peer[s] != cm->rank, "bad recv peer");
    h->recv_peer[s] = recv_peer[s];
  }
  h->send_idx = send_idx;
  h->vec_off[0] = off0;
  h->vec_off[1] = off1;
  h->ticket = cm->tickets + T_HALO;
  return 0;
}

int ensure_resid(Comm* cm, int maxiter) {
  int64_t want = (int64_t)maxiter + 2;
  int cap = (int)(want < PGB_RESID_CAP ? want : PGB_RESID_CAP);
  if (cm->resid_cap < cap) {
    if (cm->resid) cudaFree(cm->resid);
    cm->resid = nullptr;
    cm->resid_cap = 0;
    PGB_CHECK(cudaMalloc((void**)&cm->resid, sizeof(double) * (size_t)(cap + 2)));
    cm->resid_cap = cap;
  }
  return 0;
}

// The waits inside a kernel are on flags raised by kernels of OTHER ranks (which depend on EARLIER kernels
// of this rank only), so no block ever waits for a block of its own grid; the grids are sized to one
// resident wave anyway (resident_grid), which also removes the tail of late blocks.
#define VGRID(kernel) resident_grid(kernel, pgb_cdiv(n, KB))

int check_err(Comm* cm, cudaStream_t s, const char* what) {
  int err = 0;
  PGB_CHECK(cudaMemcpyAsync(&err, &((Header*)cm->base)->err, sizeof(int), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (err) {
    pgb_set_error("%s: a wait on a peer rank timed out (rank %d of %d)", what, cm->rank, cm->world);
    return 4;
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// communicator: arena allocation, IPC handle exchange (done by the caller), peer mapping
// ---------------------------------------------------------------------------------------------------------
extern "C" int pgb_comm_create(int rank, int world, int64_t arena_bytes, void** comm_out) {
  PGB_REQUIRE(world >= 1 && world <= MAXW && rank >= 0 && rank < world && arena_bytes >= 0, "pgb_comm_create: bad arguments");
  Comm* cm = new Comm();
  cm->rank = rank;
  cm->world = world;
  cm->arena_bytes = ((size_t)arena_bytes + 255) / 256 * 256;
  PGB_CHECK(cudaGetDevice(&cm->device));
  PGB_CHECK(cudaMalloc((void**)&cm->base, HEADER_BYTES + cm->arena_bytes));
  PGB_CHECK(cudaMemset(cm->base, 0, HEADER_BYTES));
  PGB_CHECK(cudaMalloc((void**)&cm->tickets, 16 * sizeof(unsigned)));
  PGB_CHECK(cudaMemset(cm->tickets, 0, 16 * sizeof(unsigned)));
  PGB_CHECK(cudaHostAlloc((void**)&cm->flag_host, 64, cudaHostAllocMapped));
  PGB_CHECK(cudaHostGetDevicePointer((void**)&cm->flag_dev, cm->flag_host, 0));
  PGB_CHECK(cudaMalloc(&cm->scal, 4096));
  PGB_CHECK(cudaMalloc((void**)&cm->parts, sizeof(double) * NRED * NPART));
  PGB_CHECK(cudaMemset(cm->parts, 0, sizeof(double) * NRED * NPART));
  PGB_CHECK(cudaMalloc((void**)&cm->gsum, sizeof(double) * 2 * NRED));
  PGB_CHECK(cudaMemset(cm->gsum, 0, sizeof(double) * 2 * NRED));
  cm->peer[rank] = cm->base;
  if (world > 1) PGB_CHECK(cudaIpcGetMemHandle(&cm->handle, cm->base));
  PGB_CHECK(cudaDeviceSynchronize());
  *comm_out = cm;
  return 0;
}

extern "C" int pgb_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int pgb_comm_get_handle(void* comm, void* out_host) {
  Comm* cm = (Comm*)comm;
  memcpy(out_host, &cm->handle, sizeof(cudaIpcMemHandle_t));
  return 0;
}

/* all_handles_host: world handles in rank order (this rank's own entry is ignored) */
extern "C" int pgb_comm_connect(void* comm, const void* all_handles_host) {
  Comm* cm = (Comm*)comm;
  for (int q = 0; q < cm->world; ++q) {
    if (q == cm->rank) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)all_handles_host + (size_t)q * sizeof(hd), sizeof(hd));
    void* p = nullptr;
    PGB_CHECK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    cm->peer[q] = (char*)p;
    cm->opened[q] = true;
  }
  return 0;
}

extern "C" void* pgb_comm_arena(void* comm) { return ((Comm*)comm)->base + HEADER_BYTES; }
extern "C" int64_t pgb_comm_arena_bytes(void* comm) { return (int64_t)((Comm*)comm)->arena_bytes; }

extern "C" int pgb_comm_destroy(void* comm) {
  Comm* cm = (Comm*)comm;
  if (!cm) return 0;
  cudaDeviceSynchronize();
  for (int q = 0; q < cm->world; ++q)
    if (cm->opened[q]) cudaIpcCloseMemHandle(cm->peer[q]);
  cudaFree(cm->base);
  cudaFree(cm->tickets);
  cudaFree(cm->scal);
  cudaFree(cm->parts);
  cudaFree(cm->gsum);
  if (cm->resid) cudaFree(cm->resid);
  cudaFreeHost(cm->flag_host);
  delete cm;
  return 0;
}

namespace {
__global__ void kd_barrier(CommDev c, unsigned long long seq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  __threadfence_system();
  for (int q = 0; q < c.world; ++q) st_release_sys(&c.hdr[q]->bar_flag[c.rank], seq);
  Header* mine = c.hdr[c.rank];
  for (int q = 0; q < c.world; ++q) wait_flag(&mine->bar_flag[q], seq, mine);
}
}  // namespace

/* Device-side barrier over the peer flags (all ranks must call it the same number of times). */
extern "C" int pgb_comm_barrier(void* comm, void* stream) {
  Comm* cm = (Comm*)comm;
  CommDev c;
  comm_dev(cm, &c);
  kd_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(c, ++cm->bar_epoch);
  PGB_LAUNCH_CHECK();
  return check_err(cm, (cudaStream_t)stream, "pgb_comm_barrier");
}

// ---------------------------------------------------------------------------------------------------------
// distributed CG.  x0 = 0.  The two search-direction buffers live in the arena at byte offsets 0 and
// arena_bytes / 2 (the same on every rank), each [n_owned | pad | ghosts] with the ghost block starting at
// ghost_start (a multiple of 16 doubles: no 128-byte line holds owned and ghost entries together).
// ---------------------------------------------------------------------------------------------------------
extern "C" int pgb_dist_cg(void* comm, int64_t n_owned, int64_t ghost_start, int64_t n_ghost, const void* indptr,
                           int idx64, const int32_t* indices, const double* data, const double* b,
                           const double* dinv, double* x, double* work /* 2 * n_owned */, int nsend,
                           const int* send_peer_host, const int64_t* send_off_host, const int64_t* send_dst_host,
                           const int32_t* send_idx, int nrecv, const int* recv_peer_host, int64_t interior_lo,
                           int64_t interior_hi, int threads_per_row, double rtol, double atol, int maxiter,
                           int* iters_host, int* info_host, double* resid_host, void* stream) {
  Comm* cm = (Comm*)comm;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = n_owned, n_local = ghost_start + n_ghost;
  PGB_REQUIRE(n > 0 && maxiter >= 0, "pgb_dist_cg: bad arguments");
  PGB_REQUIRE(ghost_start >= n && ghost_start % 16 == 0, "pgb_dist_cg: ghost block must start at a multiple of 16");
  const size_t half = cm->arena_bytes / 2 / 256 * 256;
  PGB_REQUIRE((size_t)n_local * sizeof(double) <= half, "pgb_dist_cg: arena too small for two vectors");
  PGB_REQUIRE(interior_lo >= 0 && interior_lo <= interior_hi && interior_hi <= n, "pgb_dist_cg: bad interior range");
  if (ensure_resid(cm, maxiter)) return 1;
  CommDev c;
  comm_dev(cm, &c);
  Halo h;
  if (int rc = make_halo(cm, nsend, send_peer_host, send_off_host, send_dst_host, send_idx, nrecv, recv_peer_host, 0,
                         half, &h))
    return rc;
  double* pbuf[2] = {(double*)(cm->base + HEADER_BYTES), (double*)(cm->base + HEADER_BYTES + half)};
  double* r = work;
  double* q = work + n;
  CgScal* sc = (CgScal*)cm->scal;
  int* done = (int*)((char*)cm->scal + 1024);
  CgScal hs = {};
  hs.atol_user = atol;
  hs.rtol = rtol;
  hs.resid_cap = cm->resid_cap;
  *cm->flag_host = 0;
  PGB_CHECK(cudaMemcpyAsync(sc, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
  PGB_CHECK(cudaMemsetAsync(done, 0, sizeof(int), s));
  const unsigned long long base = cm->epoch;
  cm->epoch += 2ull * ((unsigned long long)maxiter + 8ull);
  const int tpr = threads_per_row > 0 ? threads_per_row : 4;
  if (iters_host) *iters_host = 0;
  if (info_host) *info_host = 0;

  kd_cg_init<<<VGRID(kd_cg_init), KB, 0, s>>>(n, b, dinv, r, x, cm->parts, c, base + 1, cm->tickets);
  PGB_LAUNCH_CHECK();

  DSpmv a = {};
  a.n = n; a.ia = interior_lo; a.ib = interior_hi; a.gs = (int32_t)ghost_start;
  a.indptr = indptr; a.indices = indices; a.data = data;
  a.y = q; a.parts = cm->parts + S_PQ * NPART; a.done = done;
  int launched = 0;
  for (int k = 0; k < maxiter; ++k) {
    if (*(volatile int*)cm->flag_host) break;
    const unsigned long long sq = base + 1 + (unsigned long long)k;  // rr/rho of iteration k; pq uses sq too
    double* p_new = pbuf[k & 1];
    const double* p_old = pbuf[(k + 1) & 1];
    kd_cg_p<<<VGRID(kd_cg_p), KB, 0, s>>>(n, k, r, dinv, p_old, p_new, sc, done, cm->resid, cm->flag_dev, c, h, sq, sq);
    a.x = p_new;
    int rc = idx64 ? launch_dspmv<int64_t, SPMV_CG>(a, c, h, sq, S_PQ, sq, cm->tickets + T_B, tpr, s)
                   : launch_dspmv<int32_t, SPMV_CG>(a, c, h, sq, S_PQ, sq, cm->tickets + T_B, tpr, s);
    if (rc) return rc;
    kd_cg_xr<<<VGRID(kd_cg_xr), KB, 0, s>>>(n, p_new, q, dinv, x, r, cm->parts, sc, done, c, sq, sq + 1, cm->tickets);
    PGB_LAUNCH_CHECK();
    ++launched;
  }
  if (launched == maxiter)  // ||r|| after the last update, for the history (no stopping test: scipy has none there)
    kd_cg_final<<<1, KB, 0, s>>>(sc, done, cm->resid, c, base + 1 + (unsigned long long)maxiter);
  PGB_LAUNCH_CHECK();
  PGB_CHECK(cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (int rc = check_err(cm, s, "pgb_dist_cg")) return rc;
  const int iters = hs.iters;
  if (iters_host) *iters_host = iters;
  if (info_host) *info_host = hs.done ? 0 : maxiter;
  if (resid_host) {
    int cnt = iters + 1;
    if (cnt > cm->resid_cap) cnt = cm->resid_cap;
    PGB_CHECK(cudaMemcpyAsync(resid_host, cm->resid, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// distributed MINRES (Paige-Saunders as in scipy, x0 = 0).  v lives in the arena at offset 0:
// [n_owned | pad | ghosts].  work: 6 * n_owned doubles (r1, r2, y rotate; w, w2; spare).
// ---------------------------------------------------------------------------------------------------------
extern "C" int pgb_dist_minres(void* comm, int64_t n_owned, int64_t ghost_start, int64_t n_ghost, const void* indptr,
                               int idx64, const int32_t* indices, const double* data, const double* b,
                               const double* dinv, double* x, double* work /* 5 * n_owned */, int nsend,
                               const int* send_peer_host, const int64_t* send_off_host,
                               const int64_t* send_dst_host, const int32_t* send_idx, int nrecv,
                               const int* recv_peer_host, int64_t interior_lo, int64_t interior_hi,
                               int threads_per_row, double shift, double rtol, int maxiter, int* iters_host,
                               int* info_host, double* resid_host, void* stream) {
  Comm* cm = (Comm*)comm;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = n_owned, n_local = ghost_start + n_ghost;
  PGB_REQUIRE(n > 0 && maxiter >= 0, "pgb_dist_minres: bad arguments");
  PGB_REQUIRE(ghost_start >= n && ghost_start % 16 == 0, "pgb_dist_minres: ghost block must start at a multiple of 16");
  PGB_REQUIRE((size_t)n_local * sizeof(double) <= cm->arena_bytes, "pgb_dist_minres: arena too small");
  PGB_REQUIRE(interior_lo >= 0 && interior_lo <= interior_hi && interior_hi <= n, "pgb_dist_minres: bad interior range");
  if (ensure_resid(cm, maxiter)) return 1;
  CommDev c;
  comm_dev(cm, &c);
  Halo h;
  if (int rc = make_halo(cm, nsend, send_peer_host, send_off_host, send_dst_host, send_idx, nrecv, recv_peer_host, 0,
                         0, &h))
    return rc;
  double* v = (double*)(cm->base + HEADER_BYTES);
  double* rb[3] = {work, work + n, work + 2 * n};
  double* wb[2] = {work + 3 * n, work + 4 * n};
  MinresScal* sc = (MinresScal*)cm->scal;  // sc[0], sc[1]: double-buffered state; sc[2]: final
  int* done = (int*)((char*)cm->scal + 2048);
  double* out2 = (double*)((char*)cm->scal + 2048 + 64);
  *cm->flag_host = 0;
  PGB_CHECK(cudaMemsetAsync(done, 0, sizeof(int), s));
  const unsigned long long base = cm->epoch;
  cm->epoch += 2ull * ((unsigned long long)maxiter + 8ull);
  const int tpr = threads_per_row > 0 ? threads_per_row : 4;
  if (iters_host) *iters_host = 0;
  if (info_host) *info_host = 0;

  kd_mr_init<<<VGRID(kd_mr_init), KB, 0, s>>>(n, b, dinv, rb[0], x, wb[0], wb[1], cm->parts, c, base + 1, cm->tickets);
  kd_mr_start<<<VGRID(kd_mr_start), KB, 0, s>>>(n, rb[0], dinv, v, sc, done, out2, rtol, shift, maxiter, cm->resid_cap, c, h, base + 1,
                                base + 1);
  PGB_LAUNCH_CHECK();
  double hb[2];
  PGB_CHECK(cudaMemcpyAsync(hb, out2, sizeof(hb), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (hb[0] == 0.0 || hb[1] == 0.0) {  // x0 = 0 solves the system / b == 0 (scipy returns b)
    return check_err(cm, s, "pgb_dist_minres");
  }
  double* r1 = rb[0];
  double* r2 = rb[0];
  double* y = rb[1];
  double* spare = rb[2];
  double *w_new = wb[0], *w_old = wb[1];
  DSpmv a = {};
  a.n = n; a.ia = interior_lo; a.ib = interior_hi; a.gs = (int32_t)ghost_start;
  a.indptr = indptr; a.indices = indices; a.data = data;
  a.parts = cm->parts + M_ALFA * NPART; a.done = done;
  int itn = 0;
  while (itn < maxiter) {
    if (*(volatile int*)cm->flag_host) break;
    ++itn;
    const unsigned long long sq = base + 1 + (unsigned long long)itn;  // all sums of iteration itn
    const MinresScal* s_old = sc + ((itn - 1) & 1);
    MinresScal* s_new = sc + (itn & 1);
    a.x = v; a.y = y; a.r1 = r1; a.itn = itn; a.mr = s_old;
    // halo of v for iteration itn was pushed with sequence number base + itn (start: base + 1)
    int rc = idx64 ? launch_dspmv<int64_t, SPMV_MINRES>(a, c, h, base + (unsigned long long)itn, M_ALFA, sq,
                                                        cm->tickets + T_A, tpr, s)
                   : launch_dspmv<int32_t, SPMV_MINRES>(a, c, h, base + (unsigned long long)itn, M_ALFA, sq,
                                                        cm->tickets + T_A, tpr, s);
    if (rc) return rc;
    kd_mr_c<<<VGRID(kd_mr_c), KB, 0, s>>>(n, y, r2, dinv, s_old, done, cm->parts, c, sq, sq, cm->tickets);
    double* old_r1 = r1;
    r1 = r2;
    r2 = y;
    y = (old_r1 == r1) ? spare : old_r1;
    kd_mr_d<<<VGRID(kd_mr_d), KB, 0, s>>>(n, itn, v, w_old, w_new, r2, dinv, x, s_old, s_new, done, cm->resid, cm->flag_dev,
                              cm->parts, c, h, sq, sq, sq - 1, sq, base + (unsigned long long)itn + 1, cm->tickets);
    PGB_LAUNCH_CHECK();
    double* t = w_old;
    w_old = w_new;
    w_new = t;
  }
  // final test on the state of the last launched iteration
  kd_mr_final<<<1, KB, 0, s>>>(itn + 1, sc + (itn & 1), sc + 2, done, c, base + 1 + (unsigned long long)itn);
  PGB_LAUNCH_CHECK();
  MinresScal hh;
  PGB_CHECK(cudaMemcpyAsync(&hh, sc + 2, sizeof(hh), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (int rc = check_err(cm, s, "pgb_dist_minres")) return rc;
  if (iters_host) *iters_host = hh.itn;
  if (info_host) *info_host = (hh.istop == 6) ? maxiter : 0;
  if (resid_host && hh.itn > 0) {
    const int cnt = hh.itn < cm->resid_cap - 1 ? hh.itn : cm->resid_cap - 1;
    PGB_CHECK(cudaMemcpyAsync(resid_host, cm->resid + 1, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}
