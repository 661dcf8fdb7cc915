// K3: CSR SpMV (sub-warp per row, persistent grid) and the fused vector steps of CG / MINRES.
//
// The recurrences are those of scipy.sparse.linalg.cg and .minres (scipy 1.18.1,
// _isolve/iterative.py and _isolve/minres.py), which are the oracle for this part: the reference
// itself only exposes the `solver(A, b)` plug-in point (numerics/linear_system.py:79-94).
//
// Reductions are two-stage and atomic-free: every kernel that produces a dot product writes one
// partial per block into a fixed-size array (PGB_NPART slots, unused slots zeroed), and every
// consumer kernel re-reduces that array in a fixed order.  Scalars never leave the device during
// the iteration; the host polls a mapped pinned flag to stop launching.
#include "pgb_krylov.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// generic vector helpers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KB) k_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                            double* parts) {
  __shared__ double ws[KB / 32];
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) s += x[i] * y[i];
  write_partial(parts, block_sum(s, ws));
}

__global__ void __launch_bounds__(KB) k_reduce(const double* parts, int count, double* out) {
  __shared__ double ws[KB / 32];
  double s = 0;
  for (int i = threadIdx.x; i < count; i += KB) s += parts[i];
  s = block_sum(s, ws);
  if (threadIdx.x == 0) *out = s;
}

__global__ void __launch_bounds__(KB) k_axpby(int64_t n, double a, const double* __restrict__ x, double b,
                                              double* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}


// ------------------------------------------------------------------------------------------
// CG
// ------------------------------------------------------------------------------------------
// r = b - q ; partials rr, rho = r.z (z = dinv*r), bb
__global__ void __launch_bounds__(KB) k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
                                                const double* __restrict__ dinv, double* __restrict__ r,
                                                double* p_rr, double* p_rho, double* p_bb) {
  __shared__ double ws[KB / 32];
  double rr = 0, rho = 0, bb = 0;
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
    double bi = b[i];
    double ri = bi - q[i];
    r[i] = ri;
    rr += ri * ri;
    rho += ri * (dinv ? dinv[i] * ri : ri);
    bb += bi * bi;
  }
  write_partial(p_rr, block_sum(rr, ws));
  write_partial(p_rho, block_sum(rho, ws));
  write_partial(p_bb, block_sum(bb, ws));
}

// top of iteration k: convergence check, then p = z + beta p
__global__ void __launch_bounds__(KB) k_cg_p(int64_t n, int k, const double* __restrict__ r,
                                             const double* __restrict__ dinv, double* __restrict__ p,
                                             const double* p_rr, const double* p_rho, const double* p_bb,
                                             CgScal* sc, double* resid, volatile int* host_flag) {
  __shared__ double ws[3 * (KB / 32)];
  if (sc->done) return;
  const double* const parts[3] = {p_rr, p_rho, p_bb};
  double red[3];
  sum_partials_n<3>(parts, red, ws);
  const double rr = red[0], rho = red[1], bb = red[2];
  double rnorm = sqrt(rr);
  double atol = fmax(sc->atol_user, sc->rtol * sqrt(bb));
  bool done = rnorm < atol;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (k < sc->resid_cap) resid[k] = rnorm;
    sc->bnorm2 = bb;
    if (done) {
      sc->done = 1;
      *host_flag = 1;
      __threadfence_system();
    } else {
      sc->rho = rho;
      sc->iters = k + 1;
    }
  }
  if (done) return;
  double beta = (k > 0) ? rho / sc->rho_prev : 0.0;
  const int64_t st = (int64_t)gridDim.x * KB;
  for (int64_t i0 = (int64_t)blockIdx.x * KB + threadIdx.x; i0 < n; i0 += VU * st) {
    double rv[VU], pv[VU], dv[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {  // all loads of VU grid strides first (memory-level parallelism)
      const int64_t i = i0 + u * st;
      const bool ok = i < n;
      rv[u] = ok ? r[i] : 0.0;
      pv[u] = (ok && k > 0) ? p[i] : 0.0;
      dv[u] = (ok && dinv) ? dinv[i] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = i0 + u * st;
      if (i < n) {
        double z = dinv ? dv[u] * rv[u] : rv[u];
        p[i] = (k > 0) ? z + beta * pv[u] : z;
      }
    }
  }
}

// alpha = rho / (p.q); x += alpha p; r -= alpha q; partials rr, rho
__global__ void __launch_bounds__(KB) k_cg_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ q,
                                              const double* __restrict__ dinv, double* __restrict__ x,
                                              double* __restrict__ r, const double* p_pq, double* p_rr,
                                              double* p_rho, CgScal* sc) {
  __shared__ double ws[KB / 32];
  __shared__ double ws2[2 * (KB / 32)];
  if (sc->done) return;
  double pq = sum_partials(p_pq, ws);
  double rho_cur = sc->rho;
  double alpha = rho_cur / pq;
  double rr = 0, rho = 0;
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
    for (int u = 0; u < VU; ++u) {  // same per-thread summation order as a plain grid-stride loop
      const int64_t i = i0 + u * st;
      if (i < n) {
        x[i] = xv[u] + alpha * pv[u];
        double ri = rv[u] - alpha * qv[u];
        r[i] = ri;
        rr += ri * ri;
        rho += ri * (dinv ? dv[u] * ri : ri);
      }
    }
  }
  double red[2] = {rr, rho};
  block_sum_n<2>(red, ws2);
  write_partial(p_rr, red[0]);
  write_partial(p_rho, red[1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->rho_prev = rho_cur;
}

// ------------------------------------------------------------------------------------------
// MINRES
// ------------------------------------------------------------------------------------------
// r1 = b - q ; y = psolve(r1); partials beta1^2 = r1.y, bb
__global__ void __launch_bounds__(KB) k_mr_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
                                                const double* __restrict__ dinv, double* __restrict__ r1,
                                                double* p_b1, double* p_bb) {
  __shared__ double ws[KB / 32];
  double b1 = 0, bb = 0;
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
    double bi = b[i];
    double ri = bi - q[i];
    r1[i] = ri;
    b1 += ri * (dinv ? dinv[i] * ri : ri);
    bb += bi * bi;
  }
  write_partial(p_b1, block_sum(b1, ws));
  write_partial(p_bb, block_sum(bb, ws));
}

// v = psolve(r) * s ; w = w2 = 0
__global__ void __launch_bounds__(KB) k_mr_start(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
                                                 const MinresScal* sc, double* __restrict__ v, double* __restrict__ w,
                                                 double* __restrict__ w2) {
  double s = 1.0 / sc->beta;
  for (int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x; i < n; i += (int64_t)gridDim.x * KB) {
    v[i] = (dinv ? dinv[i] * r[i] : r[i]) * s;
    w[i] = 0.0;
    w2[i] = 0.0;
  }
}

// y -= (alfa/beta) r2  (y becomes the new r2); partial beta^2 = r2new . psolve(r2new)
__global__ void __launch_bounds__(KB) k_mr_c(int64_t n, double* __restrict__ y, const double* __restrict__ r2,
                                             const double* __restrict__ dinv, const double* p_alfa, double* p_beta,
                                             MinresScal* sc) {
  __shared__ double ws[KB / 32];
  if (sc->done) return;
  double alfa = sum_partials(p_alfa, ws);
  double c = alfa / sc->beta;
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
        double yi = yv[u] - c * r2v[u];
        y[i] = yi;
        bt += yi * (dinv ? dv[u] * yi : yi);
      }
    }
  }
  write_partial(p_beta, block_sum(bt, ws));
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->alfa = alfa;
}


// one thread: (i) stopping tests of the previous iteration (needs ||x|| from its x-update),
// (ii) Lanczos / rotation scalars of the current one.
__global__ void k_mr_scalars(int itn, const double* p_beta, const double* p_xx, MinresScal* sc, double* resid,
                             volatile int* host_flag, int test_only) {
  __shared__ double ws[KB / 32];
  if (sc->done) return;
  double xx = sum_partials(p_xx, ws);
  double bt = sum_partials(p_beta, ws);
  if (threadIdx.x != 0) return;
  if (mr_scalar_step(sc, itn, bt, xx, test_only)) {
    sc->done = 1;
    *host_flag = 1;
    __threadfence_system();
    return;
  }
  if (!test_only && itn < sc->resid_cap) resid[itn] = sc->phibar;
}

// w = (v - oldeps w1 - delta w2) denom, stored over w1 (= w_{k-2}, dead afterwards); x += phi w;
// v = psolve(r2)/beta; partial x.x
__global__ void __launch_bounds__(KB) k_mr_d(int64_t n, double* __restrict__ v, double* __restrict__ w1,
                                             const double* __restrict__ w2, const double* __restrict__ r2,
                                             const double* __restrict__ dinv, double* __restrict__ x, double* p_xx,
                                             const MinresScal* sc) {
  __shared__ double ws[KB / 32];
  if (sc->done) return;
  double oldeps = sc->oldeps, delta = sc->delta, denom = sc->denom, phi = sc->phi, ib = sc->inv_beta;
  double xx = 0;
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
        double wn = (vv[u] - oldeps * w1v[u] - delta * w2v[u]) * denom;
        w1[i] = wn;
        double xi = xv[u] + phi * wn;
        x[i] = xi;
        xx += xi * xi;
        v[i] = (dinv ? dv[u] * r2v[u] : r2v[u]) * ib;
      }
    }
  }
  write_partial(p_xx, block_sum(xx, ws));
}

__global__ void k_mr_setup(MinresScal* sc, const double* p_b1, const double* p_bb, double rtol, double shift,
                           int maxiter, int resid_cap, double* out2) {
  __shared__ double ws[KB / 32];
  double b1 = sum_partials(p_b1, ws);
  double bb = sum_partials(p_bb, ws);
  if (threadIdx.x != 0) return;
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
  *sc = s;
  out2[0] = b1;
  out2[1] = bb;
}

// small persistent scratch shared by the drivers, one per (host thread, device)
struct Scratch {
  int* flag_host = nullptr;  // mapped pinned
  int* flag_dev = nullptr;
  void* scal = nullptr;  // device, 1 KB
  double* parts = nullptr;  // device, 8 * NPART doubles
  double* resid = nullptr;  // device: hist_cap history entries + 2 spare slots
  int resid_cap = 0;
};
constexpr int MAX_DEV = 64;
thread_local Scratch g_scr_dev[MAX_DEV];

// number of residual-history entries kept for a solve of at most `maxiter` iterations
inline int hist_cap(int maxiter) {
  int64_t want = (int64_t)maxiter + 2;
  return (int)(want < PGB_RESID_CAP ? want : PGB_RESID_CAP);
}

int ensure_scratch(int maxiter, Scratch** out) {
  int dev = 0;
  PGB_CHECK(cudaGetDevice(&dev));
  PGB_REQUIRE(dev >= 0 && dev < MAX_DEV, "device ordinal out of range");
  Scratch& s = g_scr_dev[dev];
  if (!s.flag_host) {
    PGB_CHECK(cudaHostAlloc((void**)&s.flag_host, 64, cudaHostAllocMapped));
    PGB_CHECK(cudaHostGetDevicePointer((void**)&s.flag_dev, s.flag_host, 0));
    PGB_CHECK(cudaMalloc(&s.scal, 1024));
    PGB_CHECK(cudaMalloc((void**)&s.parts, sizeof(double) * 8 * NPART));
  }
  const int cap = hist_cap(maxiter);
  if (s.resid_cap < cap) {  // bounded by PGB_RESID_CAP (8 MB), whatever maxiter is
    if (s.resid) cudaFree(s.resid);
    s.resid = nullptr;
    s.resid_cap = 0;
    PGB_CHECK(cudaMalloc((void**)&s.resid, sizeof(double) * (size_t)(cap + 2)));
    s.resid_cap = cap;
  }
  *out = &s;
  return 0;
}

double avg_row(int64_t n, const void* indptr, int idx64, cudaStream_t s) {
  int64_t nnz = 0;
  if (idx64) {
    cudaMemcpyAsync(&nnz, (const int64_t*)indptr + n, 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  } else {
    int32_t t = 0;
    cudaMemcpyAsync(&t, (const int32_t*)indptr + n, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    nnz = t;
  }
  return n > 0 ? (double)nnz / (double)n : 0.0;
}

// ---- step API helpers (multi-GPU drivers): move partial sums to / from a contiguous buffer ------
// red[i] = sum(parts_i) for the arrays selected by `mask`
__global__ void __launch_bounds__(KB) k_parts_to_red(double* scratch, int mask, int narr) {
  __shared__ double ws[KB / 32];
  for (int i = 0; i < narr; ++i) {
    if (!((mask >> i) & 1)) continue;
    double s = sum_partials(scratch + (size_t)i * NPART, ws);
    if (threadIdx.x == 0) scratch[(size_t)narr * NPART + i] = s;
  }
}
// parts_i = (red[i], 0, 0, ...) so that the fused kernels see the globally reduced value
__global__ void __launch_bounds__(KB) k_red_to_parts(double* scratch, int mask, int narr) {
  for (int i = 0; i < narr; ++i) {
    if (!((mask >> i) & 1)) continue;
    double* p = scratch + (size_t)i * NPART;
    double v = scratch[(size_t)narr * NPART + i];
    for (int j = threadIdx.x; j < NPART; j += KB) p[j] = (j == 0) ? v : 0.0;
  }
}

struct StepScratch {
  double *p_rr, *p_rho, *p_bb, *p_pq, *red;
  CgScal* sc;
  int* flag;
  double* resid;
};
__host__ StepScratch step_scratch(double* scratch) {
  StepScratch t;
  t.p_rr = scratch;
  t.p_rho = scratch + NPART;
  t.p_bb = scratch + 2 * NPART;
  t.p_pq = scratch + 3 * NPART;
  t.red = scratch + 4 * NPART;                       // 8 doubles
  t.flag = (int*)(scratch + 4 * NPART + 8);          // 1 double slot: done flag (int)
  t.sc = (CgScal*)(scratch + 4 * NPART + 16);        // 16 doubles reserved
  t.resid = scratch + 4 * NPART + 32;                // residual history follows
  return t;
}

}  // namespace

extern "C" int pgb_spmv(int64_t n_rows, const void* indptr, int idx64, const int32_t* indices,
                        const double* data, const double* x, double* y, int threads_per_row,
                        void* stream) {
  if (n_rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  SpmvArgs a = {};
  a.n = n_rows; a.indptr = indptr; a.indices = indices; a.data = data; a.x = x; a.y = y;
  int tpr = threads_per_row > 0 ? threads_per_row : pick_tpr(avg_row(n_rows, indptr, idx64, s));
  return launch_spmv<SPMV_PLAIN>(a, idx64, tpr, s);
}

extern "C" int pgb_dot_partials(int64_t n, const double* x, const double* y, double* partials, void* stream) {
  k_dot<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, x, y, partials);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_reduce_partials(const double* partials, int count, double* out, void* stream) {
  k_reduce<<<1, KB, 0, (cudaStream_t)stream>>>(partials, count, out);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream) {
  if (n == 0) return 0;
  k_axpby<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, a, x, b, y);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_cg(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                      const double* data, const double* b, double* x, const double* dinv, double rtol,
                      double atol, int maxiter, double* work, int* iters_host, int* info_host,
                      double* resid_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(n > 0 && maxiter >= 0, "pgb_cg: bad arguments");
  Scratch* Sp = nullptr;
  if (int rc = ensure_scratch(maxiter, &Sp)) return rc;
  Scratch& S = *Sp;
  const int cap = hist_cap(maxiter);
  double* r = work;
  double* p = work + n;
  double* q = work + 2 * n;
  double *p_rr = S.parts, *p_rho = S.parts + NPART, *p_bb = S.parts + 2 * NPART, *p_pq = S.parts + 3 * NPART;
  CgScal* sc = (CgScal*)S.scal;
  CgScal h = {};
  h.atol_user = atol;
  h.rtol = rtol;
  h.resid_cap = cap;
  *S.flag_host = 0;
  PGB_CHECK(cudaMemcpyAsync(sc, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  const int tpr = pick_tpr(avg_row(n, indptr, idx64, s));
  const int vg = vec_grid(n);
  if (iters_host) *iters_host = 0;
  if (info_host) *info_host = 0;

  // scipy returns b (= 0) at once for a zero right-hand side; without this test alpha = 0/0 and no
  // later `rnorm < atol` can ever be true
  double* bb_dev = S.resid + cap + 1;
  k_dot<<<vg, KB, 0, s>>>(n, b, b, p_bb);
  k_reduce<<<1, KB, 0, s>>>(p_bb, NPART, bb_dev);
  PGB_LAUNCH_CHECK();
  double bb_host = 1.0;
  PGB_CHECK(cudaMemcpyAsync(&bb_host, bb_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (bb_host == 0.0) {
    PGB_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * n, s));
    PGB_CHECK(cudaStreamSynchronize(s));
    return 0;
  }

  SpmvArgs a = {};
  a.n = n; a.indptr = indptr; a.indices = indices; a.data = data;
  a.x = x; a.y = q;
  if (launch_spmv<SPMV_PLAIN>(a, idx64, tpr, s)) return 1;
  k_cg_init<<<resident_grid(k_cg_init, pgb_cdiv(n, KB)), KB, 0, s>>>(n, b, q, dinv, r, p_rr, p_rho, p_bb);
  PGB_LAUNCH_CHECK();
  a.x = p; a.y = q; a.parts = p_pq; a.cg = sc;
  const int g_p = resident_grid(k_cg_p, pgb_cdiv(n, KB)), g_xr = resident_grid(k_cg_xr, pgb_cdiv(n, KB));
  int launched = 0;
  for (int k = 0; k <= maxiter; ++k) {
    if (*(volatile int*)S.flag_host) break;
    if (k == maxiter) break;
    k_cg_p<<<g_p, KB, 0, s>>>(n, k, r, dinv, p, p_rr, p_rho, p_bb, sc, S.resid, S.flag_dev);
    if (launch_spmv<SPMV_CG>(a, idx64, tpr, s)) return 1;
    k_cg_xr<<<g_xr, KB, 0, s>>>(n, p, q, dinv, x, r, p_pq, p_rr, p_rho, sc);
    PGB_LAUNCH_CHECK();
    ++launched;
  }
  // final residual (also catches convergence exactly at maxiter for the history)
  double* rr_dev = S.resid + cap;
  if (launched == maxiter) {
    k_reduce<<<1, KB, 0, s>>>(p_rr, NPART, rr_dev);
    PGB_LAUNCH_CHECK();
  }
  PGB_CHECK(cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  const int iters = h.iters;
  if (iters_host) *iters_host = iters;
  if (info_host) *info_host = h.done ? 0 : maxiter;
  if (resid_host) {  // ||r_k|| for k = 0..iters, truncated to the first PGB_RESID_CAP entries
    int cnt = h.done ? iters + 1 : iters;
    if (cnt > cap) cnt = cap;
    if (cnt > 0) PGB_CHECK(cudaMemcpyAsync(resid_host, S.resid, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    if (!h.done && launched == maxiter && iters < cap) {
      double rr = 0;
      PGB_CHECK(cudaMemcpyAsync(&rr, rr_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
      PGB_CHECK(cudaStreamSynchronize(s));
      resid_host[iters] = sqrt(rr);
    }
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

extern "C" int pgb_minres(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                          const double* data, const double* b, double* x, const double* dinv, double shift,
                          double rtol, int maxiter, double* work, int* iters_host, int* info_host,
                          double* resid_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  PGB_REQUIRE(n > 0 && maxiter >= 0, "pgb_minres: bad arguments");
  Scratch* Sp = nullptr;
  if (int rc = ensure_scratch(maxiter, &Sp)) return rc;
  Scratch& S = *Sp;
  const int cap = hist_cap(maxiter);
  double* v = work;
  double* rb[3] = {work + n, work + 2 * n, work + 3 * n};  // r1, r2, y rotate
  double* wb[2] = {work + 4 * n, work + 5 * n};
  double *p_alfa = S.parts, *p_beta = S.parts + NPART, *p_xx = S.parts + 2 * NPART, *p_b1 = S.parts + 3 * NPART,
         *p_bb = S.parts + 4 * NPART;
  double* out2 = S.parts + 5 * NPART;
  MinresScal* sc = (MinresScal*)S.scal;
  *S.flag_host = 0;
  const int tpr = pick_tpr(avg_row(n, indptr, idx64, s));
  const int vg = vec_grid(n);

  SpmvArgs a = {};
  a.n = n; a.indptr = indptr; a.indices = indices; a.data = data;
  a.x = x; a.y = rb[2];
  if (launch_spmv<SPMV_PLAIN>(a, idx64, tpr, s)) return 1;
  k_mr_init<<<vg, KB, 0, s>>>(n, b, rb[2], dinv, rb[0], p_b1, p_bb);
  k_mr_setup<<<1, KB, 0, s>>>(sc, p_b1, p_bb, rtol, shift, maxiter, cap, out2);
  PGB_LAUNCH_CHECK();
  double hb[2];
  PGB_CHECK(cudaMemcpyAsync(hb, out2, sizeof(hb), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (iters_host) *iters_host = 0;
  if (info_host) *info_host = 0;
  if (hb[0] == 0.0) return 0;  // beta1 == 0: x0 solves the system
  if (hb[1] == 0.0) {          // b == 0: scipy returns b
    PGB_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * n, s));
    return 0;
  }
  // r2 = r1 (same buffer until the first rotation: scipy aliases them)
  double* r1 = rb[0];
  double* r2 = rb[0];
  double* y = rb[1];
  double* spare = rb[2];
  double *w_new = wb[0], *w_old = wb[1];  // w_{k-1}, w_{k-2}
  k_mr_start<<<vg, KB, 0, s>>>(n, r1, dinv, sc, v, w_new, w_old);
  k_dot<<<vg, KB, 0, s>>>(n, x, x, p_xx);
  PGB_LAUNCH_CHECK();
  a.parts = p_alfa; a.mr = sc;
  const int g_c = resident_grid(k_mr_c, pgb_cdiv(n, KB)), g_d = resident_grid(k_mr_d, pgb_cdiv(n, KB));
  int itn = 0;
  while (itn < maxiter) {
    if (*(volatile int*)S.flag_host) break;
    ++itn;
    a.x = v; a.y = y; a.r1 = r1; a.itn = itn;
    if (launch_spmv<SPMV_MINRES>(a, idx64, tpr, s)) return 1;
    k_mr_c<<<g_c, KB, 0, s>>>(n, y, r2, dinv, p_alfa, p_beta, sc);
    // rotate: r1 <- r2, r2 <- y, y <- a free buffer
    double* old_r1 = r1;
    r1 = r2;
    r2 = y;
    y = (old_r1 == r1) ? spare : old_r1;
    k_mr_scalars<<<1, KB, 0, s>>>(itn, p_beta, p_xx, sc, S.resid, S.flag_dev, 0);
    // scipy: w1 = w2; w2 = w; w = (v - oldeps w1 - delta w2) denom.  The new w overwrites w_{k-2}.
    k_mr_d<<<g_d, KB, 0, s>>>(n, v, w_old, w_new, r2, dinv, x, p_xx, sc);
    PGB_LAUNCH_CHECK();
    double* t = w_old;
    w_old = w_new;
    w_new = t;
  }
  k_mr_scalars<<<1, KB, 0, s>>>(itn + 1, p_beta, p_xx, sc, S.resid, S.flag_dev, 1);
  PGB_LAUNCH_CHECK();
  MinresScal h;
  PGB_CHECK(cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  if (iters_host) *iters_host = h.itn;
  if (info_host) *info_host = (h.istop == 6) ? maxiter : 0;
  if (resid_host && h.itn > 0) {  // phibar of iterations 1.., truncated to PGB_RESID_CAP - 1 entries
    const int cnt = h.itn < cap - 1 ? h.itn : cap - 1;
    PGB_CHECK(cudaMemcpyAsync(resid_host, S.resid + 1, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    PGB_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CG step API: the same fused kernels as pgb_cg, launched one by one so that a multi-GPU driver
// can put the halo exchange (before the SpMV) and the all-reduce of the partial sums (after the
// kernels that produce them) in between.  All scalars stay on the device.
// scratch layout (doubles): [4*NPART partial arrays: rr, rho, bb, pq][8: red][8: flag][16: scalars]
// [maxiter+2: residual history];  pgb_cg_scratch_doubles(maxiter) gives the size.
// ---------------------------------------------------------------------------------------------
extern "C" int64_t pgb_cg_scratch_doubles(int maxiter) { return 4 * (int64_t)NPART + 32 + hist_cap(maxiter); }

extern "C" int pgb_cg_step_init(int64_t n, const double* b, const double* q, const double* dinv, double* r,
                                double rtol, double atol, double* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  StepScratch t = step_scratch(scratch);
  CgScal h = {};
  h.atol_user = atol;
  h.rtol = rtol;
  h.resid_cap = PGB_RESID_CAP;  // the caller sized the scratch with pgb_cg_scratch_doubles(maxiter)
  PGB_CHECK(cudaMemsetAsync(t.flag, 0, 8, s));
  PGB_CHECK(cudaMemcpyAsync(t.sc, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  k_cg_init<<<vec_grid(n), KB, 0, s>>>(n, b, q, dinv, r, t.p_rr, t.p_rho, t.p_bb);
  PGB_LAUNCH_CHECK();
  return 0;
}

/* mask bits: 1 = rr, 2 = rho, 4 = bb, 8 = pq.  to_red: parts -> scratch[4*NPART + i]; from_red: back. */
extern "C" int pgb_cg_step_reduce(double* scratch, int mask, int from_red, void* stream) {
  if (from_red) k_red_to_parts<<<1, KB, 0, (cudaStream_t)stream>>>(scratch, mask, 4);
  else k_parts_to_red<<<1, KB, 0, (cudaStream_t)stream>>>(scratch, mask, 4);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_cg_step_p(int64_t n, int k, const double* r, const double* dinv, double* p, double* scratch,
                             void* stream) {
  StepScratch t = step_scratch(scratch);
  k_cg_p<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, k, r, dinv, p, t.p_rr, t.p_rho, t.p_bb, t.sc, t.resid,
                                                         t.flag);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_cg_step_spmv(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                                const double* data, const double* p, double* q, int threads_per_row,
                                double* scratch, void* stream) {
  StepScratch t = step_scratch(scratch);
  SpmvArgs a = {};
  a.n = n; a.indptr = indptr; a.indices = indices; a.data = data;
  a.x = p; a.y = q; a.parts = t.p_pq; a.cg = t.sc;
  return launch_spmv<SPMV_CG>(a, idx64, threads_per_row > 0 ? threads_per_row : 4, (cudaStream_t)stream);
}

extern "C" int pgb_cg_step_xr(int64_t n, const double* p, const double* q, const double* dinv, double* x,
                              double* r, double* scratch, void* stream) {
  StepScratch t = step_scratch(scratch);
  k_cg_xr<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, p, q, dinv, x, r, t.p_pq, t.p_rr, t.p_rho, t.sc);
  PGB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// MINRES step API (same kernels as pgb_minres; the host rotates the r / w buffers exactly as
// pgb_minres does).  scratch layout (doubles): [5*NPART partial arrays: alfa, beta, xx, b1, bb]
// [8: red][8: done flag][64: scalars][2: (beta1^2, b.b)][maxiter+2: |r| estimates (phibar)].
// mask bits for step_reduce: 1 = v.y (alfa), 2 = r2.y (beta^2), 4 = x.x, 8 = r1.y (beta1^2), 16 = b.b.
// ---------------------------------------------------------------------------------------------
namespace {
struct MrScratch {
  double *p_alfa, *p_beta, *p_xx, *p_b1, *p_bb, *red, *out2, *resid;
  int* flag;
  MinresScal* sc;
};
MrScratch mr_scratch(double* s) {
  MrScratch t;
  t.p_alfa = s;
  t.p_beta = s + NPART;
  t.p_xx = s + 2 * NPART;
  t.p_b1 = s + 3 * NPART;
  t.p_bb = s + 4 * NPART;
  t.red = s + 5 * NPART;
  t.flag = (int*)(s + 5 * NPART + 8);
  t.sc = (MinresScal*)(s + 5 * NPART + 16);
  t.out2 = s + 5 * NPART + 80;
  t.resid = s + 5 * NPART + 82;
  return t;
}
static_assert(sizeof(MinresScal) <= 64 * sizeof(double), "MinresScal must fit its scratch slot");
}  // namespace

extern "C" int64_t pgb_mr_scratch_doubles(int maxiter) { return 5 * (int64_t)NPART + 82 + hist_cap(maxiter); }

extern "C" int pgb_mr_step_reduce(double* scratch, int mask, int from_red, void* stream) {
  if (from_red) k_red_to_parts<<<1, KB, 0, (cudaStream_t)stream>>>(scratch, mask, 5);
  else k_parts_to_red<<<1, KB, 0, (cudaStream_t)stream>>>(scratch, mask, 5);
  PGB_LAUNCH_CHECK();
  return 0;
}

/* r1 = b - q (q = A x0); partials beta1^2 and b.b */
extern "C" int pgb_mr_step_init(int64_t n, const double* b, const double* q, double* r1, double* scratch,
                                void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MrScratch t = mr_scratch(scratch);
  PGB_CHECK(cudaMemsetAsync(t.flag, 0, 8, s));
  k_mr_init<<<vec_grid(n), KB, 0, s>>>(n, b, q, nullptr, r1, t.p_b1, t.p_bb);
  k_dot<<<1, KB, 0, s>>>(0, b, b, t.p_xx);  // x.x = 0 for x0 = 0 (zero-fills the partial array)
  PGB_LAUNCH_CHECK();
  return 0;
}

/* after the all-reduce of (beta1^2, b.b): scalar state + v = r1/beta1, w = w2 = 0.
 * scratch[5*NPART+80..81] = (beta1^2, b.b) for the host's early-exit tests. */
extern "C" int pgb_mr_step_start(int64_t n, const double* r1, double* v, double* w, double* w2, double rtol,
                                 double shift, int maxiter, double* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MrScratch t = mr_scratch(scratch);
  k_mr_setup<<<1, KB, 0, s>>>(t.sc, t.p_b1, t.p_bb, rtol, shift, maxiter, hist_cap(maxiter), t.out2);
  k_mr_start<<<vec_grid(n), KB, 0, s>>>(n, r1, nullptr, t.sc, v, w, w2);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_mr_step_spmv(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                                const double* data, const double* v, double* y, const double* r1, int itn,
                                int threads_per_row, double* scratch, void* stream) {
  MrScratch t = mr_scratch(scratch);
  SpmvArgs a = {};
  a.n = n; a.indptr = indptr; a.indices = indices; a.data = data;
  a.x = v; a.y = y; a.parts = t.p_alfa; a.mr = t.sc; a.r1 = r1; a.itn = itn;
  return launch_spmv<SPMV_MINRES>(a, idx64, threads_per_row > 0 ? threads_per_row : 4, (cudaStream_t)stream);
}

extern "C" int pgb_mr_step_c(int64_t n, double* y, const double* r2, double* scratch, void* stream) {
  MrScratch t = mr_scratch(scratch);
  k_mr_c<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, y, r2, nullptr, t.p_alfa, t.p_beta, t.sc);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_mr_step_scalars(int itn, int test_only, double* scratch, void* stream) {
  MrScratch t = mr_scratch(scratch);
  k_mr_scalars<<<1, KB, 0, (cudaStream_t)stream>>>(itn, t.p_beta, t.p_xx, t.sc, t.resid, t.flag, test_only);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_mr_step_d(int64_t n, double* v, double* w_old, const double* w_new, const double* r2,
                             double* x, double* scratch, void* stream) {
  MrScratch t = mr_scratch(scratch);
  k_mr_d<<<vec_grid(n), KB, 0, (cudaStream_t)stream>>>(n, v, w_old, w_new, r2, nullptr, x, t.p_xx, t.sc);
  PGB_LAUNCH_CHECK();
  return 0;
}

/* (itn, istop) of the device-side state, after a stream synchronisation */
extern "C" int pgb_mr_state(const double* scratch, int* itn_host, int* istop_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MinresScal h;
  PGB_CHECK(cudaMemcpyAsync(&h, scratch + 5 * NPART + 16, sizeof(h), cudaMemcpyDeviceToHost, s));
  PGB_CHECK(cudaStreamSynchronize(s));
  *itn_host = h.itn;
  *istop_host = h.istop;
  return 0;
}
