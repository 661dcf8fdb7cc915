// Shared device code of the Krylov drivers (single GPU: pgb_krylov.cu; NVLink peer-memory multi-GPU:
// pgb_dist.cu): block reductions, the scalar state of CG / MINRES, the sub-warp-per-row SpMV.
#pragma once
#include <math.h>
#include <map>
#include <mutex>
#include "pgb_common.cuh"

namespace {

constexpr int KB = 256;  // threads per block for all Krylov kernels
constexpr unsigned FULL = 0xffffffffu;
constexpr int NPART = PGB_NPART;
#ifndef PGB_VU
#define PGB_VU 4
#endif
constexpr int VU = PGB_VU;  // grid strides whose loads a vector kernel issues before it computes

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// fixed-order block sum, result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* ws /*[KB/32]*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) ws[w] = v;
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int i = 0; i < KB / 32; ++i) s += ws[i];
  return s;
}

__device__ __forceinline__ double sum_partials(const double* __restrict__ parts, double* ws) {
  double s = 0;
  for (int i = threadIdx.x; i < NPART; i += KB) s += parts[i];
  return block_sum(s, ws);
}

// N sums at once: the same per-value order as block_sum / sum_partials (bit-identical results), but
// one pair of barriers and overlapped loads instead of N of each.  ws: N * KB/32 doubles.
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* ws) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = warp_sum(v[j]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) ws[j * (KB / 32) + w] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < KB / 32; ++i) s += ws[j * (KB / 32) + i];
    v[j] = s;
  }
}

template <int N>
__device__ __forceinline__ void sum_partials_n(const double* const (&parts)[N], double (&out)[N], double* ws) {
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = 0;
  for (int i = threadIdx.x; i < NPART; i += KB) {
#pragma unroll
    for (int j = 0; j < N; ++j) out[j] += parts[j][i];
  }
  block_sum_n<N>(out, ws);
}

__device__ __forceinline__ void write_partial(double* parts, double v) {
  if (threadIdx.x == 0) parts[blockIdx.x] = v;
  if (blockIdx.x == 0)
    for (int i = gridDim.x + threadIdx.x; i < NPART; i += KB) parts[i] = 0.0;
}

// ------------------------------------------------------------------------------------------
// SpMV
// ------------------------------------------------------------------------------------------
struct MinresScal {
  // persistent state (scipy variable names)
  double beta1, beta, oldb, dbar, epsln, phibar, rhs1, rhs2, tnorm2, gmax, gmin, cs, sn;
  double Anorm, Acond, rnorm, root, ynorm;
  // per-iteration coefficients produced by the scalar kernel
  double alfa, oldeps, delta, denom, phi, inv_beta;
  double rtol, shift;
  int itn, istop, pending, done, maxiter, resid_cap;
};

struct CgScal {
  double rho, rho_prev, atol_user, rtol, bnorm2;
  int done, iters, resid_cap;
};

enum { SPMV_PLAIN = 0, SPMV_CG = 1, SPMV_MINRES = 2 };

struct SpmvArgs {
  int64_t n;
  const void* indptr;
  const int32_t* indices;
  const double* data;
  const double* x;
  double* y;
  double* parts;         // CG / MINRES: partial dot(x, y)
  const CgScal* cg;      // SPMV_CG
  const MinresScal* mr;  // SPMV_MINRES
  const double* r1;      // SPMV_MINRES
  int itn;               // SPMV_MINRES: 1-based iteration
};

// Grid of a grid-stride kernel with KB threads per block: as many blocks as are resident at once
// (occupancy x SMs: one wave, no tail of late blocks), at most NPART (one partial sum per block) and at
// most `work_blocks`.  The occupancy of every kernel is looked up once.
template <typename Kern>
int resident_grid(Kern kernel, int64_t work_blocks) {
  static std::mutex mu;
  static std::map<const void*, int> cache;
  int cap = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find((const void*)kernel);
    if (it != cache.end()) cap = it->second;
  }
  if (!cap) {
    int occ = 0, dev = 0, sms = PGB_SMS;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, KB, 0) != cudaSuccess || occ < 1) occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cap = occ * sms;
    if (cap > NPART) cap = NPART;
    std::lock_guard<std::mutex> lock(mu);
    cache[(const void*)kernel] = cap;
  }
  int64_t g = work_blocks < cap ? work_blocks : cap;
  return (int)(g < 1 ? 1 : g);
}

constexpr int SPMV_UR = 4;  // rows per sub-warp per outer iteration (loads of all rows issued together)

template <int TPR, typename IP, int MODE>
__global__ void __launch_bounds__(KB) k_spmv(SpmvArgs a) {
  __shared__ double ws[KB / 32];
  if (MODE == SPMV_CG && a.cg->done) return;
  if (MODE == SPMV_MINRES && a.mr->done) return;
  constexpr int RPB = KB / TPR;  // rows per block-tile
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
  double dot = 0.0;
  // row ids are int32: 32-bit row arithmetic (at 2-4 lanes per row the kernel is close to issue bound)
  const uint32_t n32 = (uint32_t)a.n;
  const uint32_t ntiles = (n32 + RPB - 1) / RPB;
  for (uint32_t t = blockIdx.x * UR; t < ntiles; t += gridDim.x * UR) {
    uint32_t row[UR];
    IP lo[UR], hi[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      row[u] = (t + u) * RPB + grp;
      bool ok = row[u] < n32;
      lo[u] = ok ? ip[row[u]] : (IP)0;
      hi[u] = ok ? ip[row[u] + 1] : (IP)0;
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
    for (int u = 0; u < UR; ++u) s[u] = (ci[u] >= 0) ? dv[u] * __ldg(a.x + ci[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < UR; ++u)
      for (IP j = lo[u] + sub + TPR; j < hi[u]; j += TPR) s[u] += data[j] * __ldg(a.x + indices[j]);
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
          if (MODE == SPMV_MINRES) {
            double xv = a.x[row[u]];
            if (shift != 0.0) sv -= shift * xv;
            if (a.itn >= 2) sv -= coef * a.r1[row[u]];
            dot += xv * sv;
          } else if (MODE == SPMV_CG) {
            dot += a.x[row[u]] * sv;
          }
          a.y[row[u]] = sv;
        }
      }
    }
  }
  if (MODE != SPMV_PLAIN) {
    double tot = block_sum(dot, ws);
    write_partial(a.parts, tot);
  }
}

template <int TPR, typename IP, int MODE>
int launch_spmv_t(const SpmvArgs& a, cudaStream_t s) {
  int64_t ntiles = ((a.n + (KB / TPR) - 1) / (KB / TPR) + SPMV_UR - 1) / SPMV_UR;
  int g = resident_grid(k_spmv<TPR, IP, MODE>, ntiles);  // one resident wave (all B200s of a node are alike)
  k_spmv<TPR, IP, MODE><<<g, KB, 0, s>>>(a);
  PGB_LAUNCH_CHECK();
  return 0;
}

template <typename IP, int MODE>
int launch_spmv_ip(const SpmvArgs& a, int tpr, cudaStream_t s) {
  switch (tpr) {
    case 2: return launch_spmv_t<2, IP, MODE>(a, s);
    case 4: return launch_spmv_t<4, IP, MODE>(a, s);
    case 8: return launch_spmv_t<8, IP, MODE>(a, s);
    case 16: return launch_spmv_t<16, IP, MODE>(a, s);
    default: return launch_spmv_t<32, IP, MODE>(a, s);
  }
}

template <int MODE>
int launch_spmv(const SpmvArgs& a, int idx64, int tpr, cudaStream_t s) {
  return idx64 ? launch_spmv_ip<int64_t, MODE>(a, tpr, s) : launch_spmv_ip<int32_t, MODE>(a, tpr, s);
}

// threads per row, tuned on B200 with 4 rows in flight per sub-warp (scripts/spmv_bench.py):
// RT0 (7 nnz/row) best at 2, P1 (14.8) at 4, Nedelec (16.2) at 4-8
int pick_tpr(double avg) {
  if (avg <= 9.0) return 2;
  if (avg <= 18.0) return 4;
  if (avg <= 40.0) return 8;
  if (avg <= 100.0) return 16;
  return 32;
}

int vec_grid(int64_t n) {
  int64_t g = pgb_cdiv(n, KB);
  if (g > NPART) g = NPART;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ inline void mr_tests(MinresScal* sc, double ynorm) {
  const double eps = 2.220446049250313e-16;
  sc->ynorm = ynorm;
  double Anorm = sc->Anorm;
  double epsx = Anorm * ynorm * eps;
  double test1 = (ynorm == 0.0 || Anorm == 0.0) ? INFINITY : sc->rnorm / (Anorm * ynorm);
  double test2 = (Anorm == 0.0) ? INFINITY : sc->root / Anorm;
  int istop = sc->pending;
  if (istop == 0) {
    double t1 = 1.0 + test1, t2 = 1.0 + test2;
    if (t2 <= 1.0) istop = 2;
    if (t1 <= 1.0) istop = 1;
    if (sc->itn >= sc->maxiter) istop = 6;
    if (sc->Acond >= 0.1 / eps) istop = 4;
    if (epsx >= sc->beta1) istop = 3;
    if (test2 <= sc->rtol) istop = 2;
    if (test1 <= sc->rtol) istop = 1;
  }
  sc->istop = istop;
}

// Stopping tests of the previous iteration (needs ||x||^2 = xx from its x-update), then the Lanczos /
// rotation scalars of iteration itn (scipy's minres loop body); sc->alfa holds v.y, bt = r2.psolve(r2).
// Works on a global or a thread-local copy of the state.  Returns true when the iteration stops.
__device__ inline bool mr_scalar_step(MinresScal* sc, int itn, double bt, double xx, int test_only) {
  const double eps = 2.220446049250313e-16;
  if (itn >= 2 || test_only) {
    mr_tests(sc, sqrt(xx));
    if (sc->istop != 0) return true;
  }
  if (test_only) return false;
  double alfa = sc->alfa;
  double oldb = sc->beta;
  double beta = sqrt(bt);  // bt < 0 cannot happen for an SPD preconditioner
  sc->tnorm2 += alfa * alfa + oldb * oldb + beta * beta;
  if (itn == 1 && beta / sc->beta1 <= 10 * eps) sc->pending = -1;
  double oldeps = sc->epsln;
  double delta = sc->cs * sc->dbar + sc->sn * alfa;
  double gbar = sc->sn * sc->dbar - sc->cs * alfa;
  double epsln = sc->sn * beta;
  double dbar = -sc->cs * beta;
  double root = hypot(gbar, dbar);
  double gamma = fmax(hypot(gbar, beta), eps);
  double cs = gbar / gamma, sn = beta / gamma;
  double phi = cs * sc->phibar;
  double phibar = sn * sc->phibar;
  sc->oldeps = oldeps;
  sc->delta = delta;
  sc->denom = 1.0 / gamma;
  sc->phi = phi;
  sc->inv_beta = 1.0 / beta;
  sc->epsln = epsln;
  sc->dbar = dbar;
  sc->cs = cs;
  sc->sn = sn;
  sc->phibar = phibar;
  sc->gmax = fmax(sc->gmax, gamma);
  sc->gmin = fmin(sc->gmin, gamma);
  double z = sc->rhs1 / gamma;
  sc->rhs1 = sc->rhs2 - delta * z;
  sc->rhs2 = -epsln * z;
  sc->Anorm = sqrt(sc->tnorm2);
  sc->rnorm = phibar;
  sc->root = root;
  sc->Acond = sc->gmax / sc->gmin;
  sc->oldb = oldb;
  sc->beta = beta;
  sc->itn = itn;
  return false;
}

}  // namespace
