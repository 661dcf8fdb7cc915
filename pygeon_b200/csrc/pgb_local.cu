// K1: per-cell local matrices, fp64, one thread per cell.
//
// Loads: cell->node ids as one 16-byte vector per cell (3D), sign / aux bit words, optional
// weight and tensor (SoA over cells -> coalesced), node coordinates gathered as 16-byte vectors
// from the padded [nn][4] (3D) / [nn][2] (2D) table.  Stores: each thread writes its L*L values
// to a shared-memory tile laid out [k][thread] (conflict-free), then the block streams the tile
// to vals[cell][k] with fully coalesced 8-byte stores.
//
// Closed forms: SURVEY.md section 8(a); each kind cites the reference code it replaces in pgb.h.
#include "pgb_common.cuh"
#include "pgb_rt1_table.cuh"

namespace {

constexpr int BLOCK = 128;

struct Params {
  int64_t nc;
  const int32_t* cell_nodes;
  const uint8_t* sign_bits;
  const uint32_t* aux_bits;
  const double* coords;
  const double* K;
  const double* w;
  const uint32_t* dst_pos;  // nullable: destination row of every local row (row-sorted layout)
  double* vals;
  double R[9];
  int r_ident;
};

template <int D>
struct Geo {
  double X[D + 1][D];
  double G[D + 1][D];
  double vol;
};

template <int D>
__device__ __forceinline__ void load_nodes(const Params& p, int64_t c, double (&X)[D + 1][D]) {
  if constexpr (D == 3) {
    int4 cn = reinterpret_cast<const int4*>(p.cell_nodes)[c];
    int ids[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#ifndef PGB_K1_LD128
      // one padded node = one 32-byte sector = one 256-bit load (one L1 wavefront per lane instead of two)
      const double* q = p.coords + 4 * (int64_t)ids[a];
      double pad;
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(X[a][0]), "=d"(X[a][1]), "=d"(X[a][2]), "=d"(pad) : "l"(q));
#else
      const double2* q = reinterpret_cast<const double2*>(p.coords) + 2 * (int64_t)ids[a];
      double2 xy = __ldg(q);
      double2 z_ = __ldg(q + 1);
      X[a][0] = xy.x;
      X[a][1] = xy.y;
      X[a][2] = z_.x;
#endif
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int id = p.cell_nodes[c * 3 + a];
      double2 xy = __ldg(reinterpret_cast<const double2*>(p.coords) + id);
      X[a][0] = xy.x;
      X[a][1] = xy.y;
    }
  }
}

// |c| and barycentric gradients from the node coordinates alone
template <int D>
__device__ __forceinline__ void gradients(Geo<D>& g) {
  if constexpr (D == 3) {
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      e1[i] = g.X[1][i] - g.X[0][i];
      e2[i] = g.X[2][i] - g.X[0][i];
      e3[i] = g.X[3][i] - g.X[0][i];
    }
    double c23[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
    double c31[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
    double c12[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    double det = e1[0] * c23[0] + e1[1] * c23[1] + e1[2] * c23[2];
    double inv = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.G[1][i] = c23[i] * inv;
      g.G[2][i] = c31[i] * inv;
      g.G[3][i] = c12[i] * inv;
      g.G[0][i] = -(g.G[1][i] + g.G[2][i] + g.G[3][i]);
    }
    g.vol = fabs(det) * (1.0 / 6.0);
  } else {
    double a = g.X[1][0] - g.X[0][0], b = g.X[1][1] - g.X[0][1];
    double c = g.X[2][0] - g.X[0][0], d = g.X[2][1] - g.X[0][1];
    double det = a * d - b * c;
    double inv = 1.0 / det;
    g.G[1][0] = d * inv;
    g.G[1][1] = -c * inv;
    g.G[2][0] = -b * inv;
    g.G[2][1] = a * inv;
    g.G[0][0] = -(g.G[1][0] + g.G[2][0]);
    g.G[0][1] = -(g.G[1][1] + g.G[2][1]);
    g.vol = fabs(det) * 0.5;
  }
}

// Kr = R K R^T (d x d)
template <int D>
__device__ __forceinline__ void load_tensor(const Params& p, int64_t c, double (&Kr)[D][D]) {
  if (p.r_ident) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) Kr[i][j] = p.K[(int64_t)(i * 3 + j) * p.nc + c];
  } else {
    double K3[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) K3[i][j] = p.K[(int64_t)(i * 3 + j) * p.nc + c];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) s += p.R[i * 3 + a] * K3[a][b] * p.R[j * 3 + b];
        Kr[i][j] = s;
      }
  }
}

// v^T Kr u helpers: KV[n] = Kr V[n]
template <int D, int N, bool HAS_K>
__device__ __forceinline__ void apply_tensor(const double (&Kr)[D][D], const double (&V)[N][D], double (&KV)[N][D]) {
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if constexpr (HAS_K) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < D; ++j) s += Kr[i][j] * V[n][j];
        KV[n][i] = s;
      } else {
        KV[n][i] = V[n][i];
      }
    }
}

template <int D>
__device__ __forceinline__ double dotd(const double (&a)[D], const double (&b)[D]) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) s += a[i] * b[i];
  return s;
}

template <int D>
__device__ __forceinline__ double mhat(int i, int j) {
  constexpr double base = 1.0 / ((D + 1) * (D + 2));
  return (i == j) ? 2.0 * base : base;
}

// RT0 local mass into A[(D+1)][(D+1)]:  w s_a s_b/(d^2|c|) [ T/((d+1)(d+2)) + y_a^T Kr y_b ]
template <int D, bool HAS_K>
__device__ __forceinline__ void rt0_local(const Geo<D>& g, const double (&Kr)[D][D], unsigned sbits, double w,
                                          double (&A)[D + 1][D + 1]) {
  double Y[D + 1][D], KY[D + 1][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double m = 0;
#pragma unroll
    for (int a = 0; a <= D; ++a) m += g.X[a][i];
    m *= 1.0 / (D + 1);
#pragma unroll
    for (int a = 0; a <= D; ++a) Y[a][i] = g.X[a][i] - m;
  }
  apply_tensor<D, D + 1, HAS_K>(Kr, Y, KY);
  double T = 0;
#pragma unroll
  for (int a = 0; a <= D; ++a) T += dotd<D>(Y[a], KY[a]);
  T *= 1.0 / ((D + 1) * (D + 2));
  double coef = w / ((double)(D * D) * g.vol);
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int b = 0; b <= D; ++b) {
      double s = (((sbits >> a) ^ (sbits >> b)) & 1u) ? -coef : coef;
      A[a][b] = s * (dotd<D>(Y[a], KY[b]) + T);
    }
}

__device__ __forceinline__ void pair3(int e, int& p, int& q) {
  constexpr int P[6] = {0, 0, 0, 1, 1, 2};
  constexpr int Q[6] = {1, 2, 3, 2, 3, 3};
  p = P[e];
  q = Q[e];
}

// local edge e of a D-simplex joins slots (p, q): (0,1),(0,2)[,(0,3)],(1,2)[,(1,3),(2,3)]
template <int D>
__device__ __forceinline__ void pair_of(int e, int& p, int& q) {
  if constexpr (D == 3) {
    pair3(e, p, q);
  } else {
    constexpr int P[3] = {0, 0, 1};
    constexpr int Q[3] = {1, 2, 2};
    p = P[e];
    q = Q[e];
  }
}

// (phi_i, phi_j) of the P2 basis on a D-simplex of measure 1 (Lagrange2.assemble_local_mass,
// h1.py:352-392): nodes first, then the edges in pair_of order
template <int D>
__device__ __forceinline__ double l2_mass_entry(int i, int j) {
  constexpr int N = D + 1;
  constexpr double nn_diag = (D == 2) ? 1.0 / 30 : 1.0 / 70, nn_off = (D == 2) ? -1.0 / 180 : 1.0 / 420;
  constexpr double ne_adj = (D == 2) ? 0.0 : -1.0 / 105, ne_opp = (D == 2) ? -1.0 / 45 : -1.0 / 70;
  constexpr double ee_diag = (D == 2) ? 8.0 / 45 : 8.0 / 105, ee_share = (D == 2) ? 4.0 / 45 : 4.0 / 105;
  constexpr double ee_disj = 2.0 / 105;
  if (i < N && j < N) return i == j ? nn_diag : nn_off;
  if (i < N || j < N) {
    int node = i < N ? i : j, p, q;
    pair_of<D>((i < N ? j : i) - N, p, q);
    return (node == p || node == q) ? ne_adj : ne_opp;
  }
  int p, q, r, s;
  pair_of<D>(i - N, p, q);
  pair_of<D>(j - N, r, s);
  int common = (p == r) + (p == s) + (q == r) + (q == s);
  return common == 2 ? ee_diag : common == 1 ? ee_share : ee_disj;
}

template <int KIND, int D>
struct LocalSize {
  static constexpr int L = (KIND == PGB_BDM1_MASS || KIND == PGB_BDM1_DIVDIV || KIND == PGB_BDM1_LUMPED) ? D * (D + 1)
                           : (KIND == PGB_N0_MASS || KIND == PGB_N0_CURLCURL) ? D * (D + 1) / 2
                           : (KIND == PGB_P0_MASS)                            ? 1
                           : (KIND == PGB_DARCY_SADDLE)                       ? D + 2
                           : (KIND == PGB_L2_MASS || KIND == PGB_L2_STIFF)    ? (D + 1) + D * (D + 1) / 2
                           : (KIND == PGB_RT1_MASS || KIND == PGB_RT1_DIVDIV) ? D * (D + 2)
                           : (KIND == PGB_RT1_LUMPED_CELL)                    ? D
                                                                              : D + 1;
};

template <int KIND, int D, bool HAS_K, bool HAS_W>
__global__ void __launch_bounds__(BLOCK) k_local(Params p) {
  constexpr int L = LocalSize<KIND, D>::L;
  constexpr int LL = L * L;
  constexpr bool STAGED = (LL <= 36);
  __shared__ double tile[STAGED ? LL * (BLOCK + 1) : 1];

  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * BLOCK;
  const int64_t c = c0 + tid;
  const bool live = c < p.nc;

  // Row-sorted output (p.dst_pos): every kind emits its values in row-major order, so a local row
  // is complete when k % L == L-1; it is then stored with vector stores (STG.E.256 for L % 4 == 0)
  // at its destination row.  Otherwise: shared-memory staging + coalesced stream-out.
  // Rows are stored with pitch LP = L rounded up to a multiple of 4: whole 32-byte sectors, written
  // with STG.E.256 (partial-sector rows of 40 / 48 bytes were measured 3x slower: L2 has to merge
  // sector halves written by different cells).
  const bool rowmajor = p.dst_pos != nullptr;
  constexpr int LP = (L + 3) & ~3;
  double rowbuf[LP];
#pragma unroll
  for (int b = L; b < LP; ++b) rowbuf[b] = 0.0;
  auto flush_row = [&](int row) {
    double* dst = p.vals + (uint64_t)p.dst_pos[c * L + row] * LP;
    if constexpr (LP % 4 == 0) {
#pragma unroll
      for (int b = 0; b < LP; b += 4)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + b), "d"(rowbuf[b]), "d"(rowbuf[b + 1]),
                     "d"(rowbuf[b + 2]), "d"(rowbuf[b + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int b = 0; b < LP; b += 2) *reinterpret_cast<double2*>(dst + b) = make_double2(rowbuf[b], rowbuf[b + 1]);
    }
  };
  auto emit = [&](int k, double v) {
    if (rowmajor) {
      rowbuf[k % L] = v;
      if (k % L == L - 1) flush_row(k / L);
    } else if constexpr (STAGED) {
      tile[k * (BLOCK + 1) + tid] = v;
    } else {
      p.vals[c * LL + k] = v;
    }
  };

  if (live) {
    Geo<D> g;
    load_nodes<D>(p, c, g.X);
    gradients<D>(g);
    double w = 1.0;
    if constexpr (HAS_W) w = p.w[c];
    double Kr[D][D];
    if constexpr (HAS_K) load_tensor<D>(p, c, Kr);

    if constexpr (KIND == PGB_P0_MASS) {
      emit(0, w / g.vol);
    } else if constexpr (KIND == PGB_DARCY_SADDLE) {
      double A[D + 1][D + 1];
      unsigned sb = p.sign_bits[c];
      rt0_local<D, HAS_K>(g, Kr, sb, w, A);
      double bv = w / g.vol;  // P0 mass (w/|c|) times the +-1 of the divergence
#pragma unroll
      for (int a = 0; a <= D; ++a) {
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, A[a][b]);
        emit(a * L + (D + 1), ((sb >> a) & 1u) ? bv : -bv);  // -B^T entry: -(s_a w/|c|)
      }
#pragma unroll
      for (int a = 0; a <= D; ++a) emit((D + 1) * L + a, ((sb >> a) & 1u) ? bv : -bv);  // -B row
      emit((D + 1) * L + (D + 1), 0.0);
    } else if constexpr (KIND == PGB_L1_LUMPED) {
      double wv = w * g.vol / (D + 1);
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, a == b ? wv : 0.0);
    } else if constexpr (KIND == PGB_RT0_LUMPED) {
      // face slot a = the face opposite node a: centre = mean of the other nodes, measure from them
      double cc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double m = 0;
#pragma unroll
        for (int t = 0; t <= D; ++t) m += g.X[t][i];
        cc[i] = m / (D + 1);
      }
#pragma unroll
      for (int a = 0; a <= D; ++a) {
        double dist[D], o[D][D];  // o: the D nodes of the face
        int k = 0;
#pragma unroll
        for (int t = 0; t <= D; ++t) {
          if (t != a) {
#pragma unroll
            for (int i = 0; i < D; ++i) o[k][i] = g.X[t][i];
            ++k;
          }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double fc = 0;
#pragma unroll
          for (int t = 0; t < D; ++t) fc += o[t][i];
          dist[i] = fc / D - cc[i];
        }
        double area;
        if constexpr (D == 3) {
          double u[3] = {o[1][0] - o[0][0], o[1][1] - o[0][1], o[1][2] - o[0][2]};
          double v[3] = {o[2][0] - o[0][0], o[2][1] - o[0][1], o[2][2] - o[0][2]};
          double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
          area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        } else {
          double ux = o[1][0] - o[0][0], uy = o[1][1] - o[0][1];
          area = sqrt(ux * ux + uy * uy);
        }
        double q = 0, nd = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          nd += dist[i] * dist[i];
#pragma unroll
          for (int j = 0; j < D; ++j) q += dist[i] * (HAS_K ? Kr[i][j] : (i == j ? 1.0 : 0.0)) * dist[j];
        }
        nd = sqrt(nd);
        double val = (nd > 0.0) ? q / nd / area : 0.0;
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, a == b ? val : 0.0);
      }
    } else if constexpr (KIND == PGB_RT0_DIVDIV) {
      unsigned sb = p.sign_bits[c];
      double wv = w / g.vol;
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, (((sb >> a) ^ (sb >> b)) & 1u) ? -wv : wv);
    } else if constexpr (KIND == PGB_BDM1_DIVDIV) {
      unsigned sb = p.sign_bits[c];
      double wv = w / (g.vol * (double)(D * D));
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int q = 0; q < L; ++q) emit(r * L + q, (((sb >> (r / D)) ^ (sb >> (q / D))) & 1u) ? -wv : wv);
    } else if constexpr (KIND == PGB_L1_MASS) {
      double wv = w * g.vol;
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, wv * mhat<D>(a, b));
    } else if constexpr (KIND == PGB_L1_STIFF || KIND == PGB_L1_STIFF_ROT) {
      double Gv[D + 1][D], KG[D + 1][D];
#pragma unroll
      for (int a = 0; a <= D; ++a) {
        if constexpr (KIND == PGB_L1_STIFF_ROT && D == 2) {
          Gv[a][0] = g.G[a][1];
          Gv[a][1] = -g.G[a][0];
        } else {
#pragma unroll
          for (int i = 0; i < D; ++i) Gv[a][i] = g.G[a][i];
        }
      }
      apply_tensor<D, D + 1, HAS_K>(Kr, Gv, KG);
      double wv = w * g.vol;
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, wv * dotd<D>(Gv[a], KG[b]));
    } else if constexpr (KIND == PGB_L2_MASS) {
      double wv = w * g.vol;
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) emit(i * L + j, wv * l2_mass_entry<D>(i, j));
    } else if constexpr (KIND == PGB_L2_STIFF) {
      // The gradients of the P2 basis are linear; their nodal values are combinations of the
      // grad lambda, integrated with the P1 mass Mhat.  The reference's entry (i, k) is
      // Psi_i^T K Psi_k; the CSR built here is handed to scipy as CSC, i.e. it must hold the
      // transpose: the formulas below are evaluated with Gm[a][b] = grad lambda_b^T K grad lambda_a.
      constexpr int N = D + 1;
      double Gv[N][D], KG[N][D], Gm[N][N];
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) Gv[a][i] = g.G[a][i];
      apply_tensor<D, N, HAS_K>(Kr, Gv, KG);
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) Gm[a][b] = g.vol * dotd<D>(Gv[b], KG[a]);
      constexpr double c1 = 1.0 / N;
#pragma unroll
      for (int i = 0; i < L; ++i) {
#pragma unroll
        for (int k = 0; k < L; ++k) {
          double v;
          if (i < N && k < N) {
            v = Gm[i][k] * (16.0 * mhat<D>(i, k) - 8.0 * c1 + 1.0);
          } else if (i < N) {
            int r, s;
            pair_of<D>(k - N, r, s);
            v = 4.0 * (Gm[i][s] * (4.0 * mhat<D>(i, r) - c1) + Gm[i][r] * (4.0 * mhat<D>(i, s) - c1));
          } else if (k < N) {
            int p_, q_;
            pair_of<D>(i - N, p_, q_);
            v = 4.0 * (Gm[q_][k] * (4.0 * mhat<D>(p_, k) - c1) + Gm[p_][k] * (4.0 * mhat<D>(q_, k) - c1));
          } else {
            int p_, q_, r, s;
            pair_of<D>(i - N, p_, q_);
            pair_of<D>(k - N, r, s);
            v = 16.0 * (mhat<D>(p_, r) * Gm[q_][s] + mhat<D>(p_, s) * Gm[q_][r] + mhat<D>(q_, r) * Gm[p_][s] +
                        mhat<D>(q_, s) * Gm[p_][r]);
          }
          emit(i * L + k, v);
        }
      }
    } else if constexpr (KIND == PGB_RT1_MASS) {
      // Nodes in rank order (ascending id), Gram matrix Q[n][m] = (K y_n) . y_m of the edge vectors
      // y_n = x_n - x_0, then the constant contraction of pgb_rt1_table.cuh.  The reference's entry
      // (p, q) is Psi_p^T K Psi_q; the CSR built here is read as CSC, i.e. it holds the transpose,
      // which is what (K y_n) . y_m (instead of y_n . K y_m) produces.
      constexpr int N = D + 1;
      const unsigned bits = p.aux_bits[c];
      double Xr[N][D];
#pragma unroll
      for (int r = 0; r < N; ++r) {
        const unsigned slot = (bits >> (2 * r)) & 3u;
#pragma unroll
        for (int a = 0; a < N; ++a)
          if (slot == (unsigned)a) {
#pragma unroll
            for (int i = 0; i < D; ++i) Xr[r][i] = g.X[a][i];
          }
      }
      double Y[D][D], KY[D][D], Q[D][D];
#pragma unroll
      for (int n = 0; n < D; ++n)
#pragma unroll
        for (int i = 0; i < D; ++i) Y[n][i] = Xr[n + 1][i] - Xr[0][i];
      apply_tensor<D, D, HAS_K>(Kr, Y, KY);
#pragma unroll
      for (int n = 0; n < D; ++n)
#pragma unroll
        for (int m = 0; m < D; ++m) Q[n][m] = dotd<D>(KY[n], Y[m]);
      const double scale = 1.0 / (D * D * g.vol);
#pragma unroll
      for (int pp = 0; pp < L; ++pp) {
        // sign of the face the dof lives on (cell functions: +)
        const bool negp = pp < D * N && ((bits >> (8 + (D * N - 1 - pp) % N)) & 1u);
        const double sp = negp ? -scale : scale;
#pragma unroll
        for (int qq = 0; qq < L; ++qq) {
          const bool negq = qq < D * N && ((bits >> (8 + (D * N - 1 - qq) % N)) & 1u);
          double v = 0.0;
#pragma unroll
          for (int n = 0; n < D; ++n)
#pragma unroll
            for (int m = 0; m < D; ++m) v += pgb_rt1::table<D>()[((pp * L + qq) * D + n) * D + m] * Q[n][m];
          emit(pp * L + qq, negq ? -(sp * v) : sp * v);
        }
      }
    } else if constexpr (KIND == PGB_RT1_LUMPED_CELL) {
      // cell block of the lumped RT1 mass (hdiv.py:970-1028): the d cell functions at the cell centre are
      // P_i = (sum_j x_j - (d+1) x_i) / ((d+1)^2 d |c|), nodes in rank order; block = |c| (d+1)/(d+2) P^T K P
      constexpr int N = D + 1;
      const unsigned bits = p.aux_bits[c];
      double S[D], P[D][D], KP[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        S[i] = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a) S[i] += g.X[a][i];
      }
      const double sc = 1.0 / ((D + 1) * (D + 1) * D * g.vol);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        const unsigned slot = (bits >> (2 * r)) & 3u;
#pragma unroll
        for (int a = 0; a < N; ++a)
          if (slot == (unsigned)a) {
#pragma unroll
            for (int i = 0; i < D; ++i) P[r][i] = (S[i] - (D + 1) * g.X[a][i]) * sc;
          }
      }
      apply_tensor<D, D, HAS_K>(Kr, P, KP);
      const double f = g.vol * (D + 1) / (D + 2);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) emit(i * L + j, f * dotd<D>(KP[i], P[j]));
    } else if constexpr (KIND == PGB_RT1_DIVDIV) {
      // RT1 stiffness B^T M_P1 B (hdiv.py:813-865 + l2.py:63-86,457-470): the nodal values of the
      // divergence of local function q are s_q Dv[r][q] / (d |c|) with the constant table
      // Dv[r][q] = 1 + (d+1)([r == i_q] - [r == j_q]) (face function at node i of the face opposite node j;
      // i = q / d, j = (d(d+1) - 1 - q) % (d+1)) and (d+1)[r == k] - 1 (cell function k), nodes in rank
      // order; with Mhat = (1 1^T + I) / ((d+1)(d+2)) the local matrix is
      // w / (d^2 |c| (d+1)(d+2)) s_p s_q (colsum_p colsum_q + sum_r Dv[r][p] Dv[r][q]).
      constexpr int N = D + 1;
      const unsigned bits = p.aux_bits[c];
      const double scale = w / (D * D * g.vol * (D + 1) * (D + 2));
      auto dv = [](int r, int q) -> double {
        if (q < D * N) {
          const int i = q / D, j = (D * N - 1 - q) % N;
          return 1.0 + N * ((r == i ? 1.0 : 0.0) - (r == j ? 1.0 : 0.0));
        }
        return N * (r == q - D * N ? 1.0 : 0.0) - 1.0;
      };
#pragma unroll
      for (int pp = 0; pp < L; ++pp) {
        const bool negp = pp < D * N && ((bits >> (8 + (D * N - 1 - pp) % N)) & 1u);
#pragma unroll
        for (int qq = 0; qq < L; ++qq) {
          const bool negq = qq < D * N && ((bits >> (8 + (D * N - 1 - qq) % N)) & 1u);
          double sp = 0.0, sq = 0.0, pq = 0.0;  // folded at compile time: pp, qq, r are unrolled constants
#pragma unroll
          for (int r = 0; r < N; ++r) {
            sp += dv(r, pp);
            sq += dv(r, qq);
            pq += dv(r, pp) * dv(r, qq);
          }
          const double v = scale * (sp * sq + pq);
          emit(pp * L + qq, negp != negq ? -v : v);
        }
      }
    } else if constexpr (KIND == PGB_RT0_MASS) {
      double A[D + 1][D + 1];
      rt0_local<D, HAS_K>(g, Kr, p.sign_bits[c], w, A);
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) emit(a * L + b, A[a][b]);
    } else if constexpr (KIND == PGB_N0_MASS) {
      double KG[D + 1][D], GKG[D + 1][D + 1];
      apply_tensor<D, D + 1, HAS_K>(Kr, g.G, KG);
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int b = 0; b <= D; ++b) GKG[a][b] = dotd<D>(g.G[a], KG[b]);
      unsigned bits = p.aux_bits[c];
      double wv = w * g.vol;
#pragma unroll
      for (int e = 0; e < L; ++e) {
        int pe, qe;
        if constexpr (D == 3) pair3(e, pe, qe);
        else { pe = (e == 0) ? 1 : 0; qe = (e == 2) ? 1 : 2; }
#pragma unroll
        for (int f = 0; f < L; ++f) {
          int rf, sf;
          if constexpr (D == 3) pair3(f, rf, sf);
          else { rf = (f == 0) ? 1 : 0; sf = (f == 2) ? 1 : 2; }
          double v = mhat<D>(pe, rf) * GKG[qe][sf] - mhat<D>(pe, sf) * GKG[qe][rf] -
                     mhat<D>(qe, rf) * GKG[pe][sf] + mhat<D>(qe, sf) * GKG[pe][rf];
          double s = (((bits >> e) ^ (bits >> f)) & 1u) ? -wv : wv;
          emit(e * L + f, s * v);
        }
      }
    } else if constexpr (KIND == PGB_N0_CURLCURL) {
      if constexpr (D == 2) {
        unsigned sb = p.sign_bits[c];
        double wv = w / g.vol;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) emit(a * 3 + b, (((sb >> a) ^ (sb >> b)) & 1u) ? -wv : wv);
      } else {
        double A[4][4];
        rt0_local<3, HAS_K>(g, Kr, p.sign_bits[c], w, A);
        unsigned bits = p.aux_bits[c];
        constexpr int FA[6] = {2, 1, 1, 0, 0, 0};
        constexpr int FB[6] = {3, 3, 2, 3, 2, 1};
#pragma unroll
        for (int e = 0; e < 6; ++e) {
          double sea = ((bits >> (8 + 2 * e)) & 1u) ? -1.0 : 1.0;
          double seb = ((bits >> (9 + 2 * e)) & 1u) ? -1.0 : 1.0;
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            double sfa = ((bits >> (8 + 2 * f)) & 1u) ? -1.0 : 1.0;
            double sfb = ((bits >> (9 + 2 * f)) & 1u) ? -1.0 : 1.0;
            double v = sea * (sfa * A[FA[e]][FA[f]] + sfb * A[FA[e]][FB[f]]) +
                       seb * (sfa * A[FB[e]][FA[f]] + sfb * A[FB[e]][FB[f]]);
            emit(e * 6 + f, v);
          }
        }
      }
    } else if constexpr (KIND == PGB_BDM1_MASS || KIND == PGB_BDM1_LUMPED) {
      unsigned bits = p.aux_bits[c];
      unsigned sb = p.sign_bits[c];
      double T[L][D], KT[L][D];
      int js[L];
      double inv = 1.0 / ((double)D * g.vol);
#pragma unroll
      for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int pos = 0; pos < D; ++pos) {
          int r = a * D + pos;
          int j = (bits >> (2 * r)) & 3u;
          js[r] = j;
          double s = ((sb >> a) & 1u) ? -inv : inv;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double xj = g.X[0][i];
#pragma unroll
            for (int t = 1; t <= D; ++t) xj = (j == t) ? g.X[t][i] : xj;
            T[r][i] = (xj - g.X[a][i]) * s;
          }
        }
      apply_tensor<D, L, HAS_K>(Kr, T, KT);
      double wv = w * g.vol;
      constexpr double base = 1.0 / ((D + 1) * (D + 2));
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int q = 0; q < L; ++q) {
          double m = (js[r] == js[q]) ? 2.0 * base : base;
          if constexpr (KIND == PGB_BDM1_LUMPED) m = (js[r] == js[q]) ? 1.0 / (D + 1) : 0.0;
          emit(r * L + q, wv * m * dotd<D>(T[r], KT[q]));
        }
    }
  }

  if (STAGED && !rowmajor) {
    __syncthreads();
    int64_t rem = p.nc - c0;
    int nb = rem < BLOCK ? (int)rem : BLOCK;
    double* out = p.vals + c0 * LL;
    for (int i = tid; i < nb * LL; i += BLOCK) {
      int cell = i / LL, k = i - cell * LL;
      out[i] = tile[k * (BLOCK + 1) + cell];
    }
  }
}

template <int KIND, int D>
int launch(const Params& p, cudaStream_t s) {
  unsigned grid = pgb_grid(p.nc, BLOCK);
  bool hk = p.K != nullptr, hw = p.w != nullptr;
  if (hk && hw) k_local<KIND, D, true, true><<<grid, BLOCK, 0, s>>>(p);
  else if (hk) k_local<KIND, D, true, false><<<grid, BLOCK, 0, s>>>(p);
  else if (hw) k_local<KIND, D, false, true><<<grid, BLOCK, 0, s>>>(p);
  else k_local<KIND, D, false, false><<<grid, BLOCK, 0, s>>>(p);
  PGB_LAUNCH_CHECK();
  return 0;
}

template <int D>
int dispatch(int kind, const Params& p, cudaStream_t s) {
  switch (kind) {
    case PGB_L1_MASS: return launch<PGB_L1_MASS, D>(p, s);
    case PGB_L1_STIFF: return launch<PGB_L1_STIFF, D>(p, s);
    case PGB_L1_STIFF_ROT: return launch<PGB_L1_STIFF_ROT, D>(p, s);
    case PGB_RT0_MASS: return launch<PGB_RT0_MASS, D>(p, s);
    case PGB_BDM1_MASS: return launch<PGB_BDM1_MASS, D>(p, s);
    case PGB_N0_MASS: return launch<PGB_N0_MASS, D>(p, s);
    case PGB_N0_CURLCURL: return launch<PGB_N0_CURLCURL, D>(p, s);
    case PGB_P0_MASS: return launch<PGB_P0_MASS, D>(p, s);
    case PGB_RT0_DIVDIV: return launch<PGB_RT0_DIVDIV, D>(p, s);
    case PGB_BDM1_DIVDIV: return launch<PGB_BDM1_DIVDIV, D>(p, s);
    case PGB_DARCY_SADDLE: return launch<PGB_DARCY_SADDLE, D>(p, s);
    case PGB_L1_LUMPED: return launch<PGB_L1_LUMPED, D>(p, s);
    case PGB_RT0_LUMPED: return launch<PGB_RT0_LUMPED, D>(p, s);
    case PGB_BDM1_LUMPED: return launch<PGB_BDM1_LUMPED, D>(p, s);
    case PGB_L2_MASS: return launch<PGB_L2_MASS, D>(p, s);
    case PGB_L2_STIFF: return launch<PGB_L2_STIFF, D>(p, s);
    case PGB_RT1_MASS: return launch<PGB_RT1_MASS, D>(p, s);
    case PGB_RT1_DIVDIV: return launch<PGB_RT1_DIVDIV, D>(p, s);
    case PGB_RT1_LUMPED_CELL: return launch<PGB_RT1_LUMPED_CELL, D>(p, s);
  }
  pgb_set_error("pgb_local_matrices: unknown kind %d", kind);
  return 2;
}

}  // namespace

extern "C" int pgb_local_size(int kind, int dim) {
  switch (kind) {
    case PGB_BDM1_MASS:
    case PGB_BDM1_LUMPED:
    case PGB_BDM1_DIVDIV: return dim * (dim + 1);
    case PGB_N0_MASS:
    case PGB_N0_CURLCURL: return dim * (dim + 1) / 2;
    case PGB_P0_MASS: return 1;
    case PGB_DARCY_SADDLE: return dim + 2;
    case PGB_L2_MASS:
    case PGB_L2_STIFF: return (dim + 1) + dim * (dim + 1) / 2;
    case PGB_RT1_MASS:
    case PGB_RT1_DIVDIV: return dim * (dim + 2);
    case PGB_RT1_LUMPED_CELL: return dim;
    default: return dim + 1;
  }
}

extern "C" int pgb_local_matrices(int kind, int dim, int64_t nc, const int32_t* cell_nodes,
                                  const uint8_t* sign_bits, const uint32_t* aux_bits,
                                  const double* coords, const double* K, const double* R_host,
                                  const double* w, const uint32_t* dst_pos, double* vals, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_local_matrices: dim must be 2 or 3");
  PGB_REQUIRE(cell_nodes && coords && vals, "pgb_local_matrices: null pointer");
  if (kind == PGB_RT0_MASS || kind == PGB_BDM1_MASS || kind == PGB_N0_CURLCURL ||
      kind == PGB_RT0_DIVDIV || kind == PGB_BDM1_DIVDIV || kind == PGB_DARCY_SADDLE || kind == PGB_BDM1_LUMPED)
    PGB_REQUIRE(sign_bits, "pgb_local_matrices: sign_bits required");
  if (kind == PGB_BDM1_MASS || kind == PGB_BDM1_LUMPED || kind == PGB_N0_MASS ||
      (kind == PGB_N0_CURLCURL && dim == 3) || kind == PGB_RT1_MASS || kind == PGB_RT1_DIVDIV || kind == PGB_RT1_LUMPED_CELL)
    PGB_REQUIRE(aux_bits, "pgb_local_matrices: aux_bits required");
  if (nc == 0) return 0;
  Params p;
  p.nc = nc;
  p.cell_nodes = cell_nodes;
  p.sign_bits = sign_bits;
  p.aux_bits = aux_bits;
  p.coords = coords;
  p.K = K;
  p.w = w;
  p.dst_pos = dst_pos;
  p.vals = vals;
  bool ident = true;
  for (int i = 0; i < 9; ++i) {
    double r = (R_host && i < dim * 3) ? R_host[i] : ((i / 3 == i % 3 && i < dim * 3) ? 1.0 : 0.0);
    p.R[i] = r;
    if (i < dim * 3 && r != ((i / 3 == i % 3) ? 1.0 : 0.0)) ident = false;
  }
  p.r_ident = ident ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  return dim == 3 ? dispatch<3>(kind, p, s) : dispatch<2>(kind, p, s);
}
