// Grid packing kernels: raw (cell_faces, face_nodes, face_ridges, ridge_peaks) index arrays
// -> per-cell tables consumed by the local-matrix kernels.  Integer only, one thread per cell.
#include "pgb_common.cuh"

namespace {

// ---- cell_nodes: node opposite every face slot --------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_pack_cells(int64_t nc, const int32_t* __restrict__ cf_idx,
                                                    const int8_t* __restrict__ cf_sgn,
                                                    const int32_t* __restrict__ fn_idx,
                                                    int32_t* __restrict__ cell_nodes,
                                                    uint8_t* __restrict__ sign_bits,
                                                    int32_t* __restrict__ err_flag) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t f[D + 1];
  int32_t fn[D + 1][D];
  unsigned bits = 0;
#pragma unroll
  for (int a = 0; a <= D; ++a) {
    f[a] = cf_idx[c * (D + 1) + a];
    if (cf_sgn && cf_sgn[c * (D + 1) + a] < 0) bits |= 1u << a;
  }
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int k = 0; k < D; ++k) fn[a][k] = __ldg(fn_idx + (int64_t)f[a] * D + k);

  // the D+1 nodes of the cell: the D nodes of face 0 plus the node of face 1 not on face 0
  int32_t nodes[D + 1];
#pragma unroll
  for (int k = 0; k < D; ++k) nodes[k] = fn[0][k];
  int32_t extra = -1;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    bool on0 = false;
#pragma unroll
    for (int j = 0; j < D; ++j) on0 |= (fn[1][k] == fn[0][j]);
    if (!on0) extra = fn[1][k];
  }
  nodes[D] = extra;
  int64_t total = 0;
#pragma unroll
  for (int k = 0; k <= D; ++k) total += nodes[k];
  bool bad = (extra < 0);
#pragma unroll
  for (int a = 0; a <= D; ++a) {
    int64_t s = 0;
    int members = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      s += fn[a][k];
#pragma unroll
      for (int j = 0; j <= D; ++j) members += (fn[a][k] == nodes[j]);
    }
    bad |= (members != D);
    int64_t opp = total - s;
    bool found = false;
#pragma unroll
    for (int j = 0; j <= D; ++j) found |= (opp == nodes[j]);
    bad |= !found;
    cell_nodes[c * (D + 1) + a] = (int32_t)opp;
  }
  if (sign_bits) sign_bits[c] = (uint8_t)bits;
  if (bad) *err_flag = 1;
}

// sign bits alone (spaces that do not need them, e.g. Lagrange1, never upload cell_faces.data)
template <int D>
__global__ void __launch_bounds__(256) k_pack_signs(int64_t nc, const int8_t* __restrict__ cf_sgn,
                                                    uint8_t* __restrict__ sign_bits) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  unsigned bits = 0;
#pragma unroll
  for (int a = 0; a <= D; ++a)
    if (cf_sgn[c * (D + 1) + a] < 0) bits |= 1u << a;
  sign_bits[c] = (uint8_t)bits;
}

// ---- BDM1 dofs -----------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_pack_bdm1(int64_t nc, int64_t nf,
                                                   const int32_t* __restrict__ cf_idx,
                                                   const int32_t* __restrict__ fn_idx,
                                                   const int32_t* __restrict__ cell_nodes,
                                                   int32_t* __restrict__ cell_dofs,
                                                   uint32_t* __restrict__ jslot_bits) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t cn[D + 1];
#pragma unroll
  for (int a = 0; a <= D; ++a) cn[a] = cell_nodes[c * (D + 1) + a];
  uint32_t bits = 0;
#pragma unroll
  for (int a = 0; a <= D; ++a) {
    int32_t f = cf_idx[c * (D + 1) + a];
#pragma unroll
    for (int pos = 0; pos < D; ++pos) {
      int32_t node = __ldg(fn_idx + (int64_t)f * D + pos);
      int j = 0;
#pragma unroll
      for (int s = 0; s <= D; ++s)
        if (cn[s] == node) j = s;
      bits |= (uint32_t)j << (2 * (a * D + pos));
      cell_dofs[c * (D * (D + 1)) + a * D + pos] = (int32_t)(f + pos * nf);
    }
  }
  jslot_bits[c] = bits;
}

// ---- Nedelec0 edges, 3D ---------------------------------------------------------------------
// local pairs in the order (01)(02)(03)(12)(13)(23); the two faces containing edge (p,q) are the
// slots {0,1,2,3} \ {p,q} (a face slot is named after its opposite node).
__constant__ int c_ep[6] = {0, 0, 0, 1, 1, 2};
__constant__ int c_eq[6] = {1, 2, 3, 2, 3, 3};
__constant__ int c_fa[6] = {2, 1, 1, 0, 0, 0};  // first face containing the edge
__constant__ int c_fb[6] = {3, 3, 2, 3, 2, 1};  // second face containing the edge

__global__ void __launch_bounds__(256) k_pack_edges3d(int64_t nc, const int32_t* __restrict__ cf_idx,
                                                      const int32_t* __restrict__ fr_idx,
                                                      const int8_t* __restrict__ fr_sgn,
                                                      const int32_t* __restrict__ rp_idx,
                                                      const int32_t* __restrict__ cell_nodes,
                                                      int32_t* __restrict__ cell_edges,
                                                      uint32_t* __restrict__ edge_bits,
                                                      int32_t* __restrict__ err_flag) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t cn[4], f[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    cn[a] = cell_nodes[c * 4 + a];
    f[a] = cf_idx[c * 4 + a];
  }
  // ridges of every face with their end nodes
  int32_t rid[4][3], r0[4][3], r1[4][3];
  int8_t rs[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int32_t e = __ldg(fr_idx + (int64_t)f[a] * 3 + k);
      rid[a][k] = e;
      rs[a][k] = __ldg(fr_sgn + (int64_t)f[a] * 3 + k);
      r0[a][k] = __ldg(rp_idx + (int64_t)e * 2);
      r1[a][k] = __ldg(rp_idx + (int64_t)e * 2 + 1);
    }
  uint32_t bits = 0;
  bool bad = false;
#pragma unroll
  for (int e = 0; e < 6; ++e) {
    int32_t np_ = cn[c_ep[e]], nq_ = cn[c_eq[e]];
    int32_t eid = -1;
    bool flip = false;
    int sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int a = c_fa[e], b = c_fb[e];
      bool ha = (r0[a][k] == np_ && r1[a][k] == nq_) || (r0[a][k] == nq_ && r1[a][k] == np_);
      if (ha) {
        eid = rid[a][k];
        flip = (r0[a][k] != np_);
        sa = rs[a][k];
      }
      bool hb = (r0[b][k] == np_ && r1[b][k] == nq_) || (r0[b][k] == nq_ && r1[b][k] == np_);
      if (hb) sb = rs[b][k];
    }
    bad |= (eid < 0) || sa == 0 || sb == 0;
    cell_edges[c * 6 + e] = eid;
    if (flip) bits |= 1u << e;
    if (sa < 0) bits |= 1u << (8 + 2 * e);
    if (sb < 0) bits |= 1u << (9 + 2 * e);
  }
  edge_bits[c] = bits;
  if (bad) *err_flag = 1;
}

// 2D: edge of face slot a joins the two other slots (lo = smaller slot index).
__global__ void __launch_bounds__(256) k_pack_edges2d(int64_t nc, const int32_t* __restrict__ cf_idx,
                                                      const int32_t* __restrict__ fr_idx,
                                                      const int32_t* __restrict__ cell_nodes,
                                                      uint32_t* __restrict__ edge_bits) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  uint32_t bits = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int lo = (a == 0) ? 1 : 0;
    int32_t f = cf_idx[c * 3 + a];
    int32_t n0 = __ldg(fr_idx + (int64_t)f * 2);
    if (n0 != cell_nodes[c * 3 + lo]) bits |= 1u << a;
  }
  edge_bits[c] = bits;
}

// ---- coordinates ------------------------------------------------------------------------------
struct Rot {
  double r[9];
};
template <int D>
__global__ void __launch_bounds__(256) k_prep_coords(int64_t nn, const double* __restrict__ nodes,
                                                     Rot R, double* __restrict__ coords) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  double x = nodes[i], y = nodes[nn + i], z = nodes[2 * nn + i];
  double o[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < D; ++k) o[k] = R.r[k * 3] * x + R.r[k * 3 + 1] * y + R.r[k * 3 + 2] * z;
  if (D == 3) {
    reinterpret_cast<double2*>(coords)[2 * i] = make_double2(o[0], o[1]);
    reinterpret_cast<double2*>(coords)[2 * i + 1] = make_double2(o[2], 0.0);
  } else {
    reinterpret_cast<double2*>(coords)[i] = make_double2(o[0], o[1]);
  }
}

}  // namespace

extern "C" int pgb_pack_cells(int dim, int64_t nc, const int32_t* cf_idx, const int8_t* cf_sgn,
                              const int32_t* fn_idx, int32_t* cell_nodes, uint8_t* sign_bits,
                              int32_t* err_flag, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_pack_cells: dim must be 2 or 3");
  if (nc == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3)
    k_pack_cells<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, cf_idx, cf_sgn, fn_idx, cell_nodes, sign_bits, err_flag);
  else
    k_pack_cells<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, cf_idx, cf_sgn, fn_idx, cell_nodes, sign_bits, err_flag);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_pack_signs(int dim, int64_t nc, const int8_t* cf_sgn, uint8_t* sign_bits, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_pack_signs: dim must be 2 or 3");
  if (nc == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3) k_pack_signs<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, cf_sgn, sign_bits);
  else k_pack_signs<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, cf_sgn, sign_bits);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_pack_bdm1(int dim, int64_t nc, int64_t nf, const int32_t* cf_idx,
                             const int32_t* fn_idx, const int32_t* cell_nodes, int32_t* cell_dofs,
                             uint32_t* jslot_bits, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_pack_bdm1: dim must be 2 or 3");
  if (nc == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3)
    k_pack_bdm1<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, nf, cf_idx, fn_idx, cell_nodes, cell_dofs, jslot_bits);
  else
    k_pack_bdm1<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, nf, cf_idx, fn_idx, cell_nodes, cell_dofs, jslot_bits);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_pack_edges3d(int64_t nc, const int32_t* cf_idx, const int32_t* fr_idx,
                                const int8_t* fr_sgn, const int32_t* rp_idx,
                                const int32_t* cell_nodes, int32_t* cell_edges,
                                uint32_t* edge_bits, int32_t* err_flag, void* stream) {
  if (nc == 0) return 0;
  k_pack_edges3d<<<pgb_grid(nc, 256), 256, 0, (cudaStream_t)stream>>>(
      nc, cf_idx, fr_idx, fr_sgn, rp_idx, cell_nodes, cell_edges, edge_bits, err_flag);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_pack_edges2d(int64_t nc, const int32_t* cf_idx, const int32_t* fr_idx,
                                const int32_t* cell_nodes, uint32_t* edge_bits, void* stream) {
  if (nc == 0) return 0;
  k_pack_edges2d<<<pgb_grid(nc, 256), 256, 0, (cudaStream_t)stream>>>(nc, cf_idx, fr_idx,
                                                                      cell_nodes, edge_bits);
  PGB_LAUNCH_CHECK();
  return 0;
}

// Lagrange2 cell -> dof table: the cell's nodes in slot order, then num_nodes + the edge the
// reference attaches to the local pair (p, q) (Lagrange2.get_edge_dof_indices, fem/h1.py:507-543):
// 2D: the faces in reverse slot order; 3D: the six ridge ids sorted ascending, permuted [5,4,2,3,1,0].
template <int D>
__global__ void __launch_bounds__(256) k_pack_l2_dofs(int64_t nc, int32_t nn, const int32_t* __restrict__ cell_nodes,
                                                      const int32_t* __restrict__ cell_edges,
                                                      int32_t* __restrict__ dofs) {
  constexpr int N = D + 1, E = D * (D + 1) / 2, L = N + E;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t e[E];
#pragma unroll
  for (int k = 0; k < E; ++k) e[k] = cell_edges[c * E + k];
  int32_t out[E];
  if constexpr (D == 2) {
    out[0] = e[2], out[1] = e[1], out[2] = e[0];
  } else {
    // sorting network for 6 keys (12 compare-exchanges)
#define PGB_CE(i, j)              \
  {                               \
    int32_t lo = min(e[i], e[j]); \
    int32_t hi = max(e[i], e[j]); \
    e[i] = lo, e[j] = hi;         \
  }
    PGB_CE(0, 5) PGB_CE(1, 3) PGB_CE(2, 4) PGB_CE(1, 2) PGB_CE(3, 4) PGB_CE(0, 3) PGB_CE(2, 5) PGB_CE(0, 1)
    PGB_CE(2, 3) PGB_CE(4, 5) PGB_CE(1, 2) PGB_CE(3, 4)
#undef PGB_CE
    out[0] = e[5], out[1] = e[4], out[2] = e[2], out[3] = e[3], out[4] = e[1], out[5] = e[0];
  }
#pragma unroll
  for (int a = 0; a < N; ++a) dofs[c * L + a] = cell_nodes[c * N + a];
#pragma unroll
  for (int k = 0; k < E; ++k) dofs[c * L + N + k] = nn + out[k];
}

extern "C" int pgb_pack_l2_dofs(int dim, int64_t nc, int64_t nn, const int32_t* cell_nodes,
                                const int32_t* cell_edges, int32_t* dofs, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_pack_l2_dofs: dim must be 2 or 3");
  if (nc == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3)
    k_pack_l2_dofs<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, (int32_t)nn, cell_nodes, cell_edges, dofs);
  else
    k_pack_l2_dofs<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, (int32_t)nn, cell_nodes, cell_edges, dofs);
  PGB_LAUNCH_CHECK();
  return 0;
}

// RT1 dofs: rank the cell's nodes by id (slot a = node opposite face slot a), see pgb.h
template <int D>
__global__ void __launch_bounds__(256) k_pack_rt1(int64_t nc, int32_t nf, const int32_t* __restrict__ cf_idx,
                                                  const int32_t* __restrict__ cell_nodes,
                                                  const uint8_t* __restrict__ sign_bits, int32_t* __restrict__ dofs,
                                                  uint32_t* __restrict__ aux) {
  constexpr int N = D + 1, L = D * (D + 2);
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t node[N], face[N];
#pragma unroll
  for (int a = 0; a < N; ++a) {
    node[a] = cell_nodes[c * N + a];
    face[a] = cf_idx[c * N + a];
  }
  const unsigned sb = sign_bits[c];
  int32_t face_of_rank[N];
  unsigned bits = 0;
#pragma unroll
  for (int a = 0; a < N; ++a) {
    int r = 0;
#pragma unroll
    for (int b = 0; b < N; ++b) r += (node[b] < node[a]) ? 1 : 0;
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (q == r) face_of_rank[q] = face[a];
    bits |= (unsigned)a << (2 * r);
    bits |= ((sb >> a) & 1u) << (8 + r);
  }
  aux[c] = bits;
#pragma unroll
  for (int q = 0; q < D * N; ++q) dofs[c * L + q] = face_of_rank[D - q % N] + (q / N) * nf;
#pragma unroll
  for (int k = 0; k < D; ++k) dofs[c * L + D * N + k] = (int32_t)((int64_t)D * nf + (int64_t)k * nc + c);
}

extern "C" int pgb_pack_rt1(int dim, int64_t nc, int64_t nf, const int32_t* cf_idx, const int32_t* cell_nodes,
                            const uint8_t* sign_bits, int32_t* cell_dofs, uint32_t* aux_bits, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_pack_rt1: dim must be 2 or 3");
  PGB_REQUIRE((int64_t)dim * (nf + nc) < ((int64_t)1 << 31), "pgb_pack_rt1: dof ids exceed int32");
  if (nc == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3)
    k_pack_rt1<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, (int32_t)nf, cf_idx, cell_nodes, sign_bits, cell_dofs, aux_bits);
  else
    k_pack_rt1<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, (int32_t)nf, cf_idx, cell_nodes, sign_bits, cell_dofs, aux_bits);
  PGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int pgb_prep_coords(int dim, int64_t nn, const double* nodes, const double* R_host,
                               double* coords, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_prep_coords: dim must be 2 or 3");
  if (nn == 0) return 0;
  Rot R;
  for (int i = 0; i < 9; ++i) R.r[i] = (i < dim * 3) ? R_host[i] : 0.0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 3)
    k_prep_coords<3><<<pgb_grid(nn, 256), 256, 0, s>>>(nn, nodes, R, coords);
  else
    k_prep_coords<2><<<pgb_grid(nn, 256), 256, 0, s>>>(nn, nodes, R, coords);
  PGB_LAUNCH_CHECK();
  return 0;
}
