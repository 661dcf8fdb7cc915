// Grid geometry on the device: the metric quantities pp.Grid.compute_geometry
// and pg.Grid.compute_edge_properties (grids/grid.py:39-64,311-337) provide for simplicial grids.
// One thread per face / cell / edge; all arrays in the grid object's own layout ((3, n) row-major).
// Not on the assembly path (K1 recomputes |c| and the gradients from the coordinates); this is the
// input-provider side, so that a 10^8-cell grid never needs its geometry computed by numpy.
#include "pgb_common.cuh"

namespace {

struct P3 {
  double x, y, z;
};
__device__ __forceinline__ P3 node(const double* __restrict__ nodes, int64_t nn, int32_t i) {
  return {__ldg(nodes + i), __ldg(nodes + nn + i), __ldg(nodes + 2 * nn + i)};
}

// normal tied to the stored node order: 2D (t_y, -t_x, 0), t = x1 - x0; 3D (x1-x0) x (x2-x0) / 2
template <int D>
__global__ void __launch_bounds__(256) k_face_geometry(int64_t nf, int64_t nn, const double* __restrict__ nodes,
                                                       const int32_t* __restrict__ fn, double* __restrict__ normals,
                                                       double* __restrict__ areas, double* __restrict__ centers) {
  int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (f >= nf) return;
  P3 p[D];
#pragma unroll
  for (int k = 0; k < D; ++k) p[k] = node(nodes, nn, fn[f * D + k]);
  double nx, ny, nz;
  if constexpr (D == 2) {
    nx = p[1].y - p[0].y;
    ny = -(p[1].x - p[0].x);
    nz = 0.0;
  } else {
    double ax = p[1].x - p[0].x, ay = p[1].y - p[0].y, az = p[1].z - p[0].z;
    double bx = p[2].x - p[0].x, by = p[2].y - p[0].y, bz = p[2].z - p[0].z;
    nx = 0.5 * (ay * bz - az * by);
    ny = 0.5 * (az * bx - ax * bz);
    nz = 0.5 * (ax * by - ay * bx);
  }
  normals[f] = nx;
  normals[nf + f] = ny;
  normals[2 * nf + f] = nz;
  areas[f] = sqrt(nx * nx + ny * ny + nz * nz);
  double cx = p[0].x, cy = p[0].y, cz = p[0].z;
#pragma unroll
  for (int k = 1; k < D; ++k) {
    cx += p[k].x;
    cy += p[k].y;
    cz += p[k].z;
  }
  centers[f] = cx / D;
  centers[nf + f] = cy / D;
  centers[2 * nf + f] = cz / D;
}

// centroid and volume from the cell's nodes (summed in ascending node id, the order of
// cell_nodes().indices); err_flag is raised when a cell_faces sign disagrees with the side of the
// cell the node-order normal points to (SURVEY 8(c) "orientation trap").
template <int D>
__global__ void __launch_bounds__(256) k_cell_geometry(int64_t nc, int64_t nf, int64_t nn,
                                                       const double* __restrict__ nodes,
                                                       const int32_t* __restrict__ cell_nodes,
                                                       const int32_t* __restrict__ cf_idx,
                                                       const int8_t* __restrict__ cf_sgn,
                                                       const double* __restrict__ normals,
                                                       const double* __restrict__ fcenters,
                                                       double* __restrict__ centers, double* __restrict__ volumes,
                                                       int32_t* __restrict__ err_flag) {
  int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= nc) return;
  int32_t id[D + 1];
#pragma unroll
  for (int a = 0; a <= D; ++a) id[a] = cell_nodes[c * (D + 1) + a];
#pragma unroll
  for (int i = 0; i <= D; ++i)  // tiny sorting network (insertion, fully unrolled)
#pragma unroll
    for (int j = 0; j < D - i; ++j) {
      int32_t lo = min(id[j], id[j + 1]), hi = max(id[j], id[j + 1]);
      id[j] = lo;
      id[j + 1] = hi;
    }
  P3 p[D + 1];
#pragma unroll
  for (int a = 0; a <= D; ++a) p[a] = node(nodes, nn, id[a]);
  double cx = p[0].x, cy = p[0].y, cz = p[0].z;
#pragma unroll
  for (int a = 1; a <= D; ++a) {
    cx += p[a].x;
    cy += p[a].y;
    cz += p[a].z;
  }
  cx /= (D + 1);
  cy /= (D + 1);
  cz /= (D + 1);
  centers[c] = cx;
  centers[nc + c] = cy;
  centers[2 * nc + c] = cz;
  double ax = p[1].x - p[0].x, ay = p[1].y - p[0].y, az = p[1].z - p[0].z;
  double bx = p[2].x - p[0].x, by = p[2].y - p[0].y, bz = p[2].z - p[0].z;
  if constexpr (D == 2) {
    volumes[c] = 0.5 * fabs(ax * by - ay * bx);
  } else {
    double ex = p[3].x - p[0].x, ey = p[3].y - p[0].y, ez = p[3].z - p[0].z;
    double det = ax * (by * ez - bz * ey) + ay * (bz * ex - bx * ez) + az * (bx * ey - by * ex);
    volumes[c] = fabs(det) / 6.0;
  }
  if (cf_sgn) {
#pragma unroll
    for (int a = 0; a <= D; ++a) {
      int64_t f = cf_idx[c * (D + 1) + a];
      double out = normals[f] * (fcenters[f] - cx) + normals[nf + f] * (fcenters[nf + f] - cy) +
                   normals[2 * nf + f] * (fcenters[2 * nf + f] - cz);
      int so = (out > 0.0) - (out < 0.0);
      int ss = cf_sgn[c * (D + 1) + a];
      ss = (ss > 0) - (ss < 0);
      if (so != ss) *err_flag = 1;
    }
  }
}

// tangent = s0 x[n0] + s1 x[n1] (nodes @ edge_nodes, grid.py:331-337); sgn null: (-1, +1)
__global__ void __launch_bounds__(256) k_edge_geometry(int64_t ne, int64_t nn, const double* __restrict__ nodes,
                                                       const int32_t* __restrict__ en, const int8_t* __restrict__ sgn,
                                                       double* __restrict__ tangents, double* __restrict__ lengths) {
  int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= ne) return;
  P3 a = node(nodes, nn, en[2 * e]), b = node(nodes, nn, en[2 * e + 1]);
  double s0 = sgn ? (double)sgn[2 * e] : -1.0, s1 = sgn ? (double)sgn[2 * e + 1] : 1.0;
  double tx = s0 * a.x + s1 * b.x, ty = s0 * a.y + s1 * b.y, tz = s0 * a.z + s1 * b.z;
  tangents[e] = tx;
  tangents[ne + e] = ty;
  tangents[2 * ne + e] = tz;
  lengths[e] = sqrt(tx * tx + ty * ty + tz * tz);
}

}  // namespace

extern "C" int pgb_grid_geometry(int dim, int64_t nc, int64_t nf, int64_t nn, const double* nodes,
                                 const int32_t* cf_idx, const int8_t* cf_sgn, const int32_t* fn_idx,
                                 const int32_t* cell_nodes, double* face_normals, double* face_areas,
                                 double* face_centers, double* cell_centers, double* cell_volumes,
                                 int32_t* err_flag, void* stream) {
  PGB_REQUIRE(dim == 2 || dim == 3, "pgb_grid_geometry: dim must be 2 or 3");
  cudaStream_t s = (cudaStream_t)stream;
  if (nf > 0) {
    if (dim == 3)
      k_face_geometry<3><<<pgb_grid(nf, 256), 256, 0, s>>>(nf, nn, nodes, fn_idx, face_normals, face_areas, face_centers);
    else
      k_face_geometry<2><<<pgb_grid(nf, 256), 256, 0, s>>>(nf, nn, nodes, fn_idx, face_normals, face_areas, face_centers);
    PGB_LAUNCH_CHECK();
  }
  if (nc > 0) {
    if (dim == 3)
      k_cell_geometry<3><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, nf, nn, nodes, cell_nodes, cf_idx, cf_sgn, face_normals,
                                                          face_centers, cell_centers, cell_volumes, err_flag);
    else
      k_cell_geometry<2><<<pgb_grid(nc, 256), 256, 0, s>>>(nc, nf, nn, nodes, cell_nodes, cf_idx, cf_sgn, face_normals,
                                                          face_centers, cell_centers, cell_volumes, err_flag);
    PGB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int pgb_edge_geometry(int64_t ne, int64_t nn, const double* nodes, const int32_t* edge_nodes,
                                 const int8_t* edge_signs, double* edge_tangents, double* edge_lengths,
                                 void* stream) {
  if (ne == 0) return 0;
  k_edge_geometry<<<pgb_grid(ne, 256), 256, 0, (cudaStream_t)stream>>>(ne, nn, nodes, edge_nodes, edge_signs,
                                                                     edge_tangents, edge_lengths);
  PGB_LAUNCH_CHECK();
  return 0;
}
