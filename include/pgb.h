/*
 * pgb.h - C ABI of libpgb.so, the B200 (sm_100a) FEM assembly-and-solve hot path.
 *
 * The reference (compgeo-mox/pygeon) is pure Python and has no FFI: its "plugin"
 * boundary is the pg.Discretization method signatures and the `solver` callable
 * of pg.LinearSystem.solve.  The entry points below are what a ctypes binding
 * placed behind those methods calls; each one names the
 * reference code it replaces (paths relative to /root/reference/src/pygeon).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the only allocations made here are the Krylov drivers'
 *     residual history (at most PGB_RESID_CAP doubles per device), a few KB of scalars and a
 *     mapped pinned stop flag;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *     unless stated otherwise; return value 0 = ok, non-zero = error with
 *     pgb_last_error() describing it (thread-local);
 *   - threading: one host thread per device (the model is one process per GPU).  Scratch of the Krylov drivers
 *     and function-attribute caches are kept per (thread, device); the symbolic phase keeps the row ranges of
 *     the last pgb_csr_symbolic_sort call in process-wide state (symbolic_finish must follow on the same
 *     thread), so concurrent symbolic calls from several host threads are not supported;
 *   - ids are int32, values fp64, COO positions uint32 (n_coo < 2^32).
 *   - "slot a" of a cell is the a-th entry of its cell_faces column; local node a
 *     is the node opposite face slot a.
 */
#ifndef PGB_H
#define PGB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int pgb_version(void);
const char* pgb_last_error(void);
/* cudaLimitMaxL2FetchGranularity hint (32 / 64 / 128 bytes) for the current device. */
int pgb_set_l2_fetch_granularity(int bytes);
int pgb_get_l2_fetch_granularity(void);

/* ---- grid packing ------------------------------------------------------------
 * Replaces Grid.compute_opposite_nodes (grids/grid.py:279-309),
 * PwLinears.get_dof_lookup_array (discretizations/fem/l2.py:598-612) and the
 * fancy-index gathers of BDM1/Nedelec1.proj_to_PwPolynomials
 * (fem/hdiv.py:470-507, fem/hcurl.py:232-289).
 *
 *   cf_idx [nc*(d+1)]  cell_faces.indices          cf_sgn [nc*(d+1)] cell_faces.data as int8
 *   fn_idx [nf*d]      face_nodes.indices
 * out:
 *   cell_nodes [nc*(d+1)]  node opposite each face slot
 *   sign_bits  [nc]        bit a set <=> cell_faces sign of slot a is -1
 *   err_flag   [1]         set to 1 if some cell is not a simplex (NotImplementedError
 *                          "Grid is not simplicial", grid.py:298-301)
 */
int pgb_pack_cells(int dim, int64_t nc, const int32_t* cf_idx, const int8_t* cf_sgn,
                   const int32_t* fn_idx, int32_t* cell_nodes, uint8_t* sign_bits,
                   int32_t* err_flag, void* stream);
/* cf_sgn / sign_bits of pgb_pack_cells may be NULL; the sign bits can then be produced later by
 * pgb_pack_signs (Lagrange1 never needs them, so cell_faces.data is not even uploaded). */
int pgb_pack_signs(int dim, int64_t nc, const int8_t* cf_sgn, uint8_t* sign_bits, void* stream);

/* BDM1 dofs (fem/hdiv.py:474-480): dof of (face slot a, pos) = f_a + pos*nf, pos = slot of
 * the node in face_nodes.indices of f_a.  cell_dofs [nc*d*(d+1)], jslot_bits[nc]: 2 bits per
 * (a,pos) = local node slot of that face node. */
int pgb_pack_bdm1(int dim, int64_t nc, int64_t nf, const int32_t* cf_idx, const int32_t* fn_idx,
                  const int32_t* cell_nodes, int32_t* cell_dofs, uint32_t* jslot_bits,
                  void* stream);

/* Nedelec0 edges.  3D (fem/hcurl.py:232-246, grids/grid.py:149-216): edge id of local pair
 * (p,q), p<q in the order (01)(02)(03)(12)(13)(23), looked up in face_ridges / ridge_peaks.
 * edge_bits: bit e (0..5) set <=> ridge_peaks order is (q,p) (Whitney form flipped);
 * bits 8+4*... see pgb_pack.cu: 2 curl-sign bits per edge (sign of the edge in its two faces).
 * 2D: the edge of face slot a joins the two other slots; bit a set <=> face_ridges.indices
 * order is (hi slot, lo slot).  fr_* = face_ridges (indices, data as int8), rp_idx =
 * ridge_peaks.indices. */
int pgb_pack_edges3d(int64_t nc, const int32_t* cf_idx, const int32_t* fr_idx,
                     const int8_t* fr_sgn, const int32_t* rp_idx, const int32_t* cell_nodes,
                     int32_t* cell_edges, uint32_t* edge_bits, int32_t* err_flag, void* stream);
int pgb_pack_edges2d(int64_t nc, const int32_t* cf_idx, const int32_t* fr_idx,
                     const int32_t* cell_nodes, uint32_t* edge_bits, void* stream);

/* Lagrange2 cell -> dof table [nc][(d+1) + d(d+1)/2]: nodes in slot order, then nn + the edge ids in
 * the order of Lagrange2.get_edge_dof_indices (fem/h1.py:507-543).  cell_edges: [nc][3] face ids (2D,
 * = cell_faces indices) or [nc][6] ridge ids of the cell in any order (3D, from pgb_pack_edges3d). */
int pgb_pack_l2_dofs(int dim, int64_t nc, int64_t nn, const int32_t* cell_nodes,
                     const int32_t* cell_edges, int32_t* dofs, void* stream);
/* RT1 cell -> dof table [nc][d(d+2)] in the reference's local order (RT1.local_dofs_of_cell +
 * reorder_faces, fem/hdiv.py:534-552,629-674): with the cell's nodes ranked by ascending id, dof
 * q < d(d+1) = (face opposite the node of rank d - q % (d+1)) + (q / (d+1)) * nf, then
 * d*nf + k*nc + c for k < d.  aux_bits[c]: bits 2r..2r+1 = face slot whose opposite node has rank r,
 * bit 8+r = 1 when the cell_faces sign of that face is -1. */
int pgb_pack_rt1(int dim, int64_t nc, int64_t nf, const int32_t* cf_idx, const int32_t* cell_nodes,
                 const uint8_t* sign_bits, int32_t* cell_dofs, uint32_t* aux_bits, void* stream);
/* coords[nn][4] (3D, padded) or [nn][2] (2D) = R @ nodes, nodes given as (3, nn) row-major
 * (sd.nodes), R = sd.rotation_matrix (d x 3, host pointer). */
int pgb_prep_coords(int dim, int64_t nn, const double* nodes, const double* R_host,
                    double* coords, void* stream);

/* ---- grid connectivity / geometry on the device --------------
 * Ridges of a 3D simplicial grid, replacing Grid._compute_ridges_3d (grids/grid.py:149-216):
 * ridges = the distinct sorted node pairs of all faces numbered lexicographically
 * (coo sum_duplicates order, grid.py:187-198); face f holds the ridges (n0,n1), (n1,n2), (n2,n0) of
 * its node order in face_nodes.indices with orientation sign(next - prev) (grid.py:172-182).
 *   face_nodes      [nf][3]   face_nodes.indices
 *   face_ridges     [nf][3]   out: face_ridges.indices (slot order = face node order)
 *   face_ridge_sign [nf][3]   out: face_ridges.data as int8
 *   ridge_peaks     [3*nf][2] out (capacity): ridge_peaks.indices = (lo, hi) node of every ridge
 *                             (ridge_peaks.data = (-1, +1)); the first *num_ridges_host rows are valid
 * Synchronises the stream (the ridge count is returned to the host).  Error (return 2) when a face
 * repeats a node or a node id lies outside [0, num_nodes). */
int64_t pgb_ridges3d_workspace_bytes(int64_t nf);
int pgb_ridges3d(int64_t nf, int64_t num_nodes, const int32_t* face_nodes, int32_t* face_ridges,
                 int8_t* face_ridge_sign, int32_t* ridge_peaks, int64_t* num_ridges_host,
                 void* workspace, int64_t workspace_bytes, void* stream);
/* Metric quantities of pp.Grid.compute_geometry for simplices (called from grids/grid.py:59), all
 * (3, n) row-major like the grid attributes: face_normals (tied to the stored node order: 2D
 * (t_y, -t_x, 0), 3D (x1-x0) x (x2-x0) / 2), face_areas, face_centers, cell_centers, cell_volumes.
 * cell_nodes = the table of pgb_pack_cells.  cf_sgn may be NULL; otherwise *err_flag is set to 1 when
 * a cell_faces sign is not the side of the cell the normal points to. */
int pgb_grid_geometry(int dim, int64_t nc, int64_t nf, int64_t nn, const double* nodes,
                      const int32_t* cf_idx, const int8_t* cf_sgn, const int32_t* fn_idx,
                      const int32_t* cell_nodes, double* face_normals, double* face_areas,
                      double* face_centers, double* cell_centers, double* cell_volumes,
                      int32_t* err_flag, void* stream);
/* Grid.compute_edge_properties (grids/grid.py:311-337): edge_tangents (3, ne) = nodes @ edge_nodes,
 * edge_lengths.  edge_nodes [ne][2] = ridge_peaks.indices (3D) / face_ridges.indices (2D);
 * edge_signs [ne][2] = the matching .data as int8, or NULL for (-1, +1). */
int pgb_edge_geometry(int64_t ne, int64_t nn, const double* nodes, const int32_t* edge_nodes,
                      const int8_t* edge_signs, double* edge_tangents, double* edge_lengths,
                      void* stream);

/* ---- K1: per-cell local matrices ------------------------------------------------
 * One thread per cell, fp64.  vals is [nc][L*L] row-major (local row a, local col b).
 * Replaces the implicit local matrices of Discretization._assemble_inner_product
 * (discretizations/discretization.py:112-136: Pi.T @ M @ Pi) and assemble_stiff_matrix
 * (:154-183: B.T @ A @ B); closed forms in SURVEY.md section 8(a).
 *   K      (3,3,nc) row-major = SecondOrderTensor.values, or NULL (identity)
 *   R_host d x 3 rotation matrix (host)      w  [nc] or NULL (1)
 * The matrix produced is A[a][b] = u_a^T (R K R^T) u_b, whose CSR arrays are the CSC arrays
 * of the reference's matrix (its contraction at fem/vec_l2.py:113-114 yields R K^T R^T).
 */
enum {
  PGB_L1_MASS = 0,      /* fem/h1.py:218-237 + fem/l2.py:63-86,457-470     L = d+1      */
  PGB_L1_STIFF = 1,     /* (K grad u, grad v): 3D stiff and assemble_grad_grad_matrix h1.py:39-63 */
  PGB_L1_STIFF_ROT = 2, /* 2D assemble_stiff_matrix = (K rot grad u, rot grad v), h1.py:46-48     */
  PGB_RT0_MASS = 3,     /* fem/hdiv.py:255-274,450-509                      L = d+1      */
  PGB_BDM1_MASS = 4,    /* fem/hdiv.py:450-509                              L = d(d+1)   */
  PGB_N0_MASS = 5,      /* fem/hcurl.py:44-61,215-291                       L = d(d+1)/2 */
  PGB_N0_CURLCURL = 6,  /* fem/hcurl.py:82-106 via discretization.py:154-183             */
  PGB_P0_MASS = 7,      /* fem/l2.py:335-353: diag(w/|c|)                   L = 1        */
  PGB_RT0_DIVDIV = 8,   /* RT0 stiff = cell_faces diag(w/|c|) cell_faces^T (hdiv.py:177-192 via
                           discretization.py:154-183); also Nedelec0 stiff in 2D     L = d+1 */
  PGB_BDM1_DIVDIV = 9,  /* BDM1 stiff, div = div_RT0 @ hstack([I]*d)/d (hdiv.py:353-371)  L = d(d+1) */
  PGB_DARCY_SADDLE = 10, /* symmetrised mixed-Darcy block system [[M, -B^T], [-B, 0]] with M = RT0 mass,
                           B = P0mass @ div (tutorials/fluid_flow/darcy.ipynb cell 10,
                           tests/patch_tests/test_laplacian_mixed.py:50), assembled in one pass:
                           local dofs = (faces of the cell, num_faces + cell)        L = d+2 */
  PGB_L1_LUMPED = 11,   /* Lagrange1.assemble_lumped_matrix (discretization.py:91-110 with the lumped
                           broken-P1 mass eye/(d+1), l2.py:472-483): diag w|c|/(d+1)  L = d+1 */
  PGB_RT0_LUMPED = 12,  /* RT0.assemble_lumped_matrix (hdiv.py:140-175): diag sum_c (dist^T K dist)/|dist|
                           / |f|, dist = face centre - cell centre                   L = d+1 */
  PGB_BDM1_LUMPED = 13, /* BDM1 lumped: w|c| delta_jl/(d+1) t_aj^T K t_bl            L = d(d+1) */
  PGB_L2_MASS = 14,     /* Lagrange2 mass (h1.py:352-392 through discretization.py:112-136): w|c| (phi_i, phi_j),
                           phi = lambda_i(2 lambda_i - 1) for the d+1 nodes, 4 lambda_p lambda_q for the edges
                           (p,q) = (0,1),(0,2)[,(0,3)],(1,2)[,(1,3),(2,3)]              L = (d+1) + d(d+1)/2 */
  PGB_L2_STIFF = 15,    /* Lagrange2 stiffness (h1.py:545-599): |c| sum_jl Mhat_jl grad phi_i(x_j)^T K
                           grad phi_k(x_l); pg.WEIGHT is not read (as the reference)   L = (d+1) + d(d+1)/2 */
  PGB_RT1_MASS = 16,    /* RT1 mass (hdiv.py:555-621, a Python loop over the cells): |c| Psi (M2 x K) Psi^T with the
                           P2 local mass M2 and the nodal values Psi of the basis (hdiv.py:676-744), evaluated as a
                           constant contraction of the Gram matrix of the cell's edge vectors (pgb_rt1_table.cuh);
                           pg.WEIGHT is not read; aux_bits from pgb_pack_rt1            L = d(d+2) */
  PGB_RT1_DIVDIV = 17,  /* RT1 stiffness (discretization.py:154-183 with the divergence of hdiv.py:813-865 into
                           PwLinears and its mass, l2.py:63-86,457-470): w / (d^2 |c|) s_p s_q (Dv^T Mhat Dv)_pq with
                           the constant nodal values Dv of the local divergences; aux_bits from pgb_pack_rt1
                                                                                        L = d(d+2) */
  PGB_RT1_LUMPED_CELL = 18 /* cell block of RT1.assemble_lumped_matrix (hdiv.py:970-1028): |c| (d+1)/(d+2) P^T K P with
                           the cell functions at the cell centre P_i = (sum_j x_j - (d+1) x_i) / ((d+1)^2 d |c|);
                           dofs (k, c) -> k * nc + c; aux_bits from pgb_pack_rt1                 L = d */
};
int pgb_local_size(int kind, int dim); /* L */
/* dst_pos (nullable): [nc*L] destination row of every local row (the inc_pos of a mode-0 plan);
 * local row a of cell c is then written to vals[dst_pos[c*L+a]*LP .. +L) instead of
 * vals[(c*L+a)*L .. +L), with the row pitch LP = L rounded up to a multiple of 4 (rows are whole
 * 32-byte sectors; vals then holds nc*L*LP doubles, the slots array of the plan nc*L*LP bytes). */
int pgb_local_matrices(int kind, int dim, int64_t nc, const int32_t* cell_nodes,
                       const uint8_t* sign_bits, const uint32_t* aux_bits, const double* coords,
                       const double* K, const double* R_host, const double* w, const uint32_t* dst_pos,
                       double* vals, void* stream);

/* ---- K2: COO -> CSR, deterministic, no floating-point atomics --------------------
 * Replaces scipy's COO->CSC constructor and SpGEMM duplicate summation
 * (e.g. fem/hdiv.py:509, fem/hcurl.py:291, fem/h1.py:237, discretization.py:136).
 * COO entry i = cell*L*L + a*L + b has row cell_dofs[cell*L+a], col cell_dofs[cell*L+b].
 *
 * Two symbolic algorithms produce the same CSR pattern (indptr, indices):
 *   mode 0 (two level, default): (1) transpose of the cell->dof table by counting: integer
 *     atomicAdd per incidence (cell, a) on its dof's counter, scan -> row_start [n_rows+1], every
 *     incidence dropped into its row block; the per-row kernel of step (2) sorts each block
 *     ascending, i.e. into the order of a stable sort by dof id, so everything downstream is
 *     independent of the order in which the atomics ran.  Returned: inc_pos [nc*L] (position of
 *     incidence cell*L + a in (dof, cell) order: where K1 must put that local row) and row_start.
 *     (2) per row: gather the candidate columns of the incident cells, find the sorted distinct
 *     columns and record for every candidate the index of its column inside the row
 *     (slots, uint8, slots[(row_start[r] + q)*LP + b], LP = L rounded up to a multiple of 4).
 *     Four kernels, picked by row length (per range of 4096-row chunks, or per size-class row list
 *     when long rows are scattered; a few long rows do not slow the others down): one thread per
 *     row with a register sorting network (<= 16 incidences, <= 64 candidates), a group of 8..32
 *     lanes per row with a shared-memory hash dedupe (<= 256 candidates), a sub-warp bitonic sort
 *     of all candidates (<= 512), one block per row (<= 2048 incidences).  (3) a scan of the row
 *     counts and a compaction.
 *     Needs every dof to be shared by at most 2048 cells and every row to have at most 255
 *     entries; otherwise returns 3 and the caller retries with mode 1.
 *     Test modes: 2 = 64-bit (col, position) keys forced in the row sorts, 3 = the sub-warp bitonic
 *     row sort forced, 4 = step (1) by a stable LSD radix sort of the incidences (32-bit keys)
 *     instead of counting.  All give bit-identical inc_pos / slots / pattern.
 *   mode 1 (full sort): stable LSD radix sort of all nc*L*L packed (row,col) keys (64-bit) with
 *     the COO position as payload (perm [n_coo]), then head-flag compaction (seg [nnz+1]).
 * symbolic_sort: workspace of pgb_csr_workspace_bytes(nc, L, n_rows, mode) bytes.
 *   perm_or_inc: perm (mode 1) or inc_pos (mode 0).  nnz_host / max_row_nnz_host are returned
 *   after a stream synchronisation.
 * symbolic_finish: must be the next call after symbolic_sort on the same workspace (it reads the
 *   row ranges chosen there).  indptr (int32 if idx64==0 else int64) [n_rows+1], indices [nnz]; mode 1 also
 *   seg [nnz+1] (start of each entry's run in perm).
 * numeric_rows (mode 0 plans): K1 has written every local row to its row-sorted position
 *   (pgb_local_matrices with dst_pos = inc_pos), so a CSR row owns a contiguous block of vals_t;
 *   one thread per row streams it with 256-bit loads and adds each value into the accumulator of
 *   its column slot in shared memory.  No gathers; 9 B read per COO entry.
 * numeric (mode 1 plans): data[k] = sum over its run of vals[perm[.]]: lane-sequential partial
 *   sums combined by a warp-shuffle segmented scan.
 * Both sum in a fixed order (bit-reproducible, no floating-point atomics).
 */
int64_t pgb_csr_workspace_bytes(int64_t nc, int L, int64_t n_rows, int mode);
/* Optional CUDA-event timing of the phases of the two-level symbolic_sort (off by default):
 * out[0] counting (k_inc_count, scan, max, host read with the placement kernel queued behind it; mode 4: the
 * radix sort), [1] rest of the placement (+ k_row_canon when the row kernel does not sort the incidences itself;
 * mode 4: row offsets),
 * [2] per-row kernel, [3] scan of the row counts. */
void pgb_set_profile(int on);
void pgb_get_phase_ms(double* out, int n);
/* Names of the per-row kernels the last two-level pgb_csr_symbolic_sort launched, joined by '+',
 * e.g. "k_rowhash<32,1,10,512>+k_rowthread<8,10,64,u32>" (template arguments as in the source);
 * "" before the first call or after a mode-1 call.  For tests and profiles. */
const char* pgb_csr_last_row_kernel(void);
int pgb_csr_symbolic_sort(int64_t n_rows, int64_t nc, int L, const int32_t* cell_dofs, int mode,
                          void* workspace, int64_t workspace_bytes, uint32_t* perm_or_inc,
                          uint8_t* slots, uint32_t* row_start, int64_t* nnz_host,
                          int* max_row_nnz_host, void* stream);
/* nnz_host == NULL: pgb_csr_symbolic_sort returns without waiting for the device (apart from its one small read of
 * the chunk maxima, behind which the placement kernel is already queued); inc_pos / slots / row_start are
 * complete in stream order, so the caller may queue K1 (pgb_local_matrices with dst_pos = inc_pos) at once and
 * collect nnz / max_row_nnz afterwards with pgb_csr_symbolic_result (waits for an event recorded before that
 * extra work; same return codes as the synchronous call, 3 = fall back to mode 1).  One pending call per device. */
int pgb_csr_symbolic_result(int64_t* nnz_host, int* max_row_nnz_host);
int pgb_csr_symbolic_finish(int64_t n_rows, int64_t nc, int L, int mode, int64_t nnz,
                            const void* workspace, const uint32_t* row_start, void* indptr, int idx64,
                            int32_t* indices, uint32_t* seg, void* stream);
/* mode 0 plans: vals_t holds the local rows in row-sorted order (K1 called with dst_pos = inc_pos).
 * flags bit 0: transpose the rows' entries through shared memory and write them with coalesced
 * stores (faster when nnz is comparable to n_coo; same values either way).
 * flags bit 1: bring the contiguous local rows of every block of 128 CSR rows into shared memory with
 * the bulk-async copy engine (cp.async.bulk + mbarrier, 3-stage ring) instead of per-thread loads; same
 * summation order, bit-identical values.
 * flags bits 2-5: G = 2, 4 or 8 lanes per CSR row (0 = one thread per row): lane j sums the local rows
 * j, j + G, ... of the row, the G partial sums of every column are then added in lane order; a warp-wide
 * load covers G consecutive local rows per row (coalesced), consecutive lanes store consecutive entries.
 * A fixed, reproducible summation order that differs from the one-thread-per-row order in the last bits.
 * flags bit 6: one thread per row with 2 instead of 4 local rows in flight and 75 % occupancy (A/B variant,
 * measured slower). */
int pgb_csr_numeric_rows(int64_t n_rows, int L, int max_row_nnz, const uint32_t* row_start,
                         const uint8_t* slots, const void* indptr, int idx64, const double* vals_t,
                         double* data, int flags, void* stream);
int pgb_csr_numeric(int64_t nnz, const uint32_t* perm, const uint32_t* seg, const double* vals,
                    double* data, void* stream);

/* ---- system build around the assembled matrix, in HBM ----------------------
 * Replaces the host SpGEMMs of LinearSystem.reduce_system (numerics/linear_system.py:64-77:
 * R_0 @ A @ R_0.T with the restriction matrix of :115-127) and serves the owned-row extraction of the
 * cell-partitioned multi-GPU path.  Count -> scan -> fill: scratch is O(rows), nothing per entry.
 *   workspace: pgb_system_workspace_bytes(n) bytes for any call below working on n rows.
 * pgb_mask_to_map: newid[i] = rank of i among the entries with mask[i] != 0, or -1 (nullable);
 *   rows[rank] = i (nullable); *count_host = number kept (stream synchronised when non-NULL).
 * pgb_csr_extract_symbolic: out row i = input row rows[i] (rows == NULL: row i), entries whose
 *   column has col_map[col] < 0 dropped (col_map == NULL: keep all): writes the int64 row pointer
 *   out_indptr[n_out + 1], *nnz_host after a synchronisation.
 * pgb_csr_extract_fill: the kept entries in their original order, columns renumbered through col_map. */
int64_t pgb_system_workspace_bytes(int64_t n);
int pgb_mask_to_map(int64_t n, const uint8_t* mask, int32_t* newid, int32_t* rows, int64_t* count_host,
                    void* workspace, int64_t workspace_bytes, void* stream);
int pgb_csr_extract_symbolic(int64_t n_out, const int32_t* rows, const void* indptr, int idx64,
                             const int32_t* indices, const int32_t* col_map, int64_t* out_indptr,
                             int64_t* nnz_host, void* workspace, int64_t workspace_bytes, void* stream);
int pgb_csr_extract_fill(int64_t n_out, const int32_t* rows, const void* indptr, int idx64,
                         const int32_t* indices, const double* data, const int32_t* col_map,
                         const int64_t* out_indptr, int32_t* out_indices, double* out_data, void* stream);
/* flags[col] = 1 for every column referenced by the rows `rows` (NULL: all n_out rows): which local dofs the
 * owned rows of a partition read (ghost detection).  touches[r] = 1 when row r has a column >= first_col:
 * the interface rows that have to wait for the halo (the others are multiplied while it is in flight). */
int pgb_csr_mark_columns(int64_t n_out, const int32_t* rows, const void* indptr, int idx64, const int32_t* indices,
                         uint8_t* flags, void* stream);
int pgb_csr_rows_touching(int64_t n, const void* indptr, int idx64, const int32_t* indices, int32_t first_col,
                          uint8_t* touches, void* stream);
/* Block stacking in HBM: sps.block_array of CSR blocks (reference utils/bmat.py:9-99; the mixed-dimensional
 * operators of numerics/innerproducts.py:162-232 and numerics/differentials.py:144-192; the saddle stack
 * sps.block_array([[M, -B.T], [-B, None]]) of tutorials/fluid_flow/darcy.ipynb).  Per block row:
 *   lens = 0;  pgb_csr_block_lengths for every block  ->  pgb_scan_i64 (exclusive, out[n] = total) = indptr;
 *   cursor = copy of indptr;  pgb_csr_block_place for every block LEFT TO RIGHT (appends row r of the block,
 *   columns shifted by col_offset, values times `scale`, at cursor[r] and advances the cursor).
 * Scratch O(rows); the output rows are sorted when the blocks' rows are.  workspace of pgb_scan_i64:
 * pgb_system_workspace_bytes(0) bytes. */
int pgb_scan_i64(int64_t n, const int64_t* in, int64_t* out, int64_t* total_host, void* workspace,
                 int64_t workspace_bytes, void* stream);
int pgb_csr_block_lengths(int64_t n, const void* indptr, int idx64, int64_t* lens, void* stream);
int pgb_csr_block_place(int64_t n, const void* indptr, int idx64, const int32_t* indices, const double* data,
                        int32_t col_offset, double scale, int64_t* cursor, int32_t* out_indices, double* out_data,
                        void* stream);
/* data[(rows[i], rows[i])] += vals[i] for m DISTINCT rows (the mortar trace term S diag(1 / (|m| k_n)) S^T of
 * the face mass matrix, innerproducts.py:210-227, which is diagonal on a conforming interface).  flag_dev: one
 * int32 of device scratch.  Synchronises the stream; returns 1 when a listed row stores no diagonal entry. */
int pgb_csr_add_diagonal(int64_t m, const int32_t* rows, const double* vals, const void* indptr, int idx64,
                         const int32_t* indices, double* data, int32_t* flag_dev, void* stream);
/* CSR arrays of the transpose (scipy's A.T.tocsr(): B -> B^T of the saddle blocks): stable radix sort of the
 * column indices with the entry position as payload, so the entries of every output row are ordered by
 * input row - deterministic, no atomics.  nnz < 2^32 - 4096, dimensions < 2^31; out_indptr has n_cols + 1
 * entries of int32 (out_idx64 = 0) or int64.  Workspace: pgb_csr_transpose_workspace_bytes(nnz) (16 B / entry). */
int64_t pgb_csr_transpose_workspace_bytes(int64_t nnz);
int pgb_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const void* indptr, int idx64,
                      const int32_t* indices, const double* data, void* out_indptr, int out_idx64,
                      int32_t* out_indices, double* out_data, void* workspace, int64_t workspace_bytes,
                      void* stream);
/* diag[r] = A[r][r] (0 when the row has no diagonal entry). */
int pgb_csr_diagonal(int64_t n, const void* indptr, int idx64, const int32_t* indices, const double* data,
                     double* diag, void* stream);
/* Inverse diagonal of the SPD block preconditioner diag(diag(M), diag(B diag(M)^-1 B^T)) of the
 * symmetrised Darcy saddle matrix [[M, -B^T], [-B, 0]] (rows < num_faces: fluxes); diag_work:
 * num_faces doubles of scratch. */
int pgb_darcy_block_jacobi(int64_t n, int64_t num_faces, const void* indptr, int idx64, const int32_t* indices,
                           const double* data, double* diag_work, double* dinv, void* stream);
/* out[i] = x[idx[i]] (+ a * y[idx[i]] when y != NULL): the restriction R v of linear_system.py:72-77 and
 * the send-buffer packing of the halo exchange.  scatter_add: dst[idx[i]] += src[i] (distinct idx): R^T v. */
int pgb_gather(int64_t n, const int32_t* idx, const double* x, const double* y, double a, double* out,
               void* stream);
int pgb_scatter_add(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream);

/* ---- K3: SpMV and fused Krylov steps ----------------------------------------------
 * The reference has no iterative solver (numerics/linear_system.py:79-94 takes any
 * `solver(A, b) -> x`); these implement the recurrences of scipy.sparse.linalg.cg / minres
 * (scipy 1.18.1), which are the oracle.  indptr is int32 (idx64==0) or int64.
 */
/* threads_per_row: 2,4,8,16,32 or 0 = choose from nnz/row (costs one device->host read). */
int pgb_spmv(int64_t n_rows, const void* indptr, int idx64, const int32_t* indices,
             const double* data, const double* x, double* y, int threads_per_row, void* stream);

/* Whole CG solve on one GPU.  x (in: x0, out: solution), work: 3*n doubles.
 * dinv: optional Jacobi preconditioner (1/diag) or NULL.  resid_host[min(maxiter+1, PGB_RESID_CAP)]
 * (host, may be NULL) receives ||r_k||_2 for k = 0..iters (the first PGB_RESID_CAP of them: the
 * history is bounded whatever maxiter is).  Stops when ||r|| <= max(atol, rtol*||b||), checked
 * every iteration like scipy; b == 0 returns x = 0 at once (scipy: `return b, 0`).  Synchronises
 * the stream.  *iters_host = iterations done, *info_host = 0 converged / maxiter otherwise. */
#define PGB_RESID_CAP (1 << 20)
int pgb_cg(int64_t n, const void* indptr, int idx64, const int32_t* indices, const double* data,
           const double* b, double* x, const double* dinv, double rtol, double atol, int maxiter,
           double* work, int* iters_host, int* info_host, double* resid_host, void* stream);

/* Whole MINRES solve (Paige-Saunders as in scipy).  work: 6*n doubles.  dinv: optional positive
 * diagonal preconditioner (the `M` of scipy.sparse.linalg.minres as an inverse diagonal) or NULL. */
int pgb_minres(int64_t n, const void* indptr, int idx64, const int32_t* indices,
               const double* data, const double* b, double* x, const double* dinv, double shift,
               double rtol, int maxiter, double* work, int* iters_host, int* info_host,
               double* resid_host, void* stream);

/* ---- K4: row-distributed CG / MINRES over NVLink peer memory (one process per GPU, one node) --------
 * The reference has no distributed path; these are the multi-GPU form of pgb_cg /
 * pgb_minres for the cell-partitioned systems of SURVEY 8(e).  The halo exchange and the global sums are
 * done BY THE KRYLOV KERNELS through peer-mapped memory (P2P stores + sequence flags), see pgb_dist.cu.
 *
 * Communicator: every rank calls pgb_comm_create (allocates an arena of arena_bytes + a header of flags
 * with cudaMalloc and exports it as a cudaIpcMemHandle_t), the caller exchanges the handles
 * (pgb_comm_handle_bytes() bytes each, e.g. with torch.distributed.all_gather) and passes all of them, in
 * rank order, to pgb_comm_connect.  arena_bytes must be the same on every rank.
 *
 * System layout: the local vector is [n_owned owned | pad | n_ghost ghosts], the ghost block starting at
 * ghost_start (a multiple of 16 entries); columns of the CSR rows use that numbering.  Halo: for send
 * peer s, the owned entries send_idx[send_off[s] .. send_off[s+1]) go to entries
 * [send_dst[s], send_dst[s] + count) of the peer's local vector (= its ghost_start + the start of this
 * rank's block there).  recv_peer lists the ranks this rank receives from.  Rows [interior_lo,
 * interior_hi) have no ghost column: they are multiplied while the halo is in flight.
 * x0 = 0; x: n_owned entries out.  b == 0 returns x = 0, iterations 0 (scipy).  All ranks must call with
 * the same rtol / atol / maxiter; they stop at the same iteration (the scalars are bit-identical).
 * Returns 4 when a wait on a peer timed out. */
int pgb_comm_create(int rank, int world, int64_t arena_bytes, void** comm_out);
int pgb_comm_handle_bytes(void);
int pgb_comm_get_handle(void* comm, void* handle_out_host);
int pgb_comm_connect(void* comm, const void* all_handles_host);
void* pgb_comm_arena(void* comm);
int64_t pgb_comm_arena_bytes(void* comm);
int pgb_comm_barrier(void* comm, void* stream);
int pgb_comm_destroy(void* comm);
/* work: 2 * n_owned doubles.  The arena must hold 2 vectors of ghost_start + n_ghost doubles (the
 * search direction is double buffered), at byte offsets 0 and arena_bytes / 2. */
int pgb_dist_cg(void* comm, int64_t n_owned, int64_t ghost_start, int64_t n_ghost, const void* indptr, int idx64,
                const int32_t* indices, const double* data, const double* b, const double* dinv, double* x,
                double* work, int nsend, const int* send_peer_host, const int64_t* send_off_host,
                const int64_t* send_dst_host, const int32_t* send_idx, int nrecv, const int* recv_peer_host,
                int64_t interior_lo, int64_t interior_hi, int threads_per_row, double rtol, double atol,
                int maxiter, int* iters_host, int* info_host, double* resid_host, void* stream);
/* work: 5 * n_owned doubles; the arena holds the Lanczos vector (ghost_start + n_ghost doubles). */
int pgb_dist_minres(void* comm, int64_t n_owned, int64_t ghost_start, int64_t n_ghost, const void* indptr,
                    int idx64, const int32_t* indices, const double* data, const double* b, const double* dinv,
                    double* x, double* work, int nsend, const int* send_peer_host, const int64_t* send_off_host,
                    const int64_t* send_dst_host, const int32_t* send_idx, int nrecv, const int* recv_peer_host,
                    int64_t interior_lo, int64_t interior_hi, int threads_per_row, double shift, double rtol,
                    int maxiter, int* iters_host, int* info_host, double* resid_host, void* stream);

/* Building blocks used by the NCCL multi-GPU drivers (halo exchange / allreduce happen between
 * them on the host side through NCCL). */
/* CG step API: the fused kernels of pgb_cg one by one, so that a multi-GPU driver can interleave
 * the halo exchange (before step_spmv) and the all-reduce of the partial sums (step_reduce with
 * from_red = 0, all-reduce of scratch[4*PGB_NPART + i], step_reduce with from_red = 1).
 * mask bits: 1 = r.r, 2 = r.z, 4 = b.b, 8 = p.q.  scratch: pgb_cg_scratch_doubles(maxiter) doubles;
 * scratch[4*PGB_NPART + 8] holds the int `done` flag, the residual norms start at
 * scratch[4*PGB_NPART + 32]. */
int64_t pgb_cg_scratch_doubles(int maxiter);
int pgb_cg_step_init(int64_t n, const double* b, const double* q, const double* dinv, double* r,
                     double rtol, double atol, double* scratch, void* stream);
int pgb_cg_step_reduce(double* scratch, int mask, int from_red, void* stream);
int pgb_cg_step_p(int64_t n, int k, const double* r, const double* dinv, double* p, double* scratch,
                  void* stream);
int pgb_cg_step_spmv(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                     const double* data, const double* p, double* q, int threads_per_row,
                     double* scratch, void* stream);
int pgb_cg_step_xr(int64_t n, const double* p, const double* q, const double* dinv, double* x,
                   double* r, double* scratch, void* stream);

/* MINRES step API, same idea (kernels of pgb_minres one by one; the host rotates the r1/r2/y and
 * w buffers exactly as pgb_minres does).  mask bits of step_reduce: 1 = v.y, 2 = r2.y, 4 = x.x,
 * 8 = r1.y (beta1^2), 16 = b.b; the reduced values sit at scratch[5*PGB_NPART + i]; the int `done`
 * flag at scratch[5*PGB_NPART + 8]; (beta1^2, b.b) at scratch[5*PGB_NPART + 80]; the residual
 * estimates (phibar) of iterations 1.. at scratch[5*PGB_NPART + 82 + itn]. */
int64_t pgb_mr_scratch_doubles(int maxiter);
int pgb_mr_step_reduce(double* scratch, int mask, int from_red, void* stream);
int pgb_mr_step_init(int64_t n, const double* b, const double* q, double* r1, double* scratch,
                     void* stream);
int pgb_mr_step_start(int64_t n, const double* r1, double* v, double* w, double* w2, double rtol,
                      double shift, int maxiter, double* scratch, void* stream);
int pgb_mr_step_spmv(int64_t n, const void* indptr, int idx64, const int32_t* indices,
                     const double* data, const double* v, double* y, const double* r1, int itn,
                     int threads_per_row, double* scratch, void* stream);
int pgb_mr_step_c(int64_t n, double* y, const double* r2, double* scratch, void* stream);
int pgb_mr_step_scalars(int itn, int test_only, double* scratch, void* stream);
int pgb_mr_step_d(int64_t n, double* v, double* w_old, const double* w_new, const double* r2,
                  double* x, double* scratch, void* stream);
int pgb_mr_state(const double* scratch, int* itn_host, int* istop_host, void* stream);

int pgb_dot_partials(int64_t n, const double* x, const double* y, double* partials /*[PGB_NPART]*/,
                     void* stream);
int pgb_reduce_partials(const double* partials, int count, double* out, void* stream);
int pgb_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream);
#define PGB_NPART 1184

#ifdef __cplusplus
}
#endif
#endif /* PGB_H */
