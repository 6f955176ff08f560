/*
 * femx.h — C ABI of the B200-native finite-element assembly engine.
 *
 * This is the drop-in boundary for the ONE hot path of yuemingl/cuda-fem:
 *   mesh coordinates + connectivity  →  NVRTC-compiled symbolic element
 *   integrand  →  assembled global sparse matrix (COO triplets or CSR).
 *
 * The reference has no FFI: every program is a main().  Each entry point below
 * therefore cites the stretch of reference host/device code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C, no torch / C++ types; every pointer named d_* is a DEVICE pointer
 *     owned by the caller; h_* is a host pointer.
 *   - every call returns a femx_status (0 = ok); it never exits, aborts or
 *     throws across the ABI (the reference's NVRTC_SAFE_CALL/CUDA_SAFE_CALL
 *     macros call exit(1): fea_symbolic_nvrtc_sparse.cpp:16-35).
 *   - calls are stream-ordered on the cudaStream_t passed as `void* stream`
 *     (NULL = legacy default stream, what the reference uses:
 *     fea_symbolic_nvrtc_sparse.cpp:614).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     returns FEMX_ERR_CUDA.
 */
#ifndef FEMX_H
#define FEMX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEMX_VERSION_MAJOR 0
#define FEMX_VERSION_MINOR 1

typedef enum femx_status {
  FEMX_OK = 0,
  FEMX_ERR_INVALID = 1,     /* bad argument                                  */
  FEMX_ERR_CUDA = 2,        /* CUDA runtime/driver error, or no device       */
  FEMX_ERR_NVRTC = 3,       /* integrand failed to compile; see femx_last_error */
  FEMX_ERR_UNSUPPORTED = 4, /* size/shape outside what this build handles    */
  FEMX_ERR_NOMEM = 5
} femx_status;

typedef enum femx_dtype { FEMX_F64 = 0, FEMX_F32 = 1 } femx_dtype;

/* Built-in weak forms (the emitter that stands in for GiNaC's
 * WeakForm::build, fea_symbolic_nvrtc_sparse.cpp:307-356). FEMX_FORM_CUSTOM
 * takes the (nn*nd)^2 C-expression strings directly, i.e. the reference's
 * $integrand{li}{lj}$ placeholders (fea_symbolic_nvrtc_sparse.cpp:384-414). */
typedef enum femx_builtin {
  FEMX_FORM_CUSTOM = 0,
  FEMX_FORM_POISSON = 1,      /* grad u . grad v                  (nd = 1) */
  FEMX_FORM_POISSON_MASS = 2, /* grad u . grad v + c * u v        (nd = 1), c = params[0] (0 → 1) */
  FEMX_FORM_MASS = 3,         /* u v                               (nd = 1) */
  FEMX_FORM_ELASTICITY = 4    /* lambda div u div v + 2 mu eps(u):eps(v), nd = dim;
                                 params[0] = lambda, params[1] = mu          */
} femx_builtin;

typedef struct femx_ctx femx_ctx;         /* one per device                     */
typedef struct femx_form femx_form;       /* a JIT-compiled element integrand   */
typedef struct femx_pattern femx_pattern; /* CSR pattern + scatter map          */

/* ------------------------------------------------------------------ context */

/* Replaces cuInit/cuDeviceGet/cuCtxCreate (fea_symbolic_nvrtc_sparse.cpp:557-559);
 * binds to the device's primary context instead of creating a new one. */
int femx_ctx_create(int device, femx_ctx** out);
void femx_ctx_destroy(femx_ctx* ctx);
/* Message of the last failing call on this ctx (includes the NVRTC log for
 * FEMX_ERR_NVRTC, which the reference prints to stdout before exit(1):
 * fea_symbolic_nvrtc_sparse.cpp:534-542).  ctx == NULL → last global error. */
const char* femx_last_error(const femx_ctx* ctx);
const char* femx_version(void);
/* Tuning options of a context (experiments; the defaults are the measured configuration).  They are read
 * from the environment (FEMX_<NAME>) once, when the context is created; this call changes one afterwards.
 * No entry point reads the environment on a launch path.  name: "spec" (stencil-class pass on/off),
 * "lattice" (element-once lattice pass on/off), "tile", "carveout", "lt_tx", "lt_ty", "lt_kc", ...
 * (csrc/femx_core.cpp: kKnobs).  Options that shape generated code apply to forms compiled afterwards;
 * spec / lattice / carveout / lt_* are looked at by every femx_assemble_csr call. */
int femx_ctx_set_option(femx_ctx* ctx, const char* name, int value);

/* -------------------------------------------------------------------- forms */

typedef struct femx_form_desc {
  int dim;     /* 2 (P1 triangle, nn = 3) or 3 (P1 tetrahedron, nn = 4)          */
  int nn;      /* nodes per element                                               */
  int nd;      /* dofs per node (1 scalar, dim for elasticity)                    */
  int dtype;   /* femx_dtype of coordinates and matrix values                     */
  int builtin; /* femx_builtin                                                    */
  double params[4];
  /* FEMX_FORM_CUSTOM: n*n strings, n = nn*nd, entries[li*n + lj] is the
   * integrand of matrix entry (row = dof li, col = dof lj) = a(u=phi_lj, v=phi_li) * jac
   * as a C expression over x1..x{nn}, y1.., z1.., and the quadrature point
   * r, s, t (and u in 3-D) — the names the reference's integrand() unpacks from
   * params[] (fea_symbolic_nvrtc_sparse.cpp:386-394).  Use the type `real`
   * for casts; pow(x,2.0) is accepted.  The strings (and the prologue) are pasted into kernel scopes: besides
   * the names above, identifiers starting with femx_ / FEMX_ / lt_ / wf_ and the kernel locals e, e0, nodes, cx, cy,
   * cz, out, code, sc, np, it, ps, po, more, srow, dacc, racc are reserved.  femx_form_compile compiles both the COO and
   * the numeric-pass kernel for custom strings, so a clash fails there (FEMX_ERR_NVRTC + log), not at first use. */
  const char* const* entries;
  /* optional C statements evaluated once per element before the quadrature
   * loop (common sub-expressions such as the Jacobian); may be NULL. */
  const char* prologue;
  /* quadrature rule; nq == 0 → default: the reference's 7-point triangle
   * literals (fea_symbolic_nvrtc_sparse.cpp:380-383) in 2-D, the 4-point
   * degree-2 rule in 3-D.  qt (2-D: third barycentric; 3-D: third reference
   * coordinate) and qu (3-D: fourth barycentric) may be NULL → 1 - sum. */
  int nq;
  const double* qw;
  const double* qr;
  const double* qs;
  const double* qt;
  const double* qu;
  int fmad; /* 1 (default build) allows FMA contraction; 0 = the reference's
               --fmad=false (fea_symbolic_nvrtc_sparse.cpp:529)              */
  /* FEMX_FORM_CUSTOM only: 0 = entries are INTEGRANDS (the reference's
   * semantics: weighted by w_q and summed over the rule); 1 = entries are the
   * final element-matrix expressions, no quadrature applied. */
  int integrated;
  /* Load vector (optional).  n strings rhs_entries[li] = integrand of b[dof li] = f * phi_li * jac
   * over the same names — what WeakForm::build generates as rhs[j] and then discards
   * (fea_symbolic_nvrtc_sparse.cpp:346-351; recorded output fea_symbolic.cu:335,339,343).
   * Built-in forms without strings use the constant source rhs_vec (scalar forms: rhs_vec[0];
   * elasticity: body force per component). */
  const char* const* rhs_entries;
  double rhs_vec[3];
} femx_form_desc;

/* Replaces WeakForm::build + nvrtcCreateProgram … cuModuleGetFunction
 * (fea_symbolic_nvrtc_sparse.cpp:307-356, 506-561).  Compiles with NVRTC
 * straight to an sm_100a cubin.  The COO kernel is compiled eagerly (so a bad
 * expression fails here); other kernel variants on first use. */
int femx_form_compile(femx_ctx* ctx, const femx_form_desc* desc, femx_form** out);
void femx_form_destroy(femx_form* form);
/* Generated CUDA source of the last compiled kernel variant (what the
 * reference prints at fea_symbolic_nvrtc_sparse.cpp:354) and its NVRTC log. */
const char* femx_form_source(const femx_form* form);
const char* femx_form_log(const femx_form* form);
/* The n*n entry strings actually used (for built-ins: what the emitter
 * produced; the analogue of the csrc_float output of WeakForm::build).
 * Returns a pointer valid for the life of the form. */
const char* femx_form_entry(const femx_form* form, int li, int lj);
const char* femx_form_prologue(const femx_form* form);
/* Compile only (no device needed): emits source + cubin for inspection with
 * cuobjdump.  kernel: "coo", "csr" (node coordinates), "csr_x" (element-expanded
 * coordinates), "csr_s" (strided node coordinates).  cubin buffer is owned by the
 * form. */
int femx_form_cubin(femx_form* form, const char* kernel, const void** cubin, size_t* size);
/* As femx_form_compile but never touches the device (ctx may be NULL):
 * used by the CPU-only test tier and for offline SASS inspection. */
int femx_form_compile_offline(const femx_form_desc* desc, femx_form** out);

/* --------------------------------------------------------------------- mesh */

typedef struct femx_mesh_view {
  int dim;
  int nn;
  int64_t n_nodes;
  int64_t n_elems;
  /* connectivity, reference layout gIdx[nn*e + k]
   * (fea_symbolic_nvrtc_sparse.cpp:580-582) */
  const int32_t* d_conn;
  /* node-indexed coordinates (SoA, dtype of the form), d_node_xyz[c][node*node_stride];
   * NULL when element-expanded coordinates are given instead */
  const void* d_node_xyz[3];
  int64_t node_stride; /* in elements; 0 → 1 */
  /* element-expanded coordinates, reference layout X[nn*e + k], Y[nn*e + k]
   * (fea_symbolic_nvrtc_sparse.cpp:574-579); NULL when node-indexed */
  const void* d_elem_xyz[3];
} femx_mesh_view;

/* Device-side structured generators (replace RectangleMesh::generate,
 * fea_symbolic_nvrtc_sparse.cpp:170-216, whose per-node `new` cannot reach
 * 10^8 elements).  Node index i*(nCol+1)+j, two CCW triangles per cell
 * (n, n+1, n+nCol+1) and (n+1, n+nCol+2, n+nCol+1) — exactly the reference.
 * Only node rows [row_lo, row_hi] and cell rows [row_lo, row_hi) are produced
 * (a slab for multi-GPU); local node index = global - row_lo*(nCol+1).
 * Any output pointer may be NULL. d_flag gets the boundary flag
 * (fea_test.cu:100-103). */
int femx_mesh_rectangle(femx_ctx* ctx, double x0, double x1, double y0, double y1,
                        int64_t nRow, int64_t nCol, int64_t row_lo, int64_t row_hi,
                        int dtype, void* d_x, void* d_y, int32_t* d_flag,
                        int32_t* d_conn, void* stream);
/* Element-expanded coordinates from node coordinates (replaces the host
 * flattening loop fea_symbolic_nvrtc_sparse.cpp:571-583). */
int femx_mesh_expand(femx_ctx* ctx, int dtype, int nn, int64_t n_elems,
                     const int32_t* d_conn, const void* d_node, void* d_elem,
                     void* stream);
/* Unit-box Kuhn mesh: (nx,ny,nz) cells, 6 positively oriented tets per cell
 * around the (0,0,0)-(1,1,1) diagonal, node index (k*(ny+1)+j)*(nx+1)+i.
 * No reference counterpart (the reference is 2-D only). Slab in k:
 * node planes [k_lo, k_hi], cell layers [k_lo, k_hi). */
int femx_mesh_box(femx_ctx* ctx, double x0, double x1, double y0, double y1,
                  double z0, double z1, int64_t nx, int64_t ny, int64_t nz,
                  int64_t k_lo, int64_t k_hi, int dtype, void* d_x, void* d_y,
                  void* d_z, int32_t* d_conn, void* stream);

/* ---------------------------------------------------------- kernel ABI #1: COO */

/* Replaces fea_kernel (COO variant) + its launch
 * (fea_symbolic_nvrtc_sparse.cpp:415-480, 604-615): slot = e*n*n + li*n + lj,
 * rowA = dof of local row li, colA = dof of local col lj, A = sum_q w_q *
 * integrand(li,lj).  Duplicates are NOT merged (as in the reference). */
int femx_assemble_coo(femx_form* form, const femx_mesh_view* mesh, void* d_A,
                      int32_t* d_rowA, int32_t* d_colA, void* stream);

/* ------------------------------------- kernel ABI #2: pattern + numeric pass */

/* The symbolic pass.  Replaces Mesh::getNeighborNodesList
 * (fea_symbolic_nvrtc_sparse2.cpp:181-210): per row the ascending, duplicate-free
 * list of all nodes sharing an element with it, diagonal included.  Runs on the
 * device (histogram + scan + per-row sort/unique) and also builds the
 * element-slot → CSR-offset scatter map used by femx_assemble_csr.
 *
 * Rows are built for local nodes [row_begin, row_end) only (0, n_nodes for the
 * whole mesh); column indices are local node ids + col_base, so a slab of a
 * partitioned mesh yields rows of the GLOBAL matrix.  Every element touching
 * an owned row must be present in d_conn (ghost elements). */
int femx_pattern_build(femx_ctx* ctx, int nn, int nd, int64_t n_nodes,
                       int64_t n_elems, const int32_t* d_conn, int64_t row_begin,
                       int64_t row_end, int64_t col_base, void* stream,
                       femx_pattern** out);
void femx_pattern_destroy(femx_pattern* pat);
/* n_rows = nd*(row_end-row_begin) dof rows; nnz of the dof-level CSR;
 * max_row = longest dof row.  Any out pointer may be NULL. */
int femx_pattern_info(const femx_pattern* pat, int64_t* n_rows, int64_t* nnz,
                      int64_t* max_row);
/* Bytes of device memory held by the pattern object (scatter map included). */
int64_t femx_pattern_bytes(const femx_pattern* pat);
/* CSR of the dof-level matrix into caller buffers: row_ptr has n_rows+1
 * entries (give either the 64- or the 32-bit buffer, or both), col_idx nnz. */
int femx_pattern_export_csr(const femx_pattern* pat, int64_t* d_row_ptr64,
                            int32_t* d_row_ptr32, int32_t* d_col_idx, void* stream);
/* The reference's padded layout (fea_symbolic_nvrtc_sparse2.cpp:181-210, 437):
 * len[i] = row length, idx[i*width + j] = j-th column; entries j >= len[i]
 * are set to 0 as the reference's host zero-fill does (:644).  nd must be 1.
 * Fails with FEMX_ERR_INVALID if width < max_row. */
int femx_pattern_export_ell(const femx_pattern* pat, int width, int32_t* d_len,
                            int32_t* d_idx, void* stream);

/* Dominant stencil class of a scalar (nd = 1) pattern.  Rows whose incidence count, row length,
 * own position, scatter-code sequence and column offsets (column - own node) coincide do the same
 * arithmetic on nodes at the same relative positions — every interior row of a structured mesh.  The symbolic pass finds the most frequent class among
 * evenly spaced sample rows and flags its rows; femx_assemble_csr then runs an NVRTC-generated
 * straight-line body for them (each neighbour's coordinates read once — at own node + offset, issued
 * before any metadata arrives — the row's values accumulated in registers, no scatter codes or column
 * lists read) and the generic incidence loop for the others.  Both paths
 * perform the same floating-point operations in the same order, so the result does not depend on
 * which one a row takes.  In the reference the mesh size is baked into the kernel's -D macros
 * (fea_symbolic_nvrtc_sparse.cpp:506-530); here it is the mesh's stencil.
 * n_incid/row_len/self_pos describe the class, rows = how many rows belong to it (0 = none found),
 * h_codes / h_offsets (HOST buffers, cap entries each; may be NULL) receive the scatter codes and the
 * column offsets.  FEMX_SPEC=0 in the environment switches detection and use off. */
int femx_pattern_stencil(const femx_pattern* pat, int* n_incid, int* row_len, int* self_pos,
                         int64_t* rows, uint32_t* h_codes, int32_t* h_offsets, int cap);
/* Diagnostic: NVRTC-compiles the numeric-pass kernel specialised for an explicitly given stencil
 * class (no device needed with a form from femx_form_compile_offline) and returns the cubin. */
int femx_form_cubin_stencil(femx_form* form, int n_incid, int row_len, int self_pos,
                            const uint32_t* h_codes, const void** cubin, size_t* size);

/* Lattice structure of the mesh the pattern was built from (found and verified element by element by the
 * symbolic pass): elements come in cells of *n_per_cell consecutive elements, cell c = ci + cells[0]*(cj +
 * cells[1]*ck), each cell a translate of cell 0; vertex a of element c*P+t is lattice node (ci,cj,ck) +
 * corner (bit 0 = dx, bit 1 = dy, bit 2 = dz) h_corners[t*nn + a]; node id = *node0 + i*strides[0] + j*strides[1] +
 * k*strides[2].  This is what RectangleMesh::generate produces (fea_symbolic_nvrtc_sparse.cpp:170-216) and
 * what femx_mesh_box produces in 3-D.  *n_per_cell = 0: no lattice (unstructured mesh, or no stencil class).
 * On 3-D lattices femx_assemble_csr runs the element-once pass for symmetric built-in forms.
 * h_cells / h_strides: 3 entries each; h_corners: 8*nn entries; any pointer may be NULL. */
int femx_pattern_lattice(const femx_pattern* pat, int* n_per_cell, int64_t* h_cells, int64_t* h_strides,
                         int64_t* node0, int32_t* h_corners);
/* Host utility (no device): the closed form the lattice symbolic pass evaluates per row.  Lattice node (i, j, k),
 * 0 <= i <= h_cells[0] ..., is node node0 + i + j h_strides[1] + k h_strides[2] (strides may be wider than the lattice:
 * the nodes in between carry nothing); its class is ci + 3 cj + 9 ck with c = 0 / 1 / 2 for the low face / interior / high
 * face along an axis.  h_out[q] = sum of h_weights[class] over the lattice nodes whose id is below h_nodes[q] — with
 * weights = row length per class this is the CSR row pointer of that node, with weights = 1 for one class the number of
 * rows of that class in front of it.  (femx_pattern.cu: lat_prefix; dim = 2: h_cells / h_strides have 2 entries.) */
int femx_lattice_prefix(int dim, const int32_t* h_cells, const int64_t* h_strides, int64_t node0, const int32_t* h_weights,
                        int64_t n, const int64_t* h_nodes, int64_t* h_out);
/* Diagnostic: NVRTC-compiles the element-once lattice pass for an explicitly given lattice cell (n_per_cell
 * elements, h_corners as above), node strides and stencil class (row_len sorted column offsets h_offsets, own
 * position self_pos); no device needed with a form from femx_form_compile_offline.  h_info (6 ints, may be NULL)
 * receives tile threads x, y, threads per CTA, shared-memory field slots, shared-memory bytes, min blocks. */
int femx_form_cubin_lattice(femx_form* form, int n_per_cell, const int32_t* h_corners, int64_t stride_y,
                            int64_t stride_z, int row_len, int self_pos, const int32_t* h_offsets,
                            const void** cubin, size_t* size, int* h_info);

/* The numeric pass.  Replaces fea_kernel (ELL + global atomicAdd variant,
 * fea_symbolic_nvrtc_sparse2.cpp:475-547): d_values[k] is the sum of all element
 * contributions to CSR slot k, accumulated in ascending element order by the
 * one thread that owns the row — no atomics, bitwise reproducible.
 * d_values has nnz entries of the form's dtype and is fully overwritten; it must be 16-byte aligned
 * (the tiles leave shared memory through cp.async.bulk stores), as must d_A/d_rowA/d_colA of
 * femx_assemble_coo — FEMX_ERR_INVALID otherwise.
 * Limits: rows of more than 127 node columns and more than 2^31 - 1 incidences / node-level nonzeros per
 * device are rejected by femx_pattern_build (FEMX_ERR_UNSUPPORTED: shard the mesh); P1 simplices only. */
int femx_assemble_csr(femx_form* form, const femx_pattern* pat,
                      const femx_mesh_view* mesh, void* d_values, void* stream);
/* Load vector b[dof] = sum over incident elements of the integrated rhs entry, accumulated in
 * ascending element order by the thread owning the row (deterministic, no atomics).  d_rhs has
 * n_rows entries of the form's dtype.  No reference kernel exists (the reference drops its RHS
 * strings); the author's intended host loop is the commented block fea_kernal.cu:193-214. */
int femx_assemble_rhs(femx_form* form, const femx_pattern* pat, const femx_mesh_view* mesh,
                      void* d_rhs, void* stream);
/* CSR values → the reference's ELL value layout A[i*width + j] (zero padded). */
int femx_csr_to_ell(const femx_pattern* pat, int dtype, int width,
                    const void* d_values, void* d_ell, void* stream);

/* Dirichlet conditions by symmetric elimination (the reference sets the node
 * boundary `flag`, fea_test.cu:100-103, and never uses it).  d_flag[dof] != 0 marks a constrained
 * dof, d_g[dof] its value; both are indexed by LOCAL dof (nd * local node + comp) and must cover
 * every node of the slab (ghosts included).  For each owned row j:
 *   constrained   : row := e_j,  rhs[j] := g_j
 *   unconstrained : rhs[j] -= sum_{i constrained} A_ji g_i (ascending column order), A_ji := 0.
 * d_rhs may be NULL (matrix only). */
int femx_apply_dirichlet(const femx_pattern* pat, int dtype, const int32_t* d_flag, const void* d_g,
                         void* d_values, void* d_rhs, void* stream);

/* ------------------------------------------------- validation: SpMV and CG */

/* y = A x for the rows of this pattern; x is indexed by (column - x_base). */
int femx_spmv(const femx_pattern* pat, int dtype, const void* d_values,
              const void* d_x, int64_t x_base, void* d_y, void* stream);
/* Fused vector kernels used by the CG driver (all length n, fp64 partial sums
 * reduced in a fixed order → deterministic):
 *   dot2:  out[0] = <a,b>, out[1] = <c,d>                       */
int femx_dot2(femx_ctx* ctx, int dtype, int64_t n, const void* d_a, const void* d_b,
              const void* d_c, const void* d_d, double* d_out, void* stream);
/*   axpy:  y += alpha[0]/alpha[1] * sign * x   (alpha read on device, no host sync) */
int femx_axpy_ratio(femx_ctx* ctx, int dtype, int64_t n, const double* d_num,
                    const double* d_den, double sign, const void* d_x, void* d_y,
                    void* stream);
/*   xpby:  p = r + (num/den) p */
int femx_xpby_ratio(femx_ctx* ctx, int dtype, int64_t n, const double* d_num,
                    const double* d_den, const void* d_r, void* d_p, void* stream);

/* ------------------------------------------------------------- multi-GPU layer */

/* One process per GPU.  Assembly shards with NO communication: rank p owns node planes [r0, r1) of the
 * structured mesh (femx_dist_slab), builds the slab [lo, hi] with one ghost layer per side
 * (femx_mesh_box / femx_mesh_rectangle with k_lo/k_hi), and femx_pattern_build(row_begin, row_end,
 * col_base) + femx_assemble_csr produce its rows of the GLOBAL matrix (bit-identical to the single-GPU
 * rows).  NCCL (resolved at run time from the libnccl.so.2 already in the process, else the system's)
 * is used only by the validation solver below: grouped ncclSend/ncclRecv for the halo of the SpMV operand
 * and ONE fused 2-double ncclAllReduce per CG iteration.  No reference counterpart (the reference runs one
 * rank: job.pbs:4,24). */
typedef struct femx_dist femx_dist;        /* communicator + streams of one rank          */
typedef struct femx_dist_op femx_dist_op;  /* this rank's rows of the operator + halo plan */
#define FEMX_DIST_ID_BYTES 128
/* Rank 0 creates the NCCL unique id (h_id: FEMX_DIST_ID_BYTES host bytes) and hands it to the other ranks by
 * whatever out-of-band channel the launcher has (bench.py: a torch.distributed broadcast; C++ clients: MPI / a file). */
int femx_dist_unique_id(void* h_id);
int femx_dist_create(femx_ctx* ctx, int rank, int world, const void* h_id, femx_dist** out);  /* world == 1: h_id may be NULL */
void femx_dist_destroy(femx_dist* d);
/* *p2p_reduction = 1 when the CG's reduction runs over NVLink peer memory (CUDA IPC buffers of all ranks mapped into every
 * rank): one kernel finishes the two dot products, exchanges the partial sums with every peer and advances the CG scalars;
 * 0 = ncclAllReduce (peer mapping unavailable, or option dist_p2p = 0). */
int femx_dist_info(const femx_dist* d, int* rank, int* world, int* p2p_reduction);
/* Even split of n_planes node planes over `world` ranks: owned planes [r0, r1), slab planes [lo, hi]. */
int femx_dist_slab(int64_t n_planes, int world, int rank, int64_t* r0, int64_t* r1, int64_t* lo, int64_t* hi);
/* In-place sum (op_max = 0) or max (1) of n doubles across ranks (timing / checksums of the harness). */
int femx_dist_allreduce(femx_dist* d, double* d_buf, int n, int op_max, void* stream);
/* The rank's assembled rows as an operator.  The pattern's local node layout is [ghost | owned rows | ghost]
 * (row_begin / row_end of femx_pattern_build); the ghost zones of the operand come from ranks rank-1 / rank+1. */
int femx_dist_op_create(femx_dist* d, const femx_pattern* pat, int dtype, const void* d_values, femx_dist_op** out);
void femx_dist_op_destroy(femx_dist_op* op);
/* n_owned dof rows; operand entries below / above the owned range; node rows [interior_lo, interior_hi) read no
 * ghost column (their SpMV overlaps the halo exchange). */
int femx_dist_op_info(const femx_dist_op* op, int64_t* n_owned, int64_t* ghost_lo, int64_t* ghost_hi, int64_t* interior_lo,
                      int64_t* interior_hi);
/* *peer_halo = 1 when femx_dist_cg moves its halo through NVLink peer memory: the update kernel stores the first / last owned
 * entries of the new residual straight into the two neighbours' ghost zones (their buffers mapped through CUDA IPC) and raises
 * a flag the boundary rows' SpMV waits for — no ncclSend/ncclRecv and no second stream inside the iteration.  0 = NCCL halo
 * (peer mapping unavailable, option dist_push = 0, or world == 1).  The option must be the same on every rank. */
int femx_dist_op_peer_halo(const femx_dist_op* op, int* peer_halo);
/* y_owned = A[owned rows] x, x given by its owned part on every rank (device pointers, n_owned entries each). */
int femx_dist_spmv(femx_dist_op* op, const void* d_x_owned, void* d_y_owned, void* stream);
/* `iters` steps of unpreconditioned CG from x0 = 0 (Chronopoulos-Gear form: one SpMV, one fused update kernel and
 * ONE reduction of two doubles per iteration — over NVLink peer memory, fused into the kernel that finishes the dot products
 * and advances alpha / beta (femx_dist_info), else ncclAllReduce; the iteration is captured in a CUDA graph and replayed).
 * h_residuals (host, iters+1 entries, may be NULL) receives ||r_k||_2; *h_ms (may be NULL) the device time of the solve.
 * Synchronises the stream before returning. */
int femx_dist_cg(femx_dist_op* op, const void* d_b_owned, void* d_x_owned, int iters, double* h_residuals, float* h_ms,
                 void* stream);
/* y = A x for node rows [row_lo, row_hi) of the pattern only (femx_spmv: all rows). */
int femx_spmv_rows(const femx_pattern* pat, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                   int64_t row_lo, int64_t row_hi, void* stream);

/* Unstructured meshes: partition by owned CSR rows with duplicated ghost elements.  The rank owns the contiguous node range
 * [node_lo, node_hi) of the caller's numbering (e.g. equal ranges after an RCM / space-filling-curve ordering).  The extraction
 * keeps every element that touches an owned node (ascending element id), renumbers the touched nodes by their rank among the
 * sorted global ids (owned nodes stay contiguous: local rows [row_begin, row_end)) and keeps the local -> global map.
 * femx_pattern_build(nn, nd, n_local_nodes, n_local_elems, d_conn_local, row_begin, row_end, 0, ...) + femx_assemble_csr on
 * the gathered coordinates then give the rank's rows of the GLOBAL matrix, bit-identical to the single-GPU rows and with no
 * communication; femx_pattern_export_csr_mapped writes global column ids.  (The SpMV / CG validator handles slab partitions only.) */
typedef struct femx_part femx_part;
int femx_partition_extract(femx_ctx* ctx, int nn, int64_t n_nodes, int64_t n_elems, const int32_t* d_conn, int64_t node_lo,
                           int64_t node_hi, void* stream, femx_part** out);
void femx_part_destroy(femx_part* part);
int femx_part_info(const femx_part* part, int64_t* n_local_nodes, int64_t* n_local_elems, int64_t* row_begin, int64_t* row_end);
/* Device arrays owned by the object: local connectivity [n_local_elems * nn], local -> global node ids [n_local_nodes]
 * (ascending), global ids of the kept elements [n_local_elems] (ascending). */
int femx_part_arrays(const femx_part* part, const int32_t** d_conn_local, const int32_t** d_local_to_global,
                     const int32_t** d_elem_ids);
/* The same arrays copied into caller-owned device buffers (any pointer may be NULL). */
int femx_part_copy(const femx_part* part, int32_t* d_conn_local, int32_t* d_local_to_global, int32_t* d_elem_ids, void* stream);
/* d_local[i] = d_global[local_to_global[i]] (node coordinates of the sub-mesh). */
int femx_part_gather(const femx_part* part, int dtype, const void* d_global, void* d_local, void* stream);
/* femx_pattern_export_csr with the columns mapped through d_local_to_global instead of shifted by col_base. */
int femx_pattern_export_csr_mapped(const femx_pattern* pat, const int32_t* d_local_to_global, int64_t* d_row_ptr64,
                                   int32_t* d_row_ptr32, int32_t* d_col_idx, void* stream);

/* ------------------------------------------------ host-side I/O (no device needed) */

/* Gmsh MSH 2.x ASCII: 3-node triangles, or 4-node tetrahedra when present (the boundary triangles of
 * a volume mesh are then ignored).  Node tags are compacted to 0..n-1 in file order.  The four
 * buffers are malloc'ed by the library; release them with femx_io_free.  (The reference can only
 * print a mesh: printMesh(), fea_test.cu:53-67.) */
int femx_io_read_gmsh(const char* path, int* dim, int64_t* n_nodes, int64_t* n_elems, double** h_x,
                      double** h_y, double** h_z, int32_t** h_conn);
void femx_io_free(void* p);
/* Matrix Market "coordinate real general" (1-based) from a host CSR, %.17g values. */
int femx_io_write_matrix_market(const char* path, int64_t n_rows, int64_t n_cols,
                                const int64_t* h_row_ptr, const int32_t* h_col_idx,
                                const double* h_values);

#ifdef __cplusplus
}
#endif
#endif /* FEMX_H */
