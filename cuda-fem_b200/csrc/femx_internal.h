// Internal declarations shared by the translation units of libfemx.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "femx.h"

// Tuning knobs (experiments; the defaults are what DESIGN.md measures).  Read from the environment
// ONCE — at femx_ctx_create, or at femx_form_compile_offline for forms without a context — and
// changed at run time only through femx_ctx_set_option: no getenv in any launch path.
struct femx_knobs {
  int tile = 0;         // FEMX_TILE      node rows per CTA of the generic pass (0 = default)
  int carveout = -1;    // FEMX_CARVEOUT  shared-memory carve-out % (-1 = derived from occupancy)
  int minblocks = -1;   // FEMX_MINBLOCKS __launch_bounds__ min blocks (-1 = default)
  int midgather = 1, unroll = 1, rotinv = 1;
  int spec = 1;         // FEMX_SPEC      stencil-class detection and use
  int spec_ahead = -1, sharedfaces = 1, accf = 1, rcp3 = 0, spec_prefetch = 0, spec_pin = 1, listlast = 0;
  int rowsum = 0, chainorder = 0;
  int lattice = 1;      // FEMX_LATTICE   element-once lattice pass on structured 3-D meshes
  int lattice_pattern = 1;  // FEMX_LATTICE_PATTERN  lattice meshes: the symbolic pass writes rows from templates
  int lt_tx = 0, lt_ty = 0, lt_kc = 0, lt_minb = 0, lt_regs = 0, lt_pf = 0, lt_unroll = 2, lt_side = 1;  // FEMX_LT_*: tile shape / k-chunk / occupancy of that pass
  int dist_p2p = 1;     // FEMX_DIST_P2P    CG reduction over NVLink peer memory, fused with the dot products and the CG scalars (0: ncclAllReduce)
  int dist_graph = 1;   // FEMX_DIST_GRAPH  the CG iteration of femx_dist_cg is replayed from a CUDA graph
  int dist_push = 1;    // FEMX_DIST_PUSH   CG halo: the update kernel stores the boundary entries of r straight into the neighbours' ghost zones (NVLink peer memory) instead of ncclSend/Recv
  std::string jit_dump; // FEMX_JIT_DUMP  directory that receives the generated sources
  std::string key() const;
};
femx_knobs femx_knobs_from_env();
bool femx_knobs_set(femx_knobs* k, const char* name, int value);

struct femx_ctx {
  femx_knobs knobs;
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  double* d_scratch = nullptr;  // partial sums of femx_dot2 (2 * FEMX_DOT_BLOCKS doubles)
  cudaStream_t s_side = nullptr;   // side stream: the boundary-row kernel runs beside the lattice pass (femx_assemble_csr)
  cudaEvent_t e_fork = nullptr, e_join = nullptr;
  cudaMemPool_t pool = nullptr; // stream-ordered pool for the symbolic pass's temporaries (retains memory)
  mutable std::string err;
};

// Driver entry points resolved through cudaGetDriverEntryPoint so that the
// library has no link-time dependency on libcuda.so (it must load, and its
// symbols be listable, on a box without a driver).
struct femx_driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};
const femx_driver* femx_get_driver(std::string* why);

int femx_fail(const femx_ctx* ctx, int status, const char* fmt, ...);
void femx_set_global_error(const std::string& s);

#define FEMX_CUDA_OK(ctx, call)                                                    \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return femx_fail((ctx), FEMX_ERR_CUDA, "%s failed: %s (%s:%d)", #call,      \
                       cudaGetErrorString(e__), __FILE__, __LINE__);               \
  } while (0)

// ---- lattice structure of a mesh (femx_pattern.cu: detect_lattice) ---------------------------
// The elements come in cells of P consecutive elements, cell c = ci + cn[0]*(cj + cn[1]*ck), and every
// cell is a translate of cell 0: vertex a of element c*P + t is lattice node (ci,cj,ck) + corner[t][a],
// node id = node0 + i + j*s[1] + k*s[2].  What RectangleMesh::generate / femx_mesh_box (and any
// structured generator that numbers cell-major) produce; verified element by element on the device.
struct femx_lattice {
  bool ok = false;
  int dim = 0, P = 0;
  int cn[3] = {1, 1, 1};             // cells per axis
  long long s[3] = {1, 0, 0};        // node strides (local node ids)
  long long node0 = 0;               // local id of lattice node (0,0,0)
  unsigned char corner[8][4] = {};   // bit 0 = dx, bit 1 = dy, bit 2 = dz
};

// ---- pattern object (femx_pattern.cu) -------------------------------------
// Node-level CSR + scatter map.  All arrays are device memory owned by the
// object.  "pair" = (owned row node, adjacent element) incidence.
struct femx_pattern {
  femx_ctx* ctx = nullptr;
  int nn = 0, nd = 1;
  int64_t n_nodes = 0, n_elems = 0;
  int64_t row_begin = 0, row_end = 0, col_base = 0;
  int64_t n_rows = 0;    // node rows owned
  int64_t nnz_node = 0;  // node-level nonzeros
  int64_t n_pairs = 0;
  int max_row = 0;       // longest node-level row
  int tile_nodes = 0;    // node rows per CTA in the numeric pass
  int64_t max_tile_nnz = 0;
  int64_t max_tile_codes = 0;  // most padded incidences (SELL entries) in one tile
  int2* d_rowinfo = nullptr;       // [n_rows+1] {row_ptr, number of incidences}
  int32_t* d_col_idx = nullptr;    // [nnz_node] LOCAL node id, ascending per row (exports add col_base)
  // SELL-32 scatter map: slice s = rows [32s, 32s+32); incidence `it` of row r sits at
  // slice_ptr[s] + 32*it + r%32 (rows padded to the slice's longest incidence list)
  int32_t* d_slice_ptr = nullptr;  // [n_slices+1]
  uint32_t* d_sell_code = nullptr; // 7-bit row positions of the element's nodes | li<<28
  int32_t* d_sell_elem = nullptr;  // e*nn + li, ascending element order per row
  int64_t n_sell = 0;              // padded incidence count
  int64_t bytes = 0;
  // Dominant stencil class (femx_pattern.cu: detect_stencil_class): the (incidence count, row length,
  // own position, scatter-code sequence, column offsets from the own node) shared by most rows — every interior row of a structured
  // mesh.  Rows of the class carry FEMX_ROW_SPEC in rowinfo.y; the others are listed in d_other_rows.
  // The numeric pass may JIT a straight-line body for the class and run the listed rows separately.
  int spec_np = 0, spec_rlen = 0, spec_self = 0;
  int64_t spec_rows = 0;
  std::vector<uint32_t> spec_codes;  // host copy, spec_np entries
  std::vector<int32_t> spec_off;     // column offsets from the row's own node (local ids), spec_rlen entries
  std::string spec_key;              // identifies the class in the form's kernel cache
  int32_t* d_other_rows = nullptr;   // [n_other] rows outside the class, ascending (device)
  int64_t n_other = 0;
  int max_row_other = 0;             // longest of those rows
  bool map_complete = true;          // false: the scatter map of the class rows is not written yet (femx_pattern_complete_map)
  void* d_lat_tmpl = nullptr;        // row templates of the lattice-templated pass (device), kept for that completion
  int lat_dom = -1;
  femx_lattice lat;                  // lattice structure of the mesh, if it has one (and a class was found)
  int64_t lat_rows = 0;              // class rows = lattice-interior nodes among the owned rows (checked)
};

int femx_pattern_complete_map(const femx_pattern* p, void* stream);

// rowinfo[i].y = #incidences (bits 0-21) | (bit 22 reserved) | FEMX_ROW_SPEC | own position << 24
#define FEMX_NP_MASK 0x3fffff
#define FEMX_ROW_SPEC (1 << 23)
#define FEMX_SPEC_MAX_NP 32    // limits of a specialised stencil (register budget of the straight-line body)
#define FEMX_SPEC_MAX_RLEN 24

#define FEMX_DOT_BLOCKS 1024

// Local index of the j-th OTHER vertex of an incidence whose row node is local vertex li.
// (li, oth(li,0), oth(li,1), ...) is always an EVEN permutation of the element's vertices
// (3-cycles for triangles, XOR by li = two transpositions for tetrahedra), so the signed
// Jacobian of the re-ordered element equals the original one.
static inline __host__ __device__ int femx_oth(int nn, int li, int j) {
  return nn == 4 ? (li ^ (j + 1)) : (li + 1 + j) % 3;
}

// node rows per CTA of the generic numeric pass (knob `tile` overrides: tuning experiments only)
static inline int femx_tile_nodes_for(int nd, const femx_knobs& k) {
  if (k.tile >= 32 && k.tile <= 1024 && k.tile % 32 == 0) return k.tile;
  return nd == 1 ? 128 : 32;
}
