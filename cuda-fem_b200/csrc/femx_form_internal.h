// Private to the form translation units (femx_form.cpp, femx_lattice.cpp): the form object and its
// cache of JIT-compiled kernel variants.
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "femx_internal.h"

struct femx_variant {
  std::string source, log;
  std::vector<char> cubin;
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
  CUfunction fn2 = nullptr;  // lattice variants: femx_rowlist (the rows outside the class)
  int smem_set = 0, smem2_set = 0;
  int carveout_set = 0;
  size_t lt_smem = 0;  // lattice variants: dynamic shared memory the generated kernel needs
  int lt_nslot = 0;
};

// a stencil class handed to the JIT (femx_pattern's dominant class, or an explicit one)
struct femx_stencil_class {
  int np = 0, rlen = 0, self = 0;
  std::vector<uint32_t> codes;
  std::string key;
};

// How the lattice pass runs on one pattern (femx_lattice.cpp)
struct femx_lattice_plan {
  bool ok = false;
  int tx = 16, ty = 16;      // threads per CTA tile: (tx-1) x (ty-1) owned node columns
  int threads = 256;         // tx*ty rounded up to whole warps
  int kc = 32;               // node planes per CTA
  int minb = 1;              // __launch_bounds__ min blocks
  int regs = 0;              // --maxrregcount (0: none)
  int unroll = 1;            // unroll factor of the plane loop (2: the plane roll needs no register moves)
  int pf = 1;                // 1: the next plane is prefetched into L1 and loaded when needed; 0: loaded one cell ahead
  int nslot = 0;             // shared-memory field slots
  int rlen = 0, self = 0;    // the stencil class
  std::vector<int> pos;      // row position of offset (ox,oy,oz): index (oz+1)*9 + (oy+1)*3 + (ox+1), -1 = none
  std::vector<std::pair<int, int>> edges;  // cell edges (corner pairs, first < second)
  std::string fallback;      // set by femx_lattice_defines when the decomposition cannot be handled
  size_t smem = 0;
};
bool femx_lattice_plan_make(const struct femx_form* f, const femx_lattice& L, int rlen, int self,
                            const std::vector<int32_t>& class_off, const femx_knobs& K, femx_lattice_plan* plan,
                            std::string* why);
std::string femx_lattice_defines(const struct femx_form* f, const femx_lattice& L, femx_lattice_plan* plan);
std::string femx_lattice_key(const femx_lattice& L, const femx_lattice_plan& plan);

// what femx_assemble_csr worked out for (pattern, lattice options) on an earlier call: the hot call does not rebuild it
struct femx_lattice_cached {
  std::string opts;            // the option values the entry was made for
  femx_lattice_plan plan;
  femx_variant* variant = nullptr;
};

struct femx_form {
  femx_ctx* ctx = nullptr;
  std::map<const void*, femx_lattice_cached> lt_cache;  // by pattern
  femx_knobs knobs;  // copied at compile time (from the context, or from the environment for offline forms)
  // element-once lattice pass (femx_lattice.cpp): available for the symmetric built-in scalar forms in 3-D;
  // K_ab = (d_a . d_b) lt_W / jac + lt_moff jac (a != b), diagonal = lt_cj * (sum of jac) - (off-diagonal row sum)
  bool lt_ok = false;
  double lt_W = 0.0, lt_moff = 0.0, lt_cj = 0.0;
  std::set<std::string> lt_failed;
  int dim = 2, nn = 3, nd = 1, dtype = FEMX_F64, builtin = 0, fmad = 1;
  int n = 3;  // nn*nd
  int integrated = 0;  // entries are final element-matrix expressions (no quadrature applied)
  std::string prologue;
  std::vector<std::string> entries;  // n*n
  // accumulate form of the built-in scalar entries: acc_pre[li] declares what row li shares, acc_entries[li*n+lj]
  // is the NEW value of an accumulator written $A (one fma chain: no separate product, no separate add)
  std::vector<std::string> acc_pre, acc_entries;
  // the same chains step by step (FEMX_CHAINORDER=1, experiment): acc_steps[li] lists statements over A0..A{n-1}
  // in the order mass, x, y, z for all entries at once, so that consecutive fma share hx / hy / hz (.reuse)
  std::vector<std::string> acc_steps;
  // 3-D scalar built-ins: the prologue is [edges u2,u3,u4 from vertex 1 | d2 = u4 x u3, d3 = u2 x u4, d4 = u3 x u2 |
  // prologue_rest]; the specialised pass then computes each face's cross product once (see build_defines)
  bool shared_faces = false;
  std::string prologue_rest;
  // FEMX_ROWSUM=1 (experiment, off by default): the rows of a stiffness matrix sum to zero, so the diagonal is not
  // accumulated entry by entry but recovered at the end of the row, D = cj * (sum of the incident Jacobians) -
  // (sum of the row's off-diagonal values), cj = c * sum_b M_ab — 4 fma per incidence become one add
  bool rowsum = false;
  double rowsum_cj = 0.0;
  std::vector<std::string> rhs;      // n load-vector integrands (may be empty)
  int rhs_integrated = 0;
  // the built-in entries are invariant under even permutations of the local vertices (see build_defines)
  bool rot_ok_matrix = false, rot_ok_rhs = false;
  int nq = 0;
  std::vector<double> qw, qr, qs, qt, qu;
  std::map<std::string, femx_variant> variants;
  std::set<std::string> spec_failed;  // stencil classes whose specialised kernel did not compile
  std::string last_source, last_log;
  mutable std::string err;
};
