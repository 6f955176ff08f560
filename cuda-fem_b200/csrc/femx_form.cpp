// Integrand emitter (stands in for GiNaC's WeakForm::build), NVRTC compilation
// to an sm_100a cubin, and the launches of the two JIT kernels.
#include <nvrtc.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>
#include <sstream>

#include "femx_form_internal.h"
#include "femx_jit_src.h"

namespace {

typedef femx_variant Variant;
typedef femx_stencil_class StencilClass;

std::string num(double v) {
  char b[64];
  snprintf(b, sizeof b, "%.17g", v);
  std::string s = b;
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";  // keep it a double literal
  return "real(" + s + ")";
}

const char* AX[3] = {"x", "y", "z"};

}  // namespace

namespace {

// ---- built-in forms ---------------------------------------------------------
// Geometry follows FunctionSpace (fea_symbolic_nvrtc_sparse.cpp:239-289) and the
// shape-function derivatives sfR_deriv/sfS_deriv (:68-101):
//   jac = (x1-x3)(y2-y3) - (y1-y3)(x2-x3)
//   grad r = ((y2-y3), (x3-x2))/jac   grad s = ((y3-y1), (x1-x3))/jac
// phi = (r, s, 1-r-s); entries are a(u=phi_lj, v=phi_li)*jac (:337, Q6).
// Unlike GiNaC's fully expanded strings the common sub-expressions live in a
// prologue evaluated once per element.
// The emitter pre-integrates: for P1 simplices the Jacobian is constant, so
//   sum_q w_q (grad phi_b . grad phi_a) jac = (d_b . d_a) * (W / jac),   W = sum_q w_q,
//   sum_q w_q  phi_b phi_a jac             = M_ab * jac,                 M_ab = sum_q w_q phi_a phi_b
// with d_a = jac * grad phi_a (no division) and W, M_ab folded on the host from the
// SAME quadrature rule (the reference's 8-digit literals in 2-D), so the values
// equal the reference's quadrature sums up to rounding.
// Every multiply-add is spelled out (fma / femx_mul) so that the rounding of an entry does not
// depend on the compiler's contraction choices — the specialised and the generic numeric pass, the
// COO kernel and any slab of a partitioned mesh produce the same bits.  Edges are taken from local
// vertex 1: in the numeric pass that is the row's own node, so neighbouring incidences share them.
//   2-D: d1 = (y2-y3, x3-x2), d2 = (y3-y1, x1-x3), d3 = -(d1+d2);  jac = d2y d1x - d2x d1y
//        (= (x1-x3)(y2-y3)-(y1-y3)(x2-x3), fea_symbolic_nvrtc_sparse.cpp:258)
//   3-D: d2 = u4 x u3, d3 = u2 x u4, d4 = u3 x u2, d1 = -(d2+d3+d4);  jac = u2 . d2
//        (= det[x1-x4, x2-x4, x3-x4], the Jacobian of X = x1 r + x2 s + x3 t + x4 (1-r-s-t))
// Component k of a x b as a difference of two rounded products: swapping a and b flips the sign and
// nothing else (rn(p) - rn(q) = -(rn(q) - rn(p)) exactly), so the cross product of a face can be shared,
// negated, by the two tetrahedra on either side of it.  (Fused, fma(p, q, -rn(r s)), it could not.)
std::string cross_c(const std::string& a, const std::string& b, int k) {
  static const char* ax[3] = {"x", "y", "z"};
  const char *i = ax[(k + 1) % 3], *j = ax[(k + 2) % 3];
  std::ostringstream o;
  o << "(femx_mul(" << a << i << "," << b << j << ")-femx_mul(" << a << j << "," << b << i << "))";
  return o.str();
}

// rest (3-D pinned only): what follows the edges u2,u3,u4 and the cross products d2,d3,d4
void emit_geometry(int dim, bool pinned, std::string* pro, std::string* rest = nullptr) {
  std::ostringstream o;
  if (!pinned) {
    // vector forms (no specialised pass): plain expressions, contraction left to the compiler
    if (dim == 2) {
      o << "const real d1x = y2-y3, d1y = x3-x2;\n"
           "  const real d2x = y3-y1, d2y = x1-x3;\n"
           "  const real d3x = -(d1x+d2x), d3y = -(d1y+d2y);\n"
           "  const real jac = d2y*d1x-d2x*d1y;\n"   // (x1-x3)(y2-y3)-(y1-y3)(x2-x3)
           "  const real ijac = femx_rcp(jac);\n";
    } else {
      // X = x1 r + x2 s + x3 t + x4 (1-r-s-t); J[c][a] = dX_c/dref_a; d_a = jac * (row a of J^-1)
      o << "const real j00 = x1-x4, j01 = x2-x4, j02 = x3-x4;\n"
           "  const real j10 = y1-y4, j11 = y2-y4, j12 = y3-y4;\n"
           "  const real j20 = z1-z4, j21 = z2-z4, j22 = z3-z4;\n"
           "  const real d1x = j11*j22-j12*j21, d1y = j02*j21-j01*j22, d1z = j01*j12-j02*j11;\n"
           "  const real d2x = j12*j20-j10*j22, d2y = j00*j22-j02*j20, d2z = j02*j10-j00*j12;\n"
           "  const real d3x = j10*j21-j11*j20, d3y = j01*j20-j00*j21, d3z = j00*j11-j01*j10;\n"
           "  const real d4x = -(d1x+d2x+d3x), d4y = -(d1y+d2y+d3y), d4z = -(d1z+d2z+d3z);\n"
           "  const real jac = j00*d1x+j01*d2x+j02*d3x;\n"
           "  const real ijac = femx_rcp(jac);\n";
    }
  } else if (dim == 2) {
    o << "const real d1x = y2-y3, d1y = x3-x2;\n"
         "  const real d2x = y3-y1, d2y = x1-x3;\n"
         "  const real d3x = -(d1x+d2x), d3y = -(d1y+d2y);\n"
         "  const real jac = fma(d2y,d1x,-femx_mul(d2x,d1y));\n"   // (x1-x3)(y2-y3)-(y1-y3)(x2-x3)
         "  const real ijac = femx_rcp(jac);\n";
  } else {
    o << "const real u2x = x2-x1, u2y = y2-y1, u2z = z2-z1;\n"
         "  const real u3x = x3-x1, u3y = y3-y1, u3z = z3-z1;\n"
         "  const real u4x = x4-x1, u4y = y4-y1, u4z = z4-z1;\n";
    const char* pairs[3][3] = {{"d2", "u4", "u3"}, {"d3", "u2", "u4"}, {"d4", "u3", "u2"}};
    for (auto& pr : pairs) {
      o << "  const real ";
      for (int k = 0; k < 3; ++k) o << (k ? ", " : "") << pr[0] << "xyz"[k] << " = " << cross_c(pr[1], pr[2], k);
      o << ";\n";
    }
    const char* tail = "  const real d1x = -(d2x+d3x+d4x), d1y = -(d2y+d3y+d4y), d1z = -(d2z+d3z+d4z);\n"
                       "  const real jac = fma(u2z,d2z,fma(u2y,d2y,femx_mul(u2x,d2x)));\n"
                       "  const real ijac = femx_rcp(jac);\n";
    o << tail;
    if (rest) *rest = tail;
  }
  *pro += o.str();
}

std::string dd(int dim, int a, int b, bool pinned = true) {  // d_a . d_b (1-based names)
  std::ostringstream o;
  if (!pinned) {
    o << "(";
    for (int k = 0; k < dim; ++k) {
      if (k) o << "+";
      o << "d" << a + 1 << AX[k] << "*d" << b + 1 << AX[k];
    }
    o << ")";
    return o.str();
  }
  for (int k = dim - 1; k >= 1; --k) o << "fma(d" << a + 1 << AX[k] << ",d" << b + 1 << AX[k] << ",";
  o << "femx_mul(d" << a + 1 << AX[0] << ",d" << b + 1 << AX[0] << ")";
  for (int k = dim - 1; k >= 1; --k) o << ")";
  return o.str();
}

double phi_at(const femx_form* f, int a, int q) {
  if (a == 0) return f->qr[q];
  if (a == 1) return f->qs[q];
  if (f->dim == 2) return 1.0 - f->qr[q] - f->qs[q];
  if (a == 2) return f->qt[q];
  return 1.0 - f->qr[q] - f->qs[q] - f->qt[q];
}

int emit_builtin(femx_form* f, const femx_form_desc* d) {
  const int dim = f->dim, nn = f->nn, nd = f->nd, n = f->n;
  const bool pinned = nd == 1;  // scalar forms: every rounding fixed by the text (see emit_geometry)
  emit_geometry(dim, pinned, &f->prologue, &f->prologue_rest);
  f->shared_faces = pinned && dim == 3 && f->knobs.sharedfaces != 0;
  f->integrated = 1;
  f->entries.assign((size_t)n * n, "");
  double W = 0.0;
  for (int q = 0; q < f->nq; ++q) W += f->qw[q];
  double M[4][4];
  for (int a = 0; a < nn; ++a)
    for (int b = 0; b < nn; ++b) {
      double m = 0.0;
      for (int q = 0; q < f->nq; ++q) m += f->qw[q] * (phi_at(f, b, q) * phi_at(f, a, q));
      M[a][b] = m;
    }
  // Are the folded quadrature constants symmetric in the vertices?  (The reference's 8-digit 2-D rule
  // is not, at the 1e-9 level, so 2-D forms with a mass term / load vector keep one case per li.)
  bool msym = true, vsym = true;
  double mvec[4];
  for (int a = 0; a < nn; ++a) {
    mvec[a] = 0.0;
    for (int q = 0; q < f->nq; ++q) mvec[a] += f->qw[q] * phi_at(f, a, q);
  }
  for (int a = 0; a < nn; ++a) {
    if (std::fabs(mvec[a] - mvec[0]) > 1e-15 * std::fabs(mvec[0])) vsym = false;
    for (int b = 0; b < nn; ++b) {
      const double ref = a == b ? M[0][0] : M[0][1];
      if (std::fabs(M[a][b] - ref) > 1e-15 * std::fabs(ref)) msym = false;
    }
  }
  const bool has_mass = d->builtin == FEMX_FORM_POISSON_MASS || d->builtin == FEMX_FORM_MASS;
  f->rot_ok_matrix = !has_mass || msym;
  f->rot_ok_rhs = vsym;
  if (pinned && dim == 3 && f->rot_ok_matrix && d->builtin != FEMX_FORM_ELASTICITY) {
    // element-once lattice pass (femx_lattice.cpp): symmetric entries, diagonal from the zero row sum of the
    // stiffness part: K_aa = c (sum_b M_ab) jac - sum_{b != a} K_ab
    const double cmass = d->builtin == FEMX_FORM_POISSON ? 0.0 : (d->builtin == FEMX_FORM_MASS ? 1.0 : (d->params[0] != 0.0 ? d->params[0] : 1.0));
    f->lt_ok = true;
    f->lt_W = W;
    f->lt_moff = cmass * M[0][1];
    f->lt_cj = 0.0;
    for (int b = 0; b < nn; ++b) f->lt_cj += cmass * M[0][b];
  }
  std::ostringstream pro;
  if (pinned) {
    pro << "  const real kq = femx_mul(" << num(W) << ",ijac);\n";
    f->prologue_rest += pro.str();
  }
  else pro << "  const real kq = " << num(W) << "*ijac;\n";
  if (d->builtin == FEMX_FORM_ELASTICITY) {
    if (nd != dim) return FEMX_ERR_INVALID;
    for (int a = 0; a < nn; ++a)
      for (int b = a; b < nn; ++b)
        pro << "  const real dd" << a + 1 << b + 1 << " = " << dd(dim, a, b, false) << ";\n";
    pro << "  const real LAMq = " << num(d->params[0]) << "*kq, MUq = " << num(d->params[1]) << "*kq;\n";
  }
  f->prologue += pro.str();
  double cm = d->params[0] != 0.0 ? d->params[0] : 1.0;
  for (int li = 0; li < n; ++li)
    for (int lj = 0; lj < n; ++lj) {
      const int a = li / nd, c = li % nd, b = lj / nd, e = lj % nd;
      std::ostringstream o;
      switch (d->builtin) {
        case FEMX_FORM_POISSON:
          o << "femx_mul(" << dd(dim, b, a) << ",kq)";
          break;
        case FEMX_FORM_POISSON_MASS:
          o << "fma(" << dd(dim, b, a) << ",kq,femx_mul(" << num(cm * M[a][b]) << ",jac))";
          break;
        case FEMX_FORM_MASS:
          o << "femx_mul(" << num(M[a][b]) << ",jac)";
          break;
        case FEMX_FORM_ELASTICITY: {
          const int lo = a < b ? a : b, hi = a < b ? b : a;
          o << "LAMq*(d" << a + 1 << AX[c] << "*d" << b + 1 << AX[e] << ")+MUq*(";
          if (c == e) o << "dd" << lo + 1 << hi + 1 << "+";
          o << "d" << a + 1 << AX[e] << "*d" << b + 1 << AX[c] << ")";
          break;
        }
        default:
          return FEMX_ERR_INVALID;
      }
      f->entries[(size_t)li * n + lj] = o.str();
    }
  // Accumulate form (scalar forms): K_ab added to an accumulator A in one chain,
  //   A <- d_b . (kq d_a) + (c M_ab jac + A),  with h = kq d_a shared by the row
  // (3-D only: there the numeric pass is bound by the fp64 pipe; in 2-D it is not, and the form costs the
  //  generic kernel 8 registers = 2 resident CTAs)
  if (pinned && dim == 3 && d->builtin != FEMX_FORM_ELASTICITY && f->knobs.accf != 0) {
    f->acc_pre.assign(n, "");
    f->acc_entries.assign((size_t)n * n, "");
    f->rowsum = f->rot_ok_matrix && (d->builtin == FEMX_FORM_POISSON || d->builtin == FEMX_FORM_POISSON_MASS) &&
                f->knobs.rowsum == 1;
    if (f->rowsum && d->builtin == FEMX_FORM_POISSON_MASS)
      for (int b = 0; b < nn; ++b) f->rowsum_cj += cm * M[0][b];
    for (int a = 0; a < n; ++a) {
      std::ostringstream pre;
      if (d->builtin != FEMX_FORM_MASS) {
        pre << "const real";
        for (int k = 0; k < dim; ++k) pre << (k ? "," : "") << " h" << AX[k] << " = femx_mul(kq,d" << a + 1 << AX[k] << ")";
        pre << ";";
      }
      f->acc_pre[a] = pre.str();
      for (int b = 0; b < n; ++b) {
        std::ostringstream o;
        std::string inner = "$A";
        if (d->builtin != FEMX_FORM_POISSON) inner = "fma(" + num((d->builtin == FEMX_FORM_MASS ? 1.0 : cm) * M[a][b]) + ",jac,$A)";
        if (d->builtin == FEMX_FORM_MASS) {
          o << inner;
        } else {
          for (int k = dim - 1; k >= 0; --k) o << "fma(d" << b + 1 << AX[k] << ",h" << AX[k] << ",";
          o << inner;
          for (int k = dim - 1; k >= 0; --k) o << ")";
        }
        f->acc_entries[(size_t)a * n + b] = o.str();
        if (f->rowsum && a == b) f->acc_entries[(size_t)a * n + b] = d->builtin == FEMX_FORM_POISSON ? "$A" : "$A+jac";
      }
    }
    if (d->builtin != FEMX_FORM_MASS && f->knobs.chainorder == 1) {
      f->acc_steps.assign(n, "");
      for (int a = 0; a < n; ++a) {
        std::ostringstream o;
        auto skip = [&](int b) { return f->rowsum && a == b; };
        for (int b = 0; b < n; ++b)
          if (skip(b)) { if (d->builtin == FEMX_FORM_POISSON_MASS) o << " A" << b << " = A" << b << "+jac;"; }
          else if (d->builtin == FEMX_FORM_POISSON_MASS) o << " A" << b << " = fma(" << num(cm * M[a][b]) << ",jac,A" << b << ");";
        for (int k = 0; k < dim; ++k)
          for (int b = 0; b < n; ++b)
            if (!skip(b)) o << " A" << b << " = fma(d" << b + 1 << AX[k] << ",h" << AX[k] << ",A" << b << ");";
        f->acc_steps[a] = o.str();
      }
    }
  }
  // constant source: b[a,c] = f_c * (sum_q w_q phi_a(q)) * jac, pre-integrated like the matrix
  f->rhs.assign(n, "");
  f->rhs_integrated = 1;
  for (int a = 0; a < nn; ++a) {
    double m = 0.0;
    for (int q = 0; q < f->nq; ++q) m += f->qw[q] * phi_at(f, a, q);
    for (int c = 0; c < nd; ++c) f->rhs[(size_t)a * nd + c] = num(d->rhs_vec[c] * m) + "*jac";
  }
  return FEMX_OK;
}

void default_rule(femx_form* f) {
  if (f->dim == 2) {
    // the reference's literals, fea_symbolic_nvrtc_sparse.cpp:380-383
    static const double w[7] = {0.06296959, 0.06619708, 0.06296959, 0.06619708,
                                0.06296959, 0.06619708, 0.11250000};
    static const double r[7] = {0.10128651, 0.47014206, 0.79742699, 0.47014206,
                                0.10128651, 0.05971587, 0.33333333};
    static const double s[7] = {0.10128651, 0.05971587, 0.10128651, 0.47014206,
                                0.79742699, 0.47014206, 0.33333333};
    static const double t[7] = {0.79742698, 0.47014207, 0.1012865, 0.05971588,
                                0.1012865,  0.47014207, 0.33333334};
    f->nq = 7;
    f->qw.assign(w, w + 7);
    f->qr.assign(r, r + 7);
    f->qs.assign(s, s + 7);
    f->qt.assign(t, t + 7);
    f->qu.assign(7, 0.0);
  } else {
    const double a = 0.5854101966249685, b = 0.1381966011250105;
    f->nq = 4;
    f->qw.assign(4, 1.0 / 24.0);
    f->qr = {a, b, b, b};
    f->qs = {b, a, b, b};
    f->qt = {b, b, a, b};
    f->qu.resize(4);
    for (int q = 0; q < 4; ++q) f->qu[q] = 1.0 - f->qr[q] - f->qs[q] - f->qt[q];
  }
}

// Does a C expression mention the quadrature point (identifiers r, s, t, u)?
bool depends_on_q(const std::string& e) {
  size_t i = 0;
  while (i < e.size()) {
    unsigned char ch = (unsigned char)e[i];
    if (isalpha(ch) || ch == '_') {
      size_t j = i;
      while (j < e.size() && (isalnum((unsigned char)e[j]) || e[j] == '_')) ++j;
      std::string id = e.substr(i, j - i);
      if (id == "r" || id == "s" || id == "t" || id == "u") return true;
      i = j;
    } else if (isdigit(ch) || (ch == '.' && i + 1 < e.size() && isdigit((unsigned char)e[i + 1]))) {
      // skip a numeric literal incl. exponent / suffix so that "1.0f" or "2e3" is not an identifier
      size_t j = i;
      while (j < e.size() && (isalnum((unsigned char)e[j]) || e[j] == '.' ||
                              ((e[j] == '+' || e[j] == '-') && (e[j - 1] == 'e' || e[j - 1] == 'E'))))
        ++j;
      i = j;
    } else {
      ++i;
    }
  }
  return false;
}

// Per matrix row li:
//   FEMX_ROWC_<li>              entries that do not depend on the quadrature point:
//                               out[lj] = E (pre-integrated) or out[lj] = (sum_q w_q) * E
//   FEMX_ROWQ_<li>(R,S,T,U,W)   one quadrature-point update of the entries that do
// tile_nodes: threads per CTA of the numeric pass (the pattern's tile, or the lattice pass's CTA size)
std::string build_defines(const femx_form* f, const std::string& kernel, const StencilClass* sc, int tile_nodes) {
  std::ostringstream o;
  const int n = f->n;
  const femx_knobs& K = f->knobs;
  if (tile_nodes <= 0) tile_nodes = femx_tile_nodes_for(f->nd, K);
  o << "#define FEMX_REAL " << (f->dtype == FEMX_F32 ? "float" : "double") << "\n";
  o << "#define NN " << f->nn << "\n#define ND " << f->nd << "\n#define DIM " << f->dim
    << "\n#define FEMX_TILE_NODES " << tile_nodes << "\n";
  // (the specialised 3-D body holds 45 coordinates + 15 accumulators per thread: 512 threads per SM keep it at
  //  128 registers without spills, measured faster than the 164 the compiler takes when left alone)
  // (2-D body: 1024 threads per SM = 64 registers; measured 0.27 ms on cfg2 against 0.29-0.31 ms at the 48, 56 or
  //  68 registers other limits produce — the schedule ptxas finds at this limit, not the occupancy, makes the difference)
  const int min_blocks_default = !sc ? 0 : (f->dim == 3 ? 512 : 1024) / tile_nodes;
  o << "#define FEMX_MIN_BLOCKS " << (K.minblocks >= 0 ? K.minblocks : min_blocks_default) << "\n";
  o << "#define FEMX_MIDGATHER " << K.midgather << "\n";
  o << "#define FEMX_UNROLL " << K.unroll << "\n";
  o << "#define FEMX_ROWSUM " << (f->rowsum ? 1 : 0) << "\n#define FEMX_CJ " << num(f->rowsum_cj) << "\n";
  o << "#define FEMX_LISTLAST " << K.listlast << "\n";
  o << "#define FEMX_RCP3 " << K.rcp3 << "\n";
  o << "#define FEMX_EXPANDED " << (kernel == "csr_x" ? 1 : 0) << "\n";
  o << "#define FEMX_UNIT_STRIDE " << (kernel == "csr" ? 1 : 0) << "\n";
  std::string esc;
  for (char ch : f->prologue) {
    if (ch == '\n') esc += " \\\n"; else esc += ch;
  }
  o << "#define FEMX_PROLOGUE " << esc << "\n";
  double W = 0.0;
  for (int q = 0; q < f->nq; ++q) W += f->qw[q];
  std::vector<int> has_q(n, 0);
  for (int li = 0; li < n; ++li) {
    std::ostringstream c, qd;
    for (int lj = 0; lj < n; ++lj) {
      const std::string& e = f->entries[(size_t)li * n + lj];
      if (f->integrated)
        c << " \\\n    out[" << lj << "] = (" << e << ");";
      else if (!depends_on_q(e))
        c << " \\\n    out[" << lj << "] = " << num(W) << "*(" << e << ");";
      else {
        qd << " \\\n    out[" << lj << "] += w*(" << e << ");";
        has_q[li] = 1;
      }
    }
    o << "#define FEMX_ROWC_" << li << c.str() << "\n";
    o << "#define FEMX_ROWQ_" << li << "(R,S,T,U,W) { const real r = (R), s = (S), t = (T), u = (U), w = (W); "
         "(void)r; (void)s; (void)t; (void)u; (void)w;" << qd.str() << " }\n";
  }
  const bool accf = !f->acc_entries.empty();
  auto subst = [](std::string e, const std::string& acc) {
    for (size_t p0 = e.find("$A"); p0 != std::string::npos; p0 = e.find("$A", p0 + acc.size())) e.replace(p0, 2, acc);
    return e;
  };
  if (accf)
    for (int li = 0; li < n; ++li) {
      o << "#define FEMX_ROWA_" << li << "(";
      for (int lj = 0; lj < n; ++lj) o << (lj ? "," : "") << "A" << lj;
      o << ") { " << f->acc_pre[li];
      if (!f->acc_steps.empty()) {
        o << " \\\n   " << f->acc_steps[li];
      } else {
        for (int lj = 0; lj < n; ++lj)
          o << " \\\n    A" << lj << " = " << subst(f->acc_entries[(size_t)li * n + lj], "A" + std::to_string(lj)) << ";";
      }
      o << " }\n";
    }
  o << "#define FEMX_QUAD(M)";
  for (int q = 0; q < f->nq; ++q)
    o << " \\\n    M(" << num(f->qr[q]) << "," << num(f->qs[q]) << "," << num(f->qt[q]) << ","
      << num(f->qu[q]) << "," << num(f->qw[q]) << ")";
  o << "\n#define FEMX_ROW_CASES";
  for (int li = 0; li < n; ++li) {
    o << " \\\n    case " << li << ": FEMX_ROWC_" << li;
    if (has_q[li]) o << " FEMX_QUAD(FEMX_ROWQ_" << li << ")";
    o << " break;";
  }
  o << "\n";
  // Load vector: case per local node a; racc[c] += integrated rhs entry (a*ND + c)
  // Built-in forms are intrinsic (independent of the local vertex numbering), so every incidence
  // can be evaluated as "row 0 of the element (own node, others...)" — an even permutation of the
  // element, same signed Jacobian: one case, no switch, no divergence on unstructured meshes.
  // Custom strings name local vertices explicitly and keep one case per li.
  const bool rot_allowed = f->builtin != FEMX_FORM_CUSTOM && K.rotinv != 0;
  const bool is_rhs = kernel == "rhs";
  const bool rotinv = rot_allowed && (is_rhs ? (f->rot_ok_rhs && f->rhs_integrated) : f->rot_ok_matrix);
  o << "#define FEMX_ROTINV " << (rotinv ? 1 : 0) << "\n";
  o << "#define FEMX_RHS_CASES";
  if (!f->rhs.empty()) {
    for (int a = 0; a < (rotinv ? 1 : f->nn); ++a) {
      o << " \\\n    case " << a << ": {";
      for (int k = 0; k < f->dim; ++k) {
        const char AXU = (char)toupper("xyz"[k]);
        o << " const real " << "xyz"[k] << a + 1 << " = S" << AXU << ";";
        for (int j = 0; j < f->nn - 1; ++j)
          o << " const real " << "xyz"[k] << femx_oth(f->nn, a, j) + 1 << " = O" << AXU << "[" << j << "];";
      }
      o << " \\\n      FEMX_PROLOGUE";
      for (int c = 0; c < f->nd; ++c) {
        const std::string& e = f->rhs[(size_t)a * f->nd + c];
        if (f->rhs_integrated)
          o << " \\\n      racc[" << c << "] += (" << e << ");";
        else if (!depends_on_q(e))
          o << " \\\n      racc[" << c << "] += " << num(W) << "*(" << e << ");";
        else {
          o << " \\\n      {";
          for (int q = 0; q < f->nq; ++q)
            o << " { const real r = " << num(f->qr[q]) << ", s = " << num(f->qs[q]) << ", t = " << num(f->qt[q])
              << ", u = " << num(f->qu[q]) << "; (void)r; (void)s; (void)t; (void)u; racc[" << c << "] += "
              << num(f->qw[q]) << "*(" << e << "); }";
          o << " }";
        }
      }
      o << " } break;";
    }
  }
  o << "\n";
  // COO, thread per element: every row in straight-line code
  o << "#define FEMX_COO_ROWS";
  for (int li = 0; li < n; ++li) {
    o << " \\\n    { real out[NDOF];";
    if (has_q[li]) o << " _Pragma(\"unroll\") for (int j_ = 0; j_ < NDOF; ++j_) out[j_] = real(0);";
    o << " FEMX_ROWC_" << li;
    if (has_q[li]) o << " FEMX_QUAD(FEMX_ROWQ_" << li << ")";
    o << " FEMX_COO_STORE(" << li << ") }";
  }
  o << "\n";
  // Numeric pass: one case per dof row li = a*ND + c of the element matrix.  In case (a, c) the
  // row's own node is local node a (coordinates sx,sy,sz) and the other vertices follow in
  // cyclic order (ox[j] = local node (a+1+j) % NN), so every name binding, the diagonal
  // accumulator and the scatter positions po[j] are static inside the case.
  const int nn = f->nn, nd = f->nd;
  static const char* ax[3] = {"x", "y", "z"};
  // One thread owns a NODE row: geometry once per incidence, then the ND dof rows of the element
  // matrix one after the other (component c), each into its own dof-row segment of the image.
  o << "#define FEMX_CSR_CASES";
  for (int a = 0; a < (rotinv ? 1 : nn); ++a) {
    o << " \\\n    case " << a << ": {";
    for (int k = 0; k < f->dim; ++k) {
      o << " const real " << ax[k] << a + 1 << " = s" << ax[k] << ";";
      for (int j = 0; j < nn - 1; ++j)
        o << " const real " << ax[k] << femx_oth(nn, a, j) + 1 << " = o" << ax[k] << "[" << j << "];";
    }
    o << " \\\n      FEMX_PROLOGUE FEMX_GATHER_NEXT";
    if (accf) {
      // accumulate form: the slots' old values (0 on first touch) go through the row's fma chains
      o << " \\\n      { real t_[" << nn - 1 << "];";
      for (int j = 0; j < nn - 1; ++j) o << " t_[" << j << "] = FEMX_FIRST(" << j << ") ? real(0) : srow[po[" << j << "]];";
      o << " \\\n        FEMX_ROWA_" << a << "(";
      for (int lj = 0; lj < nn; ++lj) {
        if (lj) o << ",";
        if (lj == a) o << "dacc[0]";
        else o << "t_[" << (nn == 4 ? (lj ^ a) - 1 : (lj - a - 1 + 3) % 3) << "]";
      }
      o << ")";
      for (int j = 0; j < nn - 1; ++j) o << " srow[po[" << j << "]] = t_[" << j << "];";
      o << " }";
    }
    for (int c = 0; c < (accf ? 0 : nd); ++c) {
      const int li = a * nd + c;
      o << " \\\n      { real out[NDOF];";
      if (has_q[li]) o << " _Pragma(\"unroll\") for (int j_ = 0; j_ < NDOF; ++j_) out[j_] = real(0);";
      o << " FEMX_ROWC_" << li;
      if (has_q[li]) o << " FEMX_QUAD(FEMX_ROWQ_" << li << ")";
      for (int d = 0; d < nd; ++d) o << " \\\n        dacc[" << c * nd + d << "] += out[" << a * nd + d << "];";
      // off-diagonal blocks: the (nn-1)*nd slots are distinct (the pattern build rejects elements
      // with a repeated node), so all loads are issued before the first store — no serialised
      // load/add/store chain through possibly-aliasing shared-memory addresses
      o << " \\\n        { real* q_ = srow + " << c << " * rstride; real t_[" << (nn - 1) * nd << "];";
      for (int j = 0; j < nn - 1; ++j)
        for (int d = 0; d < nd; ++d)
          o << " t_[" << j * nd + d << "] = FEMX_FIRST(" << j << ") ? real(0) : q_[po[" << j << "] + " << d << "];";
      for (int j = 0; j < nn - 1; ++j)
        for (int d = 0; d < nd; ++d)
          o << " \\\n          q_[po[" << j << "] + " << d << "] = t_[" << j * nd + d << "] + out["
            << femx_oth(nn, a, j) * nd + d << "];";
      o << " }";
      o << " }";
    }
    o << " } break;";
  }
  o << "\n";
  // Specialised body for one stencil class (scalar forms): the scatter codes are known here, so the
  // incidence loop is unrolled on the host with every column position a literal — the coordinates of
  // each column are loaded once into named registers (from own node + the class's constant offset), the
  // row's values accumulate in registers (first touch: 0 + v, exactly as the generic path's shared-memory
  // image) and each is stored once, by `if (mine)`: the body itself runs unconditionally.
  // Per incidence the arithmetic is the generic case's: same bindings, same prologue, same row macro.
  o << "#define FEMX_SPEC " << (sc ? 1 : 0) << "\n";
  if (sc) {
    const int dim = f->dim;
    // Streaming order: a column's coordinates are loaded `ahead` incidences before its first use and its
    // accumulator is stored right after its last one, so only the columns of the incidences in flight are
    // live (7 of 14 neighbours on the Kuhn stencil) — fewer registers, more resident warps.
    // FEMX_SPEC_AHEAD >= the incidence count loads everything up front.
    // (2-D stencils are small: everything up front measured fastest there)
    const int ahead = K.spec_ahead >= 0 ? K.spec_ahead : (dim == 2 ? 99 : 2);
    std::vector<int> first(sc->rlen, sc->np), last(sc->rlen, -1);
    for (int it = 0; it < sc->np; ++it)
      for (int j = 0; j < nn - 1; ++j) {
        const int pos = (int)((sc->codes[it] >> (7 * j)) & 127);
        first[pos] = std::min(first[pos], it);
        last[pos] = it;
      }
    first[sc->self] = 0;
    const bool pin = K.spec_pin != 0;
    auto load_col = [&](int k, bool entry) {
      o << " \\\n    const i64 q" << k << "_ = (i64)min(max(node_ + soff.v[" << k << "], 0), node_max) * FEMX_CS; const real";
      for (int c = 0; c < dim; ++c)
        o << (c ? "," : "") << " c" << ax[c] << k << " = " << (entry && pin ? "femx_ldg_pinned(" : "__ldg(") << "XYZ"[c] << " + q" << k << "_)";
      o << ";";
    };
    // 3-D scalar built-ins: every face (own node, p, q) belongs to two incident tetrahedra, which need
    // u_p x u_q with opposite signs.  The cross product is rounded so that the swap is an exact negation
    // (cross_c), hence computed once per face and negated for the second tetrahedron — same bits as the
    // generic loop, which evaluates it in each tetrahedron.
    const bool faces = f->shared_faces && rotinv && dim == 3;
    std::map<std::pair<int, int>, bool> face_done;
    if (faces) {
      std::string esc2;
      for (char ch : f->prologue_rest) {
        if (ch == '\n') esc2 += " \\\n"; else esc2 += ch;
      }
      o << "#define FEMX_PROLOGUE_REST " << esc2 << "\n";
    }
    // Experiment (off: measured slower, the volatile prefetches cost spills): FEMX_SPEC_PREFETCH=1 prefetches
    // the columns loaded later into L1 at kernel entry, =2 into L2.
    const int pf = K.spec_prefetch;
    o << "#define FEMX_SPEC_LOAD";
    for (int k = 0; k < sc->rlen; ++k)
      if (first[k] <= ahead) load_col(k, true);  // issued at kernel entry, before any metadata has arrived
    if (pf)
      for (int k = 0; k < sc->rlen; ++k)
        if (first[k] > ahead) {
          o << " \\\n    { const i64 p_ = (i64)min(max(node_ + soff.v[" << k << "], 0), node_max) * FEMX_CS;";
          for (int c = 0; c < dim; ++c)
            o << " asm volatile(\"prefetch.global." << (pf == 2 ? "L2" : "L1") << " [%0];\" ::\"l\"(" << "XYZ"[c] << " + p_));";
          o << " }";
        }
    o << "\n#define FEMX_SPEC_BODY";
    o << " \\\n    real dacc0_ = real(0);";
    for (int it = 0; it < sc->np; ++it) {
      const uint32_t code = sc->codes[it];
      const int a = rotinv ? 0 : (int)((code >> 28) & 3);
      for (int k = 0; k < sc->rlen; ++k) {
        if (first[k] == it + ahead && first[k] > ahead) load_col(k, false);
        if (first[k] == it && k != sc->self) {
          o << " real a" << k << (accf ? "_ = real(0);" : "_;");
          if (faces) {  // edge from the own node
            o << " const real";
            for (int c = 0; c < dim; ++c)
              o << (c ? "," : "") << " e" << k << ax[c] << " = c" << ax[c] << k << "-c" << ax[c] << sc->self;
            o << ";";
          }
        }
      }
      int P[3] = {0, 0, 0};
      for (int j = 0; j < nn - 1 && j < 3; ++j) P[j] = (int)((code >> (7 * j)) & 127);  // positions of local vertices 2,3,4
      static const int fa[3][2] = {{2, 1}, {0, 2}, {1, 0}};  // d2 = u4 x u3, d3 = u2 x u4, d4 = u3 x u2
      if (faces)
        for (auto& pq : fa) {
          const int lo = std::min(P[pq[0]], P[pq[1]]), hi = std::max(P[pq[0]], P[pq[1]]);
          if (face_done[{lo, hi}]) continue;
          face_done[{lo, hi}] = true;
          o << " \\\n    const real";
          for (int c = 0; c < 3; ++c)
            o << (c ? "," : "") << " n" << lo << "_" << hi << ax[c] << " = "
              << cross_c("e" + std::to_string(lo), "e" + std::to_string(hi), c);
          o << ";";
        }
      o << " \\\n    {";
      if (faces) {
        for (int v = 0; v < 3; ++v) {
          o << " const real";
          for (int c = 0; c < 3; ++c) o << (c ? "," : "") << " u" << v + 2 << ax[c] << " = e" << P[v] << ax[c];
          o << ";";
        }
        for (int v = 0; v < 3; ++v) {
          const int p0 = P[fa[v][0]], q0 = P[fa[v][1]];
          o << " const real";
          for (int c = 0; c < 3; ++c)
            o << (c ? "," : "") << " d" << v + 2 << ax[c] << " = " << (p0 < q0 ? "" : "-") << "n" << std::min(p0, q0) << "_"
              << std::max(p0, q0) << ax[c];
          o << ";";
        }
      } else {
        for (int c = 0; c < dim; ++c) {
          o << " const real " << ax[c] << a + 1 << " = c" << ax[c] << sc->self << ";";
          for (int j = 0; j < nn - 1; ++j)
            o << " const real " << ax[c] << femx_oth(nn, a, j) + 1 << " = c" << ax[c] << ((code >> (7 * j)) & 127) << ";";
        }
      }
      const char* PRO = faces ? "FEMX_PROLOGUE_REST" : "FEMX_PROLOGUE";
      if (accf) {
        o << " \\\n      " << PRO << " \\\n      FEMX_ROWA_" << a << "(";
        for (int lj = 0; lj < nn; ++lj) {
          if (lj) o << ",";
          if (lj == a) o << "dacc0_";
          else o << "a" << ((code >> (7 * (nn == 4 ? (lj ^ a) - 1 : (lj - a - 1 + 3) % 3))) & 127) << "_";
        }
        o << ") }";
      } else {
        o << " \\\n      " << PRO << " \\\n      { real out[NDOF];";
        if (has_q[a]) o << " _Pragma(\"unroll\") for (int j_ = 0; j_ < NDOF; ++j_) out[j_] = real(0);";
        o << " FEMX_ROWC_" << a;
        if (has_q[a]) o << " FEMX_QUAD(FEMX_ROWQ_" << a << ")";
        o << " \\\n        dacc0_ += out[" << a << "];";
        for (int j = 0; j < nn - 1; ++j) {
          const int pos = (int)((code >> (7 * j)) & 127);
          o << " a" << pos << "_ = ";
          if ((code >> (21 + j)) & 1) o << "real(0)"; else o << "a" << pos << "_";
          o << " + out[" << femx_oth(nn, a, j) << "];";
        }
        o << " } }";
      }
      for (int k = 0; k < sc->rlen; ++k)
        if (last[k] == it && k != sc->self) o << " if (mine) srow[" << k << "] = a" << k << "_;";
    }
    if (f->rowsum) {  // the diagonal from the row sum: the off-diagonal values, read back in ascending position order
      o << " \\\n    { real S_ = real(0);";
      for (int k = 0; k < sc->rlen; ++k)
        if (k != sc->self) o << " S_ += srow[" << k << "];";
      o << " dacc0_ = fma(FEMX_CJ, dacc0_, -S_); }";
    }
    o << " \\\n    if (mine) srow[" << sc->self << "] = dacc0_;\n";

  }
  return o.str();
}

// kernel: "coo", "coo_e", "rhs", "csr" (unit node stride), "csr_s" (strided), "csr_x" (element-expanded
// coordinates); sc != NULL: the csr kernel for a pattern with that stencil class (specialised body for
// the class rows, row-list CTAs for the others); lat/plan != NULL: the element-once lattice pass for the
// class rows of a lattice mesh.  tile = threads per CTA the kernel is compiled for (0: the form's default).
int compile_variant(femx_form* f, const std::string& kernel, Variant** outv, bool load,
                    const StencilClass* sc = nullptr, int tile = 0, const femx_lattice* lat = nullptr,
                    femx_lattice_plan* plan = nullptr, const std::string& lat_key = std::string()) {
  std::string vkey = kernel;
  if (lat) vkey += "@L" + lat_key;
  else if (sc) vkey += "@" + sc->key;
  if (tile > 0) vkey += "#" + std::to_string(tile);
  auto it = f->variants.find(vkey);
  if (it == f->variants.end()) {
    Variant v;
    const bool is_csr = kernel == "csr" || kernel == "csr_x" || kernel == "csr_s";
    std::string body;
    if (kernel == "coo") body = kFemxJitCoo;
    else if (kernel == "coo_e") body = kFemxJitCooElem;
    else if (kernel == "rhs") body = kFemxJitRhs;
    else if (is_csr) body = std::string(kFemxJitCsrRow) + (lat ? kFemxJitLattice : kFemxJitCsr);
    else return femx_fail(f->ctx, FEMX_ERR_INVALID, "unknown kernel variant '%s'", kernel.c_str());
    if ((sc || lat) && kernel != "csr" && kernel != "csr_s")
      return femx_fail(f->ctx, FEMX_ERR_INVALID, "kernel '%s' has no specialised form", kernel.c_str());
    std::string defs = build_defines(f, kernel, lat ? nullptr : sc, lat ? plan->threads : tile);
    if (lat) {
      const std::string ld = femx_lattice_defines(f, *lat, plan);
      if (ld.empty())
        return femx_fail(f->ctx, FEMX_ERR_UNSUPPORTED, "lattice pass: %s", plan->fallback.c_str());
      defs += ld;
      v.lt_smem = plan->smem;
      v.lt_nslot = plan->nslot;
    }
    v.source = "// femx JIT kernel '" + vkey + "' (generated)\n" + defs + kFemxJitCommon + body;
    nvrtcProgram prog;
    std::string fname = "femx_" + kernel + (lat ? "_lattice" : (sc ? "_spec" : "")) + ".cu";
    if (!f->knobs.jit_dump.empty()) {  // keep the source on disk (ncu --import-source)
      fname = f->knobs.jit_dump + "/" + fname;
        if (FILE* fp = fopen(fname.c_str(), "w")) { fwrite(v.source.data(), 1, v.source.size(), fp); fclose(fp); }
    }
    nvrtcResult r = nvrtcCreateProgram(&prog, v.source.c_str(), fname.c_str(), 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS)
      return femx_fail(f->ctx, FEMX_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--diag-suppress=550", "--diag-suppress=177",
                                     f->fmad ? "--fmad=true" : "--fmad=false"};
    std::string regopt;
    if (lat && plan->regs > 0) {
      regopt = "--maxrregcount=" + std::to_string(plan->regs);
      opts.push_back(regopt.c_str());
    }
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    nvrtcGetProgramLogSize(prog, &ls);
    v.log.resize(ls);
    if (ls) nvrtcGetProgramLog(prog, &v.log[0]);
    f->last_source = v.source;
    f->last_log = v.log;
    if (r != NVRTC_SUCCESS) {
      nvrtcDestroyProgram(&prog);
      f->err = "NVRTC compilation of kernel '" + vkey + "' failed:\n" + v.log;
      return femx_fail(f->ctx, FEMX_ERR_NVRTC, "%s", f->err.c_str());
    }
    size_t cs = 0;
    r = nvrtcGetCUBINSize(prog, &cs);
    if (r != NVRTC_SUCCESS || cs == 0) {
      nvrtcDestroyProgram(&prog);
      return femx_fail(f->ctx, FEMX_ERR_NVRTC, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    v.cubin.resize(cs);
    nvrtcGetCUBIN(prog, v.cubin.data());
    nvrtcDestroyProgram(&prog);
    it = f->variants.emplace(vkey, std::move(v)).first;
  }
  Variant& v = it->second;
  if (lat && plan) { plan->smem = v.lt_smem; plan->nslot = v.lt_nslot; }  // (cache hit: the generator did not run)
  if (load && !v.fn) {
    if (!f->ctx)
      return femx_fail(nullptr, FEMX_ERR_CUDA,
                       "form was compiled offline (no device context); cannot launch");
    const femx_driver* drv = femx_get_driver(nullptr);
    if (!drv) return femx_fail(f->ctx, FEMX_ERR_CUDA, "CUDA driver entry points unavailable");
    cudaSetDevice(f->ctx->device);
    CUresult cr = drv->ModuleLoadData(&v.module, v.cubin.data());
    const char* es = nullptr;
    if (cr != CUDA_SUCCESS) {
      drv->GetErrorString(cr, &es);
      return femx_fail(f->ctx, FEMX_ERR_CUDA, "cuModuleLoadData: %s", es ? es : "?");
    }
    std::string entry = kernel.compare(0, 3, "coo") == 0 ? "femx_coo" : (kernel == "rhs" ? "femx_rhs" : "femx_csr");
    cr = drv->ModuleGetFunction(&v.fn, v.module, entry.c_str());
    if (cr != CUDA_SUCCESS) {
      drv->GetErrorString(cr, &es);
      return femx_fail(f->ctx, FEMX_ERR_CUDA, "cuModuleGetFunction(%s): %s", entry.c_str(),
                       es ? es : "?");
    }
    if (lat && drv->ModuleGetFunction(&v.fn2, v.module, "femx_rowlist") != CUDA_SUCCESS)
      return femx_fail(f->ctx, FEMX_ERR_CUDA, "cuModuleGetFunction(femx_rowlist) failed");
  }
  if (outv) *outv = &v;
  return FEMX_OK;
}

int make_form(femx_ctx* ctx, const femx_form_desc* d, femx_form** out) {
  if (!d || !out) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: NULL argument");
  *out = nullptr;
  if (!((d->dim == 2 && d->nn == 3) || (d->dim == 3 && d->nn == 4)))
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED,
                     "femx_form_compile: only P1 simplices (dim=2,nn=3 / dim=3,nn=4), got dim=%d nn=%d",
                     d->dim, d->nn);
  if (d->nd < 1 || d->nd > 3)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: nd=%d out of range", d->nd);
  if (d->dtype != FEMX_F64 && d->dtype != FEMX_F32)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: bad dtype %d", d->dtype);
  femx_form* f = new femx_form();
  f->ctx = ctx;
  f->knobs = ctx ? ctx->knobs : femx_knobs_from_env();
  f->dim = d->dim; f->nn = d->nn; f->nd = d->nd; f->dtype = d->dtype;
  f->builtin = d->builtin; f->fmad = d->fmad ? 1 : 0;
  f->n = d->nn * d->nd;
  f->integrated = d->integrated ? 1 : 0;
  if (d->nq > 0) {
    if (!d->qw || !d->qr || !d->qs || (d->dim == 3 && !d->qt)) {
      delete f;
      return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: incomplete quadrature rule");
    }
    f->nq = d->nq;
    f->qw.assign(d->qw, d->qw + d->nq);
    f->qr.assign(d->qr, d->qr + d->nq);
    f->qs.assign(d->qs, d->qs + d->nq);
    f->qt.resize(d->nq);
    f->qu.assign(d->nq, 0.0);
    for (int q = 0; q < d->nq; ++q) {
      if (d->dim == 2) {
        f->qt[q] = d->qt ? d->qt[q] : 1.0 - d->qr[q] - d->qs[q];
      } else {
        f->qt[q] = d->qt[q];
        f->qu[q] = d->qu ? d->qu[q] : 1.0 - d->qr[q] - d->qs[q] - d->qt[q];
      }
    }
  } else {
    default_rule(f);
  }
  if (d->builtin == FEMX_FORM_CUSTOM) {
    if (!d->entries) {
      delete f;
      return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: custom form without entries");
    }
    for (int k = 0; k < f->n * f->n; ++k) {
      if (!d->entries[k]) {
        delete f;
        return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: entries[%d] is NULL", k);
      }
      std::string s = d->entries[k];
      while (!s.empty() && (s.back() == '\n' || s.back() == ';' || s.back() == ' ')) s.pop_back();
      f->entries.push_back(s);
    }
    if (d->prologue) f->prologue = d->prologue;
  } else {
    if (d->builtin != FEMX_FORM_ELASTICITY && d->nd != 1) {
      delete f;
      return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: scalar form needs nd=1");
    }
    int st = emit_builtin(f, d);
    if (st != FEMX_OK) {
      delete f;
      return femx_fail(ctx, st, "femx_form_compile: bad built-in form %d (nd=%d, dim=%d)",
                       d->builtin, d->nd, d->dim);
    }
  }
  if (d->rhs_entries) {  // explicit load-vector integrands (reference semantics: weighted and summed)
    f->rhs.clear();
    f->rhs_integrated = f->builtin == FEMX_FORM_CUSTOM ? f->integrated : 0;
    for (int k = 0; k < f->n; ++k) {
      if (!d->rhs_entries[k]) {
        delete f;
        return femx_fail(ctx, FEMX_ERR_INVALID, "femx_form_compile: rhs_entries[%d] is NULL", k);
      }
      std::string s = d->rhs_entries[k];
      while (!s.empty() && (s.back() == '\n' || s.back() == ';' || s.back() == ' ')) s.pop_back();
      f->rhs.push_back(s);
    }
  }
  int st = compile_variant(f, f->nd == 1 ? "coo_e" : "coo", nullptr, ctx != nullptr);
  // user strings are pasted into more than one kernel scope: compile the numeric-pass variant now as well, so that a
  // name that collides with a kernel local fails HERE with the NVRTC log, not at the first femx_assemble_csr
  if (st == FEMX_OK && f->builtin == FEMX_FORM_CUSTOM) st = compile_variant(f, "csr", nullptr, false);
  if (st != FEMX_OK) {
    if (ctx) ctx->err = f->err.empty() ? ctx->err : f->err;
    delete f;
    return st;
  }
  *out = f;
  return FEMX_OK;
}

int check_mesh(const femx_form* f, const femx_mesh_view* m, bool* expanded) {
  if (!m) return femx_fail(f->ctx, FEMX_ERR_INVALID, "mesh view is NULL");
  if (m->dim != f->dim || m->nn != f->nn)
    return femx_fail(f->ctx, FEMX_ERR_INVALID, "mesh (dim=%d,nn=%d) does not match form (dim=%d,nn=%d)",
                     m->dim, m->nn, f->dim, f->nn);
  bool ex = m->d_elem_xyz[0] != nullptr;
  const void* const* c = ex ? m->d_elem_xyz : m->d_node_xyz;
  for (int k = 0; k < f->dim; ++k)
    if (!c[k]) return femx_fail(f->ctx, FEMX_ERR_INVALID, "mesh view: coordinate array %d is NULL", k);
  if (m->n_elems < 0 || m->n_nodes < 0)
    return femx_fail(f->ctx, FEMX_ERR_INVALID, "mesh view: negative size");
  *expanded = ex;
  return FEMX_OK;
}

}  // namespace

extern "C" {

int femx_form_compile(femx_ctx* ctx, const femx_form_desc* desc, femx_form** out) {
  if (!ctx) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_form_compile: ctx is NULL");
  return make_form(ctx, desc, out);
}

int femx_form_compile_offline(const femx_form_desc* desc, femx_form** out) {
  return make_form(nullptr, desc, out);
}

void femx_form_destroy(femx_form* form) {
  if (!form) return;
  const femx_driver* drv = femx_get_driver(nullptr);
  if (form->ctx) cudaSetDevice(form->ctx->device);
  for (auto& kv : form->variants)
    if (kv.second.module && drv) drv->ModuleUnload(kv.second.module);
  delete form;
}

const char* femx_form_source(const femx_form* form) { return form ? form->last_source.c_str() : ""; }
const char* femx_form_log(const femx_form* form) { return form ? form->last_log.c_str() : ""; }
const char* femx_form_prologue(const femx_form* form) { return form ? form->prologue.c_str() : ""; }

const char* femx_form_entry(const femx_form* form, int li, int lj) {
  if (!form || li < 0 || lj < 0 || li >= form->n || lj >= form->n) return nullptr;
  return form->entries[(size_t)li * form->n + lj].c_str();
}

int femx_form_cubin(femx_form* form, const char* kernel, const void** cubin, size_t* size) {
  if (!form || !kernel) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_form_cubin: NULL argument");
  Variant* v = nullptr;
  int st = compile_variant(form, kernel, &v, false);
  if (st != FEMX_OK) return st;
  form->last_source = v->source;
  form->last_log = v->log;
  if (cubin) *cubin = v->cubin.data();
  if (size) *size = v->cubin.size();
  return FEMX_OK;
}

int femx_form_cubin_stencil(femx_form* form, int n_incid, int row_len, int self_pos, const uint32_t* h_codes,
                            const void** cubin, size_t* size) {
  if (!form || !h_codes) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_form_cubin_stencil: NULL argument");
  if (form->nd != 1) return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_form_cubin_stencil: scalar forms only");
  if (n_incid < 1 || n_incid > FEMX_SPEC_MAX_NP || row_len < form->nn || row_len > FEMX_SPEC_MAX_RLEN ||
      self_pos < 0 || self_pos >= row_len)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_form_cubin_stencil: class (%d incidences, %d columns, own %d) out of range",
                     n_incid, row_len, self_pos);
  StencilClass sc;
  sc.np = n_incid; sc.rlen = row_len; sc.self = self_pos;
  sc.codes.assign(h_codes, h_codes + n_incid);
  unsigned long long h = 1469598103934665603ull;
  for (int k = 0; k < n_incid; ++k) {
    for (int j = 0; j < form->nn - 1; ++j) {
      const int pos = (int)((h_codes[k] >> (7 * j)) & 127);
      if (pos >= row_len || pos == self_pos)
        return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_form_cubin_stencil: code %d names column %d", k, pos);
    }
    h = (h ^ h_codes[k]) * 1099511628211ull;
  }
  char key[96];
  snprintf(key, sizeof key, "user%016llx_%d_%d_%d", h, n_incid, row_len, self_pos);
  sc.key = key;
  Variant* v = nullptr;
  int st = compile_variant(form, "csr", &v, false, &sc);
  if (st != FEMX_OK) return st;
  form->last_source = v->source;
  form->last_log = v->log;
  if (cubin) *cubin = v->cubin.data();
  if (size) *size = v->cubin.size();
  return FEMX_OK;
}

int femx_form_cubin_lattice(femx_form* form, int n_per_cell, const int32_t* h_corners, int64_t stride_y,
                            int64_t stride_z, int row_len, int self_pos, const int32_t* h_offsets,
                            const void** cubin, size_t* size, int* h_info) {
  if (!form || !h_corners || !h_offsets)
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_form_cubin_lattice: NULL argument");
  if (n_per_cell < 1 || n_per_cell > 8 || row_len < 2 || row_len > 32 || self_pos < 0 || self_pos >= row_len)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_form_cubin_lattice: bad sizes");
  femx_lattice L;
  L.ok = true; L.dim = form->dim; L.P = n_per_cell;
  L.cn[0] = L.cn[1] = L.cn[2] = 64;  // nominal extents: only the tile shape depends on them
  L.s[0] = 1; L.s[1] = stride_y; L.s[2] = stride_z;
  for (int t = 0; t < n_per_cell; ++t)
    for (int a = 0; a < form->nn; ++a) L.corner[t][a] = (unsigned char)(h_corners[t * form->nn + a] & 7);
  std::vector<int32_t> off(h_offsets, h_offsets + row_len);
  femx_lattice_plan plan;
  std::string why;
  if (!femx_lattice_plan_make(form, L, row_len, self_pos, off, form->knobs, &plan, &why))
    return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED, "femx_form_cubin_lattice: %s", why.c_str());
  Variant* v = nullptr;
  int st = compile_variant(form, "csr", &v, false, nullptr, 0, &L, &plan, femx_lattice_key(L, plan));
  if (st != FEMX_OK) return st;
  form->last_source = v->source;
  form->last_log = v->log;
  if (cubin) *cubin = v->cubin.data();
  if (size) *size = v->cubin.size();
  if (h_info) { h_info[0] = plan.tx; h_info[1] = plan.ty; h_info[2] = plan.threads; h_info[3] = plan.nslot; h_info[4] = (int)plan.smem; h_info[5] = plan.minb; }
  return FEMX_OK;
}

int femx_assemble_coo(femx_form* form, const femx_mesh_view* mesh, void* d_A, int32_t* d_rowA,
                      int32_t* d_colA, void* stream) {
  if (!form) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_assemble_coo: form is NULL");
  bool expanded = false;
  int st = check_mesh(form, mesh, &expanded);
  if (st != FEMX_OK) return st;
  if (mesh->n_elems == 0) return FEMX_OK;
  if (form->ctx) FEMX_CUDA_OK(form->ctx, cudaSetDevice(form->ctx->device));
  if (((uintptr_t)d_A | (uintptr_t)d_rowA | (uintptr_t)d_colA) % 16)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_coo: d_A, d_rowA and d_colA must be 16-byte aligned");
  if (!mesh->d_conn && (!expanded || d_rowA || d_colA))
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_coo: connectivity is NULL");
  // scalar forms: one thread per element, TMA bulk stores; vector forms: one thread per (element, row)
  const bool per_elem = form->nd == 1;
  Variant* v = nullptr;
  st = compile_variant(form, per_elem ? "coo_e" : "coo", &v, true);
  if (st != FEMX_OK) return st;
  const femx_driver* drv = femx_get_driver(nullptr);
  const void* const* c = expanded ? mesh->d_elem_xyz : mesh->d_node_xyz;
  const void* X = c[0]; const void* Y = c[1]; const void* Z = c[2];
  long long cs = mesh->node_stride ? mesh->node_stride : 1;
  int ex = expanded ? 1 : 0;
  long long ne = mesh->n_elems;
  const int32_t* conn = mesh->d_conn;
  void* args[] = {&conn, &X, &Y, &Z, &cs, &ex, &d_A, &d_rowA, &d_colA, &ne};
  CUresult cr;
  if (per_elem) {
    const size_t rs = form->dtype == FEMX_F32 ? 4 : 8;
    const unsigned smem = (unsigned)(128 * form->n * form->n * (rs + 8));
    if ((int)smem > v->smem_set && smem > 48 * 1024) {
      if (drv->FuncSetAttribute(v->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS)
        return femx_fail(form->ctx, FEMX_ERR_CUDA, "cuFuncSetAttribute(smem=%u) failed", smem);
      v->smem_set = (int)smem;
    }
    if (!v->carveout_set) {
      drv->FuncSetAttribute(v->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
      v->carveout_set = 1;
    }
    long long blocks = (ne + 127) / 128;
    if (blocks > 2147483647LL)
      return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED, "femx_assemble_coo: %lld blocks exceed grid limit", blocks);
    cr = drv->LaunchKernel(v->fn, (unsigned)blocks, 1, 1, 128, 1, 1, smem, (CUstream)stream, args, nullptr);
  } else {
    long long threads = ne * form->n;
    long long blocks = (threads + 255) / 256;
    if (blocks > 2147483647LL)
      return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED, "femx_assemble_coo: %lld blocks exceed grid limit", blocks);
    cr = drv->LaunchKernel(v->fn, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr);
  }
  if (cr != CUDA_SUCCESS) {
    const char* es = nullptr;
    drv->GetErrorString(cr, &es);
    return femx_fail(form->ctx, FEMX_ERR_CUDA, "femx_assemble_coo: launch failed: %s", es ? es : "?");
  }
  return FEMX_OK;
}

int femx_assemble_rhs(femx_form* form, const femx_pattern* pat, const femx_mesh_view* mesh, void* d_rhs,
                      void* stream) {
  if (!form || !pat) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_assemble_rhs: NULL argument");
  if (form->rhs.empty())
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_rhs: the form has no load-vector integrands");
  bool expanded = false;
  int st = check_mesh(form, mesh, &expanded);
  if (st != FEMX_OK) return st;
  if (pat->nn != form->nn || pat->nd != form->nd)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_rhs: pattern does not match the form");
  if (mesh->n_nodes != pat->n_nodes || mesh->n_elems != pat->n_elems)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_rhs: mesh sizes differ from the pattern's");
  if (!form->ctx || !pat->ctx || form->ctx->device != pat->ctx->device)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_rhs: form and pattern belong to different devices");
  if (pat->n_rows == 0) return FEMX_OK;
  if (!d_rhs) return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_rhs: d_rhs is NULL");
  FEMX_CUDA_OK(form->ctx, cudaSetDevice(form->ctx->device));
  st = femx_pattern_complete_map(pat, stream);   // the load vector walks every row's scatter map
  if (st != FEMX_OK) return st;
  Variant* v = nullptr;
  st = compile_variant(form, "rhs", &v, true);
  if (st != FEMX_OK) return st;
  const femx_driver* drv = femx_get_driver(nullptr);
  const void* const* c = expanded ? mesh->d_elem_xyz : mesh->d_node_xyz;
  const void* X = c[0]; const void* Y = c[1]; const void* Z = c[2];
  long long cs = mesh->node_stride ? mesh->node_stride : 1;
  int ex = expanded ? 1 : 0;
  int n_rows = (int)pat->n_rows;
  const int2* rowinfo = pat->d_rowinfo;
  const int32_t* slice_ptr = pat->d_slice_ptr;
  const int32_t* col = pat->d_col_idx;
  const uint32_t* code = pat->d_sell_code;
  const int32_t* pelem = pat->d_sell_elem;
  void* args[] = {&rowinfo, &slice_ptr, &col, &code, &pelem, &X, &Y, &Z, &cs, &ex, &d_rhs, &n_rows};
  unsigned blocks = (unsigned)((pat->n_rows + 127) / 128);
  CUresult cr = drv->LaunchKernel(v->fn, blocks, 1, 1, 128, 1, 1, 0, (CUstream)stream, args, nullptr);
  if (cr != CUDA_SUCCESS) {
    const char* es = nullptr;
    drv->GetErrorString(cr, &es);
    return femx_fail(form->ctx, FEMX_ERR_CUDA, "femx_assemble_rhs: launch failed: %s", es ? es : "?");
  }
  return FEMX_OK;
}

int femx_assemble_csr(femx_form* form, const femx_pattern* pat, const femx_mesh_view* mesh,
                      void* d_values, void* stream) {
  if (!form || !pat) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_assemble_csr: NULL argument");
  bool expanded = false;
  int st = check_mesh(form, mesh, &expanded);
  if (st != FEMX_OK) return st;
  if (pat->nn != form->nn || pat->nd != form->nd)
    return femx_fail(form->ctx, FEMX_ERR_INVALID,
                     "femx_assemble_csr: pattern (nn=%d,nd=%d) does not match form (nn=%d,nd=%d)",
                     pat->nn, pat->nd, form->nn, form->nd);
  if (mesh->n_nodes != pat->n_nodes || mesh->n_elems != pat->n_elems)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_csr: mesh sizes differ from the pattern's");
  if (!form->ctx || !pat->ctx || form->ctx->device != pat->ctx->device)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_csr: form and pattern belong to different devices");
  if (!d_values && pat->nnz_node > 0)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_csr: d_values is NULL");
  if (pat->n_rows == 0 || pat->nnz_node == 0) return FEMX_OK;
  long long cs = mesh->node_stride ? mesh->node_stride : 1;
  FEMX_CUDA_OK(form->ctx, cudaSetDevice(form->ctx->device));
  // the tile / run / element-tile stores are 16-byte bulk copies whose phase is derived from the value INDEX
  if ((uintptr_t)d_values % 16)
    return femx_fail(form->ctx, FEMX_ERR_INVALID, "femx_assemble_csr: d_values must be 16-byte aligned");
  Variant* v = nullptr;
  const femx_knobs& live = form->ctx->knobs;  // spec / lattice / carveout can be switched per call (femx_ctx_set_option)
  // The pattern's dominant stencil class gets a straight-line body when it covers most rows.  Only
  // where both paths are certain to round identically: built-in forms (every fma spelled out) or
  // strings compiled with --fmad=false.
  StencilClass sc;
  const size_t rs = form->dtype == FEMX_F32 ? 4 : 8;
  // (the row-list CTAs of that kernel keep one private value segment per thread in shared memory:
  //  patterns whose rows outside the class are very long stay on the generic kernel)
  const bool cls = pat->spec_np > 0 && form->nd == 1 && !expanded && (form->builtin != FEMX_FORM_CUSTOM || !form->fmad) &&
                   live.spec != 0;
  bool spec = cls && pat->spec_rows * 2 >= pat->n_rows && (size_t)pat->tile_nodes * pat->max_row_other * rs <= 64 * 1024 &&
              form->spec_failed.count(pat->spec_key) == 0;
  const char* kname = expanded ? "csr_x" : (cs == 1 ? "csr" : "csr_s");
  // Lattice meshes (3-D, symmetric built-in forms): the element-once pass takes the class rows.
  femx_lattice_plan plan;
  bool lattice = false;
  std::string lat_key;
  if (cls && pat->lat.ok && pat->lat_rows > 0 && pat->lat_rows == pat->spec_rows && form->lt_ok && live.lattice != 0) {
    // (plan and kernel are looked up per (pattern, lattice options): the launch path builds no strings / maps after the first call)
    // (the key also names the lattice: a destroyed pattern's address may be reused by a different one)
    char optkey[224];
    snprintf(optkey, sizeof optkey, "%d,%d,%d,%d,%d,%d,%d|%lld|%d,%d,%d,%d,%lld,%lld,%lld,%d,%d,%s", live.lt_tx, live.lt_ty, live.lt_kc,
             live.lt_minb, live.lt_regs, live.lt_pf, live.lt_unroll, (long long)pat->spec_rows, pat->lat.P, pat->lat.cn[0], pat->lat.cn[1],
             pat->lat.cn[2], pat->lat.s[1], pat->lat.s[2], pat->lat.node0, pat->spec_rlen, pat->spec_self, pat->spec_key.c_str());
    femx_lattice_cached& cc = form->lt_cache[pat];
    if (cc.variant && cc.opts == optkey) {
      plan = cc.plan;
      v = cc.variant;
      lattice = true;
    } else {
      std::string why;
      femx_knobs kk = form->knobs;
      kk.lt_tx = live.lt_tx; kk.lt_ty = live.lt_ty; kk.lt_kc = live.lt_kc; kk.lt_minb = live.lt_minb;
      kk.lt_regs = live.lt_regs; kk.lt_pf = live.lt_pf; kk.lt_unroll = live.lt_unroll;
      if (femx_lattice_plan_make(form, pat->lat, pat->spec_rlen, pat->spec_self, pat->spec_off, kk, &plan, &why)) {
        lat_key = femx_lattice_key(pat->lat, plan);
        if (!form->lt_failed.count(lat_key)) {
          st = compile_variant(form, kname, &v, true, nullptr, 0, &pat->lat, &plan, lat_key);
          if (st == FEMX_ERR_NVRTC || st == FEMX_ERR_UNSUPPORTED) form->lt_failed.insert(lat_key);  // stencil-class kernel instead
          else if (st != FEMX_OK) return st;
          else {
            lattice = true;
            cc.opts = optkey; cc.plan = plan; cc.variant = v;
          }
        }
      }
    }
  }
  const femx_driver* drv = femx_get_driver(nullptr);
  auto prepare = [&](Variant* vv, size_t bytes, int threads) -> int {
    if (!vv->carveout_set || (int)bytes > vv->smem_set) {
      // Shared-memory carve-out: just enough for the CTAs that registers/threads allow, the rest
      // stays L1 for the coordinate gathers.  (knob carveout = percent overrides: experiments.)
      int pct = 0;
      if (live.carveout >= 0) {
        pct = live.carveout;
      } else {
        int regs = 64;
        drv->FuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, vv->fn);
        const int regs_alloc = ((regs + 7) / 8) * 8;
        int ctas = 65536 / (regs_alloc * threads);
        if (ctas > 2048 / threads) ctas = 2048 / threads;
        if (ctas > 32) ctas = 32;
        if (ctas < 1) ctas = 1;
        pct = (int)((ctas * (bytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
      }
      drv->FuncSetAttribute(vv->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct);
      vv->carveout_set = 1;
    }
    if ((int)bytes > vv->smem_set && bytes > 48 * 1024) {
      CUresult cr0 = drv->FuncSetAttribute(vv->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)bytes);
      if (cr0 != CUDA_SUCCESS) return femx_fail(form->ctx, FEMX_ERR_CUDA, "cuFuncSetAttribute(smem=%zu) failed", bytes);
      vv->smem_set = (int)bytes;
    }
    return FEMX_OK;
  };
  const void* const* c = expanded ? mesh->d_elem_xyz : mesh->d_node_xyz;
  const void* X = c[0]; const void* Y = c[1]; const void* Z = c[2];
  int n_rows = (int)pat->n_rows;
  const int2* rowinfo = pat->d_rowinfo;
  const int32_t* col = pat->d_col_idx;
  const int32_t* slice_ptr = pat->d_slice_ptr;
  const uint32_t* code = pat->d_sell_code;
  const int32_t* pelem = pat->d_sell_elem;
  int row_node0 = (int)pat->row_begin, node_max = (int)pat->n_nodes - 1;
  const int32_t* rowlist = pat->d_other_rows;
  const int seg = pat->max_row_other * form->nd * form->nd;
  int seg_arg = seg;
  CUresult cr;
  if (lattice) {
    const femx_lattice& L = pat->lat;
    struct { int cnx, cny, cnz, sy, sz, node0, klo, khi, ntx, nty, kc; } lat;
    lat.cnx = L.cn[0]; lat.cny = L.cn[1]; lat.cnz = L.cn[2];
    lat.sy = (int)L.s[1]; lat.sz = (int)L.s[2]; lat.node0 = (int)L.node0;
    // node planes that hold owned class rows
    long long klo = 1, khi = L.cn[2] - 1;
    if (pat->row_begin > L.node0) klo = std::max<long long>(klo, (pat->row_begin - L.node0) / L.s[2]);
    if (pat->row_end - 1 >= L.node0) khi = std::min<long long>(khi, (pat->row_end - 1 - L.node0) / L.s[2]);
    else khi = 0;
    lat.klo = (int)klo; lat.khi = (int)khi;
    lat.ntx = (L.cn[0] - 1 + plan.tx - 2) / (plan.tx - 1);
    lat.nty = (L.cn[1] - 1 + plan.ty - 2) / (plan.ty - 1);
    // node planes per CTA: every CTA pays one extra cell layer to prime its carry, and the grid should fill whole
    // waves of (CTAs per SM) x SMs — the chunk count that minimises waves x (planes + 1) (a knob overrides)
    const long long nk = khi >= klo ? khi - klo + 1 : 0;
    if (live.lt_kc <= 0 && nk > 0) {
      const long long tiles = (long long)lat.ntx * lat.nty, slots = std::max(1, form->ctx->sm_count * plan.minb);
      long long best_cost = -1;
      for (long long nz = 1; nz <= std::max<long long>(1, nk / 4); ++nz) {
        const long long kc = (nk + nz - 1) / nz, waves = (tiles * ((nk + kc - 1) / kc) + slots - 1) / slots;
        const long long cost = waves * (kc + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; plan.kc = (int)kc; }
      }
    }
    lat.kc = plan.kc;
    const long long ntz = nk > 0 ? (nk + plan.kc - 1) / plan.kc : 0;
    int n_list = (int)pat->n_other;
    const size_t smem = plan.smem;
    if (smem > form->ctx->smem_optin)
      return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED, "femx_assemble_csr: the lattice pass needs %zu B of shared memory (> %zu)",
                       smem, form->ctx->smem_optin);
    st = prepare(v, smem, plan.threads);
    if (st != FEMX_OK) return st;
    void* args[] = {&rowinfo, &slice_ptr, &col, &code, &pelem, &X, &Y, &Z, &cs, &d_values, &n_rows, &row_node0, &lat};
    const long long blocks = (long long)lat.ntx * lat.nty * ntz;
    if (blocks > 2147483647LL) return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED, "femx_assemble_csr: %lld blocks exceed the grid limit", blocks);
    // The rows outside the class have a (small-register) kernel of their own.  It runs on the context's side stream,
    // forked from and joined to the caller's stream by events: beside the lattice pass, not behind it.
    femx_ctx* cx = form->ctx;
    bool side = n_list > 0 && blocks > 0 && live.lt_side != 0;
    if (side && !cx->s_side) {
      if (cudaStreamCreateWithFlags(&cx->s_side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&cx->e_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&cx->e_join, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        side = false;
      }
    }
    cr = CUDA_SUCCESS;
    auto launch_rows = [&](CUstream s) -> CUresult {
      const size_t smem2 = (size_t)128 * seg * rs;
      if (smem2 > 48 * 1024 && (int)smem2 > v->smem2_set) {
        if (drv->FuncSetAttribute(v->fn2, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem2) != CUDA_SUCCESS) return CUDA_ERROR_INVALID_VALUE;
        v->smem2_set = (int)smem2;
      }
      void* args2[] = {&rowinfo, &slice_ptr, &col, &code, &pelem, &X, &Y, &Z, &cs, &d_values, &rowlist, &n_list, &seg_arg};
      return drv->LaunchKernel(v->fn2, (unsigned)((n_list + 127) / 128), 1, 1, 128, 1, 1, (unsigned)smem2, s, args2, nullptr);
    };
    if (side) {
      FEMX_CUDA_OK(cx, cudaEventRecord(cx->e_fork, (cudaStream_t)stream));
      FEMX_CUDA_OK(cx, cudaStreamWaitEvent(cx->s_side, cx->e_fork, 0));
      cr = launch_rows((CUstream)cx->s_side);
      FEMX_CUDA_OK(cx, cudaEventRecord(cx->e_join, cx->s_side));
    }
    if (cr == CUDA_SUCCESS && blocks > 0)
      cr = drv->LaunchKernel(v->fn, (unsigned)blocks, 1, 1, (unsigned)plan.threads, 1, 1, (unsigned)smem, (CUstream)stream, args, nullptr);
    if (side) FEMX_CUDA_OK(cx, cudaStreamWaitEvent((cudaStream_t)stream, cx->e_join, 0));
    else if (cr == CUDA_SUCCESS && n_list > 0) cr = launch_rows((CUstream)stream);
  } else {
    if (spec) {
      sc.np = pat->spec_np; sc.rlen = pat->spec_rlen; sc.self = pat->spec_self;
      sc.codes = pat->spec_codes; sc.key = pat->spec_key;
      st = compile_variant(form, kname, &v, true, &sc, pat->tile_nodes);
      if (st == FEMX_ERR_NVRTC) {
        // a class body the compiler rejects must not take the operator down: remember it, use the generic kernel
        // (femx_form_log keeps the compiler's message)
        form->spec_failed.insert(pat->spec_key);
        spec = false;
      } else if (st != FEMX_OK) {
        return st;
      }
    }
    if (!spec) {
      st = femx_pattern_complete_map(pat, stream);   // the generic pass reads every row's scatter map
      if (st != FEMX_OK) return st;
      st = compile_variant(form, kname, &v, true, nullptr, pat->tile_nodes);
      if (st != FEMX_OK) return st;
    }
    // generic kernel: [mbarrier 128 B | codes | values (+16 B phase pad) | columns (+32 B phase pad)];
    // stencil-class kernel: [warp masks 128 B | row ends | values (+16 B phase pad)], or one value segment per thread
    // in the row-list CTAs
    const size_t img = (((size_t)pat->max_tile_nnz * form->nd * form->nd * rs + 15) / 16) * 16 + 32;
    const size_t spec_hdr = 128 + (((size_t)pat->tile_nodes * 4 + 127) / 128) * 128;  // FEMX_SPEC_HDR
    size_t smem = spec ? std::max(spec_hdr + img, (size_t)pat->tile_nodes * seg * rs)
                       : 128 + (size_t)pat->max_tile_codes * 4 + img + (size_t)pat->max_tile_nnz * 4 + 32;
    if (smem > form->ctx->smem_optin)
      return femx_fail(form->ctx, FEMX_ERR_UNSUPPORTED,
                       "femx_assemble_csr: a %d-row tile needs %zu B of shared memory (> %zu)",
                       pat->tile_nodes, smem, form->ctx->smem_optin);
    st = prepare(v, smem, pat->tile_nodes);
    if (st != FEMX_OK) return st;
    struct { int v[24]; } soff = {};
    if (spec)
      for (int k = 0; k < pat->spec_rlen; ++k) soff.v[k] = pat->spec_off[k];
    int n_list = spec ? (int)pat->n_other : 0;
    void* args[] = {&rowinfo, &slice_ptr, &col, &code, &pelem, &X, &Y, &Z, &cs, &d_values, &n_rows,
                    &row_node0, &node_max, &soff, &rowlist, &n_list, &seg_arg};
    unsigned threads = (unsigned)pat->tile_nodes;  // one thread per node row
    unsigned blocks = (unsigned)((pat->n_rows + pat->tile_nodes - 1) / pat->tile_nodes) + (n_list + threads - 1) / threads;
    cr = drv->LaunchKernel(v->fn, blocks, 1, 1, threads, 1, 1, (unsigned)smem, (CUstream)stream, args, nullptr);
  }
  if (cr != CUDA_SUCCESS) {
    const char* es = nullptr;
    drv->GetErrorString(cr, &es);
    return femx_fail(form->ctx, FEMX_ERR_CUDA, "femx_assemble_csr: launch failed: %s", es ? es : "?");
  }
  return FEMX_OK;
}

}  // extern "C"
