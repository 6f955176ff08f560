// femx_weakform.hpp — the reference's weak-form UX in C++, GiNaC-free (header only, C++11).
//
// cuda-fem writes a weak form as two lambdas over symbolic shape functions,
//     wf.build([&](ex u, ex v) { return dot(grad(u,x,y), grad(v,x,y)); },  [&](ex v) { return f*v; });
// (FunctionSpace / WeakForm, fea_symbolic_nvrtc_sparse.cpp:226-362; main :494-503), GiNaC differentiates through
// the reference map and prints one C expression per matrix entry.  For P1 simplices no computer algebra is needed:
// a shape function is one of r, s, t, (u) and its gradient is a row of J^-1, a fixed rational expression of the vertex
// coordinates.  This header provides the same surface — FunctionSpace, WeakForm::build(lhs, rhs), grad, dot, coordinate
// symbols — over a tiny expression type that records C text, and fills the femx_form_desc that femx_form_compile
// takes: n*n entry strings, n load-vector strings and a prologue holding the shared sub-expressions (Jacobian, J^-1),
// which the reference's fully expanded strings recompute in every entry.
//
//     femx::wf::FunctionSpace fs(2);                                    // P1 triangles (3: tetrahedra)
//     femx::wf::Ex f = -2.0 * (fs.x * fs.x + fs.y * fs.y) + 36.0;       // the reference's source term
//     femx::wf::WeakForm wf(fs);
//     wf.build([&](femx::wf::Fn u, femx::wf::Fn v) { return dot(grad(u), grad(v)); },
//              [&](femx::wf::Fn v) { return f * v; });
//     femx_form_desc d = wf.desc(FEMX_F64);   femx_form_compile(ctx, &d, &form);
#ifndef FEMX_WEAKFORM_HPP
#define FEMX_WEAKFORM_HPP

#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "femx.h"

namespace femx {
namespace wf {

// an expression = C text over the names the integrand may use (x1.., y1.., z1.., r, s, t, u and the prologue's names)
struct Ex {
  std::string s;
  Ex() : s("real(0)") {}
  Ex(double v) {
    char b[64];
    snprintf(b, sizeof b, "%.17g", v);
    s = std::string("real(") + b + ")";
  }
  explicit Ex(const std::string& text) : s(text) {}
};
inline Ex operator+(const Ex& a, const Ex& b) { return Ex("(" + a.s + "+" + b.s + ")"); }
inline Ex operator-(const Ex& a, const Ex& b) { return Ex("(" + a.s + "-" + b.s + ")"); }
inline Ex operator*(const Ex& a, const Ex& b) { return Ex("(" + a.s + "*" + b.s + ")"); }
inline Ex operator/(const Ex& a, const Ex& b) { return Ex("(" + a.s + "/" + b.s + ")"); }
inline Ex operator-(const Ex& a) { return Ex("(-" + a.s + ")"); }
inline Ex operator+(double a, const Ex& b) { return Ex(a) + b; }
inline Ex operator-(double a, const Ex& b) { return Ex(a) - b; }
inline Ex operator*(double a, const Ex& b) { return Ex(a) * b; }
inline Ex operator+(const Ex& a, double b) { return a + Ex(b); }
inline Ex operator-(const Ex& a, double b) { return a - Ex(b); }
inline Ex operator*(const Ex& a, double b) { return a * Ex(b); }

struct Vec { Ex c[3]; int dim; };
inline Ex dot(const Vec& a, const Vec& b) {
  Ex r = a.c[0] * b.c[0];
  for (int k = 1; k < a.dim; ++k) r = r + a.c[k] * b.c[k];
  return r;
}

// a P1 shape function: its value at the quadrature point and its (constant) physical gradient
struct Fn {
  Ex value;
  Vec gradient;
  operator Ex() const { return value; }
};
inline Vec grad(const Fn& f) { return f.gradient; }
inline Ex operator*(const Ex& a, const Fn& f) { return a * f.value; }
inline Ex operator*(const Fn& f, const Ex& a) { return f.value * a; }
inline Ex operator*(const Fn& a, const Fn& b) { return a.value * b.value; }
inline Ex operator*(double a, const Fn& f) { return Ex(a) * f.value; }

// P1 Lagrange space on triangles (dim 2) or tetrahedra (dim 3): phi = (r, s, 1-r-s) / (r, s, t, 1-r-s-t), the affine map
// X = sum_a x_a phi_a, Jacobian and J^-1 as in FunctionSpace (fea_symbolic_nvrtc_sparse.cpp:239-289, sfR_deriv/sfS_deriv :68-101)
class FunctionSpace {
 public:
  int dim, nn;
  Ex x, y, z;            // physical coordinates at the quadrature point (for coefficient functions such as f(x,y))
  Ex jac;
  std::string prologue;  // shared sub-expressions, evaluated once per element
  std::vector<Fn> phi;
  explicit FunctionSpace(int dim_) : dim(dim_), nn(dim_ + 1), jac("jac") {
    if (dim == 2) {
      x = Ex("(x1*r+x2*s+x3*t)"); y = Ex("(y1*r+y2*s+y3*t)"); z = Ex("real(0)");
      prologue =
          "const real jac = (x1-x3)*(y2-y3)-(y1-y3)*(x2-x3);\n"     // :258
          "const real wf_ij = real(1)/jac;\n"
          "const real wf_g1x = (y2-y3)*wf_ij, wf_g1y = (x3-x2)*wf_ij;\n"   // grad r
          "const real wf_g2x = (y3-y1)*wf_ij, wf_g2y = (x1-x3)*wf_ij;\n"   // grad s
          "const real wf_g3x = -(wf_g1x+wf_g2x), wf_g3y = -(wf_g1y+wf_g2y);\n";
      const char* val[3] = {"r", "s", "t"};
      for (int a = 0; a < 3; ++a) phi.push_back(make(val[a], a));
    } else {
      x = Ex("(x1*r+x2*s+x3*t+x4*u)"); y = Ex("(y1*r+y2*s+y3*t+y4*u)"); z = Ex("(z1*r+z2*s+z3*t+z4*u)");
      prologue =
          "const real wf_ax = x1-x4, wf_bx = x2-x4, wf_cx = x3-x4;\n"
          "const real wf_ay = y1-y4, wf_by = y2-y4, wf_cy = y3-y4;\n"
          "const real wf_az = z1-z4, wf_bz = z2-z4, wf_cz = z3-z4;\n"
          "const real wf_d1x = wf_by*wf_cz-wf_cy*wf_bz, wf_d1y = wf_cx*wf_bz-wf_bx*wf_cz, wf_d1z = wf_bx*wf_cy-wf_cx*wf_by;\n"
          "const real wf_d2x = wf_cy*wf_az-wf_ay*wf_cz, wf_d2y = wf_ax*wf_cz-wf_cx*wf_az, wf_d2z = wf_cx*wf_ay-wf_ax*wf_cy;\n"
          "const real wf_d3x = wf_ay*wf_bz-wf_by*wf_az, wf_d3y = wf_bx*wf_az-wf_ax*wf_bz, wf_d3z = wf_ax*wf_by-wf_bx*wf_ay;\n"
          "const real jac = wf_ax*wf_d1x+wf_bx*wf_d2x+wf_cx*wf_d3x;\n"
          "const real wf_ij = real(1)/jac;\n"
          "const real wf_g1x = wf_d1x*wf_ij, wf_g1y = wf_d1y*wf_ij, wf_g1z = wf_d1z*wf_ij;\n"
          "const real wf_g2x = wf_d2x*wf_ij, wf_g2y = wf_d2y*wf_ij, wf_g2z = wf_d2z*wf_ij;\n"
          "const real wf_g3x = wf_d3x*wf_ij, wf_g3y = wf_d3y*wf_ij, wf_g3z = wf_d3z*wf_ij;\n"
          "const real wf_g4x = -(wf_g1x+wf_g2x+wf_g3x), wf_g4y = -(wf_g1y+wf_g2y+wf_g3y), wf_g4z = -(wf_g1z+wf_g2z+wf_g3z);\n";
      const char* val[4] = {"r", "s", "t", "u"};
      for (int a = 0; a < 4; ++a) phi.push_back(make(val[a], a));
    }
  }

 private:
  Fn make(const char* value, int a) const {
    Fn f;
    f.value = Ex(value);
    f.gradient.dim = dim;
    static const char* ax[3] = {"x", "y", "z"};
    for (int k = 0; k < dim; ++k) f.gradient.c[k] = Ex("wf_g" + std::to_string(a + 1) + ax[k]);
    return f;
  }
};

// WeakForm::build (fea_symbolic_nvrtc_sparse.cpp:307-356): entry (row li, column lj) = a(u = phi_lj, v = phi_li) * jac,
// load entry li = l(v = phi_li) * jac — the strings the reference substitutes for $integrand{li}{lj}$
class WeakForm {
 public:
  explicit WeakForm(const FunctionSpace& fs) : fs_(fs) {}
  void build(const std::function<Ex(Fn, Fn)>& lhs, const std::function<Ex(Fn)>& rhs = nullptr) {
    const int n = fs_.nn;
    entries_.assign((size_t)n * n, "");
    rhs_.clear();
    for (int li = 0; li < n; ++li) {
      for (int lj = 0; lj < n; ++lj) entries_[(size_t)li * n + lj] = (lhs(fs_.phi[lj], fs_.phi[li]) * fs_.jac).s;
      if (rhs) rhs_.push_back((rhs(fs_.phi[li]) * fs_.jac).s);
    }
  }
  const std::string& entry(int li, int lj) const { return entries_[(size_t)li * fs_.nn + lj]; }
  // descriptor for femx_form_compile; the pointers stay valid as long as this object lives and is not rebuilt
  femx_form_desc desc(int dtype = FEMX_F64, int fmad = 1) {
    femx_form_desc d = femx_form_desc();
    d.dim = fs_.dim; d.nn = fs_.nn; d.nd = 1; d.dtype = dtype; d.builtin = FEMX_FORM_CUSTOM; d.fmad = fmad;
    ptr_.clear(); rptr_.clear();
    for (auto& e : entries_) ptr_.push_back(e.c_str());
    for (auto& e : rhs_) rptr_.push_back(e.c_str());
    d.entries = ptr_.data();
    d.prologue = fs_.prologue.c_str();
    d.rhs_entries = rptr_.empty() ? nullptr : rptr_.data();
    return d;
  }

 private:
  FunctionSpace fs_;
  std::vector<std::string> entries_, rhs_;
  std::vector<const char*> ptr_, rptr_;
};

}  // namespace wf
}  // namespace femx
#endif  // FEMX_WEAKFORM_HPP
