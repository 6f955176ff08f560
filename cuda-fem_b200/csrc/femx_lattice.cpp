// Code generator of the element-once "lattice" numeric pass (kernel template: femx_jit_src.h,
// kFemxJitLattice; design: DESIGN.md §3.0L).
//
// On a lattice mesh (femx_lattice: P elements per cell, every cell a translate of cell 0) a CTA owns a
// (TX-1) x (TY-1) patch of node columns and marches through the node planes.  Thread (ix, iy) evaluates
// the P elements of ONE cell per plane — each element once, not once per vertex as the owner-computes
// row loop does — and reduces them to one value per cell edge (symmetric forms: K_ab = K_ba) plus the
// Jacobian sums of the cell's corners.  A matrix entry A(p, p+o) is the sum of the cell-edge values of
// the cells around the edge (p, p+o): the part from the cell layer below is carried in registers, the
// parts of the neighbouring columns come through shared memory ("fields"), and the diagonal follows from
// the zero row sum of the stiffness part.  This file turns the lattice description into
//   - the per-cell arithmetic (shared edge vectors, shared face normals),
//   - the list of fields (deduplicated by content, so that A(p,q) and A(q,p) are the same sum), and
//   - the gather of each row position,
// for ANY cell decomposition whose cell uses all 2^dim corners (Kuhn 6-tet split: 19 cell edges, 18 fields).
#include <algorithm>
#include <cstdio>
#include <map>
#include <sstream>

#include "femx_form_internal.h"

namespace {

std::string lnum(double v) {
  char b[64];
  snprintf(b, sizeof b, "%.17g", v);
  std::string s = b;
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return "real(" + s + ")";
}

inline int cdx(int c) { return c & 1; }
inline int cdy(int c) { return (c >> 1) & 1; }
inline int cdz(int c) { return (c >> 2) & 1; }

struct Term {  // one summand of a field: the cell-edge value (or corner Jacobian sum) `id` of this layer (sz = 0) or of the layer below (sz = -1)
  int id, sz;
  bool operator<(const Term& o) const { return id != o.id ? id < o.id : sz > o.sz; }
  bool operator==(const Term& o) const { return id == o.id && sz == o.sz; }
};

struct Field {
  std::vector<Term> terms;  // sorted: the content IS the identity of a field
  bool jac = false;         // sums corner Jacobians instead of cell-edge values
  bool dbl = false;         // double-buffered by layer parity (read one layer later by the downward entries)
  int slot = -1;
};

struct Ref {  // one summand of a row entry: field f of the thread shifted by (sx, sy), this layer's or the previous one's
  int f, sx, sy;
  bool prev;
};

}  // namespace

// plan.ok == false: *why says what kept the mesh / form off the lattice pass (the caller falls back)
bool femx_lattice_plan_make(const femx_form* f, const femx_lattice& L, int rlen, int self,
                            const std::vector<int32_t>& class_off, const femx_knobs& K, femx_lattice_plan* plan,
                            std::string* why) {
  plan->ok = false;
  auto no = [&](const char* m) { if (why) *why = m; return false; };
  if (!L.ok || L.dim != 3) return no("mesh is not a 3-D lattice");
  if (!f->lt_ok) return no("form has no symmetric element-once formulation (built-in scalar forms only)");
  if ((int)class_off.size() != rlen || rlen > 32) return no("stencil class too long");
  const int nn = f->nn, P = L.P;
  // cell edges (corner pairs) and the stencil offsets they produce
  std::map<std::pair<int, int>, int> edge_id;
  for (int t = 0; t < P; ++t)
    for (int a = 0; a < nn; ++a)
      for (int b = a + 1; b < nn; ++b) {
        int ca = L.corner[t][a], cb = L.corner[t][b];
        if (ca == cb) return no("degenerate lattice element");
        if (ca > cb) std::swap(ca, cb);
        if (!edge_id.count({ca, cb})) {
          const int id = (int)edge_id.size();
          edge_id[{ca, cb}] = id;
        }
      }
  plan->edges.assign(edge_id.size(), {0, 0});
  for (auto& kv : edge_id) plan->edges[kv.second] = kv.first;
  // row position of every stencil offset o = (ox, oy, oz) in {-1,0,1}^3
  plan->rlen = rlen;
  plan->self = self;
  plan->pos.assign(27, -1);
  int found = 0;
  for (int oz = -1; oz <= 1; ++oz)
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        const long long lin = ox + oy * L.s[1] + oz * L.s[2];
        for (int k = 0; k < rlen; ++k)
          if (class_off[k] == lin) {
            plan->pos[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)] = k;
            ++found;
          }
      }
  if (found != rlen || plan->pos[13] != self) return no("class offsets are not lattice offsets");
  // every directed cell edge must land on a row position, and every row position must be reached
  std::vector<int> reached(rlen, 0);
  reached[self] = 1;
  for (auto& e : plan->edges)
    for (int dir = 0; dir < 2; ++dir) {
      const int from = dir ? e.second : e.first, to = dir ? e.first : e.second;
      const int ox = cdx(to) - cdx(from), oy = cdy(to) - cdy(from), oz = cdz(to) - cdz(from);
      const int p = plan->pos[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)];
      if (p < 0) return no("a cell edge has no column in the stencil class");
      reached[p] = 1;
    }
  for (int k = 0; k < rlen; ++k)
    if (!reached[k]) return no("a class column is not a cell edge");
  // tile shape: (TX-1) x (TY-1) owned node columns per CTA; the owned counts should divide the interior
  // node counts (cn-1) as evenly as possible, the halo fraction is 1 - (TX-1)(TY-1)/(TX TY)
  // Measured on cfg3 (profiles/r02_lattice_sweep.txt): the pass is latency-bound (3 warps per scheduler, ~160 registers),
  // so small CTAs that de-synchronise win: 128 threads (8 x 16 columns, 7 x 15 owned), three CTAs per SM, beat 256 x 2
  // and every one-CTA shape although their halo share is larger.
  int tx = K.lt_tx, ty = K.lt_ty;
  if (tx < 2 || ty < 2 || tx * ty > 1024 || 32 % tx) {
    const int budget = 128;
    double best = -1.0;
    tx = 8; ty = 16;
    for (int a = 32; a >= 4; a /= 2)   // a line of the tile lies inside one warp (runs of rows are stored warp-locally);
      for (int b = 3; b <= 32; ++b) {  // on ties the longer line wins (16 x 8 measured 4 % faster than 8 x 16: longer bulk stores)
        if (a * b > budget) continue;
        const int threads = ((a * b + 31) / 32) * 32;
        const long long tiles_x = (std::max(L.cn[0] - 1, 1) + a - 2) / (a - 1), tiles_y = (std::max(L.cn[1] - 1, 1) + b - 2) / (b - 1);
        const double eff = (double)std::max(L.cn[0] - 1, 1) * std::max(L.cn[1] - 1, 1) / ((double)tiles_x * tiles_y * threads);
        if (eff > best + 1e-9) { best = eff; tx = a; ty = b; }
      }
  }
  plan->tx = tx;
  plan->ty = ty;
  plan->threads = ((tx * ty + 31) / 32) * 32;
  plan->kc = K.lt_kc > 0 ? K.lt_kc : 64;
  plan->minb = K.lt_minb > 0 ? K.lt_minb : std::max(1, std::min(8, 384 / plan->threads));
  plan->regs = K.lt_regs;
  plan->pf = K.lt_pf;
  plan->unroll = K.lt_unroll > 0 ? K.lt_unroll : 1;
  plan->ok = true;
  return true;
}

std::string femx_lattice_key(const femx_lattice& L, const femx_lattice_plan& plan) {
  std::ostringstream k;
  k << L.dim << "." << L.P << ".";
  for (int t = 0; t < L.P; ++t)
    for (int a = 0; a <= L.dim; ++a) k << (int)L.corner[t][a];
  k << ".";
  for (int v : plan.pos) k << (v < 0 ? std::string("x") : std::to_string(v)) << ",";
  k << plan.rlen << "." << plan.self << "." << plan.tx << "x" << plan.ty << "m" << plan.minb << "r" << plan.regs << "p" << plan.pf << "u" << plan.unroll;
  return k.str();
}

// The #define block of the lattice kernel.  Names: corner c = dx | dy << 1 | dz << 2 has coordinates cx<c>, cy<c>, cz<c>;
// E<b>_<p>{x,y,z} = corner p - corner b;  N<b>_<p>_<q> = E<b>_<p> x E<b>_<q>;  Ev<a>_<b> cell-edge value;  Jc<c> corner
// Jacobian sum;  Cv / Cj the same of the layer below (carried);  Fv<f> field values.
std::string femx_lattice_defines(const femx_form* f, const femx_lattice& L, femx_lattice_plan* plan) {
  const int nn = f->nn, P = L.P;
  std::ostringstream o;
  static const char* ax[3] = {"x", "y", "z"};
  auto ename = [](int b, int p) { return "E" + std::to_string(b) + "_" + std::to_string(p); };
  auto evname = [](std::pair<int, int> e) { return "Ev" + std::to_string(e.first) + "_" + std::to_string(e.second); };
  std::map<std::pair<int, int>, int> edge_id;
  for (size_t k = 0; k < plan->edges.size(); ++k) edge_id[plan->edges[k]] = (int)k;

  // ---- per-cell arithmetic ----------------------------------------------------------------
  // base vertex of an element: its corner that most elements of the cell share (Kuhn: corner 0 for all six), so
  // that edge vectors and face normals are shared by name
  int mult[8] = {0};
  for (int t = 0; t < P; ++t)
    for (int a = 0; a < nn; ++a) ++mult[L.corner[t][a]];
  std::ostringstream edges, cell;
  std::map<std::string, bool> have;  // names already emitted
  std::vector<int> ev_started(plan->edges.size(), 0);
  int jc_started[8] = {0};
  const bool stiff = f->builtin != FEMX_FORM_MASS, mass = f->builtin != FEMX_FORM_POISSON;
  for (int t = 0; t < P; ++t) {
    int li = 0;
    for (int a = 1; a < nn; ++a)
      if (mult[L.corner[t][a]] > mult[L.corner[t][li]] ||
          (mult[L.corner[t][a]] == mult[L.corner[t][li]] && L.corner[t][a] < L.corner[t][li]))
        li = a;
    // (li, li^1, li^2, li^3) is an even permutation of the element: same signed Jacobian
    int w[4];
    for (int m = 0; m < 4; ++m) w[m] = L.corner[t][li ^ m];
    for (int m = 1; m < 4; ++m) {
      const std::string n = ename(w[0], w[m]);
      if (!have[n]) {
        have[n] = true;
        edges << " \\\n    const real";
        for (int c = 0; c < 3; ++c)
          edges << (c ? "," : "") << " " << n << ax[c] << " = c" << ax[c] << w[m] << "-c" << ax[c] << w[0];
        edges << ";";
      }
    }
    // d2 = u4 x u3, d3 = u2 x u4, d4 = u3 x u2 (u_m = edge to local vertex m; femx_form.cpp: emit_geometry)
    static const int fa[3][2] = {{3, 2}, {1, 3}, {2, 1}};
    std::string dn[3];
    int ds[3];
    for (int v = 0; v < 3; ++v) {  // face normals: cell scope, shared by the elements on either side of the face
      int p = w[fa[v][0]], q = w[fa[v][1]];
      ds[v] = 1;
      if (p > q) { std::swap(p, q); ds[v] = -1; }
      dn[v] = "N" + std::to_string(w[0]) + "_" + std::to_string(p) + "_" + std::to_string(q);
      if (!have[dn[v]]) {
        have[dn[v]] = true;
        const std::string a = ename(w[0], p), b = ename(w[0], q);
        cell << " \\\n    const real";
        for (int c = 0; c < 3; ++c) {
          const char *i1 = ax[(c + 1) % 3], *i2 = ax[(c + 2) % 3];
          cell << (c ? "," : "") << " " << dn[v] << ax[c] << " = fma(" << a << i1 << "," << b << i2 << ",-femx_mul(" << a << i2 << "," << b << i1 << "))";
        }
        cell << ";";
      }
    }
    cell << " \\\n    { /* element " << t << ": base corner " << w[0] << " */";
    // jac = u2 . d2
    {
      const std::string u2 = ename(w[0], w[1]);
      cell << " \\\n      const real jac = " << (ds[0] < 0 ? "-" : "") << "fma(" << u2 << "z," << dn[0] << "z,fma(" << u2 << "y," << dn[0]
           << "y,femx_mul(" << u2 << "x," << dn[0] << "x)));";
    }
    if (stiff) cell << " const real kq = femx_mul(" << lnum(f->lt_W) << ",femx_rcp(jac));";
    if (mass) cell << " const real mj = femx_mul(" << lnum(f->lt_moff) << ",jac);";
    // g = d2 + d3 + d4 (d1 = -g)
    if (stiff) {
      cell << " \\\n      const real";
      for (int c = 0; c < 3; ++c) {
        cell << (c ? "," : "") << " g" << ax[c] << " = ";
        for (int v = 0; v < 3; ++v) cell << (ds[v] < 0 ? "-" : (v ? "+" : "")) << dn[v] << ax[c];
      }
      cell << ";";
    }
    // the six edge values: K_ab = (d_a . d_b) kq + m jac, roles 0 (d1 = -g), 1..3 (d2..d4)
    for (int ra = 0; ra < 4; ++ra)
      for (int rb = ra + 1; rb < 4; ++rb) {
        std::pair<int, int> e = {std::min(w[ra], w[rb]), std::max(w[ra], w[rb])};
        const int id = edge_id.at(e);
        std::string val;
        if (stiff) {
          const std::string va = ra == 0 ? "g" : dn[ra - 1], vb = dn[rb - 1];
          const int s = (ra == 0 ? -1 : ds[ra - 1]) * ds[rb - 1];
          const std::string dot = "fma(" + va + "z," + vb + "z,fma(" + va + "y," + vb + "y,femx_mul(" + va + "x," + vb + "x)))";
          if (mass) val = "fma(" + std::string(s < 0 ? "-" : "") + dot + ",kq,mj)";
          else val = "femx_mul(" + std::string(s < 0 ? "-" : "") + dot + ",kq)";
        } else {
          val = "mj";
        }
        cell << " \\\n      " << evname(e) << (ev_started[id] ? " += " : " = ") << val << ";";
        ev_started[id] = 1;
      }
    for (int m = 0; m < 4; ++m) {
      cell << " Jc" << w[m] << (jc_started[w[m]] ? " += jac;" : " = jac;");
      jc_started[w[m]] = 1;
    }
    cell << " }";
  }
  o << "#define FEMX_LT_EDGES" << edges.str() << "\n";
  o << "#define FEMX_LT_CELL \\\n    real";
  for (size_t k = 0; k < plan->edges.size(); ++k) o << (k ? ", " : " ") << evname(plan->edges[k]);
  for (int c = 0; c < 8; ++c) o << ", Jc" << c << (jc_started[c] ? "" : " = real(0)");
  o << ";" << cell.str() << "\n";

  // ---- fields ---------------------------------------------------------------------------------
  // Row entry for offset o: every directed cell edge (from, to) with pos(to) - pos(from) = o contributes, out of the
  // cell whose corner `from` is the row's node, i.e. the cell shifted by s = -pos(from).  Summands of the same
  // column shift (sx, sy) are added by that column's thread (this layer + carry of the layer below) and published
  // as one field; fields are identified by CONTENT, so A(p,q) and A(q,p) read the same fields in the same order
  // and are the same bits.
  std::vector<Field> fields;
  auto field_of = [&](std::vector<Term> terms, bool jac) {
    std::sort(terms.begin(), terms.end());
    for (size_t k = 0; k < fields.size(); ++k)
      if (fields[k].jac == jac && fields[k].terms == terms) return (int)k;
    Field fl;
    fl.terms = terms;
    fl.jac = jac;
    fields.push_back(fl);
    return (int)fields.size() - 1;
  };
  // groups[o][(sx,sy)] -> terms
  std::map<int, std::map<std::pair<int, int>, std::vector<Term>>> groups;
  for (size_t k = 0; k < plan->edges.size(); ++k)
    for (int dir = 0; dir < 2; ++dir) {
      const int from = dir ? plan->edges[k].second : plan->edges[k].first, to = dir ? plan->edges[k].first : plan->edges[k].second;
      const int ox = cdx(to) - cdx(from), oy = cdy(to) - cdy(from), oz = cdz(to) - cdz(from);
      groups[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)][{-cdx(from), -cdy(from)}].push_back({(int)k, -cdz(from)});
    }
  std::vector<std::vector<Ref>> entry(plan->rlen);  // per row position
  plan->fallback.clear();
  for (auto& go : groups) {
    const int oc = go.first, oz = oc / 9 - 1, oy = (oc / 3) % 3 - 1, ox = oc % 3 - 1;
    const int p = plan->pos[oc];
    if (oz >= 0) {
      for (auto& g : go.second) entry[p].push_back({field_of(g.second, false), g.first.first, g.first.second, false});
    } else {
      // downward entry = the upward entry (-o) of the node p + o one plane below: its fields, published one layer
      // earlier (the other parity buffer), read at the column shift of that node
      const int mc = (-oz + 1) * 9 + (-oy + 1) * 3 + (-ox + 1);
      for (auto& g : groups.at(mc)) {
        const int sx = ox + g.first.first, sy = oy + g.first.second;
        if (sx < -1 || sx > 0 || sy < -1 || sy > 0) { plan->fallback = "a downward entry reaches outside the tile halo"; return ""; }
        const int fi = field_of(g.second, false);
        fields[fi].dbl = true;
        entry[p].push_back({fi, sx, sy, true});
      }
    }
  }
  // corner Jacobian sums -> the row's total (for the diagonal)
  std::vector<Ref> jref;
  {
    std::map<std::pair<int, int>, std::vector<Term>> jg;
    for (int c = 0; c < 8; ++c)
      if (jc_started[c]) jg[{-cdx(c), -cdy(c)}].push_back({c, -cdz(c)});
    for (auto& g : jg) jref.push_back({field_of(g.second, true), g.first.first, g.first.second, false});
  }
  for (auto& e : entry) std::sort(e.begin(), e.end(), [](const Ref& a, const Ref& b) { return a.f != b.f ? a.f < b.f : a.prev < b.prev; });
  std::sort(jref.begin(), jref.end(), [](const Ref& a, const Ref& b) { return a.f < b.f; });
  // shared-memory slots.  The columns of a CTA are NOT in lock step (split-phase barrier: a thread gathers plane
  // kc-1 one cell after it published it), so a field published for plane kc must survive until the slowest thread
  // has gathered plane kc: fields read only by their own plane's gather are kept 2 deep (index kc & 1), fields also
  // read by the downward entries of the next plane 3 deep (index kc % 3).  Slot = index within the kind * depth + buffer.
  int ns = 0, ndb = 0;
  for (auto& fl : fields) fl.slot = fl.dbl ? ndb++ : ns++;
  plan->nslot = 2 * ns + 3 * ndb;

  auto tname = [&](const Field& fl, const Term& tm) {
    if (fl.jac) return std::string(tm.sz ? "Cj" : "Jc") + std::to_string(tm.id);
    const auto& e = plan->edges[tm.id];
    return std::string(tm.sz ? "Cv" : "Ev") + std::to_string(e.first) + "_" + std::to_string(e.second);
  };
  // carried values: whatever a field reads from the layer below
  std::map<std::string, std::string> carry;  // carry variable -> this layer's variable
  for (auto& fl : fields)
    for (auto& tm : fl.terms)
      if (tm.sz) { Term t0 = {tm.id, 0}; carry[tname(fl, tm)] = tname(fl, t0); }
  o << "#define FEMX_LT_CARRY_DECL real";
  {
    bool first = true;
    for (auto& c : carry) { o << (first ? " " : ", ") << c.first << " = real(0)"; first = false; }
    if (first) o << " lt_unused_ = real(0)";
  }
  o << ";\n#define FEMX_LT_FIELDS";
  for (size_t k = 0; k < fields.size(); ++k) {
    o << " \\\n    const real Fv" << k << " = ";
    for (size_t q = 0; q < fields[k].terms.size(); ++q) o << (q ? " + " : "") << tname(fields[k], fields[k].terms[q]);
    o << ";";
  }
  o << " \\\n   ";
  for (auto& c : carry) o << " " << c.first << " = " << c.second << ";";
  // PS / PD: this thread's element of the 2-deep / 3-deep buffer of the plane being published
  auto slot_off = [&](const Field& fl) {
    return std::to_string(fl.slot * (fl.dbl ? 3 : 2)) + " * LT_NT";
  };
  o << "\n#define FEMX_LT_PUBLISH(PS, PD)";
  for (size_t k = 0; k < fields.size(); ++k)
    o << " \\\n    (" << (fields[k].dbl ? "PD" : "PS") << ")[" << slot_off(fields[k]) << "] = Fv" << k << ";";
  o << "\n#define FEMX_LT_PUBLISH_UP(PD)";
  for (size_t k = 0; k < fields.size(); ++k)
    if (fields[k].dbl) o << " \\\n    (PD)[" << slot_off(fields[k]) << "] = Fv" << k << ";";
  // GS / GD: the gathered plane's buffers, GP: the 3-deep buffer of the plane below it
  auto refstr = [&](const Ref& r) {
    std::ostringstream s;
    s << "(" << (r.prev ? "GP" : (fields[r.f].dbl ? "GD" : "GS")) << ")[" << slot_off(fields[r.f]);
    if (r.sx) s << " - 1";
    if (r.sy) s << " - FEMX_LT_TX";
    s << "]";
    return s.str();
  };
  o << "\n#define FEMX_LT_GATHER(GS, GD, GP)";
  for (int p = 0; p < plan->rlen; ++p) {
    if (p == plan->self) continue;
    o << " \\\n    const real v" << p << " = ";
    for (size_t q = 0; q < entry[p].size(); ++q) o << (q ? " + " : "") << refstr(entry[p][q]);
    o << "; lt_row[" << p << "] = v" << p << ";";
  }
  o << " \\\n    { const real SJ = ";
  for (size_t q = 0; q < jref.size(); ++q) o << (q ? " + " : "") << refstr(jref[q]);
  o << "; real S_ = real(0);";
  for (int p = 0; p < plan->rlen; ++p)
    if (p != plan->self) o << " S_ += v" << p << ";";
  o << " lt_row[" << plan->self << "] = fma(" << lnum(f->lt_cj) << ", SJ, -S_); }\n";
  o << "#define FEMX_LT_NS " << ns << "\n";
  o << "#define FEMX_LT_TX " << plan->tx << "\n#define FEMX_LT_TY " << plan->ty << "\n#define FEMX_LT_NSLOT " << plan->nslot
    << "\n#define FEMX_LT_RLEN " << plan->rlen << "\n#define FEMX_LT_MINB " << plan->minb << "\n#define FEMX_LT_PF " << plan->pf << "\n#define FEMX_LT_LOOP_PRAGMA _Pragma(\"unroll " << plan->unroll << "\")" << "\n#define FEMX_LATTICE 1\n";
  // bytes of dynamic shared memory: mbarrier (128 B) | fields | value image (+ alignment slack)
  const size_t rs = f->dtype == FEMX_F32 ? 4 : 8;
  plan->smem = 128 + ((size_t)plan->nslot * plan->threads + (size_t)plan->threads * plan->rlen + 4) * rs + 16;
  return o.str();
}
