// Host-side mesh / matrix I/O.  The reference only has printMesh()
// (fea_test.cu:53-67) and stdout dumps of the first matrix rows; these two formats make the
// engine usable with real meshes and let external solvers cross-check the assembled operator.
//   femx_io_read_gmsh             Gmsh MSH 2.2 ASCII: 3-node triangles (type 2) or 4-node tets (type 4)
//   femx_io_write_matrix_market   coordinate real general, 1-based, from a host CSR
// Pure host code: no CUDA calls, usable without a device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "femx_internal.h"

extern "C" {

int femx_io_read_gmsh(const char* path, int* dim_out, int64_t* n_nodes_out, int64_t* n_elems_out,
                      double** h_x, double** h_y, double** h_z, int32_t** h_conn) {
  if (!path || !dim_out || !n_nodes_out || !n_elems_out || !h_x || !h_y || !h_z || !h_conn)
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_io_read_gmsh: NULL argument");
  FILE* fp = fopen(path, "r");
  if (!fp) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_io_read_gmsh: cannot open %s", path);
  char line[512];
  std::vector<double> X, Y, Z;
  std::map<long long, int> id2idx;  // Gmsh node tags need not be contiguous
  std::vector<int32_t> tri, tet;
  bool ok_format = false;
  int status = FEMX_OK;
  std::string why;
  while (fgets(line, sizeof line, fp)) {
    if (!strncmp(line, "$MeshFormat", 11)) {
      double ver = 0; int type = -1, dsize = 0;
      if (!fgets(line, sizeof line, fp) || sscanf(line, "%lf %d %d", &ver, &type, &dsize) != 3 || type != 0 || ver < 2.0 || ver >= 3.0) {
        status = FEMX_ERR_UNSUPPORTED; why = "only MSH 2.x ASCII is supported"; break;
      }
      ok_format = true;
    } else if (!strncmp(line, "$Nodes", 6)) {
      long long n = 0;
      if (!fgets(line, sizeof line, fp) || sscanf(line, "%lld", &n) != 1 || n < 0) { status = FEMX_ERR_INVALID; why = "bad $Nodes header"; break; }
      X.reserve(n); Y.reserve(n); Z.reserve(n);
      for (long long i = 0; i < n; ++i) {
        long long tag; double x, y, z;
        if (!fgets(line, sizeof line, fp) || sscanf(line, "%lld %lf %lf %lf", &tag, &x, &y, &z) != 4) { status = FEMX_ERR_INVALID; why = "bad node line"; break; }
        id2idx[tag] = (int)X.size();
        X.push_back(x); Y.push_back(y); Z.push_back(z);
      }
      if (status) break;
    } else if (!strncmp(line, "$Elements", 9)) {
      long long n = 0;
      if (!fgets(line, sizeof line, fp) || sscanf(line, "%lld", &n) != 1 || n < 0) { status = FEMX_ERR_INVALID; why = "bad $Elements header"; break; }
      for (long long i = 0; i < n; ++i) {
        if (!fgets(line, sizeof line, fp)) { status = FEMX_ERR_INVALID; why = "truncated $Elements"; break; }
        long long num; int type, ntags; int off = 0;
        if (sscanf(line, "%lld %d %d%n", &num, &type, &ntags, &off) != 3) { status = FEMX_ERR_INVALID; why = "bad element line"; break; }
        const char* p = line + off;
        for (int t = 0; t < ntags; ++t) { long long tag; int k = 0; if (sscanf(p, "%lld%n", &tag, &k) != 1) { status = FEMX_ERR_INVALID; break; } p += k; }
        if (status) { why = "bad element tags"; break; }
        const int nn = type == 2 ? 3 : (type == 4 ? 4 : 0);
        if (!nn) continue;  // points, lines, higher-order elements: ignored
        for (int a = 0; a < nn; ++a) {
          long long tag; int k = 0;
          if (sscanf(p, "%lld%n", &tag, &k) != 1) { status = FEMX_ERR_INVALID; why = "bad element nodes"; break; }
          p += k;
          auto it = id2idx.find(tag);
          if (it == id2idx.end()) { status = FEMX_ERR_INVALID; why = "element references an unknown node"; break; }
          (nn == 3 ? tri : tet).push_back(it->second);
        }
        if (status) break;
      }
      if (status) break;
    }
  }
  fclose(fp);
  if (!status && !ok_format) { status = FEMX_ERR_INVALID; why = "no $MeshFormat section"; }
  if (status) return femx_fail(nullptr, status, "femx_io_read_gmsh(%s): %s", path, why.c_str());
  // a volume mesh carries its boundary triangles too: tets win when present
  const bool use_tets = !tet.empty();
  const std::vector<int32_t>& conn = use_tets ? tet : tri;
  const int nn = use_tets ? 4 : 3;
  const size_t nnodes = X.size();
  *dim_out = use_tets ? 3 : 2;
  *n_nodes_out = (int64_t)nnodes;
  *n_elems_out = (int64_t)(conn.size() / nn);
  *h_x = (double*)malloc(sizeof(double) * (nnodes ? nnodes : 1));
  *h_y = (double*)malloc(sizeof(double) * (nnodes ? nnodes : 1));
  *h_z = (double*)malloc(sizeof(double) * (nnodes ? nnodes : 1));
  *h_conn = (int32_t*)malloc(sizeof(int32_t) * (conn.size() ? conn.size() : 1));
  if (!*h_x || !*h_y || !*h_z || !*h_conn) return femx_fail(nullptr, FEMX_ERR_NOMEM, "femx_io_read_gmsh: out of memory");
  memcpy(*h_x, X.data(), sizeof(double) * nnodes);
  memcpy(*h_y, Y.data(), sizeof(double) * nnodes);
  memcpy(*h_z, Z.data(), sizeof(double) * nnodes);
  memcpy(*h_conn, conn.data(), sizeof(int32_t) * conn.size());
  return FEMX_OK;
}

void femx_io_free(void* p) { free(p); }

int femx_io_write_matrix_market(const char* path, int64_t n_rows, int64_t n_cols, const int64_t* h_row_ptr,
                                const int32_t* h_col_idx, const double* h_values) {
  if (!path || !h_row_ptr || (n_rows > 0 && h_row_ptr[n_rows] > 0 && (!h_col_idx || !h_values)))
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_io_write_matrix_market: NULL argument");
  FILE* fp = fopen(path, "w");
  if (!fp) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_io_write_matrix_market: cannot open %s", path);
  fprintf(fp, "%%%%MatrixMarket matrix coordinate real general\n%% written by femx\n%lld %lld %lld\n", (long long)n_rows,
          (long long)n_cols, (long long)h_row_ptr[n_rows]);
  for (int64_t i = 0; i < n_rows; ++i)
    for (int64_t k = h_row_ptr[i]; k < h_row_ptr[i + 1]; ++k)
      fprintf(fp, "%lld %d %.17g\n", (long long)(i + 1), h_col_idx[k] + 1, h_values[k]);
  if (fclose(fp)) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_io_write_matrix_market: write error on %s", path);
  return FEMX_OK;
}

}  // extern "C"
