// Context, error reporting and driver-entry-point resolution.
#include <cstdio>
#include <cstring>
#include <mutex>

#include "femx_internal.h"

static thread_local std::string g_err;

void femx_set_global_error(const std::string& s) { g_err = s; }

int femx_fail(const femx_ctx* ctx, int status, const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  if (ctx) ctx->err = buf;
  return status;
}

namespace {
struct knob_entry { const char* env; const char* name; int femx_knobs::*field; };
const knob_entry kKnobs[] = {
    {"FEMX_TILE", "tile", &femx_knobs::tile}, {"FEMX_CARVEOUT", "carveout", &femx_knobs::carveout},
    {"FEMX_MINBLOCKS", "minblocks", &femx_knobs::minblocks}, {"FEMX_MIDGATHER", "midgather", &femx_knobs::midgather},
    {"FEMX_UNROLL", "unroll", &femx_knobs::unroll}, {"FEMX_ROTINV", "rotinv", &femx_knobs::rotinv},
    {"FEMX_SPEC", "spec", &femx_knobs::spec}, {"FEMX_SPEC_AHEAD", "spec_ahead", &femx_knobs::spec_ahead},
    {"FEMX_SHAREDFACES", "sharedfaces", &femx_knobs::sharedfaces}, {"FEMX_ACCF", "accf", &femx_knobs::accf},
    {"FEMX_RCP3", "rcp3", &femx_knobs::rcp3}, {"FEMX_SPEC_PREFETCH", "spec_prefetch", &femx_knobs::spec_prefetch},
    {"FEMX_SPEC_PIN", "spec_pin", &femx_knobs::spec_pin}, {"FEMX_LISTLAST", "listlast", &femx_knobs::listlast},
    {"FEMX_ROWSUM", "rowsum", &femx_knobs::rowsum}, {"FEMX_CHAINORDER", "chainorder", &femx_knobs::chainorder},
    {"FEMX_LATTICE", "lattice", &femx_knobs::lattice}, {"FEMX_LATTICE_PATTERN", "lattice_pattern", &femx_knobs::lattice_pattern}, {"FEMX_LT_TX", "lt_tx", &femx_knobs::lt_tx},
    {"FEMX_LT_TY", "lt_ty", &femx_knobs::lt_ty}, {"FEMX_LT_KC", "lt_kc", &femx_knobs::lt_kc},
    {"FEMX_LT_MINB", "lt_minb", &femx_knobs::lt_minb}, {"FEMX_LT_REGS", "lt_regs", &femx_knobs::lt_regs}, {"FEMX_LT_PF", "lt_pf", &femx_knobs::lt_pf}, {"FEMX_LT_UNROLL", "lt_unroll", &femx_knobs::lt_unroll},
    {"FEMX_LT_SIDE", "lt_side", &femx_knobs::lt_side},
    {"FEMX_DIST_GRAPH", "dist_graph", &femx_knobs::dist_graph}, {"FEMX_DIST_P2P", "dist_p2p", &femx_knobs::dist_p2p},
    {"FEMX_DIST_PUSH", "dist_push", &femx_knobs::dist_push},
};
}  // namespace

femx_knobs femx_knobs_from_env() {
  femx_knobs k;
  for (auto& e : kKnobs)
    if (const char* v = getenv(e.env))
      if (*v) k.*(e.field) = atoi(v);
  if (const char* d = getenv("FEMX_JIT_DUMP")) k.jit_dump = d;
  return k;
}

bool femx_knobs_set(femx_knobs* k, const char* name, int value) {
  for (auto& e : kKnobs)
    if (!strcmp(name, e.name) || !strcmp(name, e.env)) { k->*(e.field) = value; return true; }
  return false;
}

std::string femx_knobs::key() const {
  std::string s;
  for (auto& e : kKnobs) s += std::to_string(this->*(e.field)) + ",";
  return s;
}

const femx_driver* femx_get_driver(std::string* why) {
  static femx_driver drv;
  static std::string err;
  static std::once_flag once;
  std::call_once(once, [] {
    struct {
      const char* name;
      void** slot;
    } syms[] = {
        {"cuModuleLoadData", (void**)&drv.ModuleLoadData},
        {"cuModuleUnload", (void**)&drv.ModuleUnload},
        {"cuModuleGetFunction", (void**)&drv.ModuleGetFunction},
        {"cuLaunchKernel", (void**)&drv.LaunchKernel},
        {"cuFuncSetAttribute", (void**)&drv.FuncSetAttribute},
        {"cuFuncGetAttribute", (void**)&drv.FuncGetAttribute},
        {"cuGetErrorString", (void**)&drv.GetErrorString},
    };
    for (auto& s : syms) {
      cudaDriverEntryPointQueryResult q;
      cudaError_t e = cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q);
      if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.slot) {
        err = std::string("cudaGetDriverEntryPoint(") + s.name +
              ") failed: " + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return;
      }
    }
    drv.ok = true;
  });
  if (!drv.ok) {
    if (why) *why = err;
    return nullptr;
  }
  return &drv;
}

extern "C" {

const char* femx_version(void) { return "femx 0.1 (sm_100a)"; }

const char* femx_last_error(const femx_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_err.c_str();
}

int femx_ctx_create(int device, femx_ctx** out) {
  if (!out) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return femx_fail(nullptr, FEMX_ERR_CUDA,
                     "femx_ctx_create: no CUDA device (%s); this engine has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= n)
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_ctx_create: device %d out of range [0,%d)",
                     device, n);
  FEMX_CUDA_OK(nullptr, cudaSetDevice(device));
  FEMX_CUDA_OK(nullptr, cudaFree(0));  // force primary-context creation
  cudaDeviceProp p;
  FEMX_CUDA_OK(nullptr, cudaGetDeviceProperties(&p, device));
  if (p.major < 10)
    return femx_fail(nullptr, FEMX_ERR_UNSUPPORTED,
                     "femx_ctx_create: device %d is sm_%d%d; this build targets sm_100a only",
                     device, p.major, p.minor);
  std::string why;
  if (!femx_get_driver(&why))
    return femx_fail(nullptr, FEMX_ERR_CUDA, "femx_ctx_create: %s", why.c_str());
  femx_ctx* c = new femx_ctx();
  c->knobs = femx_knobs_from_env();
  c->device = device;
  c->sm_count = p.multiProcessorCount;
  c->smem_optin = p.sharedMemPerBlockOptin;
  {
    // private pool: temporaries of repeated symbolic passes are recycled instead of going back to the driver
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    if (cudaMemPoolCreate(&c->pool, &pp) == cudaSuccess) {
      unsigned long long thr = ~0ULL;
      cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else {
      (void)cudaGetLastError();
      c->pool = nullptr;
    }
  }
  *out = c;
  return FEMX_OK;
}

int femx_ctx_set_option(femx_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_ctx_set_option: NULL argument");
  if (!femx_knobs_set(&ctx->knobs, name, value))
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_ctx_set_option: unknown option '%s'", name);
  return FEMX_OK;
}

void femx_ctx_destroy(femx_ctx* ctx) {
  if (!ctx) return;
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
  if (ctx->e_fork) cudaEventDestroy(ctx->e_fork);
  if (ctx->e_join) cudaEventDestroy(ctx->e_join);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
}

}  // extern "C"
