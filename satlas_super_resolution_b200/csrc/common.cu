#include "common.cuh"

#include <atomic>
#include <vector>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

namespace ssr {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}
bool check_last(const char* what) { return check_cuda(cudaGetLastError(), what); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else cudaGetLastError();
  }
  return fn;
}

bool encode_tmap_tiled(CUtensorMap* out, CUtensorMapDataType dtype, uint32_t rank, const void* gaddr,
                       const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       CUtensorMapSwizzle swizzle, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU)");
    return false;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dtype, rank, const_cast<void*>(gaddr), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%u dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return false;
  }
  return true;
}

const char* last_error();
int64_t launch_count();

// ------------------------------------------------------------------ per-kernel-class device timing (bench.py roofline)
static bool g_prof_on = false;
struct ProfRec { cudaEvent_t a, b; int cls; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t pool_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

bool prof_enabled() { return g_prof_on; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
void prof_before(int cls, cudaStream_t s) {
  if (!g_prof_on) return;
  ProfRec r{pool_event(), pool_event(), cls};
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void prof_after(cudaStream_t s) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, s);
}

}  // namespace ssr

extern "C" {
const char* ssr_last_error(void) { return ssr::last_error(); }
int ssr_abi_version(void) { return 1; }
int64_t ssr_launch_count(void) { return ssr::launch_count(); }

int ssr_profile_start(void) {
  ssr::g_prof.clear();
  ssr::g_prof_on = true;
  return SSR_OK;
}
// ms[c] / count[c] per profile class (see ssr_b200.h); synchronises the device
int ssr_profile_stop(double* ms, int64_t* count, int32_t n_classes) {
  ssr::g_prof_on = false;
  if (!ssr::check_cuda(cudaDeviceSynchronize(), "profile sync")) return SSR_E_CUDA;
  for (int i = 0; i < n_classes; ++i) { ms[i] = 0.0; count[i] = 0; }
  for (auto& r : ssr::g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cls < n_classes) {
      ms[r.cls] += t;
      count[r.cls] += 1;
    }
    ssr::g_event_pool.push_back(r.a);
    ssr::g_event_pool.push_back(r.b);
  }
  ssr::g_prof.clear();
  cudaGetLastError();
  return SSR_OK;
}
}
