// C-ABI bookkeeping entry points of libmmd_b200.so (see include/mmd.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace mmd {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMD_NO_PDL");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

static thread_local bool g_pdl_fence = false;
void pdl_fence_next() { g_pdl_fence = true; }
bool pdl_take() {
  const bool fenced = g_pdl_fence;
  g_pdl_fence = false;
  return pdl_enabled() && !fenced;
}

// ---- per-device launch configuration (see common.cuh) ----------------------------------------------------------------
static std::mutex g_cfg_mu;
struct CfgDone { const void* kernel; int dev; size_t bytes; };
static std::vector<CfgDone> g_cfg_done;   // largest dynamic shared memory size configured per (kernel, device)
static int g_sm_count[64] = {0};
int device_sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  int n = g_sm_count[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return n;
}
int ensure_dynamic_smem(const void* kernel, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  CfgDone* hit = nullptr;
  for (auto& e : g_cfg_done)
    if (e.kernel == kernel && e.dev == dev) hit = &e;
  if (hit != nullptr && hit->bytes >= bytes) return 0;      // (a kernel may be launched with several sizes: keep the max)
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  if (hit != nullptr) hit->bytes = bytes;
  else g_cfg_done.push_back(CfgDone{kernel, dev, bytes});
  return 0;
}

// ---- profiler: event pairs around individual launches, summed per kernel kind on collect --------------------
static const char* kProfNames[PK_COUNT] = {"mta_pool", "mta_level", "mta_finish", "mta_bwd", "node_fwd", "proj_fwd",
                                           "bnapply", "node_bwd_a", "node_bwd_b", "proj_bwd", "pull", "slot",
                                           "poolfuse", "node_fwd<16,8>", "node_bwd_a<16,8>", "node_bwd_b<16,8>", "chain_fwd", "chain_bwd", "head_glue", "focal", "pseudo", "adam"};
struct ProfRec { cudaEvent_t a, b; int kind; double bytes; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
bool prof_enabled() { return g_prof_on; }
void prof_begin(int kind, double algo_bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  if (!g_prof_pool.empty()) {
    r.a = g_prof_pool.back().first;
    r.b = g_prof_pool.back().second;
    g_prof_pool.pop_back();
  } else {
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
  }
  r.kind = kind;
  r.bytes = algo_bytes;
  cudaEventRecord(r.a, s);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().b, s);
}
}  // namespace mmd

extern "C" void mmd_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(mmd::g_prof_mu);
  mmd::g_prof_on = on != 0;
}
extern "C" int mmd_prof_num_kinds(void) { return mmd::PK_COUNT; }
extern "C" const char* mmd_prof_kind_name(int k) { return (k >= 0 && k < mmd::PK_COUNT) ? mmd::kProfNames[k] : ""; }
// Synchronises the device, adds every recorded launch to ms[kind] / launches[kind] / bytes[kind] and clears the log.
extern "C" int mmd_prof_collect(double* ms, long long* launches, double* bytes) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    mmd::set_error("mmd_prof_collect: %s", cudaGetErrorString(e));
    return (int)e;
  }
  std::lock_guard<std::mutex> lk(mmd::g_prof_mu);
  for (auto& r : mmd::g_prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.kind] += t;
      launches[r.kind] += 1;
      bytes[r.kind] += r.bytes;
    }
    mmd::g_prof_pool.emplace_back(r.a, r.b);
  }
  mmd::g_prof_recs.clear();
  return 0;
}

namespace mmd {
void set_chain_fwd(int on);
void set_mta_fast(int on);
void set_proj_tma(int on);
}
// Runtime switches (each also has an environment default, read once): returns 0, or MMD_E_ARG for an unknown name.
extern "C" int mmd_set_option(const char* name, int value) {
  if (name != nullptr && strcmp(name, "chain_fwd") == 0) {
    mmd::set_chain_fwd(value);
    return 0;
  }
  if (name != nullptr && strcmp(name, "mta_fast") == 0) {
    mmd::set_mta_fast(value);
    return 0;
  }
  if (name != nullptr && strcmp(name, "proj_tma") == 0) {
    mmd::set_proj_tma(value);
    return 0;
  }
  mmd::set_error("mmd_set_option: unknown option '%s'", name ? name : "(null)");
  return MMD_E_ARG;
}

extern "C" int mmd_version(void) { return MMD_VERSION; }
extern "C" const char* mmd_last_error(void) { return mmd::g_err; }
extern "C" unsigned long long mmd_launch_count(void) { return mmd::g_launches.load(std::memory_order_relaxed); }
extern "C" size_t mmd_sizeof_op(void) { return sizeof(MmdOp); }
extern "C" size_t mmd_sizeof_mta_args(void) { return sizeof(MmdMtaArgs); }
extern "C" size_t mmd_sizeof_focal_args(void) { return sizeof(MmdFocalArgs); }
