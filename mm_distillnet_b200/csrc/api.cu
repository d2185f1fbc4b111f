// C-ABI bookkeeping entry points of libmmd_b200.so (see include/mmd.h).
#include <stdarg.h>

#include <atomic>

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
}  // namespace mmd

extern "C" int mmd_version(void) { return MMD_VERSION; }
extern "C" const char* mmd_last_error(void) { return mmd::g_err; }
extern "C" unsigned long long mmd_launch_count(void) { return mmd::g_launches.load(std::memory_order_relaxed); }
extern "C" size_t mmd_sizeof_op(void) { return sizeof(MmdOp); }
extern "C" size_t mmd_sizeof_mta_args(void) { return sizeof(MmdMtaArgs); }
