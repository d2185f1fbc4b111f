// Shared device/host helpers for libmmd_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmd.h"

namespace mmd {

// ---- error plumbing (thread-local message, no exceptions across the ABI) ------------------------------------
void set_error(const char* fmt, ...);
void count_launch();

#define MMD_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      ::mmd::set_error(__VA_ARGS__);            \
      return MMD_E_ARG;                         \
    }                                           \
  } while (0)

#define MMD_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ::mmd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return (int)e_;                                                                     \
    }                                                                                     \
  } while (0)

#define MMD_LAUNCH_CHECK()                       \
  do {                                           \
    ::mmd::count_launch();                       \
    MMD_CUDA(cudaGetLastError());                \
  } while (0)

// ---- per-device launch configuration ------------------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel and the SM count is a per-device
// number: both are tracked per device ordinal, so a process that touches several GPUs (one thread per GPU, or a test
// that moves between devices) configures every kernel on each of them.
int device_sm_count();                                    // SM count of the CURRENT device
int ensure_dynamic_smem(const void* kernel, size_t bytes); // 0 or a cudaError_t; sets the attribute once per (kernel, device)
#define MMD_SMEM(kernel, bytes)                                                          \
  do {                                                                                   \
    int rc_ = ::mmd::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (bytes)); \
    if (rc_ != 0) {                                                                      \
      ::mmd::set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu) failed: %s (%s:%d)", (size_t)(bytes), \
                       cudaGetErrorString((cudaError_t)rc_), __FILE__, __LINE__);         \
      return rc_;                                                                        \
    }                                                                                    \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The step is a chain of ~200 short dependent kernels.  Kernels launched through launch_pdl() may be scheduled while
// their predecessor still runs: everything up to pdl_wait() (TMEM allocation, mbarrier init, the bulk copy of the packed
// parameter block) overlaps the predecessor's tail; pdl_wait() returns once the predecessor has completed and its
// writes are visible.  pdl_trigger() (issued right after the wait, so that at most two kernels of the chain are ever in
// flight) lets the NEXT kernel start being scheduled.  Without the launch attribute both are no-ops.  MMD_NO_PDL=1
// disables the attribute.
bool pdl_enabled();
// mmd_bifpn_prep rewrites the packed parameter blocks that the next forward kernel bulk-copies in its prologue, i.e.
// BEFORE its griddepcontrol.wait: that one launch must not overlap its predecessor.  pdl_fence_next() (called by the
// prep entry point after its launches) makes the next launch_pdl() of this thread a plain stream-ordered launch.
void pdl_fence_next();
bool pdl_take();   // pdl_enabled() and no pending fence (consumes the fence)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_take() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- optional per-kernel CUDA-event timing (mmd_prof_*), used by bench.py for the live roofline number ---------
enum ProfKind {
  PK_MTA_POOL = 0, PK_MTA_LEVEL, PK_MTA_FINISH, PK_MTA_BWD, PK_NODE_FWD, PK_PROJ_FWD, PK_BNAPPLY, PK_NODE_BWD_A,
  PK_NODE_BWD_B, PK_PROJ_BWD, PK_PULL, PK_SLOT, PK_POOLFUSE,
  // the <16,8> instantiations (P3 / P4 of the D2 pyramid: 85 % of the bytes) are kernels of their own in the launch
  // list and are timed as such; the kinds above then hold the remaining tile shapes
  PK_NODE_FWD_16x8, PK_NODE_BWD_A_16x8, PK_NODE_BWD_B_16x8,
  // persistent small-level chains (several P5-P7 nodes per launch)
  PK_CHAIN_FWD, PK_CHAIN_BWD,
  // element-wise glue of the detection heads (heads.cu)
  PK_HEAD,
  // detection loss (focal.cu)
  PK_FOCAL,
  // pseudo-label generation (pseudo.cu): the two passes over the anchors
  PK_PSEUDO,
  // optimizer step (adam.cu)
  PK_ADAM, PK_COUNT
};
bool prof_enabled();
void prof_begin(int kind, double algo_bytes, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {  // brackets exactly one kernel launch on stream `s`
  cudaStream_t s;
  bool on;
  ProfScope(int kind, double algo_bytes, cudaStream_t s_) : s(s_), on(prof_enabled()) {
    if (on) prof_begin(kind, algo_bytes, s);
  }
  ~ProfScope() {
    if (on) prof_end(s);
  }
};

// ---- 4-channel vector access: fp32 = 16 B, bf16 = 8 B ------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st1(T* p, float v);
template <>
__device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_axpy(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---- reductions ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// SAME-padding "before" amount for a 3x3 stride-2 pool over `size` elements
// (src/YetAnotherEfficientNet.py:93-99): extra = (ceil(size/2)-1)*2 - size + 3; before = extra / 2.
__host__ __device__ __forceinline__ int pool_pad_before(int size) {
  int extra = ((size + 1) / 2 - 1) * 2 - size + 3;
  return extra / 2;
}

// fast normalised fusion weight k of n (src/YetAnotherEfficientDet.py:338-339)
__device__ __forceinline__ float fusion_weight(const float* fw, int n, int k, float eps) {
  if (fw == nullptr) return 1.f;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += fmaxf(fw[j], 0.f);
  return fmaxf(fw[k], 0.f) / (s + eps);
}

// resolve base+offset references on the host
struct Bases {
  void* const* b;
  int n;
  template <typename T>
  T* get(const MmdRef& r) const {
    if (r.base < 0 || r.base >= n) return nullptr;
    return reinterpret_cast<T*>(reinterpret_cast<char*>(b[r.base]) + r.off);
  }
};

}  // namespace mmd
