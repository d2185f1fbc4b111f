// Detection-head glue ops (Regressor / Classifier, src/YetAnotherEfficientDet.py:445-532).  The conv towers and the header
// convolutions run on the BiFPN node kernels (a head layer is a one-input, unweighted node: [BN-on-load + swish] ->
// depthwise 3x3 -> pointwise 1x1); what is left is element-wise and HBM-bound:
//   act_fwd / act_bwd          the `alignment` output swish(bn(x)) of the last level (:472, :487) and its gradient
//   head_gather / head_scatter permute(0,2,3,1) + view(B,-1,k) + cat(dim=1) [+ sigmoid] of the per-level header outputs
//                              (:475-482, :520-530): NHWC tensors padded to C channels -> one [B][sum HW][K] result
//   copy_f32                   refresh of the zero-padded staging copy of a header's pointwise weight / bias
// One thread per 4 (act) / 1 (gather, scatter) elements, fully coalesced on the wide side.
#include "bifpn.cuh"

namespace mmd {
namespace hd {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4<T>(x + 4 * i);
    st4<T>(y + 4 * i, make_float4(v.x / (1.f + expf(-v.x)), v.y / (1.f + expf(-v.y)), v.z / (1.f + expf(-v.z)),
                                  v.w / (1.f + expf(-v.w))));
  }
}

__device__ __forceinline__ float swish_grad(float x) {   // d/dx x*sigmoid(x)  (MemoryEfficientSwish backward, YetAnotherEfficientNet.py:126-137)
  const float s = 1.f / (1.f + expf(-x));
  return s * (1.f + x * (1.f - s));
}

template <typename T>
__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, T* __restrict__ dx,
                                                           long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4<T>(x + 4 * i), gg = ld4<T>(g + 4 * i);
    st4<T>(dx + 4 * i, make_float4(gg.x * swish_grad(v.x), gg.y * swish_grad(v.y), gg.z * swish_grad(v.z), gg.w * swish_grad(v.w)));
  }
}

struct HeadP {
  const void* in[2];   // gather: padded header outputs; scatter: unused
  void* dst[2];        // scatter: padded gradients
  void* out;           // the concatenated result [B][tot][K] (gather: written, scatter: read for sigmoid')
  const void* gout;    // scatter: dL/d(result)
  int B, HW, C, K, tot, off, act, n_half;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) head_gather_kernel(const __grid_constant__ HeadP P) {
  const long long total = (long long)P.B * P.HW * P.K;
  T* __restrict__ out = reinterpret_cast<T*>(P.out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % P.K);
    const long long bp = i / P.K;
    const int p = (int)(bp % P.HW), b = (int)(bp / P.HW);
    const T* src = reinterpret_cast<const T*>(P.in[k / P.C]);
    float v = ld1<T>(src + bp * P.C + (k % P.C));
    if (P.act) v = 1.f / (1.f + expf(-v));
    st1<T>(out + ((long long)b * P.tot + P.off + p) * P.K + k, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) head_scatter_kernel(const __grid_constant__ HeadP P) {
  // one thread per 4 consecutive channels of the zero-padded gradient (8 / 16-byte stores); C % 4 == 0, so a group never
  // straddles the two halves
  const int CQ = P.n_half * P.C / 4;
  const long long total = (long long)P.B * P.HW * CQ;
  const T* __restrict__ y = reinterpret_cast<const T*>(P.out);
  const T* __restrict__ g = reinterpret_cast<const T*>(P.gout);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = 4 * (int)(i % CQ);
    const long long bp = i / CQ;
    const int p = (int)(bp % P.HW), b = (int)(bp / P.HW);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < P.K) {
      const long long o = ((long long)b * P.tot + P.off + p) * P.K + c;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < P.K) {
          v[j] = ld1<T>(g + o + j);
          if (P.act) {
            const float s = ld1<T>(y + o + j);
            v[j] *= s * (1.f - s);
          }
        }
    }
    st4<T>(reinterpret_cast<T*>(P.dst[c / P.C]) + bp * P.C + (c % P.C), make_float4(v[0], v[1], v[2], v[3]));
  }
}

__global__ void __launch_bounds__(kThreads) copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

static unsigned grid_for(long long items) {
  long long g = (items + kThreads - 1) / kThreads;
  const long long cap = 8LL * device_sm_count();
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace hd

int launch_act_fwd(const void* x, void* y, long long n, int dtype, cudaStream_t s) {
  MMD_CHECK_ARG(x && y && n > 0 && n % 4 == 0, "act_fwd: bad arguments");
  const size_t es = dtype == MMD_F32 ? 4 : 2;
  ProfScope prof(PK_HEAD, 2.0 * n * es, s);
  if (dtype == MMD_F32)
    hd::act_fwd_kernel<float><<<hd::grid_for(n / 4), hd::kThreads, 0, s>>>((const float*)x, (float*)y, n / 4);
  else
    hd::act_fwd_kernel<__nv_bfloat16><<<hd::grid_for(n / 4), hd::kThreads, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 4);
  MMD_LAUNCH_CHECK();
  return 0;
}

int launch_act_bwd(const void* x, const void* g, void* dx, long long n, int dtype, cudaStream_t s) {
  MMD_CHECK_ARG(x && g && dx && n > 0 && n % 4 == 0, "act_bwd: bad arguments");
  const size_t es = dtype == MMD_F32 ? 4 : 2;
  ProfScope prof(PK_HEAD, 3.0 * n * es, s);
  if (dtype == MMD_F32)
    hd::act_bwd_kernel<float><<<hd::grid_for(n / 4), hd::kThreads, 0, s>>>((const float*)x, (const float*)g, (float*)dx, n / 4);
  else
    hd::act_bwd_kernel<__nv_bfloat16><<<hd::grid_for(n / 4), hd::kThreads, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)g,
                                                                                  (__nv_bfloat16*)dx, n / 4);
  MMD_LAUNCH_CHECK();
  return 0;
}

// `scatter` = 0: gather (in -> out); 1: scatter (gout [, out] -> dst)
int launch_head_move(int scatter, const void* const* in, void* const* dst, void* out, const void* gout, int B, int HW, int C, int K,
                     int tot, int off, int act, int dtype, cudaStream_t s) {
  const int n_half = (K + C - 1) / C;
  MMD_CHECK_ARG(B >= 1 && HW >= 1 && K >= 1 && n_half <= 2 && off >= 0 && off + HW <= tot, "head op: B=%d HW=%d K=%d tot=%d off=%d",
                B, HW, K, tot, off);
  hd::HeadP P;
  for (int j = 0; j < 2; ++j) {
    P.in[j] = (!scatter && j < n_half) ? in[j] : nullptr;
    P.dst[j] = (scatter && j < n_half) ? dst[j] : nullptr;
    if (j < n_half) MMD_CHECK_ARG(scatter ? P.dst[j] != nullptr : P.in[j] != nullptr, "head op: missing tensor %d", j);
  }
  MMD_CHECK_ARG(out != nullptr && (!scatter || gout != nullptr), "head op: missing result / gradient");
  P.out = out; P.gout = gout;
  P.B = B; P.HW = HW; P.C = C; P.K = K; P.tot = tot; P.off = off; P.act = act; P.n_half = n_half;
  const size_t es = dtype == MMD_F32 ? 4 : 2;
  const long long items = (long long)B * HW * (scatter ? n_half * C / 4 : K);
  ProfScope prof(PK_HEAD, (double)B * HW * (n_half * C + (scatter && act ? 2 : 1) * K) * es, s);
  if (!scatter) {
    if (dtype == MMD_F32) hd::head_gather_kernel<float><<<hd::grid_for(items), hd::kThreads, 0, s>>>(P);
    else hd::head_gather_kernel<__nv_bfloat16><<<hd::grid_for(items), hd::kThreads, 0, s>>>(P);
  } else {
    if (dtype == MMD_F32) hd::head_scatter_kernel<float><<<hd::grid_for(items), hd::kThreads, 0, s>>>(P);
    else hd::head_scatter_kernel<__nv_bfloat16><<<hd::grid_for(items), hd::kThreads, 0, s>>>(P);
  }
  MMD_LAUNCH_CHECK();
  return 0;
}

int launch_copy_f32(const float* src, float* dst, long long n, cudaStream_t s) {
  MMD_CHECK_ARG(src && dst && n > 0, "copy op: bad arguments");
  hd::copy_f32_kernel<<<hd::grid_for(n), hd::kThreads, 0, s>>>(src, dst, n);
  MMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmd
