// Detection loss on device: YetAnotherFocalLoss.forward (src/loss/YetAnotherFocalLoss.py:27-190) for the
// whole batch in one launch, and its gradient in one more.  The reference loops over the samples in Python, builds an
// [N, M] IoU matrix, boolean-indexes five temporaries per sample and copies the labels to the device every step.
//
// One CTA = 256 consecutive anchors of one sample.
//   phase A (thread = anchor): the sample's boxes sit in shared memory; IoU against each (calc_iou's operation order with
//            round-to-nearest intrinsics, no FMA contraction: the >= 0.5 / < 0.4 decisions are bit-for-bit the fp32
//            reference's), arg-max (first maximum), the anchor's state and — for a positive — the smooth-L1 terms of its 4
//            box deltas.  The state goes to shared memory.
//   phase B (thread = element): the CTA's 256 x K classification scores are read with consecutive threads on consecutive
//            addresses, each against its anchor's state.
//   block reduction in double, three double atomics per CTA; a 1-CTA finaliser forms the two batch means.
// HBM-bound: B*N*(K + 4) elements read once (forward), read once + written once (backward).
#include "common.cuh"

namespace mmd {
namespace fl {

constexpr int kThreads = 256;
constexpr int kIgnore = -2, kNegative = -1;

struct Assign {
  int state;        // kIgnore, kNegative, or the arg-max box (index among the sample's VALID boxes) for a positive
  int cls;          // class of that box (positives)
  float t[4];       // regression targets (dy, dx, dh, dw) (positives)
};

// calc_iou (:6-20) for one anchor (y1, x1, y2, x2) and one box (x1, y1, x2, y2), every operation rounded on its own
__device__ __forceinline__ float iou_anchor_box(const float4 a, const float* b) {
  float iw = __fsub_rn(fminf(a.w, b[2]), fmaxf(a.y, b[0]));
  float ih = __fsub_rn(fminf(a.z, b[3]), fmaxf(a.x, b[1]));
  iw = fmaxf(iw, 0.f);
  ih = fmaxf(ih, 0.f);
  // no overlap: the reference's quotient is 0 / ua = +0 exactly (ua >= 1e-8); most (anchor, box) pairs end here, without
  // the areas and the IEEE division — with the labels of three teachers a sample carries up to a few hundred boxes
  if (iw == 0.f || ih == 0.f) return 0.f;
  const float area = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
  const float inter = __fmul_rn(iw, ih);
  float ua = __fsub_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), area), inter);
  ua = fmaxf(ua, 1e-8f);
  return __fdiv_rn(inter, ua);
}

__device__ __forceinline__ Assign assign_anchor(const float4 a, const float* s_box, int M, int K) {
  Assign r;
  r.state = kNegative;     // a sample without valid boxes: every anchor is a negative (:69-101)
  r.cls = -1;
  r.t[0] = r.t[1] = r.t[2] = r.t[3] = 0.f;
  float best = -1.f;
  int arg = -1, valid = 0, argrow = -1;
  for (int m = 0; m < M; ++m) {
    const float* b = s_box + 5 * m;
    if (b[4] == -1.f) continue;            // padding row (:65)
    const float v = iou_anchor_box(a, b);
    if (arg < 0 || v > best) { best = v; arg = valid; argrow = m; }
    ++valid;
  }
  if (valid == 0) return r;
  if (best >= 0.5f) {
    const float* b = s_box + 5 * argrow;
    r.state = arg;
    const int c = (int)b[4];
    r.cls = (c >= 0 && c < K) ? c : -1;
    const float aw = a.w - a.y, ah = a.z - a.x;
    const float acx = a.y + 0.5f * aw, acy = a.x + 0.5f * ah;
    float gw = b[2] - b[0], gh = b[3] - b[1];
    const float gcx = b[0] + 0.5f * gw, gcy = b[1] + 0.5f * gh;
    gw = fmaxf(gw, 1.f);
    gh = fmaxf(gh, 1.f);
    r.t[0] = __fdiv_rn(gcy - acy, ah);
    r.t[1] = __fdiv_rn(gcx - acx, aw);
    r.t[2] = logf(__fdiv_rn(gh, ah));
    r.t[3] = logf(__fdiv_rn(gw, aw));
  } else if (!(best < 0.4f)) {
    r.state = kIgnore;
  }
  return r;
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < kThreads / 32) t = s_red[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

struct FocalP {
  int B, N, K, M;
  float alpha;
  const void *cls, *reg;
  const float *anchors, *boxes;
  double* acc;
  int* assign;
  float* loss;
  const float *g_reg_loss, *g_cls_loss;
  void *grad_cls, *grad_reg;
};

template <typename T, bool BWD>
__global__ void __launch_bounds__(kThreads) focal_kernel(const __grid_constant__ FocalP P) {
  extern __shared__ __align__(16) float s_box[];          // [M][5]
  __shared__ int s_state[kThreads];                        // kIgnore / kNegative / class of the positive (-1: class out of range)
  __shared__ double s_red[kThreads / 32];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int n0 = blockIdx.x * kThreads, n = n0 + tid;
  for (int i = tid; i < 5 * P.M; i += kThreads) s_box[i] = P.boxes[(size_t)b * P.M * 5 + i];
  __shared__ int s_any;        // backward: does ANY sample of the batch have a valid box (:61-62, :181-188)
  __shared__ int s_mend;       // rows behind the last valid one are padding: a [B][256][5] tensor of device-made labels
  if (tid == 0) s_mend = 0;    // with a handful of boxes costs what the boxes cost
  if (BWD && tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < P.B; ++i) tot += P.acc[4 * i + 3];
    s_any = tot > 0.0;
  }
  __syncthreads();
  for (int m = tid; m < P.M; m += kThreads)
    if (s_box[5 * m + 4] != -1.f) atomicMax(&s_mend, m + 1);
  __syncthreads();
  if (!BWD && blockIdx.x == 0 && tid == 0) {     // acc[b][3] = valid rows of this sample (labels may be device-made)
    int nv = 0;
    for (int m = 0; m < P.M; ++m) nv += s_box[5 * m + 4] != -1.f;
    P.acc[4 * b + 3] = (double)nv;
  }

  // ---- phase A ----
  float reg_part = 0.f, pos_part = 0.f;
  float gr_scale = 0.f, gc_scale = 0.f;
  if (BWD) {
    const double npos = P.acc[4 * b + 2];
    const float g_r = P.g_reg_loss ? *P.g_reg_loss : 0.f, g_c = P.g_cls_loss ? *P.g_cls_loss : 0.f;
    gc_scale = (float)((double)g_c / P.B / (npos > 1.0 ? npos : 1.0));
    gr_scale = npos > 0.0 ? (float)((double)g_r / P.B / (4.0 * npos)) : 0.f;
    if (!s_any) gc_scale = gr_scale = 0.f;
  }
  int state = kIgnore;
  if (n < P.N) {
    const float4 a = *reinterpret_cast<const float4*>(P.anchors + 4 * (size_t)n);
    const Assign as = assign_anchor(a, s_box, s_mend, P.K);
    state = as.state >= 0 ? (as.cls >= 0 ? as.cls : 1 << 20) : as.state;   // 1 << 20: positive without a valid class
    if (!BWD && P.assign) P.assign[(size_t)b * P.N + n] = as.state;
    const size_t ro = ((size_t)b * P.N + n) * 4;
    if (as.state >= 0) {
      const float4 r = ld4<T>(reinterpret_cast<const T*>(P.reg) + ro);
      const float rv[4] = {r.x, r.y, r.z, r.w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = as.t[j] - rv[j];
        const float d = fabsf(e);
        reg_part += (d <= 1.f / 9.f) ? 0.5f * 9.f * d * d : d - 0.5f / 9.f;       // :171-175
        const float sg = (e > 0.f) ? -1.f : (e < 0.f ? 1.f : 0.f);                 // d|t - r| / dr
        g[j] = gr_scale * sg * ((d <= 1.f / 9.f) ? 9.f * d : 1.f);
      }
      pos_part = 1.f;
      if (BWD) st4<T>(reinterpret_cast<T*>(P.grad_reg) + ro, make_float4(g[0], g[1], g[2], g[3]));
    } else if (BWD) {
      st4<T>(reinterpret_cast<T*>(P.grad_reg) + ro, f4_zero());
    }
  }
  s_state[tid] = state;
  __syncthreads();

  // ---- phase B: the CTA's classification elements, consecutive threads on consecutive addresses ----
  const int rows = min(kThreads, P.N - n0);
  const int total = rows * P.K;
  const T* cbase = reinterpret_cast<const T*>(P.cls) + ((size_t)b * P.N + n0) * P.K;
  T* gbase = BWD ? reinterpret_cast<T*>(P.grad_cls) + ((size_t)b * P.N + n0) * P.K : nullptr;
  float cls_part = 0.f;
  const float alpha = P.alpha;
  // one score: clamp (:67), focal BCE term against target 1 / 0 (:113-125) and its derivative w.r.t. the raw score
  auto element = [&](float raw, bool target1, bool counted, float& grad) -> float {
    const float c = fminf(fmaxf(raw, 1e-4f), 1.f - 1e-4f);
    const bool inside = raw >= 1e-4f && raw <= 1.f - 1e-4f;      // torch.clamp passes the gradient inside the range only
    const float om = 1.f - c;
    float term, g;
    if (target1) {
      const float lg = logf(c);
      term = alpha * om * om * (-lg);
      g = alpha * (2.f * om * lg - __fdividef(om * om, c));
    } else {
      const float lg = logf(om);
      term = (1.f - alpha) * c * c * (-lg);
      g = (1.f - alpha) * (-2.f * c * lg + __fdividef(c * c, om));
    }
    grad = (counted && inside) ? gc_scale * g : 0.f;
    return counted ? term : 0.f;
  };
  if ((P.K & 3) == 0) {     // 4 consecutive classes of one anchor per thread: 8 / 16-byte accesses, one row lookup per group
    const int KQ = P.K >> 2;
    for (int i = tid; i < rows * KQ; i += kThreads) {
      const int row = i / KQ, k0 = 4 * (i - row * KQ);
      const int st = s_state[row];
      const float4 v = ld4<T>(cbase + 4 * (size_t)i);
      const bool counted = st != kIgnore;
      float4 g;
      cls_part += element(v.x, st == k0, counted, g.x) + element(v.y, st == k0 + 1, counted, g.y) +
                  element(v.z, st == k0 + 2, counted, g.z) + element(v.w, st == k0 + 3, counted, g.w);
      if (BWD) st4<T>(gbase + 4 * (size_t)i, g);
    }
  } else {
    for (int i = tid; i < total; i += kThreads) {
      const int row = i / P.K, k = i - row * P.K;
      const int st = s_state[row];
      float g;
      cls_part += element(ld1<T>(cbase + i), st == k, st != kIgnore, g);
      if (BWD) st1<T>(gbase + i, g);
    }
  }
  if (!BWD) {
    const double cs = block_sum((double)cls_part, s_red);
    const double rs = block_sum((double)reg_part, s_red);
    const double ps = block_sum((double)pos_part, s_red);
    if (tid == 0) {
      atomicAdd(P.acc + 4 * b + 0, cs);
      if (rs != 0.0) atomicAdd(P.acc + 4 * b + 1, rs);
      if (ps != 0.0) atomicAdd(P.acc + 4 * b + 2, ps);
    }
  }
}

__global__ void focal_finish_kernel(const double* __restrict__ acc, int B, float* __restrict__ loss) {
  // regression_loss = mean_b (npos_b > 0 ? reg_sum_b / (4 npos_b) : 0); classification_loss = mean_b cls_sum_b / max(npos_b, 1)
  if (threadIdx.x == 0) {
    double r = 0.0, c = 0.0, any = 0.0;
    for (int b = 0; b < B; ++b) any += acc[4 * b + 3];
    for (int b = 0; b < B && any > 0.0; ++b) {
      const double npos = acc[4 * b + 2];
      c += acc[4 * b + 0] / (npos > 1.0 ? npos : 1.0);
      if (npos > 0.0) r += acc[4 * b + 1] / (4.0 * npos);
    }
    loss[0] = (float)(r / B);
    loss[1] = (float)(c / B);
  }
}

static int check_args(const MmdFocalArgs* a) {
  MMD_CHECK_ARG(a != nullptr, "focal: null arguments");
  MMD_CHECK_ARG(a->B >= 1 && a->N >= 1 && a->K >= 1 && a->M >= 1 && a->M <= MMD_FOCAL_MAX_BOXES, "focal: B=%d N=%d K=%d M=%d", a->B,
                a->N, a->K, a->M);
  MMD_CHECK_ARG(a->dtype == MMD_F32 || a->dtype == MMD_BF16, "focal: dtype %d", a->dtype);
  MMD_CHECK_ARG(a->gamma == 2.0f, "focal: gamma must be 2 (the reference's constant), got %g", (double)a->gamma);
  MMD_CHECK_ARG(a->cls && a->reg && a->anchors && a->boxes && a->acc && a->loss, "focal: null tensor");
  MMD_CHECK_ARG((((uintptr_t)a->anchors) & 15u) == 0 && (((uintptr_t)a->reg) & 15u) == 0 && (((uintptr_t)a->cls) & 15u) == 0,
                "focal: anchors / cls / reg must be 16-byte aligned");
  return 0;
}

static FocalP make_params(const MmdFocalArgs* a) {
  FocalP P;
  P.B = a->B; P.N = a->N; P.K = a->K; P.M = a->M;
  P.alpha = a->alpha;
  P.cls = a->cls; P.reg = a->reg; P.anchors = a->anchors; P.boxes = a->boxes;
  P.acc = a->acc; P.assign = a->assign; P.loss = a->loss;
  P.g_reg_loss = P.g_cls_loss = nullptr;
  P.grad_cls = P.grad_reg = nullptr;
  return P;
}

}  // namespace fl
}  // namespace mmd

using namespace mmd;

extern "C" int mmd_focal_fwd(const MmdFocalArgs* a, mmd_stream_t stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  int rc = fl::check_args(a);
  if (rc) return rc;
  const fl::FocalP P = fl::make_params(a);
  const dim3 grid((a->N + fl::kThreads - 1) / fl::kThreads, a->B);
  const size_t smem = (size_t)a->M * 5 * sizeof(float);
  const size_t es = a->dtype == MMD_F32 ? 4 : 2;
  {
    ProfScope prof(PK_FOCAL, (double)a->B * a->N * (a->K + 4) * es, s);
    if (a->dtype == MMD_F32) fl::focal_kernel<float, false><<<grid, fl::kThreads, smem, s>>>(P);
    else fl::focal_kernel<__nv_bfloat16, false><<<grid, fl::kThreads, smem, s>>>(P);
    MMD_LAUNCH_CHECK();
  }
  fl::focal_finish_kernel<<<1, 32, 0, s>>>(a->acc, a->B, a->loss);
  MMD_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmd_focal_bwd(const MmdFocalArgs* a, const float* grad_reg_loss, const float* grad_cls_loss, void* grad_cls,
                             void* grad_reg, mmd_stream_t stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  int rc = fl::check_args(a);
  if (rc) return rc;
  MMD_CHECK_ARG(grad_cls && grad_reg && (((uintptr_t)grad_reg) & 15u) == 0 && (((uintptr_t)grad_cls) & 15u) == 0,
                "focal bwd: missing / misaligned gradient tensors");
  fl::FocalP P = fl::make_params(a);
  P.g_reg_loss = grad_reg_loss; P.g_cls_loss = grad_cls_loss;
  P.grad_cls = grad_cls; P.grad_reg = grad_reg;
  const dim3 grid((a->N + fl::kThreads - 1) / fl::kThreads, a->B);
  const size_t smem = (size_t)a->M * 5 * sizeof(float);
  const size_t es = a->dtype == MMD_F32 ? 4 : 2;
  ProfScope prof(PK_FOCAL, 2.0 * a->B * a->N * (a->K + 4) * es, s);
  if (a->dtype == MMD_F32) fl::focal_kernel<float, true><<<grid, fl::kThreads, smem, s>>>(P);
  else fl::focal_kernel<__nv_bfloat16, true><<<grid, fl::kThreads, smem, s>>>(P);
  MMD_LAUNCH_CHECK();
  return 0;
}
