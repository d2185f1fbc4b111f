// Pseudo-label generation on device: the teachers' (classification, regression) outputs become the padded
// annotation tensor of the detection loss without a host round trip.  Restates, for all teachers and samples at once,
//   EfficientDet_post_processing + logits_to_ground_truth   src/utils/utils.py:144-231, :234-324
//   YetAnotherEfficientDetBBoxTransform / ClipBoxes          src/YetAnotherEfficientDet.py:574-602, src/utils/utils.py:123-141
//   torchvision batched_nms (coordinate trick) / nms         the reference's dependency (greedy, stable score order)
//   the cross-teacher integration + nms of the step wrappers src/optimization/train_methods.py:360-411
// The reference runs these per sample in Python with a .cpu() per sample and teacher.
//
// Four launches:
//   pl_score_kernel    grid (ceil(N/256), B, T): the CTA's 256 x K scores arrive through shared memory (consecutive threads
//                      on consecutive 16-byte vectors), thread = anchor: max score / first arg-max / threshold / valid-class
//                      test; writes 5 bytes per anchor (score or -1, class) and the CTA's two counts.  HBM-bound: this is the
//                      only pass that reads the [B][N][K] tensors.
//   pl_compact_kernel  same grid: ORDER-PRESERVING compaction (the reference's boolean indexing keeps anchor order, and the
//                      stable sort of the NMS breaks score ties by it): offset = sum of the preceding CTAs' counts, rank =
//                      ballot prefix; over-threshold anchors -> score list, those of a valid class -> decoded + clipped box,
//                      score, class.
//   pl_nms_kernel      grid (B, T), one CTA per (sample, teacher): max coordinate, 64-bit keys (~score bits, index) sorted by a
//                      shared-memory bitonic network (= stable descending score order), greedy suppression with the whole CTA
//                      testing the remaining boxes of each kept box, rows emitted in NMS order.
//   pl_merge_kernel    grid (B): teachers' rows concatenated in teacher order, the same sort + greedy NMS, padded labels out.
// Every operation that decides an index is a single rounded fp32 operation (__fadd_rn / __fmul_rn / __fdiv_rn: no FMA
// contraction), in the operation order of the reference / torchvision's CPU kernel.
#include "common.cuh"

namespace mmd {
namespace pl {

constexpr int kThreads = 256;       // score / compact: anchors per CTA
constexpr int kNmsThreads = 1024;

struct Ws {                          // carved from MmdPseudoArgs.workspace
  float* sc;                         // [T][B][N] score of an over-threshold anchor, -1 otherwise
  uint8_t* cl;                       // [T][B][N] arg-max class
  int2* blk;                         // [T][B][nblk] (over-threshold, over-threshold with a valid class) per CTA
  float* over_score;                 // [T][B][cap]
  float4* cand_box;                  // [T][B][cap] (x1, y1, x2, y2) decoded + clipped
  float* cand_score;                 // [T][B][cap]
  int* cand_cls;                     // [T][B][cap]
  size_t bytes;
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static Ws carve(const MmdPseudoArgs* a) {
  Ws w;
  const size_t TB = (size_t)a->T * a->B, nblk = (a->N + kThreads - 1) / kThreads;
  char* p = reinterpret_cast<char*>(a->workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  w.sc = reinterpret_cast<float*>(take(TB * a->N * sizeof(float)));
  w.cl = reinterpret_cast<uint8_t*>(take(TB * a->N));
  w.blk = reinterpret_cast<int2*>(take(TB * nblk * sizeof(int2)));
  w.over_score = reinterpret_cast<float*>(take(TB * a->cap * sizeof(float)));
  w.cand_box = reinterpret_cast<float4*>(take(TB * a->cap * sizeof(float4)));
  w.cand_score = reinterpret_cast<float*>(take(TB * a->cap * sizeof(float)));
  w.cand_cls = reinterpret_cast<int*>(take(TB * a->cap * sizeof(int)));
  w.bytes = off;
  return w;
}

struct P {
  int B, N, K, T, cap, max_rows, max_labels, raw_rows, n_ignore, nblk, merge01;
  int ignore[MMD_PL_MAX_IGNORE];
  float conf, size;
  double nms_thr, merge_thr;
  const void* cls[MMD_PL_MAX_TEACHERS];
  const void* reg[MMD_PL_MAX_TEACHERS];
  const float* anchors;
  const int* label_of;
  Ws w;
  float* teacher_rows;
  int* teacher_counts;
  float* labels;
  int* counts;
};

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- pass 1: score, class, flags, per-CTA counts -----------------------------------------------------------------------
// Row maximum of K values that sit in 8-byte (bf16) / 16-byte (fp32) groups of 4: one max instruction per value (fp32) or
// per pair (bf16 HMNMX2); the arg-max is only looked up for the few anchors above the threshold.
template <typename T>
__device__ __forceinline__ float row_max4(const T* row, int K);
template <>
__device__ __forceinline__ float row_max4<float>(const float* row, int K) {
  float4 m = *reinterpret_cast<const float4*>(row);
  for (int k = 4; k < K; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + k);
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
  return fmaxf(fmaxf(m.x, m.y), fmaxf(m.z, m.w));
}
template <>
__device__ __forceinline__ float row_max4<__nv_bfloat16>(const __nv_bfloat16* row, int K) {
  uint2 raw = *reinterpret_cast<const uint2*>(row);
  __nv_bfloat162 m0 = *reinterpret_cast<__nv_bfloat162*>(&raw.x), m1 = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  for (int k = 4; k < K; k += 4) {
    raw = *reinterpret_cast<const uint2*>(row + k);
    m0 = __hmax2(m0, *reinterpret_cast<__nv_bfloat162*>(&raw.x));
    m1 = __hmax2(m1, *reinterpret_cast<__nv_bfloat162*>(&raw.y));
  }
  m0 = __hmax2(m0, m1);
  return fmaxf(__low2float(m0), __high2float(m0));
}

// The CTA's 256 x K scores are copied to shared memory AS THEY LIE in HBM (16-byte vectors, consecutive threads on
// consecutive addresses, no index arithmetic); thread = anchor then reads its row with 8 / 16-byte shared loads.  For
// K = 20 the row stride is 10 (bf16) / 20 (fp32) words: the 16 (8) threads of one shared-memory phase hit disjoint banks.
template <typename T>
__global__ void __launch_bounds__(kThreads) pl_score_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(16) unsigned char s_bytes[];
  T* s_sc = reinterpret_cast<T*>(s_bytes);           // [256][K] in the tensor's own layout
  const int t = blockIdx.z, b = blockIdx.y, tid = threadIdx.x;
  const int n0 = blockIdx.x * kThreads;
  const int rows = min(kThreads, p.N - n0);
  const int total = rows * p.K;
  const size_t e0 = ((size_t)b * p.N + n0) * p.K;
  const T* src = reinterpret_cast<const T*>(p.cls[t]) + e0;
  constexpr int V = 16 / sizeof(T);
  if (e0 % V == 0) {                                 // 16-byte vectors (the tensor base is 16-byte aligned)
    const int nv = total / V;
    for (int i = tid; i < nv; i += kThreads)
      reinterpret_cast<uint4*>(s_bytes)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (int idx = nv * V + tid; idx < total; idx += kThreads) s_sc[idx] = src[idx];
  } else {
    for (int idx = tid; idx < total; idx += kThreads) s_sc[idx] = src[idx];
  }
  __syncthreads();
  bool over = false, valid = false;
  float best = -1.f;
  int arg = 0;
  if (tid < rows) {
    const T* row = s_sc + tid * p.K;
    if ((p.K & 3) == 0) {
      best = row_max4<T>(row, p.K);
    } else {
      best = to_f<T>(row[0]);
      for (int k = 1; k < p.K; ++k) best = fmaxf(best, to_f<T>(row[k]));
    }
    over = best > p.conf;                           // utils.py:178-179
    if (over) {
      while (to_f<T>(row[arg]) != best) ++arg;      // torch.max(dim): the FIRST maximum
      valid = p.label_of[arg] >= 0;                 // :197-204
    }
  }
  const int c_over = __syncthreads_count(over);
  const int c_valid = __syncthreads_count(valid);
  if (tid == 0) p.w.blk[((size_t)t * p.B + b) * p.nblk + blockIdx.x] = make_int2(c_over, c_valid);
  if (c_over > 0 && tid < rows) {                   // pass 2 only looks at the CTAs that counted something
    const size_t o = ((size_t)t * p.B + b) * p.N + n0 + tid;
    p.w.sc[o] = over ? best : -1.f;
    p.w.cl[o] = (uint8_t)arg;
  }
}

// ---- pass 2: order-preserving compaction + box decode ------------------------------------------------------------------
// One CTA owns kSpan consecutive 256-anchor blocks of one (teacher, sample) and visits only those whose pass-1 count is
// non-zero (a trained detector fires on a few hundred of 110 484 anchors: almost every block is skipped unread).
constexpr int kSpan = 16;
template <typename T>
__global__ void __launch_bounds__(kThreads) pl_compact_kernel(const __grid_constant__ P p) {
  __shared__ int s_red[2][kThreads / 32];
  __shared__ int s_off[2];
  __shared__ int2 s_cnt[kSpan];
  const int t = blockIdx.z, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t tb = (size_t)t * p.B + b;
  const int blk0 = blockIdx.x * kSpan, nspan = min(kSpan, p.nblk - blk0);
  if (tid < nspan) s_cnt[tid] = p.w.blk[tb * p.nblk + blk0 + tid];
  __syncthreads();
  int any = 0;
  for (int i = 0; i < nspan; ++i) any |= s_cnt[i].x;
  if (any == 0) return;
  // offsets of this CTA = counts of the blocks before its span
  int so = 0, sv = 0;
  for (int i = tid; i < blk0; i += kThreads) {
    const int2 c = p.w.blk[tb * p.nblk + i];
    so += c.x;
    sv += c.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    so += __shfl_xor_sync(0xffffffffu, so, o);
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
  }
  if (lane == 0) { s_red[0][warp] = so; s_red[1][warp] = sv; }
  __syncthreads();
  if (tid == 0) {
    int a0 = 0, a1 = 0;
    for (int w = 0; w < kThreads / 32; ++w) { a0 += s_red[0][w]; a1 += s_red[1][w]; }
    s_off[0] = a0;
    s_off[1] = a1;
  }
  __syncthreads();
  int off_o = s_off[0], off_v = s_off[1];
  for (int i = 0; i < nspan; ++i) {
    const int2 cnt = s_cnt[i];
    if (cnt.x == 0) continue;
    const int n = (blk0 + i) * kThreads + tid;
    float score = -1.f;
    int cls = 0;
    if (n < p.N) {
      score = p.w.sc[tb * p.N + n];
      cls = p.w.cl[tb * p.N + n];
    }
    const bool over = score >= 0.f;
    const bool valid = over && p.label_of[cls] >= 0;
    const unsigned m_o = __ballot_sync(0xffffffffu, over), m_v = __ballot_sync(0xffffffffu, valid);
    __syncthreads();
    if (lane == 0) { s_red[0][warp] = __popc(m_o); s_red[1][warp] = __popc(m_v); }
    __syncthreads();
    int base_o = off_o, base_v = off_v;
    for (int w = 0; w < warp; ++w) { base_o += s_red[0][w]; base_v += s_red[1][w]; }
    const unsigned below = (1u << lane) - 1u;
    const int pos_o = base_o + __popc(m_o & below), pos_v = base_v + __popc(m_v & below);
    if (over && pos_o < p.cap) p.w.over_score[tb * p.cap + pos_o] = score;
    if (valid && pos_v < p.cap) {
      // YetAnotherEfficientDetBBoxTransform.forward (YetAnotherEfficientDet.py:586-602), one rounding per operation
      const float4 a = *reinterpret_cast<const float4*>(p.anchors + 4 * (size_t)n);      // y1, x1, y2, x2
      const T* rp = reinterpret_cast<const T*>(p.reg[t]) + ((size_t)b * p.N + n) * 4;
      const float r0 = to_f<T>(rp[0]), r1 = to_f<T>(rp[1]), r2 = to_f<T>(rp[2]), r3 = to_f<T>(rp[3]);
      const float yca = __fdiv_rn(__fadd_rn(a.x, a.z), 2.f), xca = __fdiv_rn(__fadd_rn(a.y, a.w), 2.f);
      const float ha = __fsub_rn(a.z, a.x), wa = __fsub_rn(a.w, a.y);
      const float w = __fmul_rn((float)exp((double)r3), wa), h = __fmul_rn((float)exp((double)r2), ha);
      const float yc = __fadd_rn(__fmul_rn(r0, ha), yca), xc = __fadd_rn(__fmul_rn(r1, wa), xca);
      const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
      float4 bx;
      bx.x = fmaxf(__fsub_rn(xc, hw), 0.f);            // ClipBoxes (utils.py:134-138)
      bx.y = fmaxf(__fsub_rn(yc, hh), 0.f);
      bx.z = fminf(__fadd_rn(xc, hw), p.size);
      bx.w = fminf(__fadd_rn(yc, hh), p.size);
      p.w.cand_box[tb * p.cap + pos_v] = bx;
      p.w.cand_score[tb * p.cap + pos_v] = score;
      p.w.cand_cls[tb * p.cap + pos_v] = cls;
    }
    off_o += cnt.x;
    off_v += cnt.y;
  }
}

// ---- stable descending sort + greedy NMS of one list inside one CTA ---------------------------------------------------
// s_key[i] = (~score bits) << 32 | i for i < n (scores are positive floats: their bit patterns order like the values), all
// ones for the padding up to the next power of two; an ascending sort gives descending scores, ties by ascending index =
// the stable sort of torchvision's kernel.
__device__ __forceinline__ void bitonic_sort(unsigned long long* s_key, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int x = i ^ j;
        if (x > i) {
          const unsigned long long a = s_key[i], b = s_key[x];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_key[i] = b; s_key[x] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// torchvision nms_kernel_impl's test of box j against the kept box i (both (x1, y1, x2, y2))
__device__ __forceinline__ bool suppresses(const float4 bi, const float iarea, const float4 bj, const double thr) {
  const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y), xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
  return (double)ovr > thr;
}

// Greedy pass over the sorted list s_box[0..n) (s_dead zeroed).  Calls emit(rank) from thread 0 for every kept box, in order.
template <typename Emit>
__device__ __forceinline__ void greedy_nms(const float4* s_box, uint8_t* s_dead, int n, double thr, int* s_next, Emit emit) {
  int cur = 0;
  while (cur < n) {
    if (threadIdx.x == 0) {
      emit(cur);
      *s_next = n;
    }
    __syncthreads();
    const float4 bi = s_box[cur];
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    int first = n;
    for (int j = cur + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (s_dead[j]) continue;
      if (suppresses(bi, iarea, s_box[j], thr)) s_dead[j] = 1;
      else if (first == n) first = j;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if ((threadIdx.x & 31) == 0 && first < n) atomicMin(s_next, first);
    __syncthreads();
    cur = *s_next;
    __syncthreads();
  }
}

__device__ __forceinline__ float block_max(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s_red[w]);
  __syncthreads();
  return r;
}

__device__ __forceinline__ int next_pow2(int n) {
  int r = 1;
  while (r < n) r <<= 1;
  return r;
}

// shared memory of both NMS kernels: keys [cap2] | boxes [cap] | dead [cap]
__device__ __forceinline__ void nms_smem(unsigned char* base, int cap, unsigned long long*& key, float4*& box, uint8_t*& dead) {
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  key = reinterpret_cast<unsigned long long*>(base);
  box = reinterpret_cast<float4*>(base + (size_t)cap2 * 8);
  dead = reinterpret_cast<uint8_t*>(base + (size_t)cap2 * 8 + (size_t)cap * 16);
}
static size_t nms_smem_bytes(int cap) {
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  return (size_t)cap2 * 8 + (size_t)cap * 16 + (size_t)cap;
}

// ---- pass 3: class-wise NMS of one (teacher, sample) ------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads) pl_nms_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ float s_red[kNmsThreads / 32];
  __shared__ int s_next, s_rows;
  unsigned long long* s_key;
  float4* s_box;
  uint8_t* s_dead;
  nms_smem(s_raw, p.cap, s_key, s_box, s_dead);
  const int b = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const size_t tb = (size_t)t * p.B + b;
  // totals of the two pass-1 counts over the sample's blocks
  __shared__ int s_tot[2];
  if (tid < 2) s_tot[tid] = 0;
  __syncthreads();
  {
    int so = 0, sv = 0;
    for (int i = tid; i < p.nblk; i += kNmsThreads) {
      const int2 c = p.w.blk[tb * p.nblk + i];
      so += c.x;
      sv += c.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      so += __shfl_xor_sync(0xffffffffu, so, o);
      sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((tid & 31) == 0 && (so | sv) != 0) {
      atomicAdd(&s_tot[0], so);
      atomicAdd(&s_tot[1], sv);
    }
  }
  __syncthreads();
  if (tid == 0 && s_tot[0] > p.cap) atomicOr(p.counts + p.B, 1);
  const int n = min(s_tot[1], p.cap);
  if (n == 0) {
    if (tid == 0) p.teacher_counts[tb] = 0;
    return;
  }
  const float4* box = p.w.cand_box + tb * p.cap;
  const float* score = p.w.cand_score + tb * p.cap;
  const int* cls = p.w.cand_cls + tb * p.cap;
  // batched_nms' coordinate trick: offsets = class * (boxes.max() + 1)
  float mx = -INFINITY;
  for (int i = tid; i < n; i += kNmsThreads) {
    const float4 v = box[i];
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  mx = block_max(mx, s_red);
  const float step = __fadd_rn(mx, 1.f);
  const int n2 = next_pow2(n);
  for (int i = tid; i < n2; i += kNmsThreads)
    s_key[i] = i < n ? ((unsigned long long)(~__float_as_uint(score[i])) << 32) | (unsigned)i : ~0ull;
  __syncthreads();
  bitonic_sort(s_key, n2);
  for (int r = tid; r < n; r += kNmsThreads) {
    const int i = (int)(s_key[r] & 0xffffffffu);
    const float4 v = box[i];
    const float off = __fmul_rn((float)cls[i], step);
    s_box[r] = make_float4(__fadd_rn(v.x, off), __fadd_rn(v.y, off), __fadd_rn(v.z, off), __fadd_rn(v.w, off));
    s_dead[r] = 0;
  }
  if (tid == 0) s_rows = 0;
  __syncthreads();
  float* rows = p.teacher_rows + tb * p.max_rows * 6;
  const float* over_score = p.w.over_score + tb * p.cap;
  greedy_nms(s_box, s_dead, n, p.nms_thr, &s_next, [&](int r) {
    const int i = (int)(s_key[r] & 0xffffffffu);
    const int c = cls[i];
    for (int q = 0; q < p.n_ignore; ++q)
      if (c == p.ignore[q]) return;                      // utils.py:212-215 (dropped after the NMS: it still suppressed)
    if (s_rows >= p.max_rows) {
      atomicOr(p.counts + p.B, 2);
      return;
    }
    const float4 v = box[i];
    float* o = rows + 6 * s_rows;
    if (p.raw_rows) {
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      o[5] = (float)c;
    } else {                                             // logits_to_ground_truth :293-296: int(max(.,0)) / int(min(.,S))
      o[0] = truncf(fmaxf(v.x, 0.f));
      o[1] = truncf(fmaxf(v.y, 0.f));
      o[2] = truncf(fminf(v.z, p.size));
      o[3] = truncf(fminf(v.w, p.size));
      o[5] = (float)p.label_of[c];
    }
    o[4] = over_score[i];                                // the reference's score quirk (:195 vs :202-209)
    ++s_rows;
  });
  if (tid == 0) p.teacher_counts[tb] = s_rows;
}

// ---- pass 4: cross-teacher integration (train_methods.py:360-411) -------------------------------------------------------
// The list of sample b = the teachers' rows in teacher order; with merge01 (the augmented step, :384-386) sample 1's list
// is sample 0's list followed by its own when both are non-empty.
__global__ void __launch_bounds__(kNmsThreads) pl_merge_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int kSeg = 2 * MMD_PL_MAX_TEACHERS;
  __shared__ int s_next, s_rows, s_nseg, s_start[kSeg + 1], s_src[kSeg];     // s_src: (sample, teacher) segment -> tb index
  unsigned long long* s_key;
  float4* s_box;
  uint8_t* s_dead;
  const int capm = (p.merge01 ? 2 : 1) * p.T * p.max_rows;
  nms_smem(s_raw, capm, s_key, s_box, s_dead);
  float* s_label = reinterpret_cast<float*>(s_dead + ((capm + 15) & ~15));      // [capm] label of sorted row r
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0, nseg = 0, own = 0, first = 0;
    for (int t = 0; t < p.T; ++t) own += p.teacher_counts[(size_t)t * p.B + b];
    if (p.merge01 && b == 1 && own > 0) {
      for (int t = 0; t < p.T; ++t) first += p.teacher_counts[(size_t)t * p.B + 0];
      if (first > 0)
        for (int t = 0; t < p.T; ++t) {
          s_start[nseg] = acc;
          s_src[nseg++] = t * p.B + 0;
          acc += p.teacher_counts[(size_t)t * p.B + 0];
        }
    }
    for (int t = 0; t < p.T; ++t) {
      s_start[nseg] = acc;
      s_src[nseg++] = t * p.B + b;
      acc += p.teacher_counts[(size_t)t * p.B + b];
    }
    s_start[nseg] = acc;
    s_nseg = nseg;
    s_rows = 0;
  }
  __syncthreads();
  const int nseg = s_nseg, n = s_start[nseg];
  float* out = p.labels + (size_t)b * p.max_labels * 5;
  auto row_of = [&](int i) -> const float* {
    int g = 0;
    while (i >= s_start[g + 1]) ++g;
    return p.teacher_rows + ((size_t)s_src[g] * p.max_rows + (i - s_start[g])) * 6;
  };
  if (n > 0) {
    const int n2 = next_pow2(n);
    for (int i = tid; i < n2; i += kNmsThreads)
      s_key[i] = i < n ? ((unsigned long long)(~__float_as_uint(row_of(i)[4])) << 32) | (unsigned)i : ~0ull;
    __syncthreads();
    bitonic_sort(s_key, n2);
    for (int r = tid; r < n; r += kNmsThreads) {
      const float* row = row_of((int)(s_key[r] & 0xffffffffu));
      s_box[r] = make_float4(row[0], row[1], row[2], row[3]);
      s_label[r] = row[5];
      s_dead[r] = 0;
    }
    __syncthreads();
    greedy_nms(s_box, s_dead, n, p.merge_thr, &s_next, [&](int r) {
      if (s_rows >= p.max_labels) {
        atomicOr(p.counts + p.B, 4);
        return;
      }
      const float4 v = s_box[r];
      float* o = out + 5 * s_rows;
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      o[4] = s_label[r];
      ++s_rows;
    });
  }
  __syncthreads();
  const int kept = s_rows;
  for (int i = kept * 5 + tid; i < p.max_labels * 5; i += kNmsThreads) out[i] = -1.f;     // annot_padded's fill value
  if (tid == 0) p.counts[b] = kept;
}

__global__ void pl_clear_kernel(int* flag) { *flag = 0; }

static int check_args(const MmdPseudoArgs* a) {
  MMD_CHECK_ARG(a != nullptr, "pseudo: null arguments");
  MMD_CHECK_ARG(a->B >= 1 && a->B <= 65535 && a->N >= 1 && a->K >= 1 && a->K <= 255 && a->T >= 1 && a->T <= MMD_PL_MAX_TEACHERS,
                "pseudo: B=%d N=%d K=%d T=%d", a->B, a->N, a->K, a->T);
  MMD_CHECK_ARG(a->dtype == MMD_F32 || a->dtype == MMD_BF16, "pseudo: dtype %d", a->dtype);
  MMD_CHECK_ARG(a->cap >= 1 && a->cap <= MMD_PL_MAX_CAP && a->max_rows >= 1 && (long long)a->T * a->max_rows <= MMD_PL_MAX_CAP &&
                    a->max_labels >= 1,
                "pseudo: cap=%d max_rows=%d max_labels=%d (cap, T * max_rows <= %d)", a->cap, a->max_rows, a->max_labels, MMD_PL_MAX_CAP);
  MMD_CHECK_ARG(a->merge01 == 0 || (a->B >= 2 && 2LL * a->T * a->max_rows <= MMD_PL_MAX_CAP),
                "pseudo: merge01 needs B >= 2 and 2 * T * max_rows <= %d (B=%d T=%d max_rows=%d)", MMD_PL_MAX_CAP, a->B, a->T, a->max_rows);
  MMD_CHECK_ARG(a->n_ignore >= 0 && a->n_ignore <= MMD_PL_MAX_IGNORE, "pseudo: n_ignore=%d", a->n_ignore);
  MMD_CHECK_ARG(a->conf_threshold >= 0.f, "pseudo: conf_threshold %g must be >= 0 (scores are probabilities)", (double)a->conf_threshold);
  MMD_CHECK_ARG(a->anchors && a->label_of && a->workspace && a->teacher_rows && a->teacher_counts && a->labels && a->counts,
                "pseudo: null tensor");
  MMD_CHECK_ARG((((uintptr_t)a->anchors) & 15u) == 0 && (((uintptr_t)a->workspace) & 15u) == 0, "pseudo: anchors / workspace must be 16-byte aligned");
  for (int t = 0; t < a->T; ++t)
    MMD_CHECK_ARG(a->cls[t] && a->reg[t] && (((uintptr_t)a->cls[t]) & 15u) == 0 && (((uintptr_t)a->reg[t]) & 7u) == 0,
                  "pseudo: teacher %d: missing / misaligned cls / reg", t);
  return 0;
}

}  // namespace pl
}  // namespace mmd

using namespace mmd;

extern "C" size_t mmd_sizeof_pseudo_args(void) { return sizeof(MmdPseudoArgs); }

extern "C" size_t mmd_pseudo_workspace_bytes(const MmdPseudoArgs* a) {
  if (a == nullptr || a->B < 1 || a->N < 1 || a->T < 1 || a->cap < 1) return 0;
  MmdPseudoArgs c = *a;
  c.workspace = nullptr;
  return pl::carve(&c).bytes;
}

extern "C" int mmd_pseudo_labels(const MmdPseudoArgs* a, mmd_stream_t stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  int rc = pl::check_args(a);
  if (rc) return rc;
  const int capm = (a->merge01 ? 2 : 1) * a->T * a->max_rows;
  const size_t smem4 = pl::nms_smem_bytes(capm) + 16 + (size_t)capm * sizeof(float);
  MMD_CHECK_ARG(smem4 <= 232448, "pseudo: %d concatenated rows per sample need %zu bytes of shared memory in the merge kernel "
                "(limit 232448): lower max_rows", capm, smem4);
  pl::P p;
  p.B = a->B; p.N = a->N; p.K = a->K; p.T = a->T; p.cap = a->cap; p.max_rows = a->max_rows; p.max_labels = a->max_labels;
  p.raw_rows = a->raw_rows; p.n_ignore = a->n_ignore; p.merge01 = a->merge01 ? 1 : 0;
  p.nblk = (a->N + pl::kThreads - 1) / pl::kThreads;
  for (int i = 0; i < MMD_PL_MAX_IGNORE; ++i) p.ignore[i] = a->ignore[i];
  p.conf = a->conf_threshold; p.size = a->image_size;
  p.nms_thr = a->nms_threshold; p.merge_thr = a->merge_iou;
  for (int t = 0; t < MMD_PL_MAX_TEACHERS; ++t) { p.cls[t] = t < a->T ? a->cls[t] : nullptr; p.reg[t] = t < a->T ? a->reg[t] : nullptr; }
  p.anchors = a->anchors; p.label_of = a->label_of;
  p.w = pl::carve(a);
  p.teacher_rows = a->teacher_rows; p.teacher_counts = a->teacher_counts; p.labels = a->labels; p.counts = a->counts;

  pl::pl_clear_kernel<<<1, 1, 0, s>>>(a->counts + a->B);
  MMD_LAUNCH_CHECK();
  const dim3 grid(p.nblk, a->B, a->T);
  const size_t es = a->dtype == MMD_F32 ? 4 : 2;
  const size_t smem1 = (((size_t)pl::kThreads * a->K * es) + 15) & ~(size_t)15;
  {
    ProfScope prof(PK_PSEUDO, (double)a->T * a->B * a->N * (a->K * es), s);
    if (a->dtype == MMD_F32) {
      MMD_SMEM(pl::pl_score_kernel<float>, smem1);
      pl::pl_score_kernel<float><<<grid, pl::kThreads, smem1, s>>>(p);
    } else {
      MMD_SMEM(pl::pl_score_kernel<__nv_bfloat16>, smem1);
      pl::pl_score_kernel<__nv_bfloat16><<<grid, pl::kThreads, smem1, s>>>(p);
    }
    MMD_LAUNCH_CHECK();
  }
  {
    ProfScope prof(PK_PSEUDO, 0.0, s);
    const dim3 grid2((p.nblk + pl::kSpan - 1) / pl::kSpan, a->B, a->T);
    if (a->dtype == MMD_F32) pl::pl_compact_kernel<float><<<grid2, pl::kThreads, 0, s>>>(p);
    else pl::pl_compact_kernel<__nv_bfloat16><<<grid2, pl::kThreads, 0, s>>>(p);
    MMD_LAUNCH_CHECK();
  }
  const size_t smem3 = pl::nms_smem_bytes(a->cap);
  MMD_SMEM(pl::pl_nms_kernel, smem3);
  pl::pl_nms_kernel<<<dim3(a->B, a->T), pl::kNmsThreads, smem3, s>>>(p);
  MMD_LAUNCH_CHECK();
  MMD_SMEM(pl::pl_merge_kernel, smem4);
  pl::pl_merge_kernel<<<a->B, pl::kNmsThreads, smem4, s>>>(p);
  MMD_LAUNCH_CHECK();
  return 0;
}
