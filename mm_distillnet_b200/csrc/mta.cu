// MTA multi-teacher alignment loss — forward and backward kernels (sm_100a).
// Reference semantics: src/loss/MTALoss.py:15-77 (see include/mmd.h).  All of this is HBM-bound streaming work:
//   mta_pool   reads every feature map once (coalesced 16-byte / 8-byte channel-vector loads, shuffle reduce
//              over the channel dimension) and writes one float per pixel;
//   mta_level  one CTA per (level, sample): norms, teacher product, two softmaxes, the KL-style sum and
//              d loss / d a in a handful of block reductions over data that sits in L2;
//   mta_bwd    re-reads f_s once and writes grad = go * (p/C) * f^(p-1) * (d loss / d a).
#include <stdlib.h>

#include "common.cuh"

namespace mmd {

constexpr int kMaxSeg = (1 + MMD_MTA_MAX_TEACHERS) * MMD_MTA_MAX_LEVELS;

struct MtaSeg {
  const void* f;  // feature map
  float* a;       // pooled map [B*HW]
  int npix;       // B*HW
  int HW;
};
struct MtaPoolP {
  static constexpr int kU = 8;   // pixel-chunks in flight per warp (NHWC kernel)
  MtaSeg seg[kMaxSeg];
  int item_begin[kMaxSeg + 1];   // NHWC kernel: prefix sums of the items (kU chunks each) per map
  int nseg;
  int C;
  int group;  // lanes cooperating on one pixel (power of two <= 32)
  float p;
};

__device__ __forceinline__ float powp(float v, float p, bool p_is_2) { return p_is_2 ? v * v : powf(v, p); }

// VEC consecutive channels: VEC = 4 (fp32: 16 B, bf16: 8 B per lane) or 8 (bf16 only: 16 B per lane).  The load and the
// reduction are separate so that a warp can have all the loads of an item in flight before it consumes the first one.
template <typename T, int VEC>
__device__ __forceinline__ uint4 ld_raw(const T* p) {
  if constexpr (VEC == 8) {
    return __ldg(reinterpret_cast<const uint4*>(p));
  } else if constexpr (sizeof(T) == 4) {
    return __ldg(reinterpret_cast<const uint4*>(p));
  } else {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    return make_uint4(r.x, r.y, 0u, 0u);
  }
}
template <typename T, int VEC>
__device__ __forceinline__ float pow_sum_raw(const uint4 r, float pw, bool p2) {
  if constexpr (sizeof(T) == 4) {
    return powp(__uint_as_float(r.x), pw, p2) + powp(__uint_as_float(r.y), pw, p2) + powp(__uint_as_float(r.z), pw, p2) +
           powp(__uint_as_float(r.w), pw, p2);
  } else {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < VEC / 2; ++e) {
      const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
      a += powp(lo, pw, p2) + powp(hi, pw, p2);
    }
    return a;
  }
}
template <typename T, int VEC>
__device__ __forceinline__ float pow_sum(const T* p, float pw, bool p2) {
  return pow_sum_raw<T, VEC>(ld_raw<T, VEC>(p), pw, p2);
}

// Persistent streaming kernel over ONE flattened work list (all feature maps of the call): an item is U consecutive
// pixel-chunks of one map (a chunk = the 32/G pixels one warp-wide load covers), items are dealt round-robin to the
// warps of a grid sized to the machine, so every warp keeps U independent 16-byte loads per lane in flight for several
// iterations instead of one short-lived block per 256 pixels.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) mta_pool_nhwc(const __grid_constant__ MtaPoolP P) {
  const int C = P.C, NQ = C / VEC, G = P.group, PPW = 32 / G;
  const bool p2 = (P.p == 2.0f);
  const int lane = threadIdx.x & 31, gl = lane % G, sub = lane / G;
  const int wpb = blockDim.x >> 5;
  const int nw = gridDim.x * wpb;
  const float invC = 1.0f / (float)C;
  constexpr int U = MtaPoolP::kU;
  const int total = P.item_begin[P.nseg];
  for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < total; item += nw) {
    int si = 0;
    for (int k = 1; k < P.nseg; ++k)
      if (item >= P.item_begin[k]) si = k;
    const MtaSeg& s = P.seg[si];
    const T* __restrict__ f = reinterpret_cast<const T*>(s.f);
    const long long chunk = (long long)(item - P.item_begin[si]) * U;
    float acc[U];
    if (NQ <= G) {   // one vector per lane and pixel (C <= 32 * VEC): all U loads are issued before the first use
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pix = (chunk + u) * PPW + sub;
        raw[u] = make_uint4(0u, 0u, 0u, 0u);
        if (pix < s.npix && gl < NQ) raw[u] = ld_raw<T, VEC>(f + pix * C + VEC * gl);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = (gl < NQ) ? pow_sum_raw<T, VEC>(raw[u], P.p, p2) : 0.f;   // 0^p of a lane without data
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = 0.f;
        long long pix = (chunk + u) * PPW + sub;
        if (pix < s.npix) {
          const T* row = f + pix * C;
          for (int q = gl; q < NQ; q += G) acc[u] += pow_sum<T, VEC>(row + VEC * q, P.p, p2);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float a = acc[u];
      for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      long long pix = (chunk + u) * PPW + sub;
      if (gl == 0 && pix < s.npix) s.a[pix] = a * invC;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mta_pool_nchw(const __grid_constant__ MtaPoolP P) {
  const MtaSeg s = P.seg[blockIdx.y];
  const T* __restrict__ f = reinterpret_cast<const T*>(s.f);
  const int C = P.C;
  const bool p2 = (P.p == 2.0f);
  const float invC = 1.0f / (float)C;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < s.npix;
       pix += (long long)gridDim.x * blockDim.x) {
    int b = (int)(pix / s.HW), i = (int)(pix % s.HW);
    const T* p0 = f + (long long)b * C * s.HW + i;
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) acc += powp(ld1<T>(p0 + (long long)c * s.HW), P.p, p2);
    s.a[pix] = acc * invC;
  }
}

// ---- per (level, sample) loss ------------------------------------------------------------------------------
struct MtaLevelP {
  const float* att;   // [(1+nt)][Btot]
  float* ga;          // [Btot] or null
  float* loss_b;      // [n_levels][B]
  long long Btot;     // B * sum HW
  int cum[MMD_MTA_MAX_LEVELS];  // prefix sum of HW
  int HW[MMD_MTA_MAX_LEVELS];
  int B, nt;
  float T;
  int separate;       // 1: blockIdx.y = independent single-teacher call (that teacher alone), see MmdMtaArgs.separate
  int n_levels;
};

constexpr int kLvlThreads = 512;

template <int K>
__device__ __forceinline__ void block_reduce(double (&v)[K], unsigned max_mask, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool mx = (max_mask >> k) & 1u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = mx ? fmax(v[k], x) : v[k] + x;
    }
  }
  __syncthreads();  // s_red may still be read from a previous call
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_red[warp * K + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double r = s_red[k];
    for (int w = 1; w < nw; ++w) {
      const double x = s_red[w * K + k];
      r = ((max_mask >> k) & 1u) ? fmax(r, x) : r + x;
    }
    v[k] = r;   // identical on every thread (fixed summation order -> deterministic)
  }
}

// All arithmetic after the channel pooling is done in double: the maps are tiny (B * 12 276 floats for D2) and both
// the loss (= -ln N - 1/N + O(1e-4)) and its gradient (s * (g - <g,s>)) are differences of nearly equal numbers.
__global__ void __launch_bounds__(kLvlThreads) mta_level_kernel(const __grid_constant__ MtaLevelP P) {
  __shared__ double s_red[32 * 5];
  // grid (sample, call, level): the CTAs of the largest level (level 0 of a pyramid) are scheduled first
  const int b = blockIdx.x, l = blockIdx.z, n = P.HW[l], call = blockIdx.y;
  const int nt = P.separate ? 1 : P.nt, t0 = P.separate ? call : 0;
  const long long off = (long long)P.B * P.cum[l] + (long long)b * n;
  const float* __restrict__ as = P.att + off;
  const float* __restrict__ at[MMD_MTA_MAX_TEACHERS];
#pragma unroll
  for (int k = 0; k < MMD_MTA_MAX_TEACHERS; ++k) at[k] = P.att + (long long)(1 + t0 + (k < nt ? k : 0)) * P.Btot + off;
  const double invT = 1.0 / (double)P.T;

  // pass 1: squared L2 norms of the student map and of every teacher map (F.normalize, eps 1e-12)
  double ss[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = as[i];
    ss[0] += x * x;
#pragma unroll
    for (int k = 0; k < MMD_MTA_MAX_TEACHERS; ++k)
      if (k < nt) {
        const double y = at[k][i];
        ss[1 + k] += y * y;
      }
  }
  block_reduce<5>(ss, 0u, s_red);
  const double nrm_s_raw = sqrt(ss[0]);
  const double inv_s = 1.0 / fmax(nrm_s_raw, 1e-12);
  double inv_t[MMD_MTA_MAX_TEACHERS];
#pragma unroll
  for (int k = 0; k < MMD_MTA_MAX_TEACHERS; ++k) inv_t[k] = 1.0 / fmax(sqrt(ss[1 + k]), 1e-12);

  auto teacher_m = [&](int i) -> double {  // product of the normalised teacher attentions (MTALoss.py:51-55)
    double m = at[0][i] * inv_t[0];
#pragma unroll
    for (int k = 1; k < MMD_MTA_MAX_TEACHERS; ++k)
      if (k < nt) m *= at[k][i] * inv_t[k];
    return m;
  };

  // pass 2: L1 norm of the teacher product (only used when nt > 1), maxima for the two softmaxes
  double r2[3] = {0.0, -INFINITY, -INFINITY};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = teacher_m(i);
    r2[0] += fabs(m);
    r2[1] = fmax(r2[1], as[i] * inv_s);
    r2[2] = fmax(r2[2], m);
  }
  block_reduce<3>(r2, 0x6u, s_red);
  const double inv_l1 = (nt > 1) ? 1.0 / fmax(r2[0], 1e-12) : 1.0;  // F.normalize(p=1) (MTALoss.py:57)
  const double max_zs = r2[1] * invT, max_zt = r2[2] * inv_l1 * invT;

  const double invB = 1.0 / (double)P.B;
  constexpr int KC = 18;   // cached variant: <= 18 positions per thread (D2: P3 has 96*96 = 18 * 512 positions)
  if (n <= KC * kLvlThreads && blockDim.x == kLvlThreads) {
    // passes 3-6 with the two exponentials of every position kept in registers (8 -> 2 double exp per position)
    double es[KC], et[KC];
    double z[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int i = threadIdx.x + k * kLvlThreads;
      es[k] = et[k] = 0.0;
      if (i < n) {
        es[k] = exp(as[i] * inv_s * invT - max_zs);
        et[k] = exp(teacher_m(i) * inv_l1 * invT - max_zt);
        z[0] += es[k];
        z[1] += et[k];
      }
    }
    block_reduce<2>(z, 0u, s_red);
    const double inv_zs = 1.0 / z[0], inv_zt = 1.0 / z[1], log_zt = log(z[1]);
    double r4[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int i = threadIdx.x + k * kLvlThreads;
      if (i < n) {
        const double s = es[k] * inv_zs, t = et[k] * inv_zt;
        const double zt = teacher_m(i) * inv_l1 * invT - max_zt;
        r4[0] += t * ((zt - log_zt) - s);
        r4[1] += -t * invB * s;
      }
    }
    block_reduce<2>(r4, 0u, s_red);
    if (threadIdx.x == 0) P.loss_b[(call * P.n_levels + l) * P.B + b] = (float)r4[0];
    if (P.ga == nullptr) return;
    const double gs = r4[1];
    double r5[1] = {0.0};
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int i = threadIdx.x + k * kLvlThreads;
      if (i < n) {
        const double ah = as[i] * inv_s;
        const double dah = (es[k] * inv_zs) * (-(et[k] * inv_zt) * invB - gs) * invT;
        es[k] = dah;   // d a^ replaces the exponential
        r5[0] += ah * dah;
      }
    }
    block_reduce<1>(r5, 0u, s_red);
    const bool clamped = !(nrm_s_raw > 1e-12);
    float* __restrict__ ga = P.ga + (long long)call * P.Btot + off;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int i = threadIdx.x + k * kLvlThreads;
      if (i < n) {
        const double ah = as[i] * inv_s;
        ga[i] = (float)(clamped ? es[k] * inv_s : (es[k] - ah * r5[0]) * inv_s);
      }
    }
    return;
  }

  // pass 3: softmax denominators
  double z[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    z[0] += exp(as[i] * inv_s * invT - max_zs);
    z[1] += exp(teacher_m(i) * inv_l1 * invT - max_zt);
  }
  block_reduce<2>(z, 0u, s_red);
  const double inv_zs = 1.0 / z[0], inv_zt = 1.0 / z[1], log_zt = log(z[1]);

  // pass 4: loss_b = sum t (log t - s)   (kl_div with a PROBABILITY input: MTALoss.py:62-72), and <g, s>
  double r4[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = exp(as[i] * inv_s * invT - max_zs) * inv_zs;
    const double zt = teacher_m(i) * inv_l1 * invT - max_zt;
    const double t = exp(zt) * inv_zt;
    r4[0] += t * ((zt - log_zt) - s);
    r4[1] += -t * invB * s;
  }
  block_reduce<2>(r4, 0u, s_red);
  if (threadIdx.x == 0) P.loss_b[(call * P.n_levels + l) * P.B + b] = (float)r4[0];
  if (P.ga == nullptr) return;

  // pass 5: <a^, d a^> with d z = s (g - <g,s>), d a^ = d z / T
  const double gs = r4[1];
  double r5[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ah = as[i] * inv_s;
    const double s = exp(ah * invT - max_zs) * inv_zs;
    const double t = exp(teacher_m(i) * inv_l1 * invT - max_zt) * inv_zt;
    const double dah = s * (-t * invB - gs) * invT;
    r5[0] += ah * dah;
  }
  block_reduce<1>(r5, 0u, s_red);
  const bool clamped = !(nrm_s_raw > 1e-12);
  // pass 6: d a = (d a^ - a^ <a^, d a^>) / ||a||  (or d a^ / eps when the norm was clamped)
  float* __restrict__ ga = P.ga + (long long)call * P.Btot + off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ah = as[i] * inv_s;
    const double s = exp(ah * invT - max_zs) * inv_zs;
    const double t = exp(teacher_m(i) * inv_l1 * invT - max_zt) * inv_zt;
    const double dah = s * (-t * invB - gs) * invT;
    ga[i] = (float)(clamped ? dah * inv_s : (dah - ah * r5[0]) * inv_s);
  }
}

__global__ void mta_finish_kernel(const float* __restrict__ loss_b, float* __restrict__ loss, int B) {
  const int l = blockIdx.x;
  double acc = 0.0;  // fixed order -> deterministic
  for (int b = threadIdx.x; b < B; b += 32) acc += (double)loss_b[l * B + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) loss[l] = (float)(acc / (double)B);
}

// ---- backward ----------------------------------------------------------------------------------------------
struct MtaBwdSeg {
  const void* f;
  void* g;
  const float* ga;
  int npix, HW;
  int level, pad;
};
struct MtaBwdP {
  MtaBwdSeg seg[MMD_MTA_MAX_LEVELS];
  const float* grad_loss;   // [ncalls][n_levels]
  long long Btot;           // stride between the ga maps of consecutive calls
  long long vec_begin[MMD_MTA_MAX_LEVELS + 1];   // bf16x8 kernel: prefix sums of the 16-byte vectors per level
  int C, ncalls, n_levels;
  float p;
};

// per-pixel factor: sum over the calls of grad_loss[call][level] * (p / C) * d loss / d a
__device__ __forceinline__ float bwd_coef(const MtaBwdP& P, const MtaBwdSeg& s, long long pix, const float (&coef)[MMD_MTA_MAX_TEACHERS]) {
  float k = coef[0] * s.ga[pix];
#pragma unroll
  for (int c = 1; c < MMD_MTA_MAX_TEACHERS; ++c)
    if (c < P.ncalls) k = fmaf(coef[c], s.ga[(long long)c * P.Btot + pix], k);
  return k;
}

// bf16 NHWC, p == 2, C % 8 == 0: 16-byte vectors, one flattened index space over all levels (persistent grid, four
// independent vectors per thread and iteration)
__global__ void __launch_bounds__(256) mta_bwd_bf16x8_kernel(const __grid_constant__ MtaBwdP P) {
  const int NV = P.C >> 3;
  const long long total = P.vec_begin[P.n_levels];
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int UB = 4;
  for (long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += UB * stride) {
    uint4 x[UB];
    float k[UB];
    long long loc[UB];
    int lv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const long long v = v0 + u * stride;
      lv[u] = -1;
      if (v < total) {
        int l = 0;
        for (int j = 1; j < P.n_levels; ++j)
          if (v >= P.vec_begin[j]) l = j;
        lv[u] = l;
        loc[u] = v - P.vec_begin[l];
        x[u] = __ldg(reinterpret_cast<const uint4*>(P.seg[l].f) + loc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (lv[u] < 0) continue;
      const MtaBwdSeg& s = P.seg[lv[u]];
      const long long pix = loc[u] / NV;
      float kk = 0.f;
      for (int c = 0; c < P.ncalls; ++c)
        kk = fmaf(P.grad_loss[c * P.n_levels + s.level], s.ga[(long long)c * P.Btot + pix], kk);
      k[u] = kk * P.p / (float)P.C;
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (lv[u] < 0) continue;
      const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        __nv_bfloat162 h = __floats2bfloat162_rn(k[u] * lo, k[u] * hi);
        o[e] = *reinterpret_cast<uint32_t*>(&h);
      }
      reinterpret_cast<uint4*>(P.seg[lv[u]].g)[loc[u]] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename T, bool NCHW>
__global__ void __launch_bounds__(256) mta_bwd_kernel(const __grid_constant__ MtaBwdP P) {
  const MtaBwdSeg s = P.seg[blockIdx.y];
  const T* __restrict__ f = reinterpret_cast<const T*>(s.f);
  T* __restrict__ g = reinterpret_cast<T*>(s.g);
  const int C = P.C;
  const bool p2 = (P.p == 2.0f);
  float coef[MMD_MTA_MAX_TEACHERS];
#pragma unroll
  for (int c = 0; c < MMD_MTA_MAX_TEACHERS; ++c)
    coef[c] = (c < P.ncalls) ? P.grad_loss[c * P.n_levels + s.level] * P.p / (float)C : 0.f;
  const long long nvec = (long long)s.npix * C / 4;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    long long e = v * 4;
    if (!NCHW) {
      float k = bwd_coef(P, s, e / C, coef);
      float4 x = ld4<T>(f + e);
      float4 r;
      if (p2) r = f4_scale(x, k);
      else r = make_float4(k * powf(x.x, P.p - 1.f), k * powf(x.y, P.p - 1.f), k * powf(x.z, P.p - 1.f), k * powf(x.w, P.p - 1.f));
      st4<T>(g + e, r);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        long long ee = e + j;
        int hw = (int)(ee % s.HW);
        int b = (int)(ee / ((long long)C * s.HW));
        float k = bwd_coef(P, s, (long long)b * s.HW + hw, coef);
        float x = ld1<T>(f + ee);
        st1<T>(g + ee, p2 ? k * x : k * powf(x, P.p - 1.f));
      }
    }
  }
}

// ---- fast path of the D2 product configuration: bf16, NHWC, C = 112, p = 2 ---------------------------------------------
// A pixel is 14 16-byte vectors (224 B).  The generic kernel gives 16 lanes to a pixel (14 busy), reduces each pixel with
// four shuffle steps and looks its segment up per item: ~0.11 warp instructions per byte, i.e. issue-bound at ~half of the
// HBM rate (ncu: profiles/r2_ncu_summary.md).  Here a warp owns SUPER-CHUNKS of 16 pixels = 224 vectors = 7 fully coalesced
// warp-wide loads (all lanes busy), two super-chunks per iteration (14 independent 16-byte loads per lane in flight); the
// per-vector partial sums go through shared memory and lanes 0-15 / 16-31 each add the 14 partials of one pixel of
// super-chunk 0 / 1 (bank = (14 * lane + k) mod 32: conflict-free).  ~0.06 instructions per byte.
constexpr int kScPix = 16, kScVec = kScPix * 14;   // 224
struct MtaFastSeg {
  const uint4* f;
  float* a;
  int npix;
};
struct MtaPoolFastP {
  MtaFastSeg seg[kMaxSeg];
  int sc_begin[kMaxSeg + 1];   // prefix sums of the super-chunks per map
  int nseg;
};

__device__ __forceinline__ float sumsq8(const uint4 r) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
    a = fmaf(lo, lo, a);
    b = fmaf(hi, hi, b);
  }
  return a + b;
}

__global__ void __launch_bounds__(256) mta_pool_c112_kernel(const __grid_constant__ MtaPoolFastP P) {
  __shared__ float s_part[8][2][kScVec];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = gridDim.x * 8;
  const int total = P.sc_begin[P.nseg];
  const int u_of_lane = lane >> 4, pl = lane & 15;
  for (int sc0 = 2 * (blockIdx.x * 8 + warp); sc0 < total; sc0 += 2 * nw) {
    int si[2] = {0, 0};
    for (int k = 1; k < P.nseg; ++k) {
      const int b = P.sc_begin[k];
      if (sc0 >= b) si[0] = k;
      if (sc0 + 1 >= b) si[1] = k;
    }
    uint4 raw[2][7];
    int local[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const MtaFastSeg& s = P.seg[si[u]];
      local[u] = sc0 + u - P.sc_begin[si[u]];
      const int nvec = (sc0 + u < total) ? s.npix * 14 : 0;
      const int base = local[u] * kScVec;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int v = base + 32 * j + lane;
        raw[u][j] = make_uint4(0u, 0u, 0u, 0u);
        if (v < nvec) raw[u][j] = __ldg(s.f + v);
      }
    }
    __syncwarp();   // (scheduling fence: all 14 loads are issued before the first one is consumed)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int j = 0; j < 7; ++j) s_part[warp][u][32 * j + lane] = sumsq8(raw[u][j]);
    __syncwarp();
    {
      const float* q = &s_part[warp][u_of_lane][14 * pl];
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 14; ++k) a += q[k];
      const MtaFastSeg& s = P.seg[u_of_lane ? si[1] : si[0]];
      const int pix = (u_of_lane ? local[1] : local[0]) * kScPix + pl;
      if (sc0 + u_of_lane < total && pix < s.npix) s.a[pix] = a * (1.0f / 112.0f);
    }
    __syncwarp();
  }
}

struct MtaBwdFastSeg {
  const uint4* f;
  uint4* g;
  const float* ga;   // d loss / d a of call 0 for this level; call c at + c * Btot
  int npix;
  int level;
};
struct MtaBwdFastP {
  MtaBwdFastSeg seg[MMD_MTA_MAX_LEVELS];
  int sc_begin[MMD_MTA_MAX_LEVELS + 1];
  const float* grad_loss;   // [ncalls][n_levels]
  long long Btot;
  int n_levels, ncalls;
  float scale;              // p / C
};

__global__ void __launch_bounds__(256) mta_bwd_c112_kernel(const __grid_constant__ MtaBwdFastP P) {
  __shared__ float s_k[8][2][kScPix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = gridDim.x * 8;
  const int total = P.sc_begin[P.n_levels];
  const int u_of_lane = lane >> 4, pl = lane & 15;
  for (int sc0 = 2 * (blockIdx.x * 8 + warp); sc0 < total; sc0 += 2 * nw) {
    int si[2] = {0, 0};
    for (int k = 1; k < P.n_levels; ++k) {
      const int b = P.sc_begin[k];
      if (sc0 >= b) si[0] = k;
      if (sc0 + 1 >= b) si[1] = k;
    }
    uint4 raw[2][7];
    int base[2], nvec[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const MtaBwdFastSeg& s = P.seg[si[u]];
      nvec[u] = (sc0 + u < total) ? s.npix * 14 : 0;
      base[u] = (sc0 + u - P.sc_begin[si[u]]) * kScVec;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int v = base[u] + 32 * j + lane;
        raw[u][j] = make_uint4(0u, 0u, 0u, 0u);
        if (v < nvec[u]) raw[u][j] = __ldg(s.f + v);
      }
    }
    {   // per-pixel factor: sum over the calls of grad_loss[call][level] * (p / C) * d loss / d a
      const MtaBwdFastSeg& s = P.seg[u_of_lane ? si[1] : si[0]];
      const int pix = ((u_of_lane ? base[1] : base[0]) / 14) + pl;
      float k = 0.f;
      if (sc0 + u_of_lane < total && pix < s.npix) {
        for (int c = 0; c < P.ncalls; ++c)
          k = fmaf(P.grad_loss[c * P.n_levels + s.level], s.ga[(long long)c * P.Btot + pix], k);
      }
      s_k[warp][u_of_lane][pl] = k * P.scale;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      uint4* g = P.seg[si[u]].g;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int vl = 32 * j + lane;                       // vector inside the super-chunk
        const float k = s_k[warp][u][(vl * 4682) >> 16];    // vl / 14 for vl < 224
        const float2 k2 = make_float2(k, k);
        const uint32_t w[4] = {raw[u][j].x, raw[u][j].y, raw[u][j].z, raw[u][j].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
          const float2 r = __fmul2_rn(k2, x);
          __nv_bfloat162 h = __floats2bfloat162_rn(r.x, r.y);
          o[e] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (base[u] + vl < nvec[u]) g[base[u] + vl] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncwarp();
  }
}

static int g_mta_fast = -1;   // MMD_NO_MTA_FAST=1 / mmd_set_option("mta_fast", 0): A/B switch back to the generic kernels
void set_mta_fast(int on) { g_mta_fast = on ? 1 : 0; }
static bool mta_fast_enabled() {
  if (g_mta_fast < 0) {
    const char* e = getenv("MMD_NO_MTA_FAST");
    g_mta_fast = (e && e[0] == '1') ? 0 : 1;
  }
  return g_mta_fast == 1;
}

static int group_lanes(int C, int vec) {
  int nq = C / vec, g = 1;
  while (g < nq && g < 32) g <<= 1;
  return g;
}

}  // namespace mmd

using namespace mmd;

extern "C" int mmd_mta_fwd(const MmdMtaArgs* a, mmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMD_CHECK_ARG(a != nullptr, "mmd_mta_fwd: null args");
  MMD_CHECK_ARG(a->n_levels >= 1 && a->n_levels <= MMD_MTA_MAX_LEVELS, "mmd_mta_fwd: n_levels=%d out of range", a->n_levels);
  MMD_CHECK_ARG(a->n_teachers >= 1 && a->n_teachers <= MMD_MTA_MAX_TEACHERS, "mmd_mta_fwd: n_teachers=%d out of range", a->n_teachers);
  MMD_CHECK_ARG(a->B >= 1 && a->C >= 4 && a->C % 4 == 0, "mmd_mta_fwd: B=%d C=%d (C must be a multiple of 4)", a->B, a->C);
  MMD_CHECK_ARG(a->dtype == MMD_F32 || a->dtype == MMD_BF16, "mmd_mta_fwd: dtype %d", a->dtype);
  MMD_CHECK_ARG(a->layout == MMD_NHWC || a->layout == MMD_NCHW, "mmd_mta_fwd: layout %d", a->layout);
  MMD_CHECK_ARG(a->att_ws && a->loss_b && a->loss, "mmd_mta_fwd: null workspace/output");
  MMD_CHECK_ARG(a->T > 0.f, "mmd_mta_fwd: T must be positive");

  MtaPoolP pp;
  MtaLevelP lp;
  int cum = 0, maxpix = 0;
  for (int l = 0; l < a->n_levels; ++l) {
    MMD_CHECK_ARG(a->H[l] >= 1 && a->W[l] >= 1, "mmd_mta_fwd: level %d has empty spatial size", l);
    lp.cum[l] = cum;
    lp.HW[l] = a->H[l] * a->W[l];
    cum += lp.HW[l];
  }
  const long long Btot = (long long)a->B * cum;
  int nseg = 0;
  for (int t = 0; t <= a->n_teachers; ++t)
    for (int l = 0; l < a->n_levels; ++l) {
      const void* f = (t == 0) ? a->fs[l] : a->ft[t - 1][l];
      MMD_CHECK_ARG(f != nullptr, "mmd_mta_fwd: null feature pointer (tensor %d level %d)", t, l);
      MtaSeg& s = pp.seg[nseg++];
      s.f = f;
      s.a = a->att_ws + (long long)t * Btot + (long long)a->B * lp.cum[l];
      s.npix = a->B * lp.HW[l];
      s.HW = lp.HW[l];
      if (s.npix > maxpix) maxpix = s.npix;
    }
  const bool vec8 = (a->dtype == MMD_BF16 && a->layout == MMD_NHWC && a->C % 8 == 0);
  for (int i = 0; vec8 && i < nseg; ++i)
    if (((uintptr_t)pp.seg[i].f & 15u) != 0) { set_error("mmd_mta_fwd: bf16 feature maps must be 16-byte aligned"); return MMD_E_ARG; }
  pp.C = a->C;
  pp.group = group_lanes(a->C, vec8 ? 8 : 4);
  pp.p = a->p;

  double pool_bytes = 0.0;
  for (int i = 0; i < nseg; ++i) pool_bytes += (double)pp.seg[i].npix * a->C * (a->dtype == MMD_F32 ? 4 : 2);
  {
  ProfScope prof(PK_MTA_POOL, pool_bytes, stream);
  const bool fast = vec8 && a->C == 112 && a->p == 2.0f && mta_fast_enabled();
  if (fast) {
    MtaPoolFastP fp;
    int tot = 0;
    for (int i = 0; i < nseg; ++i) {
      fp.seg[i].f = reinterpret_cast<const uint4*>(pp.seg[i].f);
      fp.seg[i].a = pp.seg[i].a;
      fp.seg[i].npix = pp.seg[i].npix;
      fp.sc_begin[i] = tot;
      tot += (pp.seg[i].npix + kScPix - 1) / kScPix;
    }
    for (int i = nseg; i <= kMaxSeg; ++i) fp.sc_begin[i] = tot;
    for (int i = nseg; i < kMaxSeg; ++i) fp.seg[i] = fp.seg[0];
    fp.nseg = nseg;
    int gx = (tot + 15) / 16;            // 8 warps x 2 super-chunks per block and iteration
    if (gx > 148 * 8) gx = 148 * 8;
    mta_pool_c112_kernel<<<gx, 256, 0, stream>>>(fp);
  } else if (a->layout == MMD_NHWC) {
    const int ppi = (32 / pp.group) * MtaPoolP::kU;   // pixels per item
    int items = 0;
    for (int i = 0; i < nseg; ++i) {
      pp.item_begin[i] = items;
      items += (pp.seg[i].npix + ppi - 1) / ppi;
    }
    pp.item_begin[nseg] = items;
    pp.nseg = nseg;
    int gx = (items + 7) / 8;
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid(gx);
    if (a->dtype == MMD_F32) mta_pool_nhwc<float, 4><<<grid, 256, 0, stream>>>(pp);
    else if (vec8) mta_pool_nhwc<__nv_bfloat16, 8><<<grid, 256, 0, stream>>>(pp);
    else mta_pool_nhwc<__nv_bfloat16, 4><<<grid, 256, 0, stream>>>(pp);
  } else {
    int gx = (maxpix + 255) / 256;
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid(gx, nseg);
    if (a->dtype == MMD_F32) mta_pool_nchw<float><<<grid, 256, 0, stream>>>(pp);
    else mta_pool_nchw<__nv_bfloat16><<<grid, 256, 0, stream>>>(pp);
  }
  }
  MMD_LAUNCH_CHECK();

  lp.att = a->att_ws;
  lp.ga = a->ga_ws;
  lp.loss_b = a->loss_b;
  lp.Btot = Btot;
  lp.B = a->B;
  lp.nt = a->n_teachers;
  lp.T = a->T;
  lp.separate = a->separate ? 1 : 0;
  lp.n_levels = a->n_levels;
  const int ncalls = a->separate ? a->n_teachers : 1;
  {
    ProfScope prof(PK_MTA_LEVEL, 0.0, stream);
    mta_level_kernel<<<dim3(a->B, ncalls, a->n_levels), kLvlThreads, 0, stream>>>(lp);
  }
  MMD_LAUNCH_CHECK();
  {
    ProfScope prof(PK_MTA_FINISH, 0.0, stream);
    mta_finish_kernel<<<a->n_levels * ncalls, 32, 0, stream>>>(a->loss_b, a->loss, a->B);
  }
  MMD_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmd_mta_bwd(const MmdMtaArgs* a, const float* grad_loss, void* const* grad_fs, mmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMD_CHECK_ARG(a != nullptr && grad_loss != nullptr && grad_fs != nullptr, "mmd_mta_bwd: null args");
  MMD_CHECK_ARG(a->n_levels >= 1 && a->n_levels <= MMD_MTA_MAX_LEVELS, "mmd_mta_bwd: n_levels=%d out of range", a->n_levels);
  MMD_CHECK_ARG(a->ga_ws != nullptr, "mmd_mta_bwd: forward was run without ga_ws");
  MMD_CHECK_ARG(a->C % 4 == 0, "mmd_mta_bwd: C must be a multiple of 4");
  MtaBwdP bp;
  int cum = 0;
  long long maxvec = 0;
  for (int l = 0; l < a->n_levels; ++l) {
    MMD_CHECK_ARG(grad_fs[l] != nullptr && a->fs[l] != nullptr, "mmd_mta_bwd: null pointer at level %d", l);
    MtaBwdSeg& s = bp.seg[l];
    s.f = a->fs[l];
    s.g = grad_fs[l];
    s.HW = a->H[l] * a->W[l];
    s.npix = a->B * s.HW;
    s.ga = a->ga_ws + (long long)a->B * cum;
    s.level = l;
    cum += s.HW;
    long long nv = (long long)s.npix * a->C / 4;
    if (nv > maxvec) maxvec = nv;
  }
  bp.grad_loss = grad_loss;
  bp.C = a->C;
  bp.p = a->p;
  bp.ncalls = a->separate ? a->n_teachers : 1;
  bp.n_levels = a->n_levels;
  bp.Btot = (long long)a->B * cum;
  bool vec8 = (a->dtype == MMD_BF16 && a->layout == MMD_NHWC && a->C % 8 == 0 && a->p == 2.0f);
  for (int l = 0; vec8 && l < a->n_levels; ++l)
    if ((((uintptr_t)bp.seg[l].f | (uintptr_t)bp.seg[l].g) & 15u) != 0) vec8 = false;
  long long gx = (maxvec + 255) / 256;
  if (gx > 148 * 16) gx = 148 * 16;
  dim3 grid((unsigned)gx, a->n_levels);
  if (vec8) {
    long long tot = 0;
    for (int l = 0; l < a->n_levels; ++l) {
      bp.vec_begin[l] = tot;
      tot += (long long)bp.seg[l].npix * (a->C / 8);
    }
    bp.vec_begin[a->n_levels] = tot;
    for (int l = a->n_levels + 1; l <= MMD_MTA_MAX_LEVELS; ++l) bp.vec_begin[l] = tot;
    gx = (tot + 4 * 256 - 1) / (4 * 256);
    if (gx > 148 * 8) gx = 148 * 8;
    grid = dim3((unsigned)gx);
  }
  double bwd_bytes = 0.0;
  for (int l = 0; l < a->n_levels; ++l) bwd_bytes += 2.0 * bp.seg[l].npix * a->C * (a->dtype == MMD_F32 ? 4 : 2);
  ProfScope prof(PK_MTA_BWD, bwd_bytes, stream);
  if (vec8 && a->C == 112 && mta_fast_enabled()) {
    MtaBwdFastP fp;
    int tot = 0;
    for (int l = 0; l < a->n_levels; ++l) {
      fp.seg[l].f = reinterpret_cast<const uint4*>(bp.seg[l].f);
      fp.seg[l].g = reinterpret_cast<uint4*>(bp.seg[l].g);
      fp.seg[l].ga = bp.seg[l].ga;
      fp.seg[l].npix = bp.seg[l].npix;
      fp.seg[l].level = l;
      fp.sc_begin[l] = tot;
      tot += (bp.seg[l].npix + kScPix - 1) / kScPix;
    }
    for (int l = a->n_levels; l <= MMD_MTA_MAX_LEVELS; ++l) fp.sc_begin[l] = tot;
    for (int l = a->n_levels; l < MMD_MTA_MAX_LEVELS; ++l) fp.seg[l] = fp.seg[0];
    fp.grad_loss = grad_loss;
    fp.Btot = bp.Btot;
    fp.n_levels = a->n_levels;
    fp.ncalls = bp.ncalls;
    fp.scale = a->p / (float)a->C;
    int gxf = (tot + 15) / 16;
    if (gxf > 148 * 8) gxf = 148 * 8;
    mta_bwd_c112_kernel<<<gxf, 256, 0, stream>>>(fp);
  } else if (a->layout == MMD_NHWC) {
    if (a->dtype == MMD_F32) mta_bwd_kernel<float, false><<<grid, 256, 0, stream>>>(bp);
    else if (vec8) mta_bwd_bf16x8_kernel<<<grid, 256, 0, stream>>>(bp);
    else mta_bwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(bp);
  } else {
    if (a->dtype == MMD_F32) mta_bwd_kernel<float, true><<<grid, 256, 0, stream>>>(bp);
    else mta_bwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(bp);
  }
  MMD_LAUNCH_CHECK();
  return 0;
}
