// BiFPN backward kernels (sm_100a).  Gather ("pull") formulation: every node writes exactly one activation-sized
// tensor, dL/du (gradient of its pre-swish fused sum); the producer of each input later GATHERS its gradient from
// the dL/du of its (<= 3) consumers, undoing the resampling on load:
//     SAME : w * du[y, x]          UP2 : w * sum of the 2x2 block          POOL : w * du[window] for the windows whose
//     recorded arg-max is this element (indices saved by the forward, overlapping windows accumulate).
// There are no atomics on activations and no read-modify-write of gradient maps.  Train-mode BatchNorm needs
// sum(G) and sum(G*xhat) over the whole batch before any dL/dx can be formed; the CONSUMER accumulates them per
// input edge while it has du and the input in hand ("slots", double atomics), so no extra reduction pass exists.
//
//   node_bwd_a : G gather -> BN backward (dy = A*G + Bc*y + Cc) -> dL/dd = dy * W_pw (FFMA GEMM), dW_pw += dy^T d,
//                db += sum dy   (d = depthwise output saved by the forward)
//   node_bwd_b : dL/dv = depthwise^T(dL/dd) (flipped taps over a halo tile), dK_dw, rebuild u from the inputs,
//                dL/du = dL/dv * swish'(u) -> written once; per-input slots; the last CTA forms the fusion-weight
//                gradient (relu / normalise backward, src/YetAnotherEfficientDet.py:338-339).
//   proj_bwd   : first-cell projections; pull_kernel / slot_kernel: P6/P7 synthesis and the stack boundary.
#include <stdlib.h>

#include "bifpn_bwd_common.cuh"

namespace mmd {

// ---- node backward, part A -----------------------------------------------------------------------------------
template <int C>
struct BwdASmem {
  static constexpr int LDD = C + 4;
  static constexpr int kGy = kTileP * LDD;
  static constexpr int kD = kTileP * LDD;
  static constexpr int kW = C * C;
  static constexpr int kCoef = 3 * C;
  static constexpr int kFloats = kGy + kD + kW + kCoef;
};

template <typename T, int C>
__global__ void __launch_bounds__(kThreads, 1) node_bwd_a_kernel(const __grid_constant__ NodeBwdP P) {
  using S = BwdASmem<C>;
  constexpr int NQ = C / 4, NJ = C / 16, LDD = S::LDD;
  extern __shared__ __align__(16) float smem[];
  float* s_gy = smem;
  float* s_d = s_gy + S::kGy;
  float* s_w = s_d + S::kD;
  float* s_coef = s_w + S::kW;

  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const TileGeom g = P.g;
  float cw[3];
  cons_weights(P, cw);
  bn_bwd_coefs<C>(P, cw, s_coef);
  for (int idx = tid; idx < C * C; idx += kThreads) s_w[idx] = P.pw_w[idx];  // native [o][i] is already [k][n] here
  __syncthreads();

  const T* __restrict__ yraw = reinterpret_cast<const T*>(P.out);
  const T* __restrict__ dsave = reinterpret_cast<const T*>(P.save_d);
  T* __restrict__ ddout = reinterpret_cast<T*>(P.dd);

  float accW[NJ][NJ];
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int bb = 0; bb < NJ; ++bb) accW[a][bb] = 0.f;
  float accB = 0.f;

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int b = tile / (g.tiles_x * g.tiles_y);
    const int rem = tile - b * (g.tiles_x * g.tiles_y);
    const int ty0 = (rem / g.tiles_x) * g.TH, tx0 = (rem % g.tiles_x) * g.TW;
    const int th = min(g.TH, g.H - ty0), tw = min(g.TW, g.W - tx0);

    for (int idx = tid; idx < kTileP * NQ; idx += kThreads) {
      const int p = idx / NQ, q = idx - p * NQ;
      const int ty = p / g.TW, tx = p - ty * g.TW;
      float4 gy = f4_zero(), d = f4_zero();
      if (ty < th && tx < tw) {
        const int y = ty0 + ty, x = tx0 + tx;
        const long long off = (((long long)b * g.H + y) * g.W + x) * C + 4 * q;
        const float4 G = pull_grad<T, C>(P, cw, b, y, x, q);
        const float4 yr = ld4<T>(yraw + off);
        const float4 A = *reinterpret_cast<const float4*>(s_coef + 4 * q);
        const float4 Bc = *reinterpret_cast<const float4*>(s_coef + C + 4 * q);
        const float4 Cc = *reinterpret_cast<const float4*>(s_coef + 2 * C + 4 * q);
        gy = f4_fma(A, G, f4_fma(Bc, yr, Cc));
        d = ld4<T>(dsave + off);
      }
      *reinterpret_cast<float4*>(s_gy + p * LDD + 4 * q) = gy;
      *reinterpret_cast<float4*>(s_d + p * LDD + 4 * q) = d;
    }
    __syncthreads();

    // dL/dd[p][i] = sum_o dy[p][o] * W[o][i]
    float acc[8][NJ];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < C; ++k) {
      float a[8], w[NJ];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = s_gy[(tm + 16 * r) * LDD + k];
#pragma unroll
      for (int j = 0; j < NJ; ++j) w[j] = s_w[k * C + tn + 16 * j];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
    }
    // dW[o][i] += sum_p dy[p][o] * d[p][i]   (o = tm + 16a, i = tn + 16bb) ; db[o] += sum_p dy[p][o]
#pragma unroll 2
    for (int p = 0; p < kTileP; ++p) {
      float go[NJ], di[NJ];
#pragma unroll
      for (int a = 0; a < NJ; ++a) go[a] = s_gy[p * LDD + tm + 16 * a];
#pragma unroll
      for (int bb = 0; bb < NJ; ++bb) di[bb] = s_d[p * LDD + tn + 16 * bb];
#pragma unroll
      for (int a = 0; a < NJ; ++a)
#pragma unroll
        for (int bb = 0; bb < NJ; ++bb) accW[a][bb] = fmaf(go[a], di[bb], accW[a][bb]);
    }
    if (tid < C) {
      float s = 0.f;
      for (int p = 0; p < kTileP; ++p) s += s_gy[p * LDD + tid];
      accB += s;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < NJ; ++j) s_d[(tm + 16 * r) * LDD + tn + 16 * j] = acc[r][j];
    __syncthreads();
    for (int idx = tid; idx < g.TH * g.TW * NQ; idx += kThreads) {
      const int p = idx / NQ, q = idx - p * NQ;
      const int ty = p / g.TW, tx = p - ty * g.TW;
      if (ty < th && tx < tw)
        st4<T>(ddout + (((long long)b * g.H + ty0 + ty) * g.W + tx0 + tx) * C + 4 * q,
               *reinterpret_cast<const float4*>(s_d + p * LDD + 4 * q));
    }
    __syncthreads();
  }
  // every thread owns a [NJ][NJ] block of dW (o = tm + 16a, i = tn + 16bb)
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int bb = 0; bb < NJ; ++bb) atomicAdd(P.g_pw + (tm + 16 * a) * C + tn + 16 * bb, accW[a][bb]);
  if (tid < C && P.g_pb) atomicAdd(P.g_pb + tid, accB);
}

// ---- node backward, part B -----------------------------------------------------------------------------------
template <int C>
struct BwdBSmem {
  static constexpr int kDd = kHaloMax * C;
  static constexpr int kV = kHaloMax * C;
  static constexpr int kK = 9 * C;
  static constexpr int kIn = 3 * 4 * C;
  static constexpr int kFloats = kDd + kV + kK + kIn;
};

template <typename T, int C>
__global__ void __launch_bounds__(kThreads, 1) node_bwd_b_kernel(const __grid_constant__ NodeBwdP P) {
  using S = BwdBSmem<C>;
  constexpr int NQ = C / 4;
  constexpr int ROWS = kThreads / NQ;  // 9 rows of 28 channel-quads
  constexpr int NACC = 36 + 12 + 12;   // dK[9][4], S1[3][4], S2[3][4]
  extern __shared__ __align__(16) float smem[];
  float* s_dd = smem;
  float* s_v = s_dd + S::kDd;
  float* s_k = s_v + S::kV;
  float* s_in = s_k + S::kK;  // [input][scale, shift, mean, invstd][C]
  __shared__ int s_flag;
  __shared__ float s_gw[3];

  const int tid = threadIdx.x;
  const TileGeom g = P.g;
  const int HW2 = g.TW + 2;
  const bool active = tid < ROWS * NQ;
  const int q = tid % NQ, prow = tid / NQ;

  float wgt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) wgt[i] = (i < P.n_in) ? fusion_weight(P.fw, P.n_in, i, P.fw_eps) : 0.f;
  for (int idx = tid; idx < 9 * C; idx += kThreads) {
    const int c = idx / 9, tap = idx - c * 9;
    s_k[tap * C + c] = P.dw_w[idx];
  }
  for (int idx = tid; idx < 3 * C; idx += kThreads) {
    const int i = idx / C, c = idx - i * C;
    const float* bn = (i < P.n_in) ? P.in[i].bn : nullptr;
    s_in[(4 * i + 0) * C + c] = bn ? bn[c] : 1.f;
    s_in[(4 * i + 1) * C + c] = bn ? bn[C + c] : 0.f;
    s_in[(4 * i + 2) * C + c] = bn ? bn[2 * C + c] : 0.f;
    s_in[(4 * i + 3) * C + c] = bn ? bn[3 * C + c] : 1.f;
  }
  __syncthreads();

  float4 dK[9], S1[3], S2[3];
#pragma unroll
  for (int t = 0; t < 9; ++t) dK[t] = f4_zero();
#pragma unroll
  for (int i = 0; i < 3; ++i) { S1[i] = f4_zero(); S2[i] = f4_zero(); }

  const T* __restrict__ ddin = reinterpret_cast<const T*>(P.dd);
  T* __restrict__ duout = reinterpret_cast<T*>(P.du);

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int b = tile / (g.tiles_x * g.tiles_y);
    const int rem = tile - b * (g.tiles_x * g.tiles_y);
    const int ty0 = (rem / g.tiles_x) * g.TH, tx0 = (rem % g.tiles_x) * g.TW;
    const int th = min(g.TH, g.H - ty0), tw = min(g.TW, g.W - tx0);
    const int nh = (g.TH + 2) * HW2;

    for (int idx = tid; idx < nh * NQ; idx += kThreads) {
      const int hp = idx / NQ, qq = idx - hp * NQ;
      const int hy = hp / HW2, hx = hp - hy * HW2;
      const int y = ty0 - 1 + hy, x = tx0 - 1 + hx;
      float4 u = f4_zero(), dd = f4_zero();
      if (y >= 0 && y < g.H && x >= 0 && x < g.W) {
        dd = ld4<T>(ddin + (((long long)b * g.H + y) * g.W + x) * C + 4 * qq);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < P.n_in) {
            const float4 sc = *reinterpret_cast<const float4*>(s_in + (4 * i) * C + 4 * qq);
            const float4 sh = *reinterpret_cast<const float4*>(s_in + (4 * i + 1) * C + 4 * qq);
            float4 val, raw;
            unsigned arg;
            load_input<T, C>(P.in[i], P.mode[i], b, y, x, qq, sc, sh, val, raw, arg);
            u = f4_axpy(wgt[i], val, u);
          }
        }
        if (P.swish) {
          u.x *= sigmoidf_(u.x); u.y *= sigmoidf_(u.y); u.z *= sigmoidf_(u.z); u.w *= sigmoidf_(u.w);
        }
      }
      *reinterpret_cast<float4*>(s_v + hp * C + 4 * qq) = u;
      *reinterpret_cast<float4*>(s_dd + hp * C + 4 * qq) = dd;
    }
    __syncthreads();

    if (active) {
      for (int p = prow; p < g.TH * g.TW; p += ROWS) {
        const int ty = p / g.TW, tx = p - ty * g.TW;
        if (ty >= th || tx >= tw) continue;
        const int y = ty0 + ty, x = tx0 + tx;
        // dL/dv = sum_taps K[tap] * dd[y - dy + 1][x - dx + 1]; dK[tap] += dd[y][x] * v[y + dy - 1][x + dx - 1]
        float4 dv = f4_zero();
        const float4 ddc = *reinterpret_cast<const float4*>(s_dd + ((ty + 1) * HW2 + tx + 1) * C + 4 * q);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 k = *reinterpret_cast<const float4*>(s_k + (dy * 3 + dx) * C + 4 * q);
            const float4 dn = *reinterpret_cast<const float4*>(s_dd + ((ty + 2 - dy) * HW2 + tx + 2 - dx) * C + 4 * q);
            dv = f4_fma(k, dn, dv);
            const float4 vn = *reinterpret_cast<const float4*>(s_v + ((ty + dy) * HW2 + tx + dx) * C + 4 * q);
            dK[dy * 3 + dx] = f4_fma(ddc, vn, dK[dy * 3 + dx]);
          }
        // rebuild u at the centre from the inputs (L1/L2 hits: the halo pass just read them)
        float4 u = f4_zero();
        float4 xh[3];
        float4 msk[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          xh[i] = f4_zero();
          msk[i] = f4_zero();
          if (i < P.n_in) {
            const float4 sc = *reinterpret_cast<const float4*>(s_in + (4 * i) * C + 4 * q);
            const float4 sh = *reinterpret_cast<const float4*>(s_in + (4 * i + 1) * C + 4 * q);
            float4 val, raw;
            unsigned arg;
            load_input<T, C>(P.in[i], P.mode[i], b, y, x, q, sc, sh, val, raw, arg);
            u = f4_axpy(wgt[i], val, u);
            msk[i] = make_float4((arg & 0xffu) == 9u ? 0.f : 1.f, ((arg >> 8) & 0xffu) == 9u ? 0.f : 1.f,
                                 ((arg >> 16) & 0xffu) == 9u ? 0.f : 1.f, ((arg >> 24) & 0xffu) == 9u ? 0.f : 1.f);
            if (P.in[i].bn != nullptr) {
              const float4 mu = *reinterpret_cast<const float4*>(s_in + (4 * i + 2) * C + 4 * q);
              const float4 is = *reinterpret_cast<const float4*>(s_in + (4 * i + 3) * C + 4 * q);
              xh[i] = make_float4((raw.x - mu.x) * is.x, (raw.y - mu.y) * is.y, (raw.z - mu.z) * is.z, (raw.w - mu.w) * is.w);
            } else {
              xh[i] = val;  // final tensor: the slot carries sum(du * x) for the fusion-weight gradient
            }
          }
        }
        float4 du = dv;
        if (P.swish) {
          const float sx = sigmoidf_(u.x), sy = sigmoidf_(u.y), sz = sigmoidf_(u.z), sw = sigmoidf_(u.w);
          du.x *= sx * (1.f + u.x * (1.f - sx));
          du.y *= sy * (1.f + u.y * (1.f - sy));
          du.z *= sz * (1.f + u.z * (1.f - sz));
          du.w *= sw * (1.f + u.w * (1.f - sw));
        }
        st4<T>(duout + (((long long)b * g.H + y) * g.W + x) * C + 4 * q, du);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < P.n_in) {
            const float4 dm = f4_mul(du, msk[i]);
            S1[i] = f4_add(S1[i], dm);
            S2[i] = f4_fma(dm, xh[i], S2[i]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- block reduction of the per-thread partials over the ROWS threads that share a channel quad
  float* s_red = s_dd;  // [ROWS*NQ][NACC]
  if (active) {
    float* r = s_red + (prow * NQ + q) * NACC;
#pragma unroll
    for (int t = 0; t < 9; ++t) { r[4 * t] = dK[t].x; r[4 * t + 1] = dK[t].y; r[4 * t + 2] = dK[t].z; r[4 * t + 3] = dK[t].w; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      r[36 + 4 * i] = S1[i].x; r[37 + 4 * i] = S1[i].y; r[38 + 4 * i] = S1[i].z; r[39 + 4 * i] = S1[i].w;
      r[48 + 4 * i] = S2[i].x; r[49 + 4 * i] = S2[i].y; r[50 + 4 * i] = S2[i].z; r[51 + 4 * i] = S2[i].w;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < NQ * NACC; idx += kThreads) {
    const int qq = idx / NACC, e = idx - qq * NACC;
    float s = 0.f;
    for (int rr = 0; rr < ROWS; ++rr) s += s_red[(rr * NQ + qq) * NACC + e];
    const int c = 4 * qq + (e & 3);
    if (e < 36) {
      if (P.g_dw) atomicAdd(P.g_dw + c * 9 + (e >> 2), s);
    } else if (e < 48) {
      const int i = (e - 36) >> 2;
      if (i < P.n_in && P.in_slot[i]) atomicAdd(P.in_slot[i] + c, (double)s);
    } else {
      const int i = (e - 48) >> 2;
      if (i < P.n_in && P.in_slot[i]) atomicAdd(P.in_slot[i] + C + c, (double)s);
    }
  }

  // ---- the last CTA turns the slots into the fusion-weight gradient
  if (P.fw == nullptr || P.g_fw == nullptr) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned ticket = atomicAdd(P.counter, 1u);
    s_flag = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < P.n_in) {
    const int i = warp;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const double s1 = __ldcg(P.in_slot[i] + c), s2 = __ldcg(P.in_slot[i] + C + c);
      if (P.in[i].bn != nullptr) acc += (float)((double)P.in_bn_w[i][c] * s2 + (double)P.in_bn_b[i][c] * s1);
      else acc += (float)s2;
    }
    acc = warp_sum(acc);
    if (lane == 0) s_gw[i] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    float ssum = 0.f;
    for (int j = 0; j < P.n_in; ++j) ssum += fmaxf(P.fw[j], 0.f);
    const float denom = ssum + P.fw_eps;
    float dot = 0.f;
    for (int j = 0; j < P.n_in; ++j) dot += fmaxf(P.fw[j], 0.f) / denom * s_gw[j];
    for (int k = 0; k < P.n_in; ++k) P.g_fw[k] = (P.fw[k] > 0.f) ? (s_gw[k] - dot) / denom : 0.f;
    *P.counter = 0u;
  }
}

// ---- first-cell projection backward ---------------------------------------------------------------------------
constexpr int kProjBC = 64;  // input-channel chunk per blockIdx.y
template <int C>
struct ProjBSmem {
  static constexpr int LDD = C + 4;
  static constexpr int LDX = kProjBC + 4;
  static constexpr int kGy = kTileP * LDD;
  static constexpr int kX = kTileP * LDX;
  static constexpr int kW = C * kProjBC;
  static constexpr int kCoef = 3 * C;
  static constexpr int kFloats = kGy + kX + kW + kCoef;
};

template <typename T, int C>
__global__ void __launch_bounds__(kThreads, 1) proj_bwd_kernel(const __grid_constant__ NodeBwdP P) {
  using S = ProjBSmem<C>;
  constexpr int NQ = C / 4, NJ = C / 16, LDD = S::LDD, LDX = S::LDX, BC = kProjBC, NI = BC / 16;
  extern __shared__ __align__(16) float smem[];
  float* s_gy = smem;
  float* s_x = s_gy + S::kGy;
  float* s_w = s_x + S::kX;
  float* s_coef = s_w + S::kW;

  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const TileGeom g = P.g;
  const int Cin = P.Cin, k0 = blockIdx.y * BC;
  float cw[3];
  cons_weights(P, cw);
  bn_bwd_coefs<C>(P, cw, s_coef);
  for (int idx = tid; idx < C * BC; idx += kThreads) {
    const int o = idx / BC, i = idx - o * BC;
    s_w[idx] = (k0 + i < Cin) ? P.pw_w[(long long)o * Cin + k0 + i] : 0.f;
  }
  __syncthreads();

  const T* __restrict__ yraw = reinterpret_cast<const T*>(P.out);
  const T* __restrict__ xin = reinterpret_cast<const T*>(P.in[0].data);
  T* __restrict__ dx = reinterpret_cast<T*>(P.dx);

  float accW[NJ][NI];
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int bb = 0; bb < NI; ++bb) accW[a][bb] = 0.f;
  float accB = 0.f;

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int b = tile / (g.tiles_x * g.tiles_y);
    const int rem = tile - b * (g.tiles_x * g.tiles_y);
    const int ty0 = (rem / g.tiles_x) * g.TH, tx0 = (rem % g.tiles_x) * g.TW;
    const int th = min(g.TH, g.H - ty0), tw = min(g.TW, g.W - tx0);

    for (int idx = tid; idx < kTileP * NQ; idx += kThreads) {
      const int p = idx / NQ, q = idx - p * NQ;
      const int ty = p / g.TW, tx = p - ty * g.TW;
      float4 gy = f4_zero();
      if (ty < th && tx < tw) {
        const int y = ty0 + ty, x = tx0 + tx;
        const float4 G = pull_grad<T, C>(P, cw, b, y, x, q);
        const float4 yr = ld4<T>(yraw + (((long long)b * g.H + y) * g.W + x) * C + 4 * q);
        const float4 A = *reinterpret_cast<const float4*>(s_coef + 4 * q);
        const float4 Bc = *reinterpret_cast<const float4*>(s_coef + C + 4 * q);
        const float4 Cc = *reinterpret_cast<const float4*>(s_coef + 2 * C + 4 * q);
        gy = f4_fma(A, G, f4_fma(Bc, yr, Cc));
      }
      *reinterpret_cast<float4*>(s_gy + p * LDD + 4 * q) = gy;
    }
    for (int idx = tid; idx < kTileP * (BC / 4); idx += kThreads) {
      const int p = idx / (BC / 4), kq = idx - p * (BC / 4);
      const int ty = p / g.TW, tx = p - ty * g.TW;
      float4 v = f4_zero();
      if (ty < th && tx < tw && k0 + 4 * kq < Cin)
        v = ld4<T>(xin + (((long long)b * g.H + ty0 + ty) * g.W + tx0 + tx) * Cin + k0 + 4 * kq);
      *reinterpret_cast<float4*>(s_x + p * LDX + 4 * kq) = v;
    }
    __syncthreads();

    // dW[o][i] += sum_p dy[p][o] * x[p][i]
#pragma unroll 2
    for (int p = 0; p < kTileP; ++p) {
      float go[NJ], xi[NI];
#pragma unroll
      for (int a = 0; a < NJ; ++a) go[a] = s_gy[p * LDD + tm + 16 * a];
#pragma unroll
      for (int bb = 0; bb < NI; ++bb) xi[bb] = s_x[p * LDX + tn + 16 * bb];
#pragma unroll
      for (int a = 0; a < NJ; ++a)
#pragma unroll
        for (int bb = 0; bb < NI; ++bb) accW[a][bb] = fmaf(go[a], xi[bb], accW[a][bb]);
    }
    if (blockIdx.y == 0 && tid < C) {
      float s = 0.f;
      for (int p = 0; p < kTileP; ++p) s += s_gy[p * LDD + tid];
      accB += s;
    }
    if (dx != nullptr) {
      // dx[p][i] = sum_o dy[p][o] * W[o][i]
      float acc[8][NI];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[r][j] = 0.f;
#pragma unroll 4
      for (int o = 0; o < C; ++o) {
        float a[8], w[NI];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = s_gy[(tm + 16 * r) * LDD + o];
#pragma unroll
        for (int j = 0; j < NI; ++j) w[j] = s_w[o * BC + tn + 16 * j];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int j = 0; j < NI; ++j) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < NI; ++j) s_x[(tm + 16 * r) * LDX + tn + 16 * j] = acc[r][j];
      __syncthreads();
      for (int idx = tid; idx < g.TH * g.TW * (BC / 4); idx += kThreads) {
        const int p = idx / (BC / 4), kq = idx - p * (BC / 4);
        const int ty = p / g.TW, tx = p - ty * g.TW;
        if (ty < th && tx < tw && k0 + 4 * kq < Cin) {
          T* dst = dx + (((long long)b * g.H + ty0 + ty) * g.W + tx0 + tx) * Cin + k0 + 4 * kq;
          float4 v = *reinterpret_cast<const float4*>(s_x + p * LDX + 4 * kq);
          if (P.accumulate_dx) v = f4_add(v, ld4<T>(dst));
          st4<T>(dst, v);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int bb = 0; bb < NI; ++bb) {
      const int i = k0 + tn + 16 * bb;
      if (i < Cin) atomicAdd(P.g_pw + (long long)(tm + 16 * a) * Cin + i, accW[a][bb]);
    }
  if (blockIdx.y == 0 && tid < C && P.g_pb) atomicAdd(P.g_pb + tid, accB);
}

// ---- materialise a gathered gradient -----------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(kThreads) pull_kernel(const __grid_constant__ NodeBwdP P) {
  constexpr int NQ = C / 4;
  const TileGeom g = P.g;
  float cw[3];
  cons_weights(P, cw);
  const long long total = (long long)g.B * g.H * g.W * NQ;
  T* dx = reinterpret_cast<T*>(P.dx);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % NQ);
    long long pos = idx / NQ;
    const int x = (int)(pos % g.W);
    pos /= g.W;
    const int y = (int)(pos % g.H);
    const int b = (int)(pos / g.H);
    float4 G = pull_grad<T, C>(P, cw, b, y, x, q);
    T* dst = dx + (((long long)b * g.H + y) * g.W + x) * C + 4 * q;
    if (P.accumulate_dx) G = f4_add(G, ld4<T>(dst));
    st4<T>(dst, G);
  }
}

// ---- slot for a deferred tensor consumed by a BNAPPLY op: (sum G, sum G * xhat) -------------------------------
// P.g = geometry of G (the BNAPPLY output); P.in[0] = the deferred source, P.mode[0] how it was read; P.cons[0].du = G.
template <typename T, int C>
__device__ __forceinline__ void slot_body(const NodeBwdP& P);

template <typename T, int C>
__global__ void __launch_bounds__(kThreads) slot_kernel(const __grid_constant__ NodeBwdP P) {
  slot_body<T, C>(P);
}
// the independent SLOT ops that open a backward in one launch: blockIdx.y selects the op
template <typename T, int C>
__global__ void __launch_bounds__(kThreads) slot_group_kernel(const __grid_constant__ NodeBwdGroup GROUP) {
  slot_body<T, C>(GROUP.p[blockIdx.y]);
}

template <typename T, int C>
__device__ __forceinline__ void slot_body(const NodeBwdP& P) {
  constexpr int NQ = C / 4;
  constexpr int ROWS = kThreads / NQ;
  __shared__ float s_red[ROWS * NQ * 8];
  const TileGeom g = P.g;
  const int tid = threadIdx.x;
  // a group launch sizes the grid for its largest op: the blocks an op does not need leave at once (every block ends
  // with 2*C double atomics onto the same addresses)
  const long long npos = (long long)g.B * g.H * g.W;
  const long long want = (npos + 9 * 8 - 1) / (9 * 8);
  const int nblk = (int)(want < (long long)gridDim.x ? (want < 1 ? 1 : want) : (long long)gridDim.x);
  if ((int)blockIdx.x >= nblk) return;
  const bool active = tid < ROWS * NQ;
  const int q = tid % NQ, prow = tid / NQ;
  const T* G = reinterpret_cast<const T*>(P.cons[0].du);
  const T* src = reinterpret_cast<const T*>(P.in[0].data);
  const float* bn = P.in[0].bn;
  const int Hs = P.in[0].H, Ws = P.in[0].W;
  float4 S1 = f4_zero(), S2 = f4_zero();
  if (active) {
    const float4 mu = *reinterpret_cast<const float4*>(bn + 2 * C + 4 * q);
    const float4 is = *reinterpret_cast<const float4*>(bn + 3 * C + 4 * q);
    const int top = pool_pad_before(Hs), left = pool_pad_before(Ws);
    for (long long pos = (long long)blockIdx.x * ROWS + prow; pos < npos; pos += (long long)nblk * ROWS) {
      const int x = (int)(pos % g.W);
      const int y = (int)((pos / g.W) % g.H);
      const int b = (int)(pos / ((long long)g.W * g.H));
      const float4 gg = ld4<T>(G + pos * C + 4 * q);
      if (P.mode[0] == MMD_IN_SAME) {
        const float4 r = ld4<T>(src + pos * C + 4 * q);
        S1 = f4_add(S1, gg);
        S2.x = fmaf(gg.x, (r.x - mu.x) * is.x, S2.x);
        S2.y = fmaf(gg.y, (r.y - mu.y) * is.y, S2.y);
        S2.z = fmaf(gg.z, (r.z - mu.z) * is.z, S2.z);
        S2.w = fmaf(gg.w, (r.w - mu.w) * is.w, S2.w);
      } else {
        const unsigned packed = *reinterpret_cast<const unsigned*>(P.pidx[0] + pos * C + 4 * q);
        const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
        const float muv[4] = {mu.x, mu.y, mu.z, mu.w};
        const float isv[4] = {is.x, is.y, is.z, is.w};
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned id = (packed >> (8 * j)) & 0xffu;
          if (id < 9u) {
            const int fy = 2 * y - top + (int)(id / 3u), fx = 2 * x - left + (int)(id % 3u);
            const float r = ld1<T>(src + (((long long)b * Hs + fy) * Ws + fx) * C + 4 * q + j);
            s1[j] = gv[j];
            s2[j] = gv[j] * (r - muv[j]) * isv[j];
          }
        }
        S1 = f4_add(S1, make_float4(s1[0], s1[1], s1[2], s1[3]));
        S2 = f4_add(S2, make_float4(s2[0], s2[1], s2[2], s2[3]));
      }
    }
    float* r = s_red + (prow * NQ + q) * 8;
    r[0] = S1.x; r[1] = S1.y; r[2] = S1.z; r[3] = S1.w;
    r[4] = S2.x; r[5] = S2.y; r[6] = S2.z; r[7] = S2.w;
  }
  __syncthreads();
  for (int idx = tid; idx < NQ * 8; idx += kThreads) {
    const int qq = idx >> 3, e = idx & 7;
    float s = 0.f;
    for (int rr = 0; rr < ROWS; ++rr) s += s_red[(rr * NQ + qq) * 8 + e];
    const int c = 4 * qq + (e & 3);
    atomicAdd(P.in_slot[0] + (e < 4 ? c : C + c), (double)s);
  }
}

// ---- host launchers ----------------------------------------------------------------------------------------
static int sm_count() { return device_sm_count(); }

template <typename T>
static int launch_node_bwd_t(const NodeBwdP& p, cudaStream_t s) {
  constexpr int C = 112;
  const size_t smem_a = BwdASmem<C>::kFloats * sizeof(float), smem_b = BwdBSmem<C>::kFloats * sizeof(float);
  MMD_SMEM((node_bwd_a_kernel<T, C>), smem_a);
  MMD_SMEM((node_bwd_b_kernel<T, C>), smem_b);
  const int grid = p.g.ntiles < sm_count() ? p.g.ntiles : sm_count();
  const double bytes = node_algo_bytes(p.in, p.n_in, p.g, C, sizeof(T));
  static int no_v4 = -1;
  if (no_v4 < 0) {
    const char* e = getenv("MMD_NO_BWD_V4");
    no_v4 = (e && e[0] == '1') ? 1 : 0;
  }
  if (sizeof(T) == 2 && !tc_disabled() && !no_v4 && bwd_v4_usable(p)) return launch_node_bwd_v4(p, C, s);
  if (sizeof(T) == 2 && !tc_disabled()) {
    int rc = launch_node_bwd_a_tc(p, C, s);
    if (rc) return rc;
  } else {
    {
      ProfScope prof(PK_NODE_BWD_A, bytes, s);
      node_bwd_a_kernel<T, C><<<grid, kThreads, smem_a, s>>>(p);
    }
    MMD_LAUNCH_CHECK();
  }
  {
    ProfScope prof(PK_NODE_BWD_B, bytes, s);
    node_bwd_b_kernel<T, C><<<grid, kThreads, smem_b, s>>>(p);
  }
  MMD_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_proj_bwd_t(const NodeBwdP& p, cudaStream_t s) {
  constexpr int C = 112;
  if (sizeof(T) == 2 && !tc_disabled() && proj_bwd_v4_usable(p)) return launch_proj_bwd_v4(p, C, s);
  const size_t smem = ProjBSmem<C>::kFloats * sizeof(float);
  MMD_SMEM((proj_bwd_kernel<T, C>), smem);
  const int gx = p.g.ntiles < sm_count() ? p.g.ntiles : sm_count();
  const int gy = (p.Cin + kProjBC - 1) / kProjBC;
  ProfScope prof(PK_PROJ_BWD, 2.0 * p.g.B * p.g.H * p.g.W * (p.Cin + C) * sizeof(T), s);
  proj_bwd_kernel<T, C><<<dim3(gx, gy), kThreads, smem, s>>>(p);
  MMD_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_pull_t(const NodeBwdP& p, cudaStream_t s) {
  constexpr int C = 112;
  long long total = (long long)p.g.B * p.g.H * p.g.W * (C / 4);
  long long grid = (total + kThreads - 1) / kThreads;
  if (grid > 8LL * sm_count()) grid = 8LL * sm_count();
  ProfScope prof(PK_PULL, 2.0 * p.g.B * p.g.H * p.g.W * C * sizeof(T), s);
  pull_kernel<T, C><<<(unsigned)grid, kThreads, 0, s>>>(p);
  MMD_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_slot_group_t(const NodeBwdP* p, int n, cudaStream_t s) {
  constexpr int C = 112;
  NodeBwdGroup group;
  double bytes = 0.0;
  long long maxpos = 0;
  for (int i = 0; i < kMaxGroupOps; ++i) group.p[i] = p[i < n ? i : 0];
  for (int i = 0; i < n; ++i) {
    const long long npos = (long long)p[i].g.B * p[i].g.H * p[i].g.W;
    bytes += 2.0 * npos * C * sizeof(T);
    if (npos > maxpos) maxpos = npos;
  }
  long long grid = (maxpos + 9 * 8 - 1) / (9 * 8);
  if (grid > 4LL * sm_count()) grid = 4LL * sm_count();
  if (grid < 1) grid = 1;
  ProfScope prof(PK_SLOT, bytes, s);
  slot_group_kernel<T, C><<<dim3((unsigned)grid, n), kThreads, 0, s>>>(group);
  MMD_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_slot_t(const NodeBwdP& p, cudaStream_t s) {
  constexpr int C = 112;
  long long npos = (long long)p.g.B * p.g.H * p.g.W;
  long long grid = (npos + 9 * 8 - 1) / (9 * 8);  // ~8 positions per thread-row
  if (grid > 4LL * sm_count()) grid = 4LL * sm_count();
  if (grid < 1) grid = 1;
  ProfScope prof(PK_SLOT, 2.0 * p.g.B * p.g.H * p.g.W * C * sizeof(T), s);
  slot_kernel<T, C><<<(unsigned)grid, kThreads, 0, s>>>(p);
  MMD_LAUNCH_CHECK();
  return 0;
}

#define MMD_DISPATCH(fn)                                                             \
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C); \
  if (dtype == MMD_F32) return fn<float>(p, s);                                      \
  if (dtype == MMD_BF16) return fn<__nv_bfloat16>(p, s);                             \
  set_error("unsupported dtype %d", dtype);                                          \
  return MMD_E_ARG;

int launch_node_bwd(const NodeBwdP& p, int C, int dtype, cudaStream_t s) { MMD_DISPATCH(launch_node_bwd_t) }
int launch_proj_bwd(const NodeBwdP& p, int C, int dtype, cudaStream_t s) { MMD_DISPATCH(launch_proj_bwd_t) }
int launch_pull(const NodeBwdP& p, int C, int dtype, cudaStream_t s) { MMD_DISPATCH(launch_pull_t) }
int launch_slot(const NodeBwdP& p, int C, int dtype, cudaStream_t s) { MMD_DISPATCH(launch_slot_t) }
int launch_slot_group(const NodeBwdP* p, int n, int C, int dtype, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  MMD_CHECK_ARG(n >= 1 && n <= kMaxGroupOps, "slot group of %d ops", n);
  if (dtype == MMD_F32) return launch_slot_group_t<float>(p, n, s);
  if (dtype == MMD_BF16 && !tc_disabled() && slot_same_bf16_usable(p, n)) return launch_slot_same_bf16(p, n, s);
  if (dtype == MMD_BF16) return launch_slot_group_t<__nv_bfloat16>(p, n, s);
  set_error("unsupported dtype %d", dtype);
  return MMD_E_ARG;
}

}  // namespace mmd
