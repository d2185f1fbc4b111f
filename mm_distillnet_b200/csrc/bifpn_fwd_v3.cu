// BiFPN fusion node, forward, bf16 storage — v3 (sm_100a).
//
// Same fusion as bifpn_fwd_tc.cu (BN-on-load + resample + fast-normalised weighted sum + swish -> depthwise 3x3 ->
// pointwise 1x1 on tcgen05 -> bias / BatchNorm statistics), restructured around the async copy engine:
//   * parameters come from the packed block written by mmd_bifpn_prep (prep.cu): ONE cp.async.bulk of 29.5 KB per CTA
//     (B operand in UMMA layout, bias, taps) instead of every CTA re-reading and re-formatting 50 KB of fp32 weights;
//   * the raw halo tiles of the SAME / nearest-x2 inputs are fetched by cp.async.bulk row copies (one per halo row, a
//     halo row of an NHWC tensor is contiguous) that complete on an mbarrier: no registers, no per-element LDG, the
//     whole tile's HBM latency is paid once; pooled inputs (3x3 stride-2 windows) still use vector loads;
//   * region 0 of shared memory holds raw input 0 -> v (in place) -> the bf16 output staging tile; region 1 holds raw
//     input 1 -> the UMMA A operand (channel-group stride padded to 2064 B: the depthwise stage's stores are spread
//     over the banks instead of 14-way conflicting);
//   * packed fp32x2 arithmetic (FFMA2) for BN / fusion / swish / depthwise, a 2-column register block in the depthwise
//     stage, BatchNorm statistics taken from the staging tile by 224 threads in parallel.
// Two CTAs per SM (110 KB each): while one waits for its copies or its MMA the other computes.
#include "bifpn.cuh"
#include "tc.cuh"

namespace mmd {

typedef __nv_bfloat16 bf16;
using tc::add2;
using tc::bf2_to_f2;
using tc::f2_to_bf2;
using tc::fma2;
using tc::mul2;

template <int C>
struct FwdV3 {
  static constexpr int NG = C / 8;
  static constexpr int kPos = C * 2;                         // bytes per position (224)
  static constexpr int kRegion = kHaloMax * kPos;            // 40 320
  static constexpr int kAStride = kTileP * 16 + 16;          // 2064: bytes between channel groups of the A operand
  static constexpr int kABytes = NG * kAStride;
  static constexpr int LDS = C + 8;                          // staging row, bf16 elements
  static constexpr int kWBytes = C * C * 2;
  static constexpr int kPackBytes = kWBytes + C * 4 + 9 * C * 4;
  static constexpr int offR0 = 0;
  static constexpr int offR1 = kRegion;
  static constexpr int offPack = 2 * kRegion;
  static constexpr int offCoef = offPack + kPackBytes;
  static constexpr int kCoefFloats = 5 * C;                  // a[3][C] | shsum[C] | pool shift[C]
  static constexpr int offBar = offCoef + kCoefFloats * 4;
  static constexpr int kBytes = offBar + 32;
  static constexpr int kLanesP1 = kThreads / NG;             // 18 halo positions in flight in phase 1
  static_assert(kTileP * LDS * 2 <= kRegion, "staging tile must fit in region 0");
  static_assert(kABytes <= kRegion, "A operand must fit in region 1");
  static_assert(offR1 % 128 == 0 && offPack % 128 == 0 && offCoef % 16 == 0 && offBar % 8 == 0, "alignment");
  static_assert(2 * (kBytes + 1024) <= 233472, "two CTAs per SM");
};

// ---- an input that is NOT staged by bulk copies (pooled inputs; a third SAME / UP2 input) -------------------------
// Adds w * resample(scale * x + shift) for 8 channels.  `coef` = this input's scale row (already multiplied by the
// fusion weight for non-pooled inputs, whose shift lives in the common shift row), `poolsh` = shift row for pooling.
template <int C>
__device__ __forceinline__ void direct_input(const TensorP& t, int mode, int b, int y, int x, int cg, const float* coef,
                                             const float* poolsh, float w, float2 (&u)[4], unsigned char* pidx_dst) {
  const bf16* base = reinterpret_cast<const bf16*>(t.data);
  float2 sc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) sc[e] = *reinterpret_cast<const float2*>(coef + 8 * cg + 2 * e);
  if (mode != MMD_IN_POOL) {
    const int sy = (mode == MMD_IN_UP2) ? (y >> 1) : y, sx = (mode == MMD_IN_UP2) ? (x >> 1) : x;
    const uint4 r = *reinterpret_cast<const uint4*>(base + (((long long)b * t.H + sy) * t.W + sx) * C + 8 * cg);
    u[0] = fma2(bf2_to_f2(r.x), sc[0], u[0]);
    u[1] = fma2(bf2_to_f2(r.y), sc[1], u[1]);
    u[2] = fma2(bf2_to_f2(r.z), sc[2], u[2]);
    u[3] = fma2(bf2_to_f2(r.w), sc[3], u[3]);
    return;
  }
  float2 sh[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) sh[e] = *reinterpret_cast<const float2*>(poolsh + 8 * cg + 2 * e);
  // MaxPool2dStaticSamePadding (YetAnotherEfficientNet.py:90-104): zero padding takes part in the max, the first
  // maximum in row-major window order wins (index 9 = padding)
  const int top = pool_pad_before(t.H), left = pool_pad_before(t.W);
  float m[8];
  unsigned a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; a[e] = 9u; }
#pragma unroll
  for (int wy = 0; wy < 3; ++wy) {
    const int fy = 2 * y - top + wy;
#pragma unroll
    for (int wx = 0; wx < 3; ++wx) {
      const int fx = 2 * x - left + wx;
      const bool inside = (fy >= 0) && (fy < t.H) && (fx >= 0) && (fx < t.W);
      float2 v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = make_float2(0.f, 0.f);
      if (inside) {
        const uint4 r = *reinterpret_cast<const uint4*>(base + (((long long)b * t.H + fy) * t.W + fx) * C + 8 * cg);
        v[0] = fma2(bf2_to_f2(r.x), sc[0], sh[0]);
        v[1] = fma2(bf2_to_f2(r.y), sc[1], sh[1]);
        v[2] = fma2(bf2_to_f2(r.z), sc[2], sh[2]);
        v[3] = fma2(bf2_to_f2(r.w), sc[3], sh[3]);
      }
      const unsigned id = inside ? (unsigned)(wy * 3 + wx) : 9u;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (v[e].x > m[2 * e]) { m[2 * e] = v[e].x; a[2 * e] = id; }
        if (v[e].y > m[2 * e + 1]) { m[2 * e + 1] = v[e].y; a[2 * e + 1] = id; }
      }
    }
  }
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int e = 0; e < 4; ++e) u[e] = fma2(w2, make_float2(m[2 * e], m[2 * e + 1]), u[e]);
  if (pidx_dst != nullptr) {
    uint2 pk;
    pk.x = a[0] | (a[1] << 8) | (a[2] << 16) | (a[3] << 24);
    pk.y = a[4] | (a[5] << 8) | (a[6] << 16) | (a[7] << 24);
    *reinterpret_cast<uint2*>(pidx_dst) = pk;
  }
}

template <int C>
__global__ void __launch_bounds__(kThreads, 2) node_fwd_v3_kernel(const __grid_constant__ NodeFwdP P) {
  using S = FwdV3<C>;
  constexpr int NG = S::NG, NQ = C / 4, LDS = S::LDS, POS = S::kPos;
  constexpr uint32_t kTmemCols = 128;
  constexpr uint32_t kIdesc = tc::make_idesc_bf16(128, C, false, false);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* r0 = smem + S::offR0;
  unsigned char* r1 = smem + S::offR1;
  unsigned char* s_pack = smem + S::offPack;
  const float* s_bias = reinterpret_cast<const float*>(s_pack + S::kWBytes);
  const float* s_k = s_bias + C;
  float* s_coef = reinterpret_cast<float*>(smem + S::offCoef);
  uint64_t* bar_pack = reinterpret_cast<uint64_t*>(smem + S::offBar);
  uint64_t* bar_in = bar_pack + 1;
  uint64_t* bar_mma = bar_pack + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_pack + 3);
  __shared__ int s_flag;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TileGeom g = P.g;
  const bool train = P.train != 0;
  bf16* __restrict__ out = reinterpret_cast<bf16*>(P.out);
  const bool staged1 = (P.n_in >= 2) && (P.mode[1] != MMD_IN_POOL);
  const int first_direct = staged1 ? 2 : 1;
  const int mode1 = P.mode[1];

  // ---- per-CTA setup ---------------------------------------------------------------------------------------------
  if (warp == 0) tc::tmem_alloc(s_tmem, kTmemCols);
  if (tid == 32) {
    tc::mbar_init(bar_pack, 1);
    tc::mbar_init(bar_in, 1);
    tc::mbar_init(bar_mma, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(bar_pack, S::kPackBytes);
    tc::bulk_g2s(s_pack, P.packed, S::kPackBytes, bar_pack);
  }
  float wgt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) wgt[i] = (i < P.n_in) ? fusion_weight(P.fw, P.n_in, i, P.fw_eps) : 0.f;
  if (tid < C) {
    float shsum = 0.f, poolsh = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float a = 0.f;
      if (i < P.n_in) {
        const float* bn = P.in[i].bn;
        const float sc = bn ? bn[tid] : 1.f, sh = bn ? bn[C + tid] : 0.f;
        if (P.mode[i] == MMD_IN_POOL) {   // the weight is applied after the max
          a = sc;
          poolsh = sh;
        } else {
          a = sc * wgt[i];
          shsum = fmaf(sh, wgt[i], shsum);
        }
      }
      s_coef[i * C + tid] = a;
    }
    s_coef[3 * C + tid] = shsum;
    s_coef[4 * C + tid] = poolsh;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t a_addr = tc::smem_u32(r1), b_addr = tc::smem_u32(s_pack);

  // phase-1 role: 8 channels (cg) of every 18th halo position; phase-2 role: 4 channels (q) of a column pair
  const int cg = tid % NG, j1 = tid / NG;
  float2 a0[4], a1[4], shs[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    a0[e] = *reinterpret_cast<const float2*>(s_coef + 8 * cg + 2 * e);
    a1[e] = *reinterpret_cast<const float2*>(s_coef + C + 8 * cg + 2 * e);
    shs[e] = *reinterpret_cast<const float2*>(s_coef + 3 * C + 8 * cg + 2 * e);
  }
  const int q2 = tid % NQ, cp2 = tid / NQ;
  // statistics role: channel pair sp of row slice ss (32 staging rows)
  const int sp = tid % (C / 2), ss = tid / (C / 2);
  double st[4] = {0.0, 0.0, 0.0, 0.0};

  const int HW2 = g.TW + 2, HH2 = g.TH + 2, nh = HH2 * HW2;
  const bf16* __restrict__ in0 = reinterpret_cast<const bf16*>(P.in[0].data);
  const bf16* __restrict__ in1 = reinterpret_cast<const bf16*>(P.in[1].data);
  uint32_t ph_in = 0, ph_mma = 0;
  bool pack_ready = false;

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int b = tile / (g.tiles_x * g.tiles_y);
    const int rem = tile - b * (g.tiles_x * g.tiles_y);
    const int ty0 = (rem / g.tiles_x) * g.TH, tx0 = (rem % g.tiles_x) * g.TW;
    const int th = min(g.TH, g.H - ty0), tw = min(g.TW, g.W - tx0);
    // halo rectangle clipped to the image, and the source rectangle of a nearest-x2 input 1
    const int y_lo = max(ty0 - 1, 0), y_hi = min(ty0 + g.TH, g.H - 1);
    const int x_lo = max(tx0 - 1, 0), x_hi = min(tx0 + g.TW, g.W - 1);
    const int sy_lo = y_lo >> 1, sx_lo = x_lo >> 1;
    const int pitch1 = (x_hi >> 1) - sx_lo + 1;

    // ---- (1) bulk row copies of the raw inputs (warp 1: one halo row per lane)
    if (warp == 1) {
      const int rows0 = y_hi - y_lo + 1;
      const uint32_t rb0 = (uint32_t)(x_hi - x_lo + 1) * POS;
      int rows1 = 0;
      uint32_t rb1 = 0;
      if (staged1) {
        if (mode1 == MMD_IN_SAME) { rows1 = rows0; rb1 = rb0; }
        else { rows1 = (y_hi >> 1) - sy_lo + 1; rb1 = (uint32_t)pitch1 * POS; }
      }
      if (lane == 0) tc::mbar_expect_tx(bar_in, rows0 * rb0 + rows1 * rb1);
      __syncwarp();
      if (lane < rows0) {
        const int y = y_lo + lane;
        tc::bulk_g2s(r0 + ((y - (ty0 - 1)) * HW2 + (x_lo - (tx0 - 1))) * POS,
                     in0 + (((long long)b * g.H + y) * g.W + x_lo) * C, rb0, bar_in);
      }
      if (lane < rows1) {
        if (mode1 == MMD_IN_SAME) {
          const int y = y_lo + lane;
          tc::bulk_g2s(r1 + ((y - (ty0 - 1)) * HW2 + (x_lo - (tx0 - 1))) * POS,
                       in1 + (((long long)b * g.H + y) * g.W + x_lo) * C, rb1, bar_in);
        } else {
          const int sy = sy_lo + lane;
          tc::bulk_g2s(r1 + (lane * pitch1) * POS,
                       in1 + (((long long)b * P.in[1].H + sy) * P.in[1].W + sx_lo) * C, rb1, bar_in);
        }
      }
    }
    tc::mbar_wait(bar_in, ph_in);
    ph_in ^= 1u;

    // ---- (2) phase 1: v = swish(sum_i w_i * resample_i(bn_i(x_i))), written in place over raw input 0
    if (j1 < S::kLanesP1) {
      int hy = j1 / HW2, hx = j1 - hy * HW2;
      for (int hp = j1; hp < nh; hp += S::kLanesP1) {
        const int y = ty0 - 1 + hy, x = tx0 - 1 + hx;
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
        unsigned char* cell = r0 + hp * POS + cg * 16;
        if (y >= 0 && y < g.H && x >= 0 && x < g.W) {
          float2 u[4];
          {
            const uint4 r = *reinterpret_cast<const uint4*>(cell);
            u[0] = fma2(bf2_to_f2(r.x), a0[0], shs[0]);
            u[1] = fma2(bf2_to_f2(r.y), a0[1], shs[1]);
            u[2] = fma2(bf2_to_f2(r.z), a0[2], shs[2]);
            u[3] = fma2(bf2_to_f2(r.w), a0[3], shs[3]);
          }
          if (staged1) {
            const int i1 = (mode1 == MMD_IN_SAME) ? hp : ((y >> 1) - sy_lo) * pitch1 + ((x >> 1) - sx_lo);
            const uint4 r = *reinterpret_cast<const uint4*>(r1 + i1 * POS + cg * 16);
            u[0] = fma2(bf2_to_f2(r.x), a1[0], u[0]);
            u[1] = fma2(bf2_to_f2(r.y), a1[1], u[1]);
            u[2] = fma2(bf2_to_f2(r.z), a1[2], u[2]);
            u[3] = fma2(bf2_to_f2(r.w), a1[3], u[3]);
          }
          for (int i = first_direct; i < P.n_in; ++i) {
            const bool center = (hy >= 1 && hy <= th && hx >= 1 && hx <= tw);
            unsigned char* pd = (center && P.mode[i] == MMD_IN_POOL && P.pidx[i] != nullptr)
                                    ? P.pidx[i] + (((long long)b * g.H + y) * g.W + x) * C + 8 * cg
                                    : nullptr;
            direct_input<C>(P.in[i], P.mode[i], b, y, x, cg, s_coef + i * C, s_coef + 4 * C, wgt[i], u, pd);
          }
          if (P.swish) {
#pragma unroll
            for (int e = 0; e < 4; ++e) u[e] = tc::swish2(u[e]);
          }
          packed.x = f2_to_bf2(u[0]); packed.y = f2_to_bf2(u[1]);
          packed.z = f2_to_bf2(u[2]); packed.w = f2_to_bf2(u[3]);
        }
        *reinterpret_cast<uint4*>(cell) = packed;
        hx += S::kLanesP1;
        while (hx >= HW2) { hx -= HW2; ++hy; }
      }
    }
    __syncthreads();
    if (!pack_ready) {   // taps / bias / B operand have landed (first tile only)
      tc::mbar_wait(bar_pack, 0u);
      pack_ready = true;
    }

    // ---- (3) phase 2: depthwise 3x3, two adjacent columns per thread, rolling accumulators over the halo rows
    if (2 * cp2 < g.TW) {
      const int c0 = 2 * cp2;
      float2 wk[9][2];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) {
        const float4 k4 = *reinterpret_cast<const float4*>(s_k + t9 * C + 4 * q2);
        wk[t9][0] = make_float2(k4.x, k4.y);
        wk[t9][1] = make_float2(k4.z, k4.w);
      }
      float2 acc[3][2][2];   // [rolling slot][column][channel pair]
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[s][c][0] = acc[s][c][1] = make_float2(0.f, 0.f);
      const unsigned char* vcol = r0 + c0 * POS + q2 * 8;
      unsigned char* arow = r1 + (q2 >> 1) * S::kAStride + (q2 & 1) * 8;
      bf16* dsave = (P.save_d != nullptr)
                        ? reinterpret_cast<bf16*>(P.save_d) + (((long long)b * g.H + ty0) * g.W + tx0 + c0) * C + 4 * q2
                        : nullptr;
      const bool col1 = (c0 + 1 < g.TW);
      auto row_step = [&](int r, float2 (&a_new)[2][2], float2 (&a_mid)[2][2], float2 (&a_old)[2][2]) {
        // halo row r feeds tap row 0 of output row r (a_new), tap row 1 of r-1 (a_mid), tap row 2 of r-2 (a_old)
        float2 v[4][2];
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const uint2 raw = *reinterpret_cast<const uint2*>(vcol + (r * HW2 + dx) * POS);
          v[dx][0] = bf2_to_f2(raw.x);
          v[dx][1] = bf2_to_f2(raw.y);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a_new[c][h] = mul2(v[c][h], wk[0][h]);
            a_new[c][h] = fma2(v[c + 1][h], wk[1][h], a_new[c][h]);
            a_new[c][h] = fma2(v[c + 2][h], wk[2][h], a_new[c][h]);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              a_mid[c][h] = fma2(v[c + dx][h], wk[3 + dx][h], a_mid[c][h]);
              a_old[c][h] = fma2(v[c + dx][h], wk[6 + dx][h], a_old[c][h]);
            }
          }
        const int ty = r - 2;
        if (ty >= 0 && ty < th) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c == 1 && !col1) break;
            if (c0 + c < tw) {
              uint2 pk;
              pk.x = f2_to_bf2(a_old[c][0]);
              pk.y = f2_to_bf2(a_old[c][1]);
              *reinterpret_cast<uint2*>(arow + (ty * g.TW + c0 + c) * 16) = pk;
              if (dsave) *reinterpret_cast<uint2*>(dsave + ((long long)ty * g.W + c) * C) = pk;
            }
          }
        }
      };
      for (int rr = 0; rr < HH2; rr += 3) {
        row_step(rr, acc[0], acc[2], acc[1]);
        if (rr + 1 < HH2) row_step(rr + 1, acc[1], acc[0], acc[2]);
        if (rr + 2 < HH2) row_step(rr + 2, acc[2], acc[1], acc[0]);
      }
    }
    tc::fence_async_smem();   // the A operand was written through the generic proxy
    __syncthreads();

    // ---- (4) pointwise 1x1 on the tensor cores
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {
        const uint64_t adesc = tc::make_desc(a_addr + j * 2 * S::kAStride, S::kAStride, 128);
        const uint64_t bdesc = tc::make_desc(b_addr + j * 2 * (C * 16), C * 16, 128);
        tc::umma_bf16(tmem_base, adesc, bdesc, kIdesc, j > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1u;
    tc::fence_after_sync();

    // ---- (5) epilogue: TMEM -> +bias -> bf16 staging (region 0) -> statistics + 16-byte coalesced stores
    bf16* s_y = reinterpret_cast<bf16*>(r0);
    {
      const int row = 32 * (warp & 3) + lane;
      const int col0 = (warp >> 2) * (C / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
      float acc[C / 16][8];
#pragma unroll
      for (int j = 0; j < C / 16; ++j) tc::tmem_ld8(taddr + 8 * j, acc[j]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + col0 + 8 * j);
        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + col0 + 8 * j + 4);
        uint4 pk;
        pk.x = f2_to_bf2(add2(make_float2(acc[j][0], acc[j][1]), make_float2(b0.x, b0.y)));
        pk.y = f2_to_bf2(add2(make_float2(acc[j][2], acc[j][3]), make_float2(b0.z, b0.w)));
        pk.z = f2_to_bf2(add2(make_float2(acc[j][4], acc[j][5]), make_float2(b1.x, b1.y)));
        pk.w = f2_to_bf2(add2(make_float2(acc[j][6], acc[j][7]), make_float2(b1.z, b1.w)));
        *reinterpret_cast<uint4*>(s_y + row * LDS + col0 + 8 * j) = pk;
      }
    }
    tc::fence_before_sync();   // order the TMEM reads before the next tile's MMAs
    __syncthreads();
    const int nrow = g.TH * g.TW;
    if (train && ss < 4) {
      const int p0 = ss * 32, p1 = min(p0 + 32, nrow);
      float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
      const bf16* colp = s_y + 2 * sp;
      if (th == g.TH && tw == g.TW) {
#pragma unroll 8
        for (int p = p0; p < p1; ++p) {
          const float2 v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(colp + p * LDS));
          s = add2(s, v);
          q = fma2(v, v, q);
        }
      } else {
        for (int p = p0; p < p1; ++p) {
          const int ty = p / g.TW, tx = p - ty * g.TW;
          if (ty < th && tx < tw) {
            const float2 v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(colp + p * LDS));
            s = add2(s, v);
            q = fma2(v, v, q);
          }
        }
      }
      st[0] += (double)s.x; st[1] += (double)s.y; st[2] += (double)q.x; st[3] += (double)q.y;
    }
    for (int idx = tid; idx < nrow * NG; idx += kThreads) {
      const int p = idx / NG, gq = idx - p * NG;
      const int ty = p / g.TW, tx = p - ty * g.TW;
      if (ty < th && tx < tw)
        *reinterpret_cast<uint4*>(out + (((long long)b * g.H + ty0 + ty) * g.W + tx0 + tx) * C + 8 * gq) =
            *reinterpret_cast<const uint4*>(s_y + p * LDS + 8 * gq);
    }
    tc::fence_async_smem();   // generic accesses to regions 0/1 are ordered before the next tile's bulk copies
    __syncthreads();
  }

  // ---- teardown + BatchNorm finalisation (last CTA)
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
  if (!train) return;
  double* s_red = reinterpret_cast<double*>(r1);   // [4 slices][C/2 pairs][4]
  if (ss < 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) s_red[(ss * (C / 2) + sp) * 4 + e] = st[e];
  }
  __syncthreads();
  if (tid < C) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += s_red[(k * (C / 2) + (tid >> 1)) * 4 + (tid & 1)];
      q += s_red[(k * (C / 2) + (tid >> 1)) * 4 + 2 + (tid & 1)];
    }
    atomicAdd(P.stats + tid, s);
    atomicAdd(P.stats + C + tid, q);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(P.counter, 1u);
    s_flag = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_flag == 0) return;
  __threadfence();
  if (tid < C) {
    const double n = (double)g.B * g.H * g.W;
    const double mean = __ldcg(P.stats + tid) / n;
    double var = __ldcg(P.stats + C + tid) / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)P.bn_eps));
    const float scale = P.bn_w[tid] * invstd;
    P.out_bn[tid] = scale;
    P.out_bn[C + tid] = P.bn_b[tid] - (float)mean * scale;
    P.out_bn[2 * C + tid] = (float)mean;
    P.out_bn[3 * C + tid] = invstd;
    const double unbiased = var * (n / (n > 1.0 ? n - 1.0 : 1.0));
    P.bn_rm[tid] = (1.f - P.bn_mom) * P.bn_rm[tid] + P.bn_mom * (float)mean;
    P.bn_rv[tid] = (1.f - P.bn_mom) * P.bn_rv[tid] + P.bn_mom * (float)unbiased;
    P.stats[tid] = 0.0;
    P.stats[C + tid] = 0.0;
  }
  if (tid == 0) {
    *P.counter = 0u;
    if (P.bn_nbt) *P.bn_nbt += 1;
  }
}

bool fwd_v3_usable(const NodeFwdP& p) {
  if (p.packed == nullptr || p.n_in < 1 || p.mode[0] != MMD_IN_SAME) return false;
  for (int i = 0; i < p.n_in; ++i)
    if (((uintptr_t)p.in[i].data & 15u) != 0) return false;
  return true;
}

int launch_node_fwd_v3(const NodeFwdP& p, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  constexpr int CC = 112;
  const size_t smem = FwdV3<CC>::kBytes;
  static bool configured = false;
  if (!configured) {
    MMD_CUDA(cudaFuncSetAttribute(node_fwd_v3_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.g.ntiles < 2 * sms ? p.g.ntiles : 2 * sms;
  ProfScope prof(PK_NODE_FWD, node_algo_bytes(p.in, p.n_in, p.g, C, 2), s);
  node_fwd_v3_kernel<CC><<<grid, kThreads, smem, s>>>(p);
  MMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmd
