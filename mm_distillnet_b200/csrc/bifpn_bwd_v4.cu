// BiFPN fusion node, backward, bf16 storage — v4 (sm_100a): compile-time tile geometry, two kernels per node.
//
//   node_bwd_a4 : per TW x TH tile: G = gather of dL/du from the (<= 3) consumers (identity / 2x2 block sum / arg-max
//                 match against the bytes the forward pre-pass recorded), BatchNorm backward dy = A*G + Bc*y + Cc, and
//                 both contractions on tcgen05:  dL/dd = dy * W (accumulator columns [0,128) of TMEM, written straight
//                 to HBM from the epilogue registers) and dW += dy^T * [d | 1] (columns [128,256), kept in TMEM across
//                 all tiles of the persistent CTA; the extra all-ones column makes the bias gradient fall out of the
//                 same MMA), flushed once with 16-byte vector reductions.
//   node_bwd_b4 : warp-specialised, double-buffered.  A producer warp streams, per tile, the halo tile of dL/dd and the
//                 centre tiles of the node's inputs into shared memory with bulk copies (mbarrier full / ready pairs)
//                 while 14 compute warps run the previous tile: depthwise^T over a 3-row register window of dL/dd,
//                 u rebuilt from the inputs, swish', dL/du written IN PLACE over the raw input-0 tile and sent to HBM by
//                 bulk stores; depthwise weight gradient, per-input-edge BatchNorm sums (the "slots" of the producers'
//                 BN backward) and, by the last CTA, the fusion-weight gradient.
// Thread = one channel pair x two columns; every shared-memory offset is a compile-time constant and both kernels are
// fully unrolled over the tile rows (see bifpn_fwd_v4.cu for the measurements that motivated this).
#include <stdlib.h>
#include "bifpn_bwd_common.cuh"
#include "tc.cuh"

namespace mmd {
namespace b4 {

typedef __nv_bfloat16 bf16;
using tc::add2;
using tc::bf2_to_f2;
using tc::f2_to_bf2;
using tc::fma2;
using tc::mul2;

constexpr int C = 112, NG = C / 8, NPR = C / 2, POS = C * 2;
constexpr int X1_NONE = 0, X1_SAME = 1, X1_UP2 = 2, X1_AUX = 3;

constexpr int up128(int a) { return (a + 127) / 128 * 128; }
constexpr int up32(int a) { return (a + 31) / 32 * 32; }

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_gmem)),
               "r"(tc::smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

struct TilePos {
  int b, ty0, tx0;
};
// tile -> (sample, first row, first column).  The two divisions are by loop-invariant small numbers: ncu's source view
// charged ~5 % of the forward kernel's instructions to them, so they are multiply-high by precomputed reciprocals
// (exact for tile * divisor < 2^32).
struct TileDiv {
  int per, tiles_x;
  uint32_t m_per, m_tx;
};
__device__ __forceinline__ TileDiv make_tile_div(int tiles_x, int tiles_y) {
  TileDiv d;
  d.per = tiles_x * tiles_y;
  d.tiles_x = tiles_x;
  d.m_per = 0xFFFFFFFFu / (uint32_t)d.per + 1u;
  d.m_tx = 0xFFFFFFFFu / (uint32_t)tiles_x + 1u;
  return d;
}
__device__ __forceinline__ TilePos tile_pos(int tile, const TileDiv& d, int TW, int TH) {
  TilePos t;
  t.b = (d.per == 1) ? tile : (int)__umulhi((uint32_t)tile, d.m_per);
  const int rem = tile - t.b * d.per;
  const int ry = (d.tiles_x == 1) ? rem : (int)__umulhi((uint32_t)rem, d.m_tx);
  t.ty0 = ry * TH;
  t.tx0 = (rem - ry * d.tiles_x) * TW;
  return t;
}

// =====================================================================================================================
// part A
// =====================================================================================================================
template <int TW, int TH>
struct CfgA {
  static constexpr int NP = TW * TH;
  // one channel group of an operand tile: [128 rows][8] bf16, + 16 bytes of padding: the tile loader's lanes write
  // consecutive channel groups of one row (2048-byte stride = the same banks for all 14 lanes: ncu counted 3.1 M
  // conflict wavefronts of 3.8 M at P3)
  static constexpr int kGroup = 128 * 16 + 16;
  static constexpr int offGy = 0;                          // dy tile, 14 groups
  static constexpr int offD = NG * kGroup;                 // d tile, 16 groups (14: all-ones column, 15: zeros); must follow dy
  static constexpr int offB = up128(offD + 16 * kGroup);
  static constexpr int offCoef = offB + up128(C * C * 2);
  static constexpr int offBar = offCoef + 3 * C * 4;
  static constexpr int kBytes = offBar + 64;
  static constexpr int kLoadThreads = NG * TW;
  // rows per register chunk of the load phase.  Large levels (16-wide tiles, several tiles per CTA, 2 CTAs / SM at <= 128
  // registers): 4.  Small levels (P5-P7: ONE tile per CTA, fewer CTAs than SMs): the whole tile at once — their time is the
  // latency of the gather's dependent load groups (ncu: 7 - 17 warps stalled on the long scoreboard per issue), which chunking
  // multiplies; the kernel is then compiled for one CTA per SM so that the TH x 4 gather accumulators fit without spills.
  static constexpr bool kSmall = TW < 16;
  static constexpr int RC = kSmall ? TH : ((TH % 4 == 0) ? 4 : 2);
  static constexpr int kMinCtas = kSmall ? 1 : 2;
  static_assert(NP <= 128 && kLoadThreads <= kThreads && TH % RC == 0, "tile shape");
  static_assert(2 * (kBytes + 1024) <= 233472, "two CTAs per SM");
};

// accumulate cw * (8 bf16 channels in w) into acc
__device__ __forceinline__ void acc8(float2 (&acc)[4], const uint4 w, const float2 cw2) {
  acc[0] = fma2(bf2_to_f2(w.x), cw2, acc[0]);
  acc[1] = fma2(bf2_to_f2(w.y), cw2, acc[1]);
  acc[2] = fma2(bf2_to_f2(w.z), cw2, acc[2]);
  acc[3] = fma2(bf2_to_f2(w.w), cw2, acc[3]);
}
// one pooling window: keep the channels whose recorded arg-max byte equals `id`
__device__ __forceinline__ void acc8_match(float2 (&acc)[4], uint4 w, const uint2 pidx, const uint32_t id, const float2 cw2) {
  const uint32_t m0 = __vcmpeq4(pidx.x, id * 0x01010101u), m1 = __vcmpeq4(pidx.y, id * 0x01010101u);
  w.x &= __byte_perm(m0, 0u, 0x1100);
  w.y &= __byte_perm(m0, 0u, 0x3322);
  w.z &= __byte_perm(m1, 0u, 0x1100);
  w.w &= __byte_perm(m1, 0u, 0x3322);
  acc8(acc, w, cw2);
}

// G[r] += sum over the consumers of (edge weight) * (their dL/du as seen by rows ty0 + tyc + r, column x, channel group cg)
template <int RC>
__device__ __forceinline__ void gather_rows(const NodeBwdP& P, const float (&cw)[3], const TilePos t, const int tyc, const int x,
                                            const int cg, const int H, const int W, float2 (&G)[RC][4]) {
  for (int c = 0; c < P.n_cons; ++c) {
    const ConsP& cs = P.cons[c];
    const bf16* du = reinterpret_cast<const bf16*>(cs.du);
    const float2 cw2 = make_float2(cw[c], cw[c]);
    if (cs.mode == MMD_CONS_SAME) {
      const bf16* src = du + (((long long)t.b * H + t.ty0 + tyc) * W + x) * C + 8 * cg;
#pragma unroll
      for (int r = 0; r < RC; ++r) acc8(G[r], ldg16(src + (long long)r * W * C), cw2);
    } else if (cs.mode == MMD_CONS_UP2) {
      const int W2 = 2 * W;
      const bf16* src = du + (((long long)t.b * 2 * H + 2 * (t.ty0 + tyc)) * W2 + 2 * x) * C + 8 * cg;
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        const bf16* s0 = src + (long long)(2 * r) * W2 * C;
        acc8(G[r], ldg16(s0), cw2);
        acc8(G[r], ldg16(s0 + C), cw2);
        acc8(G[r], ldg16(s0 + (long long)W2 * C), cw2);
        acc8(G[r], ldg16(s0 + (long long)W2 * C + C), cw2);
      }
    } else {
      // consumer pooled this tensor (3x3 stride 2, even sizes: no top / left padding): position (y, x) lies in
      // window rows i = y/2 (wy 0) and y/2 - 1 (wy 2) when y is even, (y-1)/2 (wy 1) when odd; same for columns
      const int cH = H >> 1, cW = W >> 1;
      const bool xodd = (x & 1) != 0;
      const int j0 = x >> 1;
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        const int y = t.ty0 + tyc + r;
        const bool yodd = ((tyc + r) & 1) != 0;   // ty0 is even
        const int i0 = y >> 1;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (a == 1 && (yodd || i0 == 0)) continue;
          const int i = i0 - a;
          const uint32_t wy = yodd ? 1u : (a == 0 ? 0u : 2u);
          const long long rowoff = ((long long)t.b * cH + i) * cW;
          {
            const long long off = (rowoff + j0) * C + 8 * cg;
            acc8_match(G[r], ldg16(du + off), __ldg(reinterpret_cast<const uint2*>(cs.pidx + off)),
                       wy * 3u + (xodd ? 1u : 0u), cw2);
          }
          if (!xodd && j0 > 0) {
            const long long off = (rowoff + j0 - 1) * C + 8 * cg;
            acc8_match(G[r], ldg16(du + off), __ldg(reinterpret_cast<const uint2*>(cs.pidx + off)), wy * 3u + 2u, cw2);
          }
        }
      }
    }
  }
}

template <int TW, int TH>
__global__ void __launch_bounds__(kThreads, CfgA<TW, TH>::kMinCtas) node_bwd_a4_kernel(const __grid_constant__ NodeBwdP P, int packed_off_bwd) {
  using S = CfgA<TW, TH>;
  constexpr int RC = S::RC;
  constexpr uint32_t kTmemCols = 256;   // [0,128): dL/dd accumulator, [128,256): [dW | db] accumulator
  constexpr uint32_t kIdesc1 = tc::make_idesc_bf16(128, C, false, false);
  constexpr uint32_t kIdesc2 = tc::make_idesc_bf16(128, 128, true, true);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_gy = smem + S::offGy;
  unsigned char* s_d = smem + S::offD;
  unsigned char* s_b = smem + S::offB;
  float* s_coef = reinterpret_cast<float*>(smem + S::offCoef);
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + S::offBar);
  uint64_t* bar_w = bar_mma + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.g.H, W = P.g.W;
  const int tiles_x = W / TW, tiles_y = H / TH, ntiles = P.g.B * tiles_x * tiles_y;
  const TileDiv tdiv = make_tile_div(tiles_x, tiles_y);

  if (warp == 0) tc::tmem_alloc(s_tmem, kTmemCols);
  if (tid == 32) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(bar_w, C * C * 2);
    tc::bulk_g2s(s_b, P.packed + packed_off_bwd, C * C * 2, bar_w);
  }
  pdl_wait();
  pdl_trigger();
  float cw[3];
  cons_weights(P, cw);
  bn_bwd_coefs<C>(P, cw, s_coef);
  // constant parts of the operand tiles: rows >= NP are zero (they take part in the position contraction of GEMM 2),
  // channel group 14 of the d tile is the all-ones column (bias gradient), group 15 is zero
  for (int idx = tid; idx < 16 * 128; idx += kThreads) {
    const int grp = idx >> 7, row = idx & 127;
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (grp < NG) {
      if (row >= S::NP) {
        *reinterpret_cast<uint4*>(s_gy + grp * S::kGroup + row * 16) = z;
        *reinterpret_cast<uint4*>(s_d + grp * S::kGroup + row * 16) = z;
      }
    } else {
      if (grp == NG && row < S::NP) z.x = 0x00003f80u;   // bf16 1.0 in element 0 (channel 112)
      *reinterpret_cast<uint4*>(s_d + grp * S::kGroup + row * 16) = z;
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t gy_addr = tc::smem_u32(s_gy), d_addr = tc::smem_u32(s_d), b_addr = tc::smem_u32(s_b);

  const bf16* __restrict__ yraw = reinterpret_cast<const bf16*>(P.out);
  const bf16* __restrict__ dsave = reinterpret_cast<const bf16*>(P.save_d);
  bf16* __restrict__ ddout = reinterpret_cast<bf16*>(P.dd);

  // load-phase role: 8 channels (cg) of tile column tx, all rows
  const bool loader = tid < S::kLoadThreads;
  const int cg = tid % NG, tx = tid / NG;
  float2 cA[4], cB[4], cC[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    cA[e] = *reinterpret_cast<const float2*>(s_coef + 8 * cg + 2 * e);
    cB[e] = *reinterpret_cast<const float2*>(s_coef + C + 8 * cg + 2 * e);
    cC[e] = *reinterpret_cast<const float2*>(s_coef + 2 * C + 8 * cg + 2 * e);
  }
  uint32_t phase = 0;
  int iter = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const TilePos t = tile_pos(tile, tdiv, TW, TH);
    if (loader) {
      const int x = t.tx0 + tx;
#pragma unroll
      for (int tyc = 0; tyc < TH; tyc += RC) {
        float2 G[RC][4];
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) G[r][e] = make_float2(0.f, 0.f);
        gather_rows<RC>(P, cw, t, tyc, x, cg, H, W, G);
        // BatchNorm backward + operand tiles
        const long long off0 = (((long long)t.b * H + t.ty0 + tyc) * W + x) * C + 8 * cg;
#pragma unroll
        for (int r = 0; r < RC; ++r) {
          const long long off = off0 + (long long)r * W * C;
          const uint4 yr = ldg16(yraw + off);
          const uint4 dv = ldg16(dsave + off);
          uint4 pk;
          pk.x = f2_to_bf2(fma2(cA[0], G[r][0], fma2(cB[0], bf2_to_f2(yr.x), cC[0])));
          pk.y = f2_to_bf2(fma2(cA[1], G[r][1], fma2(cB[1], bf2_to_f2(yr.y), cC[1])));
          pk.z = f2_to_bf2(fma2(cA[2], G[r][2], fma2(cB[2], bf2_to_f2(yr.z), cC[2])));
          pk.w = f2_to_bf2(fma2(cA[3], G[r][3], fma2(cB[3], bf2_to_f2(yr.w), cC[3])));
          const int p = (tyc + r) * TW + tx;
          *reinterpret_cast<uint4*>(s_gy + cg * S::kGroup + p * 16) = pk;
          *reinterpret_cast<uint4*>(s_d + cg * S::kGroup + p * 16) = dv;
        }
      }
    }
    tc::fence_async_smem();
    __syncthreads();

    if (tid == 0) {
      if (iter == 0) tc::mbar_wait(bar_w, 0u);   // W^T has landed (only the MMAs read it)
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {   // GEMM 1: K runs over the output channels o
        const uint64_t adesc = tc::make_desc(gy_addr + j * 2 * S::kGroup, S::kGroup, 128);
        const uint64_t bdesc = tc::make_desc(b_addr + j * 2 * (C * 16), C * 16, 128);
        tc::umma_bf16(tmem_base, adesc, bdesc, kIdesc1, j > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int s = 0; s < 128 / 16; ++s) {   // GEMM 2: K runs over the 128 tile rows, 16 per MMA (MN-major views)
        const uint64_t adesc = tc::make_desc(gy_addr + s * 256, 128, S::kGroup);
        const uint64_t bdesc = tc::make_desc(d_addr + s * 256, 128, S::kGroup);
        tc::umma_bf16(tmem_base + 128, adesc, bdesc, kIdesc2, (iter > 0 || s > 0) ? 1u : 0u);
      }
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, phase);
    phase ^= 1u;
    tc::fence_after_sync();

    // epilogue: dL/dd accumulator -> bf16 -> HBM (row = tile position, 56 channels per thread)
    {
      const int row = 32 * (warp & 3) + lane;
      const int col0 = (warp >> 2) * (C / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
      float acc[C / 16][8];
      tc::tmem_ld56(taddr, acc);
      if (row < S::NP) {
        const int ty = row / TW, txx = row - ty * TW;
        bf16* dst = ddout + (((long long)t.b * H + t.ty0 + ty) * W + t.tx0 + txx) * C + col0;
#pragma unroll
        for (int j = 0; j < C / 16; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = tc::pack8_bf16(acc[j]);
      }
    }
    tc::fence_before_sync();   // TMEM reads ordered before the next tile's MMAs (issued after the next barrier)
  }

  // ---- flush [dW | db]: TMEM lane = output channel o, column = input channel i (column 112 = bias gradient)
  __syncthreads();
  tc::fence_after_sync();
  if (iter > 0) {
    const int o = 32 * (warp & 3) + lane;
    const int col0 = (warp >> 2) * 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 128u + (uint32_t)col0;
    // every CTA adds into the same 112 x 113 gradient entries: each starts at another column group (rotation by the
    // block index) so that concurrent CTAs hit different addresses of the L2 atomic units
    const int rot = (int)(blockIdx.x & 7u);
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::tmem_ld8(taddr + 8 * ((j + rot) & 7), acc[j]);
    tc::tmem_ld_wait();
    if (o < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i0 = col0 + 8 * ((j + rot) & 7);
        if (i0 < C) {
          red_add_v4(P.g_pw + o * C + i0, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
          red_add_v4(P.g_pw + o * C + i0 + 4, acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        } else if (i0 == C && P.g_pb) {
          atomicAdd(P.g_pb + o, acc[j][0]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// first-cell projection backward (1x1 conv Cin -> C + BatchNorm), tensor cores
// =====================================================================================================================
// Same skeleton as part A: dy tile from the gathered gradient, then  dx[p][i] = sum_o dy[p][o] W[o][i]  (GEMM 1, N = one
// chunk of NC input channels) and  dW[o][i] += sum_p dy[p][o] x[p][i]  (GEMM 2, the x tile carries an all-ones column for
// the bias gradient).  blockIdx.y = input-channel chunk (see proj_chunks()).
template <int TW, int TH, int NC>
struct CfgP {
  static constexpr int NP = TW * TH;
  static constexpr int kGroup = 128 * 16 + 16;             // dy tile: padded like CfgA::kGroup
  static constexpr int kXStride = 128 * 16 + 16;           // padded: the tile loader writes consecutive channel groups
  static constexpr int NXG = NC / 8 + 2;                   // x tile groups: NC/8 data, one all-ones column group, one zero
  static constexpr int N2 = NC + 16;
  static constexpr int offGy = 0;
  static constexpr int offX = NG * kGroup;                 // must follow dy (GEMM 2 reads 16 channel groups of "dy")
  static constexpr int offB = up128(offX + NXG * kXStride);
  static constexpr int kBBytes = (C / 8) * NC * 16;
  static constexpr int offCoef = up128(offB + kBBytes);
  static constexpr int offBar = offCoef + 3 * C * 4;
  static constexpr int kBytes = offBar + 64;
  static constexpr uint32_t kTmemCols = (NC + N2 <= 128) ? 128u : 512u;
  static constexpr uint32_t kColW = kTmemCols / 2;         // [0, NC): dx accumulator, [kColW, kColW + N2): [dW | db]
  static constexpr int kLoadThreads = NG * TW;
  static constexpr int RC = (TH % 4 == 0) ? 4 : 2;
  static_assert(NP <= 128 && NC % 16 == 0 && N2 <= 256 && NC <= (int)kColW && N2 <= (int)kColW, "shape");
  static_assert(kBytes <= 232448, "shared memory");
};

template <int TW, int TH, int NC>
__global__ void __launch_bounds__(kThreads, 1) proj_bwd4_kernel(const __grid_constant__ NodeBwdP P, int packed_off_bwd) {
  using S = CfgP<TW, TH, NC>;
  constexpr int RC = S::RC;
  constexpr uint32_t kIdesc1 = tc::make_idesc_bf16(128, NC, false, false);
  constexpr uint32_t kIdesc2 = tc::make_idesc_bf16(128, S::N2, true, true);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_gy = smem + S::offGy;
  unsigned char* s_x = smem + S::offX;
  unsigned char* s_b = smem + S::offB;
  float* s_coef = reinterpret_cast<float*>(smem + S::offCoef);
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + S::offBar);
  uint64_t* bar_w = bar_mma + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.g.H, W = P.g.W, Cin = P.Cin;
  const int chunk = blockIdx.y, cbase = chunk * NC;
  const int valid = (Cin - cbase < NC) ? Cin - cbase : NC;     // real input channels of this chunk (multiple of 8)
  const int tiles_x = W / TW, tiles_y = H / TH, ntiles = P.g.B * tiles_x * tiles_y;
  const TileDiv tdiv = make_tile_div(tiles_x, tiles_y);

  if (warp == 0) tc::tmem_alloc(s_tmem, S::kTmemCols);
  if (tid == 32) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(bar_w, S::kBBytes);
    tc::bulk_g2s(s_b, P.packed + packed_off_bwd + (size_t)chunk * S::kBBytes, S::kBBytes, bar_w);
  }
  pdl_wait();
  pdl_trigger();
  float cw[3];
  cons_weights(P, cw);
  bn_bwd_coefs<C>(P, cw, s_coef);
  // constant parts: dy rows >= NP are zero; x tile: rows >= NP zero, channels >= valid zero, ones column, zero group
  for (int idx = tid; idx < NG * 128; idx += kThreads) {
    const int grp = idx >> 7, row = idx & 127;
    if (row >= S::NP) *reinterpret_cast<uint4*>(s_gy + grp * S::kGroup + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int idx = tid; idx < S::NXG * 128; idx += kThreads) {
    const int grp = idx >> 7, row = idx & 127;
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (grp == NC / 8 && row < S::NP) z.x = 0x00003f80u;   // bf16 1.0 in column NC
    if (grp >= valid / 8 || row >= S::NP) *reinterpret_cast<uint4*>(s_x + grp * S::kXStride + row * 16) = z;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t gy_addr = tc::smem_u32(s_gy), x_addr = tc::smem_u32(s_x), b_addr = tc::smem_u32(s_b);

  const bf16* __restrict__ yraw = reinterpret_cast<const bf16*>(P.out);
  const bf16* __restrict__ xin = reinterpret_cast<const bf16*>(P.in[0].data);
  bf16* __restrict__ dx = reinterpret_cast<bf16*>(P.dx);

  const bool loader = tid < S::kLoadThreads;
  const int cg = tid % NG, tx = tid / NG;
  float2 cA[4], cB[4], cC[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    cA[e] = *reinterpret_cast<const float2*>(s_coef + 8 * cg + 2 * e);
    cB[e] = *reinterpret_cast<const float2*>(s_coef + C + 8 * cg + 2 * e);
    cC[e] = *reinterpret_cast<const float2*>(s_coef + 2 * C + 8 * cg + 2 * e);
  }
  uint32_t phase = 0;
  int iter = 0;
  const int vg = valid / 8;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const TilePos t = tile_pos(tile, tdiv, TW, TH);
    // x tile: [channel group][tile position][8]
    for (int idx = tid; idx < S::NP * vg; idx += kThreads) {
      const int p = idx / vg, g = idx - p * vg;
      const int ty = p / TW, txx = p - ty * TW;
      const uint4 v = ldg16(xin + (((long long)t.b * H + t.ty0 + ty) * W + t.tx0 + txx) * Cin + cbase + 8 * g);
      *reinterpret_cast<uint4*>(s_x + g * S::kXStride + p * 16) = v;
    }
    if (loader) {
      const int x = t.tx0 + tx;
#pragma unroll
      for (int tyc = 0; tyc < TH; tyc += RC) {
        float2 G[RC][4];
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) G[r][e] = make_float2(0.f, 0.f);
        gather_rows<RC>(P, cw, t, tyc, x, cg, H, W, G);
        const long long off0 = (((long long)t.b * H + t.ty0 + tyc) * W + x) * C + 8 * cg;
#pragma unroll
        for (int r = 0; r < RC; ++r) {
          const uint4 yr = ldg16(yraw + off0 + (long long)r * W * C);
          uint4 pk;
          pk.x = f2_to_bf2(fma2(cA[0], G[r][0], fma2(cB[0], bf2_to_f2(yr.x), cC[0])));
          pk.y = f2_to_bf2(fma2(cA[1], G[r][1], fma2(cB[1], bf2_to_f2(yr.y), cC[1])));
          pk.z = f2_to_bf2(fma2(cA[2], G[r][2], fma2(cB[2], bf2_to_f2(yr.z), cC[2])));
          pk.w = f2_to_bf2(fma2(cA[3], G[r][3], fma2(cB[3], bf2_to_f2(yr.w), cC[3])));
          *reinterpret_cast<uint4*>(s_gy + cg * S::kGroup + ((tyc + r) * TW + tx) * 16) = pk;
        }
      }
    }
    tc::fence_async_smem();
    __syncthreads();

    if (tid == 0) {
      if (iter == 0) tc::mbar_wait(bar_w, 0u);
      tc::fence_after_sync();
      if (dx != nullptr) {
#pragma unroll
        for (int j = 0; j < C / 16; ++j) {   // GEMM 1: K runs over the output channels o
          const uint64_t adesc = tc::make_desc(gy_addr + j * 2 * S::kGroup, S::kGroup, 128);
          const uint64_t bdesc = tc::make_desc(b_addr + j * 2 * (NC * 16), NC * 16, 128);
          tc::umma_bf16(tmem_base, adesc, bdesc, kIdesc1, j > 0 ? 1u : 0u);
        }
      }
#pragma unroll
      for (int s = 0; s < 128 / 16; ++s) {   // GEMM 2: K runs over the 128 tile rows (MN-major views)
        const uint64_t adesc = tc::make_desc(gy_addr + s * 256, 128, S::kGroup);
        const uint64_t bdesc = tc::make_desc(x_addr + s * 256, 128, S::kXStride);
        tc::umma_bf16(tmem_base + S::kColW, adesc, bdesc, kIdesc2, (iter > 0 || s > 0) ? 1u : 0u);
      }
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, phase);
    phase ^= 1u;
    tc::fence_after_sync();

    if (dx != nullptr) {   // dx accumulator -> bf16 -> HBM (row = tile position, NC/2 input channels per thread)
      const int row = 32 * (warp & 3) + lane;
      const int col0 = (warp >> 2) * (NC / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
      float acc[NC / 16][8];
#pragma unroll
      for (int j = 0; j < NC / 16; ++j) tc::tmem_ld8(taddr + 8 * j, acc[j]);
      tc::tmem_ld_wait();
      if (row < S::NP) {
        const int ty = row / TW, txx = row - ty * TW;
        bf16* dst = dx + (((long long)t.b * H + t.ty0 + ty) * W + t.tx0 + txx) * Cin + cbase + col0;
#pragma unroll
        for (int j = 0; j < NC / 16; ++j) {
          if (col0 + 8 * j < valid) {
            if (P.accumulate_dx) {
              const uint4 old = *reinterpret_cast<const uint4*>(dst + 8 * j);
              const float2 o0 = bf2_to_f2(old.x), o1 = bf2_to_f2(old.y), o2 = bf2_to_f2(old.z), o3 = bf2_to_f2(old.w);
              acc[j][0] += o0.x; acc[j][1] += o0.y; acc[j][2] += o1.x; acc[j][3] += o1.y;
              acc[j][4] += o2.x; acc[j][5] += o2.y; acc[j][6] += o3.x; acc[j][7] += o3.y;
            }
            *reinterpret_cast<uint4*>(dst + 8 * j) = tc::pack8_bf16(acc[j]);
          }
        }
      }
    }
    tc::fence_before_sync();
  }

  // ---- flush [dW | db]: TMEM lane = output channel o, column = input channel of this chunk (column NC = bias gradient)
  __syncthreads();
  tc::fence_after_sync();
  if (iter > 0) {
    const int o = 32 * (warp & 3) + lane;
    const int col0 = (warp >> 2) * (S::N2 / 2);
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + S::kColW + (uint32_t)col0;
    float acc[S::N2 / 16][8];
#pragma unroll
    for (int j = 0; j < S::N2 / 16; ++j) tc::tmem_ld8(taddr + 8 * j, acc[j]);
    tc::tmem_ld_wait();
    if (o < C) {
#pragma unroll
      for (int j = 0; j < S::N2 / 16; ++j) {
        const int i0 = col0 + 8 * j;
        if (i0 < valid) {
          float* dst = P.g_pw + (long long)o * Cin + cbase + i0;
          red_add_v4(dst, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
          red_add_v4(dst + 4, acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        } else if (i0 == NC && chunk == 0 && P.g_pb) {
          atomicAdd(P.g_pb + o, acc[j][0]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, S::kTmemCols);
}

// =====================================================================================================================
// part B
// =====================================================================================================================
template <int TW, int TH>
struct CfgB {
  static constexpr int HW2 = TW + 2, HH2 = TH + 2;
  static constexpr int kDD = up128(HH2 * HW2 * POS);
  static constexpr int kX = up128(TW * TH * POS);
  static constexpr int kBuf = kDD + 2 * kX;                    // dL/dd halo tile | input 0 / dL/du tile | input 1 tile
  static constexpr int offTaps = 2 * kBuf;
  static constexpr int offCoef = offTaps + 9 * C * 4;
  static constexpr int offBar = offCoef + 3 * C * 4;
  static constexpr int kBytes = offBar + 64;
  static constexpr int kActive = NPR * (TW / 2);               // channel pairs x column pairs
  static constexpr int kCT = up32(kActive);                    // compute threads (whole warps)
  static constexpr int kBlock = kCT + 32;                      // + the producer warp
  static constexpr int NACC = 18 + 12;                         // dK[9][2] | S1, S2(x0), S2(x1), S1p, S2p, S2mid
  static_assert(kBytes <= 232448, "shared memory");
  static_assert(kActive * NACC * 4 <= kBuf, "reduction scratch must fit in buffer 0");
};

template <int TW, int TH>
__device__ __forceinline__ void issue_tile_b(unsigned char* buf, const NodeBwdP& P, int x1mode, const void* x1src,
                                             const TilePos t, int H, int W, int lane, uint64_t* bar) {
  using S = CfgB<TW, TH>;
  const bf16* dd = reinterpret_cast<const bf16*>(P.dd);
  const bf16* x0 = reinterpret_cast<const bf16*>(P.in[0].data);
  const bf16* x1 = reinterpret_cast<const bf16*>(x1src);
  const int rows_lo = (t.ty0 == 0) ? 1 : 0, rows_hi = (t.ty0 + TH >= H) ? S::HH2 - 2 : S::HH2 - 1;
  const int col_lo = (t.tx0 == 0) ? 1 : 0, col_hi = (t.tx0 + TW >= W) ? S::HW2 - 2 : S::HW2 - 1;
  const uint32_t rb = (uint32_t)(col_hi - col_lo + 1) * POS;
  const int nrows = rows_hi - rows_lo + 1;
  uint32_t total = (uint32_t)nrows * rb + TH * TW * POS;
  if (x1mode == X1_SAME || x1mode == X1_AUX) total += TH * TW * POS;
  else if (x1mode == X1_UP2) total += (TH / 2) * (TW / 2) * POS;
  if (lane == 0) tc::mbar_expect_tx(bar, total);
  __syncwarp();
  if (lane < nrows) {
    const int r = rows_lo + lane;
    tc::bulk_g2s(buf + (r * S::HW2 + col_lo) * POS, dd + (((long long)t.b * H + t.ty0 - 1 + r) * W + t.tx0 - 1 + col_lo) * C,
                 rb, bar);
  }
  if (lane < TH) {
    const long long g = (((long long)t.b * H + t.ty0 + lane) * W + t.tx0) * C;
    tc::bulk_g2s(buf + S::kDD + lane * (TW * POS), x0 + g, TW * POS, bar);
    if (x1mode == X1_SAME || x1mode == X1_AUX) tc::bulk_g2s(buf + S::kDD + S::kX + lane * (TW * POS), x1 + g, TW * POS, bar);
  }
  if (x1mode == X1_UP2 && lane < TH / 2) {
    const int H1 = H >> 1, W1 = W >> 1;
    tc::bulk_g2s(buf + S::kDD + S::kX + lane * ((TW / 2) * POS),
                 x1 + (((long long)t.b * H1 + (t.ty0 >> 1) + lane) * W1 + (t.tx0 >> 1)) * C, (TW / 2) * POS, bar);
  }
}

struct AccB {
  float2 dK[9];
  float2 s1, s2a, s2b, s1p, s2p, s2m;
};

// one tile of the compute warps.  X1M: how input 1 is staged; SW: the node applies swish to the fused sum
template <int TW, int TH, int X1M, bool SW>
__device__ __forceinline__ void compute_tile_b(unsigned char* buf, const float2 (&K)[9], const float2 a0, const float2 a1,
                                               const float2 sh, AccB& A, const int pr, const int cp, const TilePos t,
                                               const int H, const int W, const bf16* mid, const bf16* praw,
                                               const unsigned char* pidx) {
  using S = CfgB<TW, TH>;
  const int c0 = 2 * cp;
  const bool zl = (cp == 0) && (t.tx0 == 0), zr = (cp == TW / 2 - 1) && (t.tx0 + TW >= W);
  const bool ztop = (t.ty0 == 0), zbot = (t.ty0 + TH >= H);
  const unsigned char* ddp = buf + c0 * POS + pr * 4;
  unsigned char* x0p = buf + S::kDD + c0 * POS + pr * 4;
  const unsigned char* x1p = buf + S::kDD + S::kX + ((X1M == X1_UP2) ? cp * POS : c0 * POS) + pr * 4;
  const long long gpos0 = (((long long)t.b * H + t.ty0) * W + t.tx0 + c0) * C + 2 * pr;   // element offset of (row 0, col c0)
  float2 win[3][4];
  float2 x1u = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < S::HH2; ++r) {
    // halo row r of dL/dd -> window slot r % 3 (positions outside the image read as zero)
    {
      const bool zrow = (r == 0 && ztop) || (r == S::HH2 - 1 && zbot);
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(ddp + (r * S::HW2 + dx) * POS);
        if (zrow || (dx == 0 && zl) || (dx == 3 && zr)) w = 0u;
        win[r % 3][dx] = bf2_to_f2(w);
      }
    }
    if (r < 2) continue;
    const int o = r - 2;   // output row; window rows: top = o % 3, mid = (o + 1) % 3, bottom = (o + 2) % 3
    uint32_t midw[2] = {0u, 0u}, praww[2] = {0u, 0u}, pidw[2] = {0x0909u, 0x0909u};
    if (X1M == X1_AUX) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const long long g = gpos0 + ((long long)o * W + c) * C;
        if (mid != nullptr) midw[c] = __ldg(reinterpret_cast<const uint32_t*>(mid + g));
        if (praw != nullptr) {
          praww[c] = __ldg(reinterpret_cast<const uint32_t*>(praw + g));
          pidw[c] = __ldg(reinterpret_cast<const unsigned short*>(pidx + g));
        }
      }
    }
    if (X1M == X1_UP2 && (o & 1) == 0)
      x1u = bf2_to_f2(*reinterpret_cast<const uint32_t*>(x1p + ((o >> 1) * (TW / 2)) * POS));
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      // dL/dv = sum_{ky,kx} K[ky][kx] * dd[o + 2 - ky][c + 2 - kx]  (halo coordinates)
      float2 dv = mul2(K[0], win[(o + 2) % 3][c + 2]);
      dv = fma2(K[1], win[(o + 2) % 3][c + 1], dv);
      dv = fma2(K[2], win[(o + 2) % 3][c], dv);
      dv = fma2(K[3], win[(o + 1) % 3][c + 2], dv);
      dv = fma2(K[4], win[(o + 1) % 3][c + 1], dv);
      dv = fma2(K[5], win[(o + 1) % 3][c], dv);
      dv = fma2(K[6], win[o % 3][c + 2], dv);
      dv = fma2(K[7], win[o % 3][c + 1], dv);
      dv = fma2(K[8], win[o % 3][c], dv);
      // rebuild u, v = swish(u), g = swish'(u)
      unsigned char* cell = x0p + (o * TW + c) * POS;
      const float2 x0 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(cell));
      float2 x1 = make_float2(0.f, 0.f);
      if (X1M == X1_SAME || X1M == X1_AUX) x1 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(x1p + (o * TW + c) * POS));
      if (X1M == X1_UP2) x1 = x1u;
      float2 u = fma2(a0, x0, sh);
      if (X1M != X1_NONE) u = fma2(a1, x1, u);
      float2 v = u, du = dv;
      if (SW) {
        const float2 h = mul2(u, make_float2(0.5f, 0.5f));
        float tx_, ty_;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx_) : "f"(h.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty_) : "f"(h.y));
        const float2 th = make_float2(tx_, ty_);
        const float2 sg = fma2(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));      // sigmoid(u)
        const float2 ns = fma2(th, make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));    // 1 - sigmoid(u)
        v = mul2(u, sg);
        du = mul2(dv, fma2(v, ns, sg));                                                    // sg * (1 + u * (1 - sg))
      }
      // depthwise weight gradient: dK[ky][kx] += dd[o + 2 - ky][c + 2 - kx] * v
      A.dK[0] = fma2(win[(o + 2) % 3][c + 2], v, A.dK[0]);
      A.dK[1] = fma2(win[(o + 2) % 3][c + 1], v, A.dK[1]);
      A.dK[2] = fma2(win[(o + 2) % 3][c], v, A.dK[2]);
      A.dK[3] = fma2(win[(o + 1) % 3][c + 2], v, A.dK[3]);
      A.dK[4] = fma2(win[(o + 1) % 3][c + 1], v, A.dK[4]);
      A.dK[5] = fma2(win[(o + 1) % 3][c], v, A.dK[5]);
      A.dK[6] = fma2(win[o % 3][c + 2], v, A.dK[6]);
      A.dK[7] = fma2(win[o % 3][c + 1], v, A.dK[7]);
      A.dK[8] = fma2(win[o % 3][c], v, A.dK[8]);
      // per-edge sums for the producers' BatchNorm backward / the fusion-weight gradient
      A.s1 = add2(A.s1, du);
      A.s2a = fma2(du, x0, A.s2a);
      if (X1M == X1_SAME || X1M == X1_UP2) A.s2b = fma2(du, x1, A.s2b);
      if (X1M == X1_AUX) {
        A.s2m = fma2(du, bf2_to_f2(midw[c]), A.s2m);
        const float2 m = make_float2((pidw[c] & 0xffu) == 9u ? 0.f : 1.f, ((pidw[c] >> 8) & 0xffu) == 9u ? 0.f : 1.f);
        const float2 dm = mul2(du, m);
        A.s1p = add2(A.s1p, dm);
        A.s2p = fma2(dm, bf2_to_f2(praww[c]), A.s2p);
      }
      *reinterpret_cast<uint32_t*>(cell) = f2_to_bf2(du);   // dL/du in place over the raw input-0 tile
    }
  }
}

template <int TW, int TH, int X1M, bool SW>
__global__ void __launch_bounds__(CfgB<TW, TH>::kBlock, 1) node_bwd_b4_kernel(const __grid_constant__ NodeBwdP P) {
  using S = CfgB<TW, TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* s_taps = reinterpret_cast<float*>(smem + S::offTaps);   // [9][C]
  float* s_coef = reinterpret_cast<float*>(smem + S::offCoef);   // a0 | a1 | shift
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + S::offBar);   // [2]
  uint64_t* bar_ready = bar_full + 2;                                   // [2] compute warps are done with a buffer
  __shared__ int s_flag;
  __shared__ float s_gw[3];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.g.H, W = P.g.W;
  const int tiles_x = W / TW, tiles_y = H / TH, ntiles = P.g.B * tiles_x * tiles_y;
  const TileDiv tdiv = make_tile_div(tiles_x, tiles_y);
  const bool producer = (warp == S::kCT / 32);

  // which inputs are what
  int pi = -1, mi = -1;   // pooled input, the other same-resolution input of a pooled node
  if (X1M == X1_AUX) {
    for (int i = 1; i < P.n_in; ++i) {
      if (P.mode[i] == MMD_IN_POOL) pi = i;
      else mi = i;
    }
  }
  const void* x1src = (X1M == X1_AUX) ? P.aux : P.in[1].data;

  float wgt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) wgt[i] = (i < P.n_in) ? fusion_weight(P.fw, P.n_in, i, P.fw_eps) : 0.f;
  if (producer && lane == 0) {   // the producer warp owns the barriers: it can start loading right after pdl_wait()
    tc::mbar_init(bar_full, 1);
    tc::mbar_init(bar_full + 1, 1);
    tc::mbar_init(bar_ready, 1);
    tc::mbar_init(bar_ready + 1, 1);
    tc::fence_mbar_init();
  }
  for (int idx = tid; idx < 9 * C; idx += S::kBlock) {
    const int c = idx / 9, tap = idx - c * 9;
    s_taps[tap * C + c] = P.dw_w[idx];
  }
  pdl_wait();
  pdl_trigger();
  if (producer) {   // first two tiles: in flight while the coefficients below are set up
    __syncwarp();
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles && k < 2; tile += gridDim.x, ++k)
      issue_tile_b<TW, TH>(smem + k * S::kBuf, P, X1M, x1src, tile_pos(tile, tdiv, TW, TH), H, W, lane, bar_full + k);
  }
  if (tid < C) {
    const float* bn0 = P.in[0].bn;
    const float sc0 = bn0 ? bn0[tid] : 1.f, sh0 = bn0 ? bn0[C + tid] : 0.f;
    float a1 = 0.f, shift = sh0 * wgt[0];
    if (X1M == X1_AUX) {
      a1 = 1.f;
    } else if (X1M != X1_NONE) {
      const float* bn1 = P.in[1].bn;
      const float sc1 = bn1 ? bn1[tid] : 1.f, sh1 = bn1 ? bn1[C + tid] : 0.f;
      a1 = sc1 * wgt[1];
      shift = fmaf(sh1, wgt[1], shift);
    }
    s_coef[tid] = sc0 * wgt[0];
    s_coef[C + tid] = a1;
    s_coef[2 * C + tid] = shift;
  }
  __syncthreads();

  AccB A;
#pragma unroll
  for (int k = 0; k < 9; ++k) A.dK[k] = make_float2(0.f, 0.f);
  A.s1 = A.s2a = A.s2b = A.s1p = A.s2p = A.s2m = make_float2(0.f, 0.f);
  const int pr = tid % NPR, cp = tid / NPR;

  if (producer) {
    // ---- producer warp: bulk loads two tiles ahead, bulk stores of finished dL/du tiles
    bf16* duout = reinterpret_cast<bf16*>(P.du);
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const TilePos t = tile_pos(tile, tdiv, TW, TH);
      tc::mbar_wait(bar_ready + s, (uint32_t)((k >> 1) & 1));
      if (lane < TH)
        bulk_s2g(duout + (((long long)t.b * H + t.ty0 + lane) * W + t.tx0) * C, smem + s * S::kBuf + S::kDD + lane * (TW * POS),
                 TW * POS);
      bulk_commit();
      bulk_wait_read0();
      __syncwarp();
      const int nxt = tile + 2 * gridDim.x;
      if (nxt < ntiles)
        issue_tile_b<TW, TH>(smem + s * S::kBuf, P, X1M, x1src, tile_pos(nxt, tdiv, TW, TH), H, W, lane, bar_full + s);
    }
  } else {
    // ---- compute warps
    const bool active = tid < S::kActive;
    float2 K[9], a0 = make_float2(0.f, 0.f), a1 = a0, sh = a0;
    if (active) {
#pragma unroll
      for (int k9 = 0; k9 < 9; ++k9) K[k9] = *reinterpret_cast<const float2*>(s_taps + k9 * C + 2 * pr);
      a0 = *reinterpret_cast<const float2*>(s_coef + 2 * pr);
      a1 = *reinterpret_cast<const float2*>(s_coef + C + 2 * pr);
      sh = *reinterpret_cast<const float2*>(s_coef + 2 * C + 2 * pr);
    }
    const bf16* mid = (X1M == X1_AUX && mi >= 0) ? reinterpret_cast<const bf16*>(P.in[mi].data) : nullptr;
    const bf16* praw = (X1M == X1_AUX && pi >= 0) ? reinterpret_cast<const bf16*>(P.praw) : nullptr;
    const unsigned char* pidx = (X1M == X1_AUX && pi >= 0) ? P.pidx[pi] : nullptr;
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      tc::mbar_wait(bar_full + s, (uint32_t)((k >> 1) & 1));
      if (active)
        compute_tile_b<TW, TH, X1M, SW>(smem + s * S::kBuf, K, a0, a1, sh, A, pr, cp, tile_pos(tile, tdiv, TW, TH), H,
                                        W, mid, praw, pidx);
      tc::fence_async_smem();   // dL/du tile (generic writes) -> visible to the bulk store engine
      asm volatile("bar.sync 1, %0;" ::"n"(S::kCT) : "memory");
      if (tid == 0) mbar_arrive(bar_ready + s);
    }
  }
  __syncthreads();

  // ---- block reduction over the column-pair threads that share a channel pair, then global accumulation
  float* s_red = reinterpret_cast<float*>(smem);   // [kActive][NACC]
  if (!producer && tid < S::kActive) {
    float* r = s_red + tid * S::NACC;
#pragma unroll
    for (int k9 = 0; k9 < 9; ++k9) { r[2 * k9] = A.dK[k9].x; r[2 * k9 + 1] = A.dK[k9].y; }
    r[18] = A.s1.x; r[19] = A.s1.y; r[20] = A.s2a.x; r[21] = A.s2a.y; r[22] = A.s2b.x; r[23] = A.s2b.y;
    r[24] = A.s1p.x; r[25] = A.s1p.y; r[26] = A.s2p.x; r[27] = A.s2p.y; r[28] = A.s2m.x; r[29] = A.s2m.y;
  }
  __syncthreads();
  for (int idx = tid; idx < NPR * S::NACC; idx += S::kBlock) {
    const int p2 = idx / S::NACC, e = idx - p2 * S::NACC;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < TW / 2; ++j) s += s_red[(j * NPR + p2) * S::NACC + e];
    s_red[S::kActive * S::NACC + idx] = s;   // totals [pair][NACC] behind the partials
  }
  __syncthreads();
  const float* tot = s_red + S::kActive * S::NACC;
  // (all CTAs add into the same addresses: each starts at another offset so that concurrent CTAs spread over the L2
  //  atomic units instead of queueing on one address)
  const int rot9 = (int)((blockIdx.x * 131u) % (unsigned)(C * 9));
  for (int k = tid; k < C * 9; k += S::kBlock) {   // depthwise weight gradient [C][3][3]
    int idx = k + rot9;
    if (idx >= C * 9) idx -= C * 9;
    const int c = idx / 9, t9 = idx - c * 9;
    if (P.g_dw) atomicAdd(P.g_dw + idx, tot[(c >> 1) * S::NACC + 2 * t9 + (c & 1)]);
  }
  if (tid < C) {
    // slots: (sum du*mask, sum du*mask*xhat) per input edge; xhat = (raw - mean) * invstd for a deferred-BN input,
    // the (final) value itself otherwise
    const int c = (tid + (int)((blockIdx.x * 29u) % (unsigned)C)) % C;
    const float* tp = tot + (c >> 1) * S::NACC + (c & 1);
    const double s1 = tp[18], s2a = tp[20], s2b = tp[22], s1p = tp[24], s2p = tp[26], s2m = tp[28];
    for (int i = 0; i < P.n_in; ++i) {
      if (P.in_slot[i] == nullptr) continue;
      double a = s1, b2;
      if (i == 0) b2 = s2a;
      else if (X1M == X1_AUX) {
        if (i == pi) { a = s1p; b2 = s2p; }
        else b2 = s2m;
      } else b2 = s2b;
      const float* bn = P.in[i].bn;
      if (bn != nullptr) b2 = (b2 - (double)bn[2 * C + c] * a) * (double)bn[3 * C + c];
      atomicAdd(P.in_slot[i] + c, a);
      atomicAdd(P.in_slot[i] + C + c, b2);
    }
  }

  // ---- the last CTA turns the slots into the fusion-weight gradient (relu / normalise backward,
  //      src/YetAnotherEfficientDet.py:338-339) — unless that is left to fwgrad_kernel at the end of the backward
  if (P.fw == nullptr || P.g_fw == nullptr || P.defer_fw) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(P.counter, 1u);
    s_flag = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
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

// ---- deferred fusion-weight gradients: one block per node --------------------------------------------------------------
__global__ void __launch_bounds__(128) fwgrad_kernel(const __grid_constant__ FwGradBatch BATCH) {
  const FwGradEntry& E = BATCH.e[blockIdx.x];
  __shared__ float s_gw[3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  if (warp < E.n_in) {
    const int i = warp;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const double s1 = __ldcg(E.slot[i] + c), s2 = __ldcg(E.slot[i] + C + c);
      if (E.in_bn_b[i] != nullptr) acc += (float)((double)E.in_bn_w[i][c] * s2 + (double)E.in_bn_b[i][c] * s1);
      else acc += (float)s2;
    }
    acc = warp_sum(acc);
    if (lane == 0) s_gw[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ssum = 0.f;
    for (int j = 0; j < E.n_in; ++j) ssum += fmaxf(E.fw[j], 0.f);
    const float denom = ssum + E.fw_eps;
    float dot = 0.f;
    for (int j = 0; j < E.n_in; ++j) dot += fmaxf(E.fw[j], 0.f) / denom * s_gw[j];
    for (int k = 0; k < E.n_in; ++k) E.g_fw[k] = (E.fw[k] > 0.f) ? (s_gw[k] - dot) / denom : 0.f;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static int sm_count() { return device_sm_count(); }

static int x1_mode(const NodeBwdP& p) {
  if (p.n_in == 1) return X1_NONE;
  for (int i = 1; i < p.n_in; ++i)
    if (p.mode[i] == MMD_IN_POOL) return X1_AUX;
  return p.mode[1] == MMD_IN_UP2 ? X1_UP2 : X1_SAME;
}

template <int TW, int TH, int X1M, bool SW>
static int launch_b(const NodeBwdP& p, cudaStream_t s) {
  using S = CfgB<TW, TH>;
  MMD_SMEM((node_bwd_b4_kernel<TW, TH, X1M, SW>), S::kBytes);
  const int ntiles = p.g.B * (p.g.H / TH) * (p.g.W / TW);
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  MMD_CUDA(launch_pdl(node_bwd_b4_kernel<TW, TH, X1M, SW>, dim3(grid), dim3(S::kBlock), S::kBytes, s, p));
  MMD_LAUNCH_CHECK();
  return 0;
}

template <int TW, int TH>
static int launch_geom(const NodeBwdP& p, cudaStream_t s) {
  using SA = CfgA<TW, TH>;
  MMD_SMEM((node_bwd_a4_kernel<TW, TH>), SA::kBytes);
  const int ntiles = p.g.B * (p.g.H / TH) * (p.g.W / TW);
  const double bytes = node_algo_bytes(p.in, p.n_in, p.g, C, 2);
  {
    const int cap = SA::kMinCtas * sm_count();   // resident CTAs (small levels: one per SM, whole-tile gather chunks)
    const int grid = ntiles < cap ? ntiles : cap;
    ProfScope prof((TW == 16 && TH == 8) ? PK_NODE_BWD_A_16x8 : PK_NODE_BWD_A, bytes, s);
    MMD_CUDA(launch_pdl(node_bwd_a4_kernel<TW, TH>, dim3(grid), dim3(kThreads), SA::kBytes, s, p,
                        packed_layout(MMD_OP_NODE_FWD, C, C).offBwd));
    MMD_LAUNCH_CHECK();
  }
  ProfScope prof((TW == 16 && TH == 8) ? PK_NODE_BWD_B_16x8 : PK_NODE_BWD_B, bytes, s);
  const bool sw = p.swish != 0;
  switch (x1_mode(p)) {
    case X1_SAME: return sw ? launch_b<TW, TH, X1_SAME, true>(p, s) : launch_b<TW, TH, X1_SAME, false>(p, s);
    case X1_UP2: return sw ? launch_b<TW, TH, X1_UP2, true>(p, s) : launch_b<TW, TH, X1_UP2, false>(p, s);
    case X1_AUX: return sw ? launch_b<TW, TH, X1_AUX, true>(p, s) : launch_b<TW, TH, X1_AUX, false>(p, s);
    default: return sw ? launch_b<TW, TH, X1_NONE, true>(p, s) : launch_b<TW, TH, X1_NONE, false>(p, s);
  }
}

template <int TW, int TH, int NC>
static int launch_proj(const NodeBwdP& p, int nchunks, cudaStream_t s) {
  using S = CfgP<TW, TH, NC>;
  MMD_SMEM((proj_bwd4_kernel<TW, TH, NC>), S::kBytes);
  const int ntiles = p.g.B * (p.g.H / TH) * (p.g.W / TW);
  int gx = sm_count() / nchunks;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  ProfScope prof(PK_PROJ_BWD, 2.0 * p.g.B * p.g.H * p.g.W * (p.Cin + C) * 2.0, s);
  MMD_CUDA(launch_pdl(proj_bwd4_kernel<TW, TH, NC>, dim3(gx, nchunks), dim3(kThreads), S::kBytes, s, p,
                      packed_layout(MMD_OP_PROJ_FWD, p.Cin, C).offBwd));
  MMD_LAUNCH_CHECK();
  return 0;
}

template <int TW, int TH>
static int launch_proj_geom(const NodeBwdP& p, cudaStream_t s) {
  const ProjChunks pc = proj_chunks(p.Cin);
  switch (pc.NC) {
    case 48: return launch_proj<TW, TH, 48>(p, pc.n, s);
    case 128: return launch_proj<TW, TH, 128>(p, pc.n, s);
    default: return launch_proj<TW, TH, 176>(p, pc.n, s);
  }
}

static int pick_geom(int H, int W) {
  if (W % 16 == 0 && H % 8 == 0) return 0;
  // A/B switch (MMD_BWD_GEOM_8x8=1): 24x24 (P5 of the D2 pyramid) as nine 8x8 tiles per sample instead of six 12x8 ones —
  // at B = 32 the student's 192 12x8 tiles are 1.3 waves of the one-CTA-per-SM small-level kernels, 288 8x8 tiles 1.95
  static const bool prefer8 = getenv("MMD_BWD_GEOM_8x8") != nullptr && atoi(getenv("MMD_BWD_GEOM_8x8")) != 0;
  if (prefer8 && W % 8 == 0 && H % 8 == 0) return 2;
  if (W % 12 == 0 && H % 8 == 0) return 1;
  if (W % 8 == 0 && H % 8 == 0) return 2;
  if (W % 12 == 0 && H % 6 == 0) return 3;
  if (W % 6 == 0 && H % 6 == 0) return 4;
  return -1;
}

}  // namespace b4


// ---- streaming SLOT (bf16, SAME mode): sum G and sum G * xhat per channel for the 5 stack outputs -----------------------
// Same thread layout as bnapply_same_bf16_kernel (16 positions x 14 channel groups, fixed channel group per thread, four
// position blocks in flight); per-thread fp32 partials, block reduction through shared memory, 2 * C double atomics per
// block.  The generic kernel (bifpn_bwd.cu) runs at ~0.2 of the HBM rate (8-byte loads, 64-bit index divisions).
namespace b4 {
constexpr int kSlotThreads = 224;
__global__ void __launch_bounds__(kSlotThreads) slot_same_bf16_kernel(const __grid_constant__ NodeBwdGroup GROUP) {
  const NodeBwdP& P = GROUP.p[blockIdx.y];
  __shared__ float s_red[16][NG][16];
  const long long npos = (long long)P.g.B * P.g.H * P.g.W;
  const long long nblk = (npos + 15) / 16;
  long long want = (nblk + 15) / 16;   // >= 16 position blocks per CTA: every CTA ends with 2 * C same-address atomics
  if (want < 1) want = 1;
  const long long nctas = want < (long long)gridDim.x ? want : (long long)gridDim.x;
  if ((long long)blockIdx.x >= nctas) return;
  pdl_wait();
  pdl_trigger();
  const int cg = threadIdx.x % NG, pl = threadIdx.x / NG;
  const float* bn = P.in[0].bn;
  float2 mu[4], is[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    mu[e] = *reinterpret_cast<const float2*>(bn + 2 * C + 8 * cg + 2 * e);
    is[e] = *reinterpret_cast<const float2*>(bn + 3 * C + 8 * cg + 2 * e);
  }
  const uint4* __restrict__ G = reinterpret_cast<const uint4*>(P.cons[0].du);
  const uint4* __restrict__ X = reinterpret_cast<const uint4*>(P.in[0].data);
  float2 s1[4], s2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) s1[e] = s2[e] = make_float2(0.f, 0.f);
  constexpr int U = 4;
  for (long long blk = blockIdx.x; blk < nblk; blk += (long long)U * nctas) {
    uint4 g[U], x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pos = (blk + (long long)u * nctas) * 16 + pl;
      g[u] = x[u] = make_uint4(0u, 0u, 0u, 0u);
      if (pos < npos) {
        g[u] = __ldg(G + pos * NG + cg);
        x[u] = __ldg(X + pos * NG + cg);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {   // (out-of-range positions carry g = 0: they add nothing)
      const uint32_t gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w}, xw[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 gg = bf2_to_f2(gw[e]);
        const float2 xh = mul2(add2(bf2_to_f2(xw[e]), make_float2(-mu[e].x, -mu[e].y)), is[e]);
        s1[e] = add2(s1[e], gg);
        s2[e] = fma2(gg, xh, s2[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s_red[pl][cg][2 * e] = s1[e].x; s_red[pl][cg][2 * e + 1] = s1[e].y;
    s_red[pl][cg][8 + 2 * e] = s2[e].x; s_red[pl][cg][8 + 2 * e + 1] = s2[e].y;
  }
  __syncthreads();
  {   // thread (cg, k = pl): channel 8 * cg + (k & 7), quantity k >> 3
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) a += s_red[r][cg][pl];
    const int c = 8 * cg + (pl & 7);
    atomicAdd(P.in_slot[0] + ((pl >> 3) ? C + c : c), (double)a);
  }
}
}  // namespace b4

bool slot_same_bf16_usable(const NodeBwdP* p, int n) {
  for (int i = 0; i < n; ++i) {
    if (p[i].mode[0] != MMD_IN_SAME || p[i].in[0].bn == nullptr || p[i].in_slot[0] == nullptr) return false;
    if ((((uintptr_t)p[i].in[0].data | (uintptr_t)p[i].cons[0].du) & 15u) != 0) return false;
  }
  return true;
}

int launch_slot_same_bf16(const NodeBwdP* p, int n, cudaStream_t s) {
  NodeBwdGroup group;
  double bytes = 0.0;
  long long maxcta = 1;
  for (int i = 0; i < kMaxGroupOps; ++i) group.p[i] = p[i < n ? i : 0];
  for (int i = 0; i < n; ++i) {
    const long long npos = (long long)p[i].g.B * p[i].g.H * p[i].g.W;
    bytes += 2.0 * npos * 112 * 2;
    const long long want = ((npos + 15) / 16 + 15) / 16;
    if (want > maxcta) maxcta = want;
  }
  const long long cap = 4LL * device_sm_count();
  if (maxcta > cap) maxcta = cap;
  ProfScope prof(PK_SLOT, bytes, s);
  MMD_CUDA(launch_pdl(b4::slot_same_bf16_kernel, dim3((unsigned)maxcta, n), dim3(b4::kSlotThreads), 0, s, group));
  MMD_LAUNCH_CHECK();
  return 0;
}

bool fwgrad_deferral_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMD_NO_DEFER_FW");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

int launch_fwgrad(const FwGradEntry* entries, int n, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  for (int i0 = 0; i0 < n; i0 += kFwGradPerLaunch) {
    const int m = (n - i0 < kFwGradPerLaunch) ? n - i0 : kFwGradPerLaunch;
    FwGradBatch batch;
    for (int k = 0; k < kFwGradPerLaunch; ++k) batch.e[k] = entries[i0 + (k < m ? k : 0)];
    MMD_CUDA(launch_pdl(b4::fwgrad_kernel, dim3(m), dim3(128), 0, s, batch));
    MMD_LAUNCH_CHECK();
  }
  return 0;
}

bool bwd_v4_usable(const NodeBwdP& p) {
  if (p.packed == nullptr || p.n_in < 1 || p.n_in > 3 || p.mode[0] != MMD_IN_SAME) return false;
  if (b4::pick_geom(p.g.H, p.g.W) < 0) return false;
  int npool = 0;
  for (int i = 1; i < p.n_in; ++i) {
    if (p.mode[i] == MMD_IN_POOL) {
      ++npool;
      if (p.in[i].H != 2 * p.g.H || p.in[i].W != 2 * p.g.W) return false;
    } else if (p.mode[i] == MMD_IN_UP2) {
      if (i != 1 || p.n_in != 2 || 2 * p.in[i].H != p.g.H || 2 * p.in[i].W != p.g.W) return false;
    }
  }
  if (npool > 1 || (npool == 1 && (p.aux == nullptr || p.praw == nullptr))) return false;
  if (npool == 0 && p.n_in == 3) return false;
  for (int c = 0; c < p.n_cons; ++c) {
    const ConsP& cs = p.cons[c];
    if (cs.mode == MMD_CONS_SAME && (cs.H != p.g.H || cs.W != p.g.W)) return false;
    if (cs.mode == MMD_CONS_UP2 && (cs.H != 2 * p.g.H || cs.W != 2 * p.g.W)) return false;
    if (cs.mode == MMD_CONS_POOL && (2 * cs.H != p.g.H || 2 * cs.W != p.g.W || cs.pidx == nullptr)) return false;
  }
  return p.out && p.out_bn && p.save_d && p.du && p.dd && p.g_pw && p.counter;
}

bool proj_bwd_v4_usable(const NodeBwdP& p) {
  if (p.packed == nullptr || p.Cin < 8 || p.Cin % 8 != 0 || b4::pick_geom(p.g.H, p.g.W) < 0) return false;
  if (p.in[0].data == nullptr || p.out == nullptr || p.out_bn == nullptr || p.g_pw == nullptr) return false;
  if ((((uintptr_t)p.g_pw) & 15u) != 0 || ((p.Cin * 4) & 15) != 0) return false;
  for (int c = 0; c < p.n_cons; ++c) {
    const ConsP& cs = p.cons[c];
    if (cs.mode == MMD_CONS_SAME && (cs.H != p.g.H || cs.W != p.g.W)) return false;
    if (cs.mode == MMD_CONS_UP2 && (cs.H != 2 * p.g.H || cs.W != 2 * p.g.W)) return false;
    if (cs.mode == MMD_CONS_POOL && (2 * cs.H != p.g.H || 2 * cs.W != p.g.W || cs.pidx == nullptr)) return false;
  }
  return true;
}

int launch_proj_bwd_v4(const NodeBwdP& p, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  switch (b4::pick_geom(p.g.H, p.g.W)) {
    case 0: return b4::launch_proj_geom<16, 8>(p, s);
    case 1: return b4::launch_proj_geom<12, 8>(p, s);
    case 2: return b4::launch_proj_geom<8, 8>(p, s);
    case 3: return b4::launch_proj_geom<12, 6>(p, s);
    case 4: return b4::launch_proj_geom<6, 6>(p, s);
  }
  set_error("proj_bwd_v4: no tile shape fits %dx%d", p.g.H, p.g.W);
  return MMD_E_ARG;
}

int launch_node_bwd_v4(const NodeBwdP& p, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  switch (b4::pick_geom(p.g.H, p.g.W)) {
    case 0: return b4::launch_geom<16, 8>(p, s);
    case 1: return b4::launch_geom<12, 8>(p, s);
    case 2: return b4::launch_geom<8, 8>(p, s);
    case 3: return b4::launch_geom<12, 6>(p, s);
    case 4: return b4::launch_geom<6, 6>(p, s);
  }
  set_error("node_bwd_v4: no tile shape fits %dx%d", p.g.H, p.g.W);
  return MMD_E_ARG;
}

}  // namespace mmd
