// BiFPN fusion node, backward part A, bf16 storage: both contractions run on tcgen05 tensor cores.
//
// Per 128-position tile:
//   load      : gather G from the consumers, dy = A*G + Bc*y + Cc (BatchNorm backward), d = saved depthwise output;
//               both are written as bf16 into the core-matrix layout [c/8][p][c%8] (see tc.cuh)
//   GEMM 1    : dL/dd[128 x 112] = dy[128 x 112(o)] * W[o][i]       A = dy tile (K-major), B = W^T resident in smem
//   GEMM 2    : dW[112(o) x 112(i)] += dy^T[o x 128(p)] * d[p x i]   A = the SAME dy tile read as an MN-major operand,
//               B = the d tile read as an MN-major operand; the accumulator stays in TMEM across all tiles of the
//               (persistent) CTA and is flushed once with atomics
//   epilogue  : dL/dd accumulator -> bf16 staging -> 16-byte coalesced stores
#include "bifpn_bwd_common.cuh"
#include "tc.cuh"

namespace mmd {

typedef __nv_bfloat16 bf16;

template <int C>
struct BwdATcSmem {
  static constexpr int kTileBytes = kTileP * C * 2;   // one operand tile [C/8][128][8] bf16 = 28 672 B
  static constexpr int LDS = C + 8;
  static constexpr int offGy = 0;
  static constexpr int offD = offGy + kTileBytes;     // must directly follow dy: GEMM 2 reads M=128 (> 112) row groups
  static constexpr int offB = offD + kTileBytes;
  static constexpr int kBBytes = C * C * 2;
  static constexpr int offCoef = offB + ((kBBytes + 127) / 128) * 128;
  static constexpr int offBar = offCoef + 3 * C * 4;
  static constexpr int kBytes = offBar + 32;
  static_assert(kTileP * LDS * 2 <= 2 * kTileBytes, "staging tile must fit in the two operand tiles");
};

template <int C>
__global__ void __launch_bounds__(kThreads, 2) node_bwd_a_tc_kernel(const __grid_constant__ NodeBwdP P, int packed_off_bwd) {
  using S = BwdATcSmem<C>;
  constexpr int NQ = C / 4, NG = C / 8, LDS = S::LDS;
  constexpr uint32_t kTmemCols = 256;   // [0,128): dL/dd accumulator, [128,256): dW accumulator
  constexpr uint32_t kIdesc1 = tc::make_idesc_bf16(128, C, false, false);
  constexpr uint32_t kIdesc2 = tc::make_idesc_bf16(128, C, true, true);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* s_gy = reinterpret_cast<bf16*>(smem_raw + S::offGy);
  bf16* s_d = reinterpret_cast<bf16*>(smem_raw + S::offD);
  bf16* s_b = reinterpret_cast<bf16*>(smem_raw + S::offB);
  bf16* s_stage = s_gy;  // dL/dd staging aliases the operand tiles once the MMAs have completed
  float* s_coef = reinterpret_cast<float*>(smem_raw + S::offCoef);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + S::offBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem_raw + S::offBar + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TileGeom g = P.g;

  if (warp == 0) tc::tmem_alloc(s_tmem, kTmemCols);
  if (tid == 32) {
    tc::mbar_init(s_bar, 1);
    tc::fence_mbar_init();
  }
  float cw[3];
  cons_weights(P, cw);
  bn_bwd_coefs<C>(P, cw, s_coef);
  // B operand of GEMM 1: B[n = i][k = o] = W[o][i]  ->  bf16 [o/8][i][o%8]: one bulk copy of the block prepared by
  // mmd_bifpn_prep, or (no packed block) converted here from the fp32 parameter
  uint64_t* s_bar_w = s_bar + 2;
  if (P.packed != nullptr) {
    if (tid == 32) {
      tc::mbar_init(s_bar_w, 1);
      tc::fence_mbar_init();
      tc::mbar_expect_tx(s_bar_w, C * C * 2);
      tc::bulk_g2s(s_b, P.packed + packed_off_bwd, C * C * 2, s_bar_w);
    }
  } else {
#pragma unroll 7
    for (int idx = tid; idx < C * C; idx += kThreads) {
      const int o = idx / C, i = idx - o * C;
      s_b[(o >> 3) * (C * 8) + i * 8 + (o & 7)] = __float2bfloat16_rn(P.pw_w[idx]);
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

  float accB = 0.f;
  uint32_t phase = 0;
  int iter = 0;

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++iter) {
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
        const float4 G = pull_grad<bf16, C>(P, cw, b, y, x, q);
        const float4 yr = ld4<bf16>(yraw + off);
        const float4 A = *reinterpret_cast<const float4*>(s_coef + 4 * q);
        const float4 Bc = *reinterpret_cast<const float4*>(s_coef + C + 4 * q);
        const float4 Cc = *reinterpret_cast<const float4*>(s_coef + 2 * C + 4 * q);
        gy = f4_fma(A, G, f4_fma(Bc, yr, Cc));
        d = ld4<bf16>(dsave + off);
      }
      const int so = (q >> 1) * (kTileP * 8) + p * 8 + (q & 1) * 4;
      st4<bf16>(s_gy + so, gy);
      st4<bf16>(s_d + so, d);
    }
    tc::fence_async_smem();
    __syncthreads();

    if (tid == 0) {
      if (iter == 0 && P.packed != nullptr) tc::mbar_wait(s_bar_w, 0u);   // W^T has landed (only the MMAs read it)
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {   // GEMM 1: K runs over the output channels o
        const uint64_t adesc = tc::make_desc(gy_addr + j * 2 * (kTileP * 16), kTileP * 16, 128);
        const uint64_t bdesc = tc::make_desc(b_addr + j * 2 * (C * 16), C * 16, 128);
        tc::umma_bf16(tmem_base, adesc, bdesc, kIdesc1, j > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int s = 0; s < kTileP / 16; ++s) {   // GEMM 2: K runs over the 128 positions, 16 per MMA
        // MN-major view of the same bytes: 8 channels contiguous (16 B), positions 16 B apart, channel groups
        // kTileP*16 B apart (SBO), groups of 8 positions 128 B apart (LBO)
        const uint64_t adesc = tc::make_desc(gy_addr + s * 256, 128, kTileP * 16);
        const uint64_t bdesc = tc::make_desc(d_addr + s * 256, 128, kTileP * 16);
        tc::umma_bf16(tmem_base + 128, adesc, bdesc, kIdesc2, (iter > 0 || s > 0) ? 1u : 0u);
      }
      tc::umma_commit(s_bar);
    }
    if (tid < C && P.g_pb) {   // db[o] = sum_p dy[p][o]  (zero in exact arithmetic for a bias feeding a train-mode BN)
      float s = 0.f;
      const bf16* col = s_gy + (tid >> 3) * (kTileP * 8) + (tid & 7);
      for (int p = 0; p < kTileP; ++p) s += __bfloat162float(col[p * 8]);
      accB += s;
    }
    tc::mbar_wait(s_bar, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    __syncthreads();   // every thread is done reading the operand tiles (db) before they become the staging tile

    {
      const int row = 32 * (warp & 3) + lane;
      const int col0 = (warp >> 2) * (C / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
      float acc[C / 16][8];
#pragma unroll
      for (int j = 0; j < C / 16; ++j) tc::tmem_ld8(taddr + 8 * j, acc[j]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < C / 16; ++j)
        *reinterpret_cast<uint4*>(s_stage + row * LDS + col0 + 8 * j) = tc::pack8_bf16(acc[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    for (int idx = tid; idx < g.TH * g.TW * NG; idx += kThreads) {
      const int p = idx / NG, gq = idx - p * NG;
      const int ty = p / g.TW, tx = p - ty * g.TW;
      if (ty < th && tx < tw)
        *reinterpret_cast<uint4*>(ddout + (((long long)b * g.H + ty0 + ty) * g.W + tx0 + tx) * C + 8 * gq) =
            *reinterpret_cast<const uint4*>(s_stage + p * LDS + 8 * gq);
    }
    __syncthreads();
  }

  // ---- flush the dW accumulator: TMEM lane = output channel o, column = input channel i
  tc::fence_after_sync();
  {
    const int o = 32 * (warp & 3) + lane;
    const int col0 = (warp >> 2) * (C / 2);
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 128u + (uint32_t)col0;
    float acc[C / 16][8];
#pragma unroll
    for (int j = 0; j < C / 16; ++j) tc::tmem_ld8(taddr + 8 * j, acc[j]);
    tc::tmem_ld_wait();
    if (o < C) {
#pragma unroll
      for (int j = 0; j < C / 16; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(P.g_pw + o * C + col0 + 8 * j + e, acc[j][e]);
    }
  }
  if (tid < C && P.g_pb) atomicAdd(P.g_pb + tid, accB);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

int launch_node_bwd_a_tc(const NodeBwdP& p, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  constexpr int CC = 112;
  const size_t smem = BwdATcSmem<CC>::kBytes;
  MMD_SMEM((node_bwd_a_tc_kernel<CC>), smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.g.ntiles < 2 * sms ? p.g.ntiles : 2 * sms;
  ProfScope prof(PK_NODE_BWD_A, node_algo_bytes(p.in, p.n_in, p.g, C, 2), s);
  node_bwd_a_tc_kernel<CC><<<grid, kThreads, smem, s>>>(p, packed_layout(MMD_OP_NODE_FWD, C, C).offBwd);
  MMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmd
