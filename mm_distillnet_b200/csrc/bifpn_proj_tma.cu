// First-cell 1x1 projections Cin -> 112 (+bias, BatchNorm statistics), forward, bf16 — warp-specialised TMA pipeline (sm_100a).
//
// The projection is a plain [positions x Cin] x [Cin x 112] contraction over the FLATTENED NHWC tensor: a tile is 128
// consecutive positions, whatever the image geometry.  Replaces proj_fwd_tc_kernel (synchronous generic loads -> st.shared
// -> MMA -> epilogue, nothing overlapped: 0.25 of the HBM rate, profiles/r2_ncu_full_tables.md):
//
//   warp 0 (one lane)  TMA producer: per tile, ceil(Cin / 64) tensor-map copies (cp.async.bulk.tensor.2d, SASS UTMALDG) of a
//                      [128 positions x 64 channels] box into a ring of kStages 16 KB buffers, SWIZZLE_128B (the canonical
//                      K-major UMMA operand layout: 8-row x 128-byte atoms, 1024 bytes per 8-row group).  Channels beyond
//                      Cin and positions beyond the tensor are zero-filled by the TMA unit (no bounds code anywhere);
//   warp 1 (one lane)  MMA issuer: tcgen05.mma 128 x 112 x 16, A from the ring (swizzled descriptor), B = the packed weight
//                      block (resident, core-matrix layout written by prep_kernel), accumulator = one of TWO 128-column TMEM
//                      buffers; tcgen05.commit releases the ring slot and signals the epilogue;
//   warps 2-5          epilogue, one accumulator row (position) per thread: tcgen05.ld x32/x16, + bias, bf16, staging rows
//                      in shared memory, BatchNorm partial sums (training), one tensor-map STORE per warp (32 rows,
//                      cp.async.bulk.tensor.2d.global.shared::cta, SASS UTMASTG: rows beyond the tensor are clipped).
// Up to four networks (student + teachers) share a launch (blockIdx.y); one persistent CTA per SM and network slice.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "bifpn.cuh"
#include "tc.cuh"

namespace mmd {
namespace ptma {

typedef __nv_bfloat16 bf16;
constexpr int C = 112;
constexpr int kRows = 128;                  // positions per tile
constexpr int kBoxK = 64;                   // channels per TMA box (128 bytes: one swizzle span)
constexpr int kStageBytes = kRows * 128;    // 16 384
constexpr int kThreadsP = 192;              // producer warp, MMA warp, 4 epilogue warps
constexpr int kEpiWarp0 = 2;
constexpr uint32_t kTmemCols = 256;         // two accumulators of 128 columns
constexpr int kStagingBytes = kRows * C * 2;   // 28 672 per accumulator

struct ProjTmaP {
  CUtensorMap in_map[kMaxBatchNets];    // [positions][Cin] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap out_map[kMaxBatchNets];   // [positions][112] bf16, box {112, 32}, no swizzle
  NodeFwdP p[kMaxBatchNets];
  int kboxes;          // ceil(Cin / 64)
  int ksteps_last;     // MMA K-steps (of 16 channels) of the last box
  int ntiles;          // ceil(B*H*W / 128)
  int npos;
  int stages;
  int w_bytes, off_bias;   // packed block: B operand bytes, offset of the bias
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(tc::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// K-major SWIZZLE_128B operand descriptor: 8-row x 128-byte atoms, SBO = 1024 bytes between 8-row groups, LBO = 16 bytes
// (the K step inside the swizzle span); layout type 2.  `byte_off` advances K inside the span (32 bytes per K = 16 step).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // leading byte offset (16 B units)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
  d |= (uint64_t)1 << 46;                    // version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kThreadsP, 1) proj_fwd_tma_kernel(const __grid_constant__ ProjTmaP Q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int net = blockIdx.y;
  const NodeFwdP& P = Q.p[net];
  const CUtensorMap* in_map = &Q.in_map[net];
  const CUtensorMap* out_map = &Q.out_map[net];
  // shared memory: ring of A stages (1024-byte aligned: SWIZZLE_128B atoms) | staging [2][128][112] bf16 | B operand | bias |
  // reduction scratch | barriers.  The launch reserves 1 KB of slack for the manual alignment.
  unsigned char* s_a = smem + ((1024u - (tc::smem_u32(smem) & 1023u)) & 1023u);
  unsigned char* s_stage = s_a + Q.stages * kStageBytes;
  unsigned char* s_b = s_stage + 2 * kStagingBytes;
  float* s_bias = reinterpret_cast<float*>(s_b + ((Q.w_bytes + 127) / 128) * 128);
  double* s_red = reinterpret_cast<double*>(s_bias + C);                 // [4 warps][C][2]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_red + 4 * C * 2);   // [stages]
  uint64_t* bar_empty = bar_full + 8;                                    // [stages]
  uint64_t* bar_tfull = bar_empty + 8;                                   // [2] accumulator ready
  uint64_t* bar_tempty = bar_tfull + 2;                                  // [2] accumulator drained
  uint64_t* bar_w = bar_tempty + 2;                                      // weights + bias landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_w + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool train = P.train != 0;

  if (warp == 1) tc::tmem_alloc(s_tmem, kTmemCols);
  if (tid == 0) {
    tma_prefetch_desc(in_map);
    tma_prefetch_desc(out_map);
    for (int s = 0; s < Q.stages; ++s) {
      tc::mbar_init(bar_full + s, 1);
      tc::mbar_init(bar_empty + s, 1);
    }
    tc::mbar_init(bar_tfull, 1);
    tc::mbar_init(bar_tfull + 1, 1);
    tc::mbar_init(bar_tempty, 4);       // one arrival per epilogue warp
    tc::mbar_init(bar_tempty + 1, 4);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(bar_w, (uint32_t)(Q.w_bytes + C * 4));
    tc::bulk_g2s(s_b, P.packed, (uint32_t)Q.w_bytes, bar_w);
    tc::bulk_g2s(s_bias, P.packed + Q.off_bias, C * 4, bar_w);
  }
  pdl_wait();
  pdl_trigger();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < Q.ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < Q.kboxes; ++kb, ++it) {
          const int s = it % Q.stages;
          const uint32_t par = (uint32_t)((it / Q.stages) & 1);
          tc::mbar_wait(bar_empty + s, par ^ 1u);          // slot free (first pass: passes immediately)
          tc::mbar_expect_tx(bar_full + s, kStageBytes);
          tma_load_2d(s_a + s * kStageBytes, in_map, kb * kBoxK, tile * kRows, bar_full + s);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t kIdesc = tc::make_idesc_bf16(128, C, false, false);
      tc::mbar_wait(bar_w, 0u);
      const uint32_t b_addr = tc::smem_u32(s_b);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < Q.ntiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        tc::mbar_wait(bar_tempty + acc, (uint32_t)(((lt >> 1) & 1) ^ 1));   // epilogue has drained this accumulator
        tc::fence_after_sync();
        for (int kb = 0; kb < Q.kboxes; ++kb, ++it) {
          const int s = it % Q.stages;
          tc::mbar_wait(bar_full + s, (uint32_t)((it / Q.stages) & 1));
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(s_a + s * kStageBytes);
          const int nk = (kb == Q.kboxes - 1) ? Q.ksteps_last : 4;
          for (int j = 0; j < nk; ++j) {
            const uint64_t adesc = make_desc_sw128(a_addr + j * 32);
            const uint64_t bdesc = tc::make_desc(b_addr + (kb * 4 + j) * 2 * (C * 16), C * 16, 128);
            tc::umma_bf16(tmem_base + acc * 128, adesc, bdesc, kIdesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
          tc::umma_commit(bar_empty + s);                   // ring slot free once these MMAs have read it
        }
        tc::umma_commit(bar_tfull + acc);                    // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int row = 32 * quad + lane;
    // statistics role of this lane: channel pairs lane and lane + 28 (lanes 0..27)
    double st[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    tc::mbar_wait(bar_w, 0u);   // bias
    int lt = 0;
    for (int tile = blockIdx.x; tile < Q.ntiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      unsigned char* stg = s_stage + acc * kStagingBytes + quad * (32 * C * 2);   // this warp's 32 staging rows
      tc::mbar_wait(bar_tfull + acc, (uint32_t)((lt >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * quad) << 16) + (uint32_t)(acc * 128);
      float v0[7][8], v1[7][8];
      tc::tmem_ld56(taddr, v0);
      tc::tmem_ld56(taddr + 56, v1);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + acc);          // the MMA warp may overwrite this accumulator
      // the tensor store issued from this staging slice two tiles ago must have read it
      if (lane == 0) bulk_wait_read_le1();
      __syncwarp();
      bf16* myrow = reinterpret_cast<bf16*>(stg) + lane * C;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float(&v)[8] = h ? v1[j] : v0[j];
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + 56 * h + 8 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + 56 * h + 8 * j + 4);
          uint4 pk;
          pk.x = tc::f2_to_bf2(make_float2(v[0] + b0.x, v[1] + b0.y));
          pk.y = tc::f2_to_bf2(make_float2(v[2] + b0.z, v[3] + b0.w));
          pk.z = tc::f2_to_bf2(make_float2(v[4] + b1.x, v[5] + b1.y));
          pk.w = tc::f2_to_bf2(make_float2(v[6] + b1.z, v[7] + b1.w));
          *reinterpret_cast<uint4*>(myrow + 56 * h + 8 * j) = pk;
        }
      }
      tc::fence_async_smem();   // staging rows -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(out_map, stg, 0, tile * kRows + 32 * quad);
        bulk_commit();
      }
      if (train) {   // per-channel sums over this warp's valid rows, from the bf16 values that were stored
        int nvalid = Q.npos - (tile * kRows + 32 * quad);
        nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
        if (lane < 28) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const bf16* col = reinterpret_cast<const bf16*>(stg) + 2 * (lane + 28 * k);
            float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
            for (int r = 0; r < nvalid; ++r) {
              const float2 x = tc::bf2_to_f2(*reinterpret_cast<const uint32_t*>(col + r * C));
              s = tc::add2(s, x);
              q = tc::fma2(x, x, q);
            }
            st[k][0] += (double)s.x; st[k][1] += (double)s.y; st[k][2] += (double)q.x; st[k][3] += (double)q.y;
          }
        }
      }
      (void)row;
    }
    if (lane == 0) bulk_wait_all();
    if (train && lane < 28) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = 2 * (lane + 28 * k);
        s_red[(quad * C + c) * 2] = st[k][0];
        s_red[(quad * C + c + 1) * 2] = st[k][1];
        s_red[(quad * C + c) * 2 + 1] = st[k][2];
        s_red[(quad * C + c + 1) * 2 + 1] = st[k][3];
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, kTmemCols);
  if (train && tid < C && (int)blockIdx.x < Q.ntiles) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      s += s_red[(w * C + tid) * 2];
      q += s_red[(w * C + tid) * 2 + 1];
    }
    atomicAdd(P.stats + tid, s);
    atomicAdd(P.stats + C + tid, q);
  }
}

// ---- host: tensor maps ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [rows][cols] (cols contiguous), box {box_cols, box_rows}
static int encode_2d(CUtensorMap* m, const void* base, long long rows, int cols, int box_cols, int box_rows, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MMD_E_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%lld x %d] bf16 tensor, box {%d, %d}", (int)r, rows, cols, box_cols, box_rows);
    return MMD_E_ARG;
  }
  return 0;
}

}  // namespace ptma

static int g_proj_tma = -1;   // MMD_NO_PROJ_TMA=1 / mmd_set_option("proj_tma", 0): back to proj_fwd_tc_kernel
void set_proj_tma(int on) { g_proj_tma = on ? 1 : 0; }
static bool proj_tma_enabled() {
  if (g_proj_tma < 0) {
    const char* e = getenv("MMD_NO_PROJ_TMA");
    g_proj_tma = (e && e[0] == '1') ? 0 : 1;
  }
  return g_proj_tma == 1;
}

bool proj_fwd_tma_usable(const NodeFwdP* ps, int n) {
  if (!proj_tma_enabled() || ptma::encode_fn() == nullptr) return false;
  for (int i = 0; i < n; ++i) {
    const NodeFwdP& p = ps[i];
    if (p.packed == nullptr || p.Cin % 8 != 0 || p.Cin < 8 || p.Cin > 512) return false;
    if ((((uintptr_t)p.in[0].data | (uintptr_t)p.out) & 15u) != 0) return false;
    if (p.train && !p.defer_bn) return false;   // the in-kernel BatchNorm finaliser lives in proj_fwd_tc_kernel only
    if (p.Cin != ps[0].Cin || p.g.B != ps[0].g.B || p.g.H != ps[0].g.H || p.g.W != ps[0].g.W) return false;
  }
  return true;
}

int launch_proj_fwd_tma(const NodeFwdP* ps, int n, int C, cudaStream_t s) {
  MMD_CHECK_ARG(C == 112, "BiFPN kernels are built for C=112 (EfficientDet-D2), got %d", C);
  MMD_CHECK_ARG(n >= 1 && n <= kMaxBatchNets, "proj_fwd_tma: %d networks in one launch", n);
  const NodeFwdP& p0 = ps[0];
  const long long npos = (long long)p0.g.B * p0.g.H * p0.g.W;
  MMD_CHECK_ARG(npos < (1LL << 31), "proj_fwd_tma: %lld positions", npos);
  ptma::ProjTmaP Q;
  const PackedLayout L = packed_layout(MMD_OP_PROJ_FWD, p0.Cin, C);
  Q.kboxes = (p0.Cin + ptma::kBoxK - 1) / ptma::kBoxK;
  Q.ksteps_last = (L.Kp - (Q.kboxes - 1) * ptma::kBoxK) / 16;
  Q.ntiles = (int)((npos + ptma::kRows - 1) / ptma::kRows);
  Q.npos = (int)npos;
  Q.w_bytes = L.Kp * C * 2;
  Q.off_bias = L.offBias;
  // shared memory: stages of 16 KB | 2 staging tiles | B | bias | reduction scratch | barriers
  const size_t fixed = 2 * ptma::kStagingBytes + ((Q.w_bytes + 127) / 128) * 128 + C * 4 + 4 * C * 2 * 8 + 32 * 8 + 16;
  int stages = (int)((227 * 1024 - 2048 - fixed) / ptma::kStageBytes);
  if (stages > 8) stages = 8;
  MMD_CHECK_ARG(stages >= 2, "proj_fwd_tma: Cin=%d leaves no room for the input ring", p0.Cin);
  Q.stages = stages;
  const size_t smem = (size_t)stages * ptma::kStageBytes + fixed + 1024;   // + alignment slack
  for (int i = 0; i < kMaxBatchNets; ++i) {
    const NodeFwdP& p = ps[i < n ? i : 0];
    Q.p[i] = p;
    int rc = ptma::encode_2d(&Q.in_map[i], p.in[0].data, npos, p.Cin, ptma::kBoxK, ptma::kRows, true);
    if (rc) return rc;
    if ((rc = ptma::encode_2d(&Q.out_map[i], p.out, npos, C, C, 32, false))) return rc;
  }
  MMD_SMEM(ptma::proj_fwd_tma_kernel, smem);
  int gx = device_sm_count() / n;
  if (gx < 1) gx = 1;
  if (gx > Q.ntiles) gx = Q.ntiles;
  ProfScope prof(PK_PROJ_FWD, (double)n * npos * (p0.Cin + C) * 2, s);
  MMD_CUDA(launch_pdl(ptma::proj_fwd_tma_kernel, dim3(gx, n), dim3(ptma::kThreadsP), smem, s, Q));
  MMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmd
