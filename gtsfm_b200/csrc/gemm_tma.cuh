// tcgen05 + TMA split-fp16 "NT" GEMM (the production linear kernel):  C[M,N] = [A1 | A2][M,K] * B[N,K]^T.
//
// Same arithmetic as gemm_tc.cuh (acc0 += Ah Bh ; acc1 += Ah Bl + Al Bh ; result acc0 + acc1 * 2^-11, two fp32
// accumulators in TMEM), different data movement: operand tiles are fetched by the Tensor Memory Accelerator
// (cp.async.bulk.tensor.2d, 128-byte swizzle, zero-filled out of bounds) straight into shared memory and signalled on
// mbarriers with complete_tx, so staging costs one instruction per tile instead of 24 per thread.
//
// Warp roles (128 threads): warp 0 lane 0 = TMA producer over a ring of stages; warp 1 = TMEM allocator, its lane 0 =
// MMA issuer (12 tcgen05.mma per 64-wide K chunk, tcgen05.commit frees the stage / publishes the accumulator); then all
// four warps run the epilogue (tcgen05.ld -> shared-memory transpose -> coalesced bias / scale / ReLU / residual ->
// fp32 and/or split-fp16 planes, row-major or attention head-major).
//
// Shared-memory tile = [rows][64 halves] with 128-byte rows, XOR-swizzled in 8-row x 128-byte atoms: the UMMA
// descriptor is layout SWIZZLE_128B, SBO = 1024 B (one atom), and one MMA K step (16 halves) advances the start address
// by 32 B.  Tiles are 1024-byte aligned.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_tc.cuh"  // GemmTcProblem-style epilogue contract, k_split_f32
#include "tc.cuh"

constexpr int TM_M = 128, TM_N = 64, TM_K = 64, TM_STAGES = 2;
constexpr int TM_A_BYTES = TM_M * TM_K * 2;                       // 16 KB per plane
constexpr int TM_B_BYTES = TM_N * TM_K * 2;                       // 8 KB per plane
constexpr int TM_STAGE_BYTES = 2 * TM_A_BYTES + 2 * TM_B_BYTES;   // 48 KB
constexpr size_t TM_GEMM_SMEM = TM_STAGES * TM_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmTmaMaps {  // 128-byte TMA descriptors, passed as a __grid_constant__ kernel parameter
  CUtensorMap a1h[2], a1l[2], a2h[2], a2l[2];  // per problem
  CUtensorMap bh, bl;
};
struct GemmTmaArgs {
  GemmTcProblem p[2];  // only resid / C / Ch / Cl / M are used (operands come through the tensor maps)
  int K1, K2, N;
  const float* bias;
  int ldr;
  float scale;
  int ldc, ldch, head_major, relu;
  int lo_unscaled;  // split outputs keep lo = fp16(x - hi) (attention q / k operands of k_flash_ws)
  int* err_flag;
};

namespace tc {
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;               // leading byte offset field (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;     // stride byte offset: one 8-row x 128-byte swizzle atom
  d |= (uint64_t)1 << 46;               // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;               // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
}  // namespace tc

static __global__ void __launch_bounds__(128, 2) k_gemm_tma(const __grid_constant__ GemmTmaMaps maps, GemmTmaArgs g) {
  extern __shared__ unsigned char tsm_raw[];
  const uint32_t raw = tc::smem_u32(tsm_raw);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;  // 1024-byte aligned tile area
  unsigned char* tsm = tsm_raw + (smem0 - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(tsm + TM_STAGES * TM_STAGE_BYTES);  // TMA bytes landed in stage s
  uint64_t* empty = full + TM_STAGES;                                              // MMAs reading stage s done
  uint64_t* accum = empty + TM_STAGES;                                             // all MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int z = blockIdx.z;
  const GemmTcProblem& pb = g.p[z];
  const int M = pb.M;
  const int m0 = blockIdx.y * TM_M, n0 = blockIdx.x * TM_N;
  if (m0 >= M) return;  // uniform; before any allocation / barrier
  const int nk = (g.K1 + g.K2) / TM_K;

  if (t == 0) {
    for (int s = 0; s < TM_STAGES; ++s) tc::mbar_init(&full[s], 1), tc::mbar_init(&empty[s], 1);
    tc::mbar_init(accum, 1);
    tc::fence_mbar_init();
    tc::tma_prefetch_desc(&maps.a1h[z]);
    tc::tma_prefetch_desc(&maps.a1l[z]);
    tc::tma_prefetch_desc(&maps.bh);
    tc::tma_prefetch_desc(&maps.bl);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * TM_N);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  bool ok = true;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TM_STAGES;
      if (kc >= TM_STAGES) ok = tc::mbar_wait(&empty[s], ((kc / TM_STAGES) - 1) & 1) && ok;
      const int k0 = kc * TM_K;
      const bool seg2 = k0 >= g.K1;
      const CUtensorMap* ah = seg2 ? &maps.a2h[z] : &maps.a1h[z];
      const CUtensorMap* al = seg2 ? &maps.a2l[z] : &maps.a1l[z];
      const int ka = seg2 ? k0 - g.K1 : k0;
      const uint32_t sA = smem0 + s * TM_STAGE_BYTES, sB = sA + 2 * TM_A_BYTES;
      tc::mbar_expect_tx(&full[s], TM_STAGE_BYTES);
      tc::tma_load_2d(sA, ah, &full[s], ka, m0);
      tc::tma_load_2d(sA + TM_A_BYTES, al, &full[s], ka, m0);
      tc::tma_load_2d(sB, &maps.bh, &full[s], k0, n0);
      tc::tma_load_2d(sB + TM_B_BYTES, &maps.bl, &full[s], k0, n0);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with uniform operands, one elected lane issues (tc.cuh) =====
    const uint32_t idesc = tc::idesc_f16(TM_M, TM_N);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TM_STAGES;
      ok = tc::mbar_wait(&full[s], (kc / TM_STAGES) & 1) && ok;
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t aH = smem0 + s * TM_STAGE_BYTES, aL = aH + TM_A_BYTES, bH = aH + 2 * TM_A_BYTES, bL = bH + TM_B_BYTES;
      const uint64_t dAh = tc::smem_desc_sw128(aH), dAl = tc::smem_desc_sw128(aL), dBh = tc::smem_desc_sw128(bH), dBl = tc::smem_desc_sw128(bL);
#pragma unroll
      for (int ks = 0; ks < TM_K / 16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K step, in 16-byte units of the start-address field
        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
        tc::umma_f16_w(tmem, dAh + adv, dBh + adv, idesc, first);         // acc0 (+)= Ah Bh
        tc::umma_f16_w(tmem + TM_N, dAh + adv, dBl + adv, idesc, first);  // acc1 (+)= Ah Bl
        tc::umma_f16_w(tmem + TM_N, dAl + adv, dBh + adv, idesc, 1u);     // acc1  += Al Bh
      }
      tc::umma_commit_w(&empty[s]);
    }
    tc::umma_commit_w(accum);
  }
  __syncwarp();
  ok = tc::mbar_wait(accum, 0) && ok;
  tc::fence_after_sync();
  if (!ok && g.err_flag) *g.err_flag = 1;
  __syncthreads();  // every warp is past the accumulator wait before the operand tiles are reused as scratch

  // ---- epilogue (identical contract to k_gemm_tc) -------------------------------------------------------------------
  float* scratch = reinterpret_cast<float*>(tsm) + warp * (32 * 33);
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const int mw = m0 + warp * 32;
  const int rows = min(32, M - mw);
  const float scale = g.scale;
  const int relu = g.relu;
#pragma unroll 1
  for (int cc = 0; cc < TM_N / 32; ++cc) {
    if (n0 + cc * 32 >= g.N) break;
    {
      float a0[32], a1[32];
      tc::tmem_ld32(lane_base + cc * 32, a0);
      tc::tmem_ld32(lane_base + TM_N + cc * 32, a1);
#pragma unroll
      for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = fmaf(a1[j], tc::LO_INV, a0[j]);
    }
    __syncwarp();
    const int n = n0 + cc * 32 + lane;
    if (n < g.N && rows > 0) {
      const float bn = g.bias ? g.bias[n] : 0.f;
      const size_t off_h = ((size_t)(n >> 6) * M + mw) * 64 + (n & 63);
      const size_t off_c = g.head_major ? off_h : (size_t)mw * g.ldc + n;
      const size_t off_s = g.head_major ? off_h : (size_t)mw * g.ldch + n;
      const int str_c = g.head_major ? 64 : g.ldc, str_s = g.head_major ? 64 : g.ldch;
      const float* rp = pb.resid ? pb.resid + (size_t)mw * g.ldr + n : nullptr;
      float* cp = pb.C ? pb.C + off_c : nullptr;
      __half* hp = pb.Ch ? pb.Ch + off_s : nullptr;
      __half* lp = pb.Ch ? pb.Cl + off_s : nullptr;
      const float* sp = scratch + lane;
      if (rp) {
        float rv[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) rv[r] = r < rows ? rp[(size_t)r * g.ldr] : 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          if (r < rows) {
            float v = (sp[r * 33] + bn) * scale;
            if (relu) v = fmaxf(v, 0.f);
            v += rv[r];
            if (cp) cp[(size_t)r * str_c] = v;
            if (hp) {
              __half hh, ll;
              if (g.lo_unscaled) tc::split_h_unscaled(v, hh, ll);
              else tc::split_h(v, hh, ll);
              hp[(size_t)r * str_s] = hh;
              lp[(size_t)r * str_s] = ll;
            }
          }
        }
      } else {
#pragma unroll 8
        for (int r = 0; r < rows; ++r) {
          float v = (sp[r * 33] + bn) * scale;
          if (relu) v = fmaxf(v, 0.f);
          if (cp) cp[(size_t)r * str_c] = v;
          if (hp) {
            __half hh, ll;
            if (g.lo_unscaled) tc::split_h_unscaled(v, hh, ll);
            else tc::split_h(v, hh, ll);
            hp[(size_t)r * str_s] = hh;
            lp[(size_t)r * str_s] = ll;
          }
        }
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 2 * TM_N);
}

// ---- host: tensor-map construction (driver entry point resolved at run time: the library does not link libcuda) ------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled tma_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D fp16 matrix [rows][ld] (K contiguous), box = 64 K-elements x box_rows rows, 128-byte swizzle, zero OOB fill.
static inline bool tma_map_2d(CUtensorMap* out, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = tma_encoder();
  if (!enc || !base) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(__half)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
