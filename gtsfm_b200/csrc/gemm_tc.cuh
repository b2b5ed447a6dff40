// tcgen05 split-fp16 "NT" GEMM: C[M,N] = [A1 | A2][M,K] * B[N,K]^T with ~fp32 accuracy on the 5th-gen tensor cores.
//
// Each fp32 operand x is split as x ~= hi + lo * 2^-11 (two fp16 values, 22 significand bits).  Three MMAs per K step,
//     acc0 += Ah * Bh            acc1 += Ah * Bl + Al * Bh          result = acc0 + acc1 * 2^-11
// with both fp32 accumulators living in TMEM (2 x 128 columns).  The dropped Al * Bl term is 2^-22 relative.
//
// One CTA = 128 threads = one 128 x 128 output tile.  Per 64-wide K chunk the CTA stages A (fp32 -> split fp16, done
// in-kernel) and B (pre-split fp16 weights, or fp32 activations for the assignment matrix) into shared memory in the
// UMMA interleaved K-major canonical layout (tc.cuh), one elected thread issues the 12 tcgen05.mma, tcgen05.commit
// signals an mbarrier, and after the last chunk every warp drains its 32 TMEM lanes with tcgen05.ld for the epilogue.
// 64 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM overlap staging with the tensor pipe.
#pragma once
#include "common.cuh"
#include "tc.cuh"

struct GemmTcArgs {
  const float* A1;
  int lda1;
  int K1;
  const float* A2;
  int lda2;
  int K2;
  const __half* Bh;  // [N][K] fp16 hi   (B_IS_F32 == false)
  const __half* Bl;  // [N][K] fp16 lo * 2^11
  const float* Bf;   // [N][K] fp32      (B_IS_F32 == true)
  int ldb;
  float* C;
  int ldc;
  int M;
  int N;
  const float* bias;
  const float* resid;
  int ldr;
  float scale;
  int head_major;
  int* err_flag;  // set to 1 if an mbarrier wait timed out (pipeline bug): results are then invalid
};

constexpr int TC_M = 128, TC_N = 128, TC_K = 64;
constexpr int TC_TILE_BYTES = TC_M * TC_K * 2;  // one fp16 operand tile: 16 KB
constexpr size_t TC_GEMM_SMEM = 4 * TC_TILE_BYTES + 1024;

template <bool B_IS_F32>
__global__ void __launch_bounds__(128) k_gemm_tc(GemmTcArgs g) {
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned char* sAh = tsm;
  unsigned char* sAl = sAh + TC_TILE_BYTES;
  unsigned char* sBh = sAl + TC_TILE_BYTES;
  unsigned char* sBl = sBh + TC_TILE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + TC_TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int t = threadIdx.x, warp = t >> 5;
  const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * TC_N;
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (t == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = tc::idesc_f16(TC_M, TC_N);
  const uint32_t lbo = (TC_M / 8) * 128;  // 2048 B between K chunks
  const int K = g.K1 + g.K2;
  uint32_t phase = 0;
  bool ok = true;

  for (int k0 = 0; k0 < K; k0 += TC_K) {
    // ---- stage A: thread t owns row m0 + t ------------------------------------------------------------------
    {
      const int m = m0 + t;
      const float* src = nullptr;
      if (m < g.M) src = (k0 < g.K1) ? g.A1 + (size_t)m * g.lda1 + k0 : g.A2 + (size_t)m * g.lda2 + (k0 - g.K1);
#pragma unroll
      for (int c = 0; c < TC_K / 8; ++c) {
        float x[8];
        if (src) {
          float4 a = *reinterpret_cast<const float4*>(src + c * 8), b = *reinterpret_cast<const float4*>(src + c * 8 + 4);
          x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const uint32_t off = tc::canon_off(t, c, TC_M);
        *reinterpret_cast<uint4*>(sAh + off) = hi;
        *reinterpret_cast<uint4*>(sAl + off) = lo;
      }
    }
    // ---- stage B: thread t owns row n0 + t --------------------------------------------------------------------
    {
      const int n = n0 + t;
#pragma unroll
      for (int c = 0; c < TC_K / 8; ++c) {
        uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
        if (n < g.N) {
          if (B_IS_F32) {
            const float* src = g.Bf + (size_t)n * g.ldb + k0 + c * 8;
            float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            tc::split8(x, hi, lo);
          } else {
            hi = *reinterpret_cast<const uint4*>(g.Bh + (size_t)n * g.ldb + k0 + c * 8);
            lo = *reinterpret_cast<const uint4*>(g.Bl + (size_t)n * g.ldb + k0 + c * 8);
          }
        }
        const uint32_t off = tc::canon_off(t, c, TC_N);
        *reinterpret_cast<uint4*>(sBh + off) = hi;
        *reinterpret_cast<uint4*>(sBl + off) = lo;
      }
    }
    tc::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (t == 0) {
      tc::fence_after_sync();
      const uint32_t aH = tc::smem_u32(sAh), aL = tc::smem_u32(sAl), bH = tc::smem_u32(sBh), bL = tc::smem_u32(sBl);
#pragma unroll
      for (int s = 0; s < TC_K / 16; ++s) {
        const uint32_t ko = 2 * s * lbo;
        const uint64_t dAh = tc::smem_desc(aH + ko, lbo), dAl = tc::smem_desc(aL + ko, lbo);
        const uint64_t dBh = tc::smem_desc(bH + ko, lbo), dBl = tc::smem_desc(bL + ko, lbo);
        const uint32_t first = (k0 == 0 && s == 0) ? 0u : 1u;
        tc::umma_f16(tmem, dAh, dBh, idesc, first);         // acc0 (+)= Ah Bh
        tc::umma_f16(tmem + TC_N, dAh, dBl, idesc, first);  // acc1 (+)= Ah Bl
        tc::umma_f16(tmem + TC_N, dAl, dBh, idesc, 1u);     // acc1  += Al Bh
      }
      tc::umma_commit(bar);
    }
    ok = tc::mbar_wait(bar, phase) && ok;  // MMAs of this chunk done: smem may be overwritten
    phase ^= 1;
  }
  tc::fence_after_sync();
  if (!ok && g.err_flag) *g.err_flag = 1;

  // ---- epilogue: thread t = accumulator row (TMEM lane) t ---------------------------------------------------------
  const int m = m0 + t;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int cc = 0; cc < TC_N / 32; ++cc) {
    float a0[32], a1[32];
    tc::tmem_ld32(lane_base + cc * 32, a0);
    tc::tmem_ld32(lane_base + TC_N + cc * 32, a1);
    if (m < g.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + cc * 32 + j;
        if (n < g.N) {
          float v = fmaf(a1[j], tc::LO_INV, a0[j]);
          if (g.bias) v += g.bias[n];
          v *= g.scale;
          if (g.resid) v += g.resid[(size_t)m * g.ldr + n];
          if (g.head_major)
            g.C[((size_t)(n >> 6) * g.M + m) * 64 + (n & 63)] = v;
          else
            g.C[(size_t)m * g.ldc + n] = v;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// fp32 [rows][cols] -> fp16 hi / lo*2^11 copies (weights, once at load time)
static __global__ void k_split_f32(const float* __restrict__ x, size_t n, __half* __restrict__ hi, __half* __restrict__ lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __half h, l;
  tc::split_h(x[i], h, l);
  hi[i] = h;
  lo[i] = l;
}
