// 3x3 convolution (pad 1) + bias + ReLU (+ fused 2x2/2 max-pool) as a TMA-fed tcgen05 implicit GEMM, split-fp16 (~fp32).
//
// The SuperPoint encoder / head convolutions (thirdparty/SuperGluePretrainedNetwork/models/superpoint.py:119-134,148-162)
// are GEMMs with M = pixels, N = output channels, K = 9 taps x Cin.  Activations live in HBM as NHWC fp16 hi / lo planes
// (x ~= hi + lo * 2^-11).  An output tile is 8 rows x 16 columns of pixels (M = 128) by 64 output channels; for each
// (tap, 64-channel chunk) the producer issues ONE 3-D TMA box load {64 ch, 16 px, 8 px} per plane at the tap-shifted
// pixel coordinate.  Out-of-image coordinates are zero-filled by the TMA unit, which *is* the convolution's zero padding:
// no im2col buffer, no halo logic, and the tile arrives 128-byte swizzled exactly as the UMMA K-major descriptor wants
// it (pixel = row, 64 channels = one 128-byte row).  Weights are a 2-D map over [Cout][9 * Cin].
// Pipeline / arithmetic are those of k_gemm_tma (producer warp, MMA-issuer warp, 2-stage full / empty mbarrier ring,
// acc0 += Ah Bh ; acc1 += Ah Bl + Al Bh in TMEM).  Epilogue: thread = pixel (TMEM lane); bias + ReLU; the 2x2 max-pool is
// two warp shuffles (a warp holds rows 2w, 2w+1 of the tile); outputs are written as fp16 planes for the next
// convolution and / or fp32 for the SIMT head kernels.
#pragma once
#include "common.cuh"
#include "tma.cuh"

constexpr int CV_TH = 8, CV_TW = 16, CV_N = 64, CV_STAGES = 2;
constexpr int CV_A_BYTES = 128 * 64 * 2;   // 16 KB per plane
constexpr int CV_B_BYTES = CV_N * 64 * 2;  // 8 KB per plane
constexpr int CV_STAGE_BYTES = 2 * CV_A_BYTES + 2 * CV_B_BYTES;
constexpr size_t CV_SMEM = CV_STAGES * CV_STAGE_BYTES + 1024 + 256;

struct ConvTmaMaps {
  CUtensorMap ah, al;  // activations: 3-D {C, W, H}
  CUtensorMap wh, wl;  // weights: 2-D {9 * Cin, Cout}
};
struct ConvTmaArgs {
  int H, W, Cin, Cout;
  int pool;           // 1: 2x2/2 max-pool fused; output is (H/2, W/2)
  const float* bias;  // [Cout]
  __half *Oh, *Ol;    // optional output planes NHWC
  float* Of;          // optional fp32 output NHWC
  int* err_flag;
};

namespace tc {
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
}  // namespace tc

static __global__ void __launch_bounds__(128, 2) k_conv_tma(const __grid_constant__ ConvTmaMaps maps, ConvTmaArgs g) {
  extern __shared__ unsigned char cv_raw[];
  const uint32_t raw = tc::smem_u32(cv_raw);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  unsigned char* sm = cv_raw + (smem0 - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + CV_STAGES * CV_STAGE_BYTES);
  uint64_t* empty = full + CV_STAGES;
  uint64_t* accum = empty + CV_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int x0 = blockIdx.x * CV_TW, y0 = blockIdx.y * CV_TH, n0 = blockIdx.z * CV_N;
  const int cchunks = g.Cin / 64, nk = 9 * cchunks;

  if (t == 0) {
    for (int s = 0; s < CV_STAGES; ++s) tc::mbar_init(&full[s], 1), tc::mbar_init(&empty[s], 1);
    tc::mbar_init(accum, 1);
    tc::fence_mbar_init();
    tc::tma_prefetch_desc(&maps.ah);
    tc::tma_prefetch_desc(&maps.al);
    tc::tma_prefetch_desc(&maps.wh);
    tc::tma_prefetch_desc(&maps.wl);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * CV_N);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  bool ok = true;

  if (warp == 0 && lane == 0) {
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % CV_STAGES;
      if (kc >= CV_STAGES) ok = tc::mbar_wait(&empty[s], ((kc / CV_STAGES) - 1) & 1) && ok;
      const int tap = kc / cchunks, cc = kc % cchunks;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const uint32_t sA = smem0 + s * CV_STAGE_BYTES, sB = sA + 2 * CV_A_BYTES;
      tc::mbar_expect_tx(&full[s], CV_STAGE_BYTES);
      tc::tma_load_3d(sA, &maps.ah, &full[s], cc * 64, x0 + dx, y0 + dy);  // zero fill outside the image = conv padding
      tc::tma_load_3d(sA + CV_A_BYTES, &maps.al, &full[s], cc * 64, x0 + dx, y0 + dy);
      tc::tma_load_2d(sB, &maps.wh, &full[s], kc * 64, n0);
      tc::tma_load_2d(sB + CV_B_BYTES, &maps.wl, &full[s], kc * 64, n0);
    }
  } else if (warp == 1) {  // whole warp, one elected lane issues (tc.cuh: warp-uniform issue)
    const uint32_t idesc = tc::idesc_f16(128, CV_N);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % CV_STAGES;
      ok = tc::mbar_wait(&full[s], (kc / CV_STAGES) & 1) && ok;
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t aH = smem0 + s * CV_STAGE_BYTES, aL = aH + CV_A_BYTES, bH = aH + 2 * CV_A_BYTES, bL = bH + CV_B_BYTES;
      const uint64_t dAh = tc::smem_desc_sw128(aH), dAl = tc::smem_desc_sw128(aL), dBh = tc::smem_desc_sw128(bH), dBl = tc::smem_desc_sw128(bL);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
        tc::umma_f16_w(tmem, dAh + adv, dBh + adv, idesc, first);
        tc::umma_f16_w(tmem + CV_N, dAh + adv, dBl + adv, idesc, first);
        tc::umma_f16_w(tmem + CV_N, dAl + adv, dBh + adv, idesc, 1u);
      }
      tc::umma_commit_w(&empty[s]);
    }
    tc::umma_commit_w(accum);
  }
  __syncwarp();
  ok = tc::mbar_wait(accum, 0) && ok;
  tc::fence_after_sync();
  if (!ok && g.err_flag) *g.err_flag = 1;

  // ---- epilogue: thread t = tile pixel (h, w) = (t / 16, t % 16) ----------------------------------------------------
  const int h = t >> 4, w = t & 15;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  int oy, ox, OH, OW;
  bool writer;
  if (g.pool) {
    OH = g.H >> 1, OW = g.W >> 1;
    oy = (y0 + h) >> 1, ox = (x0 + w) >> 1;
    writer = ((h & 1) == 0) && ((w & 1) == 0) && oy < OH && ox < OW;  // lanes 0,2,..14 of each warp's first row
  } else {
    OH = g.H, OW = g.W;
    oy = y0 + h, ox = x0 + w;
    writer = oy < OH && ox < OW;
  }
  const size_t opix = ((size_t)oy * OW + ox) * g.Cout + n0;
#pragma unroll 1
  for (int cc = 0; cc < CV_N / 32; ++cc) {
    float a0[32], a1[32];
    tc::tmem_ld32(lane_base + cc * 32, a0);
    tc::tmem_ld32(lane_base + CV_N + cc * 32, a1);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(a1[j], tc::LO_INV, a0[j]);
    if (g.pool) {  // max over the 2x2 window: partners are lane ^ 1 (w) and lane ^ 16 (h); max commutes with bias + ReLU
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 16));
      }
    }
    if (writer) {
      const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + cc * 32);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b = bp[j4];
        v[4 * j4] = fmaxf(v[4 * j4] + b.x, 0.f);
        v[4 * j4 + 1] = fmaxf(v[4 * j4 + 1] + b.y, 0.f);
        v[4 * j4 + 2] = fmaxf(v[4 * j4 + 2] + b.z, 0.f);
        v[4 * j4 + 3] = fmaxf(v[4 * j4 + 3] + b.w, 0.f);
      }
      if (g.Of) {
        float4* d = reinterpret_cast<float4*>(g.Of + opix + cc * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) d[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      }
      if (g.Oh) {
        uint4* dh = reinterpret_cast<uint4*>(g.Oh + opix + cc * 32);
        uint4* dl = reinterpret_cast<uint4*>(g.Ol + opix + cc * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) tc::split2(v[8 * c + 2 * i], v[8 * c + 2 * i + 1], hi[i], lo[i]);
          dh[c] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dl[c] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 2 * CV_N);
}

// NHWC fp16 activation plane [H][W][C] -> 3-D map {C, W, H}, box {64, 16, 8}, 128-byte swizzle, zero OOB fill
static inline bool tma_map_nhwc(CUtensorMap* out, const __half* base, int H, int W, int C) {
  PFN_encodeTiled enc = tma_encoder();
  if (!enc || !base) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t box[3] = {64, CV_TW, CV_TH};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
