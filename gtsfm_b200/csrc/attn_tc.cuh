// tcgen05 flash attention with split-fp16 operands (~fp32 accuracy): O[Nq][256] = softmax(scale * Q K^T) V per head.
//
// One CTA = 128 threads = 128 queries of one head; keys/values stream through in tiles of 64.  Per tile:
//   S  (128 x 64, two fp32 accumulators in TMEM)  = Qh Kh^T ;  Qh Kl^T + Ql Kh^T          12 tcgen05.mma (K = 64 dims)
//   every thread owns one query row (= one TMEM lane): tcgen05.ld, online softmax in registers (no shuffles),
//   P = exp(S - m) split into fp16 hi / lo*2^11 and written to shared memory as the next A operand
//   O' (128 x 64, two accumulators)               = Ph Vh ;  Ph Vl + Pl Vh                12 tcgen05.mma (K = 64 keys)
//   O = O * exp(m_old - m_new) + O'  in registers.
// Operands are converted fp32 -> split fp16 while being staged into the UMMA interleaved K-major layout (tc.cuh);
// V is transposed on the way in so that the PV product is K-major on both sides.  96 KB of shared memory and 256 TMEM
// columns per CTA: two CTAs per SM, so one CTA's softmax overlaps the other's MMAs.
#pragma once
#include "common.cuh"
#include "tc.cuh"

constexpr int AT_Q = 128, AT_KV = 64, AT_D = 64;
constexpr int AT_Q_BYTES = AT_Q * AT_D * 2;    // 16 KB  (Q hi / lo, P hi / lo)
constexpr int AT_KV_BYTES = AT_KV * AT_D * 2;  // 8 KB   (K hi / lo, V^T hi / lo)
constexpr size_t AT_SMEM = 4 * AT_Q_BYTES + 4 * AT_KV_BYTES + 1024;

static __global__ void __launch_bounds__(128, 2) k_flash_tc(const float* __restrict__ Q, const float* __restrict__ Kp,
                                                      const float* __restrict__ V, float* __restrict__ O, int Nq, int Nk,
                                                      float scale, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) unsigned char asm_[];
  unsigned char* sQh = asm_;
  unsigned char* sQl = sQh + AT_Q_BYTES;
  unsigned char* sPh = sQl + AT_Q_BYTES;
  unsigned char* sPl = sPh + AT_Q_BYTES;
  unsigned char* sKh = sPl + AT_Q_BYTES;
  unsigned char* sKl = sKh + AT_KV_BYTES;
  unsigned char* sVh = sKl + AT_KV_BYTES;
  unsigned char* sVl = sVh + AT_KV_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVl + AT_KV_BYTES);  // bar[0]: S ready, bar[1]: O' ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int t = threadIdx.x, warp = t >> 5;
  const int h = blockIdx.y, q0 = blockIdx.x * AT_Q;
  const float* Qh = Q + (size_t)h * Nq * 64;
  const float* Kh = Kp + (size_t)h * Nk * 64;
  const float* Vh = V + (size_t)h * Nk * 64;

  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (t == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_mbar_init();
  }
  // stage Q once: thread t = query row q0 + t
  {
    const int q = q0 + t;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x[8];
      if (q < Nq) {
        float4 a = *reinterpret_cast<const float4*>(Qh + (size_t)q * 64 + c * 8), b = *reinterpret_cast<const float4*>(Qh + (size_t)q * 64 + c * 8 + 4);
        x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      uint4 hi, lo;
      tc::split8(x, hi, lo);
      const uint32_t off = tc::canon_off(t, c, AT_Q);
      *reinterpret_cast<uint4*>(sQh + off) = hi;
      *reinterpret_cast<uint4*>(sQl + off) = lo;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS0 = tmem, tS1 = tmem + 64, tO0 = tmem + 128, tO1 = tmem + 192;
  const uint32_t idesc = tc::idesc_f16(AT_Q, AT_KV);  // M = 128, N = 64 for both products
  const uint32_t lboQ = (AT_Q / 8) * 128, lboK = (AT_KV / 8) * 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  float m_i = -INFINITY, l_i = 0.f;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  uint32_t phase = 0;
  bool ok = true;

  for (int k0 = 0; k0 < Nk; k0 += AT_KV) {
    // ---- stage K (threads 0..63: key rows) and V^T (threads 64..127: key rows, scattered into [dim][key]) ---------
    if (t < 64) {
      const int kk = k0 + t;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float x[8];
        if (kk < Nk) {
          float4 a = *reinterpret_cast<const float4*>(Kh + (size_t)kk * 64 + c * 8), b = *reinterpret_cast<const float4*>(Kh + (size_t)kk * 64 + c * 8 + 4);
          x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const uint32_t off = tc::canon_off(t, c, AT_KV);
        *reinterpret_cast<uint4*>(sKh + off) = hi;
        *reinterpret_cast<uint4*>(sKl + off) = lo;
      }
    } else {
      const int j = t - 64, kk = k0 + j;  // key j of the tile -> column (k index) j of V^T
      __half* vh = reinterpret_cast<__half*>(sVh);
      __half* vl = reinterpret_cast<__half*>(sVl);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk < Nk) a = *reinterpret_cast<const float4*>(Vh + (size_t)kk * 64 + c * 4);
        const float x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = c * 4 + i;  // row d of V^T, k-chunk j / 8, position j % 8
          const uint32_t off = (tc::canon_off(d, j >> 3, AT_D) >> 1) + (j & 7);
          __half hh, ll;
          tc::split_h(x[i], hh, ll);
          vh[off] = hh;
          vl[off] = ll;
        }
      }
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
      tc::fence_after_sync();
      const uint32_t qH = tc::smem_u32(sQh), qL = tc::smem_u32(sQl), kH = tc::smem_u32(sKh), kL = tc::smem_u32(sKl);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t dQh = tc::smem_desc(qH + 2 * s * lboQ, lboQ), dQl = tc::smem_desc(qL + 2 * s * lboQ, lboQ);
        const uint64_t dKh = tc::smem_desc(kH + 2 * s * lboK, lboK), dKl = tc::smem_desc(kL + 2 * s * lboK, lboK);
        tc::umma_f16(tS0, dQh, dKh, idesc, s ? 1u : 0u);
        tc::umma_f16(tS1, dQh, dKl, idesc, s ? 1u : 0u);
        tc::umma_f16(tS1, dQl, dKh, idesc, 1u);
      }
      tc::umma_commit(&bar[0]);
    }
    ok = tc::mbar_wait(&bar[0], phase) && ok;
    tc::fence_after_sync();

    // ---- online softmax on this thread's row ------------------------------------------------------------------
    float sv[64];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      float a0[32], a1[32];
      tc::tmem_ld32(tS0 + lane_off + cc * 32, a0);
      tc::tmem_ld32(tS1 + lane_off + cc * 32, a1);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = fmaf(a1[j], tc::LO_INV, a0[j]) * scale;
        sv[cc * 32 + j] = (k0 + cc * 32 + j < Nk) ? s : -INFINITY;
      }
    }
    float mx = sv[0];
#pragma unroll
    for (int j = 1; j < 64; ++j) mx = fmaxf(mx, sv[j]);
    const float m_new = fmaxf(m_i, mx);
    const float corr = expf(m_i - m_new);
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float p[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = expf(sv[c * 8 + i] - m_new);
        rs += p[i];
      }
      uint4 hi, lo;
      tc::split8(p, hi, lo);
      const uint32_t off = tc::canon_off(t, c, AT_Q);
      *reinterpret_cast<uint4*>(sPh + off) = hi;
      *reinterpret_cast<uint4*>(sPl + off) = lo;
    }
    l_i = l_i * corr + rs;
    m_i = m_new;
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc::fence_after_sync();
      const uint32_t pH = tc::smem_u32(sPh), pL = tc::smem_u32(sPl), vH = tc::smem_u32(sVh), vL = tc::smem_u32(sVl);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t dPh = tc::smem_desc(pH + 2 * s * lboQ, lboQ), dPl = tc::smem_desc(pL + 2 * s * lboQ, lboQ);
        const uint64_t dVh = tc::smem_desc(vH + 2 * s * lboK, lboK), dVl = tc::smem_desc(vL + 2 * s * lboK, lboK);
        tc::umma_f16(tO0, dPh, dVh, idesc, s ? 1u : 0u);
        tc::umma_f16(tO1, dPh, dVl, idesc, s ? 1u : 0u);
        tc::umma_f16(tO1, dPl, dVh, idesc, 1u);
      }
      tc::umma_commit(&bar[1]);
    }
    ok = tc::mbar_wait(&bar[1], phase) && ok;
    tc::fence_after_sync();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      float a0[32], a1[32];
      tc::tmem_ld32(tO0 + lane_off + cc * 32, a0);
      tc::tmem_ld32(tO1 + lane_off + cc * 32, a1);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[cc * 32 + j] = fmaf(o[cc * 32 + j], corr, fmaf(a1[j], tc::LO_INV, a0[j]));
    }
    phase ^= 1;
    tc::fence_before_sync();  // TMEM reads of this tile are ordered before the next tile's MMAs (issued after the next sync)
  }
  if (!ok && err_flag) *err_flag = 1;
  const int q = q0 + t;
  if (q < Nq) {
    const float inv = 1.0f / l_i;
    float4* dst = reinterpret_cast<float4*>(O + (size_t)q * 256 + h * 64);
#pragma unroll
    for (int c = 0; c < 16; ++c) dst[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
