// tcgen05 flash attention with split-fp16 operands (~fp32 accuracy): O[Nq][256] = softmax(scale * Q K^T) V per head.
//
// Inputs are already split (x ~= hi + lo * 2^-11, two fp16 planes, written by the producing kernels' epilogues) in
// head-major [4][N][64] layout, so operand staging is pure 16-byte cp.async traffic straight into the UMMA interleaved
// canonical layout: no register staging, no conversion, and the loads for key tile i+1 fly while tile i is in softmax.
//
// One CTA = 128 threads = 128 queries of one head of one of up to two problems (blockIdx.z: both self-attention calls, or
// both directions of the cross attention, share one launch so the grid fills the 148 SMs x 2 CTAs).  Keys/values stream
// through in tiles of 64.  Per tile:
//   S  (128 x 64, two fp32 accumulators in TMEM)  = Qh Kh^T ;  Qh Kl^T + Ql Kh^T          12 tcgen05.mma (K = 64 dims)
//   every thread owns one query row (= one TMEM lane): tcgen05.ld, online softmax in registers (no shuffles),
//   P = exp(S - m) split into fp16 hi / lo and written to shared memory as the next A operand
//   O' (128 x 64, two accumulators)               = Ph Vh ;  Ph Vl + Pl Vh                12 tcgen05.mma (K = 64 keys,
//                                                                                          V consumed MN-major as stored)
//   O = O * exp(m_old - m_new) + O'  in registers.
// 96 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM: one CTA's softmax overlaps the other's MMAs.
#pragma once
#include "common.cuh"
#include "tc.cuh"

constexpr int AT_Q = 128, AT_KV = 64, AT_D = 64;
constexpr int AT_Q_BYTES = AT_Q * AT_D * 2;    // 16 KB  (Q hi / lo, P hi / lo)
constexpr int AT_KV_BYTES = AT_KV * AT_D * 2;  // 8 KB   (K hi / lo, V hi / lo)
constexpr size_t AT_SMEM = 4 * AT_Q_BYTES + 4 * AT_KV_BYTES + 1024;

struct AttnProblem {
  const __half *Qh, *Ql, *Kh, *Kl, *Vh, *Vl;  // [4][N][64] halves
  __half *Oh, *Ol;                             // output planes [Nq][256], column h * 64 + d (the out_proj GEMM's A operand)
  int Nq, Nk;
};
struct AttnArgs {
  AttnProblem p[2];
  float scale;
  int* err_flag;
};

static __global__ void __launch_bounds__(128, 2) k_flash_tc(AttnArgs args) {
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

  const AttnProblem& pr = args.p[blockIdx.z];
  const int Nq = pr.Nq, Nk = pr.Nk;
  const int t = threadIdx.x, warp = t >> 5;
  const int h = blockIdx.y, q0 = blockIdx.x * AT_Q;
  if (q0 >= Nq) return;  // uniform per CTA
  const float c2 = args.scale * 1.4426950408889634f;  // softmax scale * log2(e): exponentials are taken in base 2
  const __half* Qh = pr.Qh + (size_t)h * Nq * 64;
  const __half* Ql = pr.Ql + (size_t)h * Nq * 64;
  const __half* Kh = pr.Kh + (size_t)h * Nk * 64;
  const __half* Kl = pr.Kl + (size_t)h * Nk * 64;
  const __half* Vh = pr.Vh + (size_t)h * Nk * 64;
  const __half* Vl = pr.Vl + (size_t)h * Nk * 64;

  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (t == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_mbar_init();
  }
  const uint32_t aQh = tc::smem_u32(sQh), aQl = tc::smem_u32(sQl), aKh = tc::smem_u32(sKh), aKl = tc::smem_u32(sKl),
                 aVh = tc::smem_u32(sVh), aVl = tc::smem_u32(sVl), aPh = tc::smem_u32(sPh), aPl = tc::smem_u32(sPl);
  // Q: 128 rows x 8 chunks x {hi, lo}; consecutive threads take consecutive rows of one chunk (conflict-free 16-B stores)
  for (int i = t; i < AT_Q * 8; i += 128) {
    const int r = i & 127, c = i >> 7;
    const int q = q0 + r;
    const uint32_t ok = q < Nq ? 16u : 0u;
    const size_t src = (size_t)(q < Nq ? q : 0) * 64 + c * 8;
    const uint32_t off = tc::canon_off(r, c, AT_Q);
    tc::cp_async16(aQh + off, Qh + src, ok);
    tc::cp_async16(aQl + off, Ql + src, ok);
  }
  auto load_K = [&](int k0) {  // 64 key rows x 8 chunks, K-major canonical
    for (int i = t; i < AT_KV * 8; i += 128) {
      const int r = i & 63, c = i >> 6;
      const int kk = k0 + r;
      const uint32_t ok = kk < Nk ? 16u : 0u;
      const size_t src = (size_t)(kk < Nk ? kk : 0) * 64 + c * 8;
      const uint32_t off = tc::canon_off(r, c, AT_KV);
      tc::cp_async16(aKh + off, Kh + src, ok);
      tc::cp_async16(aKl + off, Kl + src, ok);
    }
  };
  auto load_V = [&](int k0) {  // V stays [key][dim]: MN-major B operand; chunk (key r, dims 8c..) at c*1024 + (r/8)*128 + (r%8)*16
    for (int i = t; i < AT_KV * 8; i += 128) {
      const int r = i & 63, c = i >> 6;
      const int kk = k0 + r;
      const uint32_t ok = kk < Nk ? 16u : 0u;
      const size_t src = (size_t)(kk < Nk ? kk : 0) * 64 + c * 8;
      const uint32_t off = (uint32_t)c * 1024u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
      tc::cp_async16(aVh + off, Vh + src, ok);
      tc::cp_async16(aVl + off, Vl + src, ok);
    }
  };
  load_K(0);
  tc::cp_async_commit();  // group: Q + K(0)
  load_V(0);
  tc::cp_async_commit();  // group: V(0)

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS0 = tmem, tS1 = tmem + 64, tO0 = tmem + 128, tO1 = tmem + 192;
  const uint32_t idescS = tc::idesc_f16(AT_Q, AT_KV);                          // Q K^T: both operands K-major
  const uint32_t idescO = tc::idesc_f16(AT_Q, AT_D) | tc::IDESC_B_MN_MAJOR;    // P V: V is MN-major
  const uint32_t lboQ = (AT_Q / 8) * 128, lboK = (AT_KV / 8) * 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  float m_i = -INFINITY, l_i = 0.f;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  uint32_t phase = 0;
  bool ok = true;

  for (int k0 = 0; k0 < Nk; k0 += AT_KV) {
    tc::cp_async_wait<1>();  // everything but the newest group: K(k0) (and Q) landed
    tc::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t dQh = tc::smem_desc(aQh + 2 * s * lboQ, lboQ), dQl = tc::smem_desc(aQl + 2 * s * lboQ, lboQ);
        const uint64_t dKh = tc::smem_desc(aKh + 2 * s * lboK, lboK), dKl = tc::smem_desc(aKl + 2 * s * lboK, lboK);
        tc::umma_f16(tS0, dQh, dKh, idescS, s ? 1u : 0u);
        tc::umma_f16(tS1, dQh, dKl, idescS, s ? 1u : 0u);
        tc::umma_f16(tS1, dQl, dKh, idescS, 1u);
      }
      tc::umma_commit(&bar[0]);
    }
    ok = tc::mbar_wait(&bar[0], phase) && ok;
    tc::fence_after_sync();
    if (k0 + AT_KV < Nk) load_K(k0 + AT_KV);  // K buffer is free again: prefetch the next tile behind the softmax
    tc::cp_async_commit();

    // ---- online softmax on this thread's row, in base 2: pass 1 = row max, pass 2 = exponentials ------------------
    // (softmax scale and log2(e) are folded into one multiplier; masking only runs on the ragged last tile)
    const bool ragged = k0 + AT_KV > Nk;  // uniform
    float mx = -INFINITY;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      float a0[32], a1[32];
      tc::tmem_ld32(tS0 + lane_off + cc * 32, a0);
      tc::tmem_ld32(tS1 + lane_off + cc * 32, a1);
      if (!ragged) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(a1[j], tc::LO_INV, a0[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (k0 + cc * 32 + j < Nk) ? fmaf(a1[j], tc::LO_INV, a0[j]) : -INFINITY);
      }
    }
    const float m_new = fmaxf(m_i, mx * c2);  // c2 > 0: max commutes with the scaling
    const float corr = tc::ex2(m_i - m_new);
    float rs = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      float a0[32], a1[32];
      tc::tmem_ld32(tS0 + lane_off + cc * 32, a0);
      tc::tmem_ld32(tS1 + lane_off + cc * 32, a1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = c * 8 + 2 * i;
          float pa = tc::ex2(fmaf(fmaf(a1[j], tc::LO_INV, a0[j]), c2, -m_new));
          float pb = tc::ex2(fmaf(fmaf(a1[j + 1], tc::LO_INV, a0[j + 1]), c2, -m_new));
          if (ragged) {
            if (k0 + cc * 32 + j >= Nk) pa = 0.f;
            if (k0 + cc * 32 + j + 1 >= Nk) pb = 0.f;
          }
          rs += pa + pb;
          tc::split2(pa, pb, hi[i], lo[i]);
        }
        const uint32_t off = tc::canon_off(t, cc * 4 + c, AT_Q);
        *reinterpret_cast<uint4*>(sPh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sPl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    l_i = l_i * corr + rs;
    m_i = m_new;
    tc::cp_async_wait<1>();  // V(k0) landed (the newest group is the K prefetch)
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t dPh = tc::smem_desc(aPh + 2 * s * lboQ, lboQ), dPl = tc::smem_desc(aPl + 2 * s * lboQ, lboQ);
        const uint64_t dVh = tc::smem_desc(aVh + s * 256, 128, 1024), dVl = tc::smem_desc(aVl + s * 256, 128, 1024);
        tc::umma_f16(tO0, dPh, dVh, idescO, s ? 1u : 0u);
        tc::umma_f16(tO1, dPh, dVl, idescO, s ? 1u : 0u);
        tc::umma_f16(tO1, dPl, dVh, idescO, 1u);
      }
      tc::umma_commit(&bar[1]);
    }
    ok = tc::mbar_wait(&bar[1], phase) && ok;
    tc::fence_after_sync();
    if (k0 + AT_KV < Nk) load_V(k0 + AT_KV);  // V buffer is free again
    tc::cp_async_commit();
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      float a0[32], a1[32];
      tc::tmem_ld32(tO0 + lane_off + cc * 32, a0);
      tc::tmem_ld32(tO1 + lane_off + cc * 32, a1);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (cc == 0) o[j] = fmaf(o[j], corr, fmaf(a1[j], tc::LO_INV, a0[j]));
        else o[32 + j] = fmaf(o[32 + j], corr, fmaf(a1[j], tc::LO_INV, a0[j]));
      }
    }
    phase ^= 1;
    tc::fence_before_sync();  // TMEM reads of this tile are ordered before the next tile's MMAs (issued after the next sync)
  }
  tc::cp_async_wait<0>();
  if (!ok && args.err_flag) *args.err_flag = 1;
  const int q = q0 + t;
  if (q < Nq) {
    const float inv = 1.0f / l_i;
    uint4* dh = reinterpret_cast<uint4*>(pr.Oh + (size_t)q * 256 + h * 64);
    uint4* dl = reinterpret_cast<uint4*>(pr.Ol + (size_t)q * 256 + h * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tc::split2(o[8 * c + 2 * i] * inv, o[8 * c + 2 * i + 1] * inv, hi[i], lo[i]);
      dh[c] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dl[c] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
