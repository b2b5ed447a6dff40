// Warp-specialised tcgen05 + TMA flash attention with split-fp16 operands (~fp32 accuracy).
//
//   O[Nq][256] = softmax(scale * Q K^T) V per head,  inputs / outputs as fp16 hi / lo planes (x ~= hi + lo * 2^-11).
//
// One CTA (320 threads, one per SM) owns TWO 128-query tiles of one head and streams a range of 64-key tiles:
//   warp 8 lane 0 : TMA producer  - Q once, then K / V tiles through 2-stage rings (128-byte swizzled, zero OOB fill)
//   warp 9        : TMEM allocator (512 columns); lane 0 = MMA issuer.  Per key tile and query tile q:
//                     S_q  (128 x 64, ONE fp32 accumulator, double-buffered; q / k lo planes are unscaled)
//                                                            = Qh Kh^T + Qh Kl^T + Ql Kh^T       12 tcgen05.mma
//                     O'_q (128 x 64, two accumulators)      = Ph Vh   ; Ph Vl + Pl Vh           12 tcgen05.mma
//                   issued in the order PV_0(i), S_0(i+2), PV_1(i), S_1(i+2): the logits run two key tiles ahead of the
//                   softmax, so a warpgroup never waits for the tensor pipe in steady state.
//   warps 0-3 / 4-7: softmax warpgroup of query tile 0 / 1, one thread per query row (= TMEM lane): tcgen05.ld S, online
//                   softmax in base 2 entirely in registers, P = 2^(s - m) split to fp16 hi / lo and stored (128-byte
//                   swizzled) as the next A operand, O = (O + O'_{i-1}) * 2^(m_old - m_new) folded in one tile late so the
//                   exponentials overlap the PV product.
// All hand-offs are mbarriers: TMA complete_tx (q_full, k_full, v_full), tcgen05.commit (s_full, o_full, k_empty, v_empty),
// and 128-thread arrivals (p_full).  With `nsplit` > 1 a CTA covers a slice of the key range and writes un-normalised
// partials (O, m, l); k_attn_merge combines them (keeps the 148 SMs busy when tiles x heads is just over a wave).
#pragma once
#include "common.cuh"
#include "gemm_tma.cuh"
#include "tc.cuh"

constexpr int AW_Q = 128, AW_KV = 64, AW_D = 64;
constexpr int AW_Q_BYTES = AW_Q * AW_D * 2;    // 16 KB per plane
constexpr int AW_KV_BYTES = AW_KV * AW_D * 2;  // 8 KB per plane
// smem: Q 2 tiles x 2 planes (64 KB) | P 2 tiles x 2 planes (64 KB) | K 2 stages x 2 planes (32 KB) | V same (32 KB)
constexpr int AW_OFF_Q = 0, AW_OFF_P = 4 * AW_Q_BYTES, AW_OFF_K = 8 * AW_Q_BYTES, AW_OFF_V = AW_OFF_K + 4 * AW_KV_BYTES;
constexpr int AW_TILE_BYTES = AW_OFF_V + 4 * AW_KV_BYTES;  // 192 KB
constexpr size_t AW_SMEM = AW_TILE_BYTES + 1024 + 256;
constexpr int AW_THREADS = 320;

struct AttnWsMaps {
  CUtensorMap qh[2], ql[2], kh[2], kl[2], vh[2], vl[2];  // per problem; 2-D views [4 * N rows][64] of the head-major planes
};
struct AttnWsProblem {
  __half *Oh, *Ol;  // final output planes [Nq][256] (column h * 64 + d)            (nsplit == 1)
  float* Opart;     // partial O   [nsplit][Nq][256] fp32, un-normalised            (nsplit > 1)
  float* ml;        // partial m,l [nsplit][4][Nq][2]
  int Nq, Nk;
};
struct AttnWsArgs {
  AttnWsProblem p[2];
  float scale;
  int nsplit;
  int* err_flag;
};

namespace tc {
// MN-major operand stored [k][64 mn-elements] with 128-byte rows, 128-byte swizzle: SBO = 8-row group stride
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr) { return smem_desc_sw128(saddr); }
}  // namespace tc

static __global__ void __launch_bounds__(AW_THREADS, 1) k_flash_ws(const __grid_constant__ AttnWsMaps maps, AttnWsArgs args) {
  extern __shared__ unsigned char aw_raw[];
  const uint32_t raw = tc::smem_u32(aw_raw);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  unsigned char* sm = aw_raw + (smem0 - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AW_TILE_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [4] query tile x logits buffer
  uint64_t* p_full = bars + 13;     // [2]
  uint64_t* o_full = bars + 15;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int z = blockIdx.z / args.nsplit, split = blockIdx.z % args.nsplit;
  const AttnWsProblem& pr = args.p[z];
  const int Nq = pr.Nq, Nk = pr.Nk;
  const int h = blockIdx.y, q0 = blockIdx.x * (2 * AW_Q);
  if (q0 >= Nq) return;  // uniform
  // key-tile range of this split
  const int tiles_total = (Nk + AW_KV - 1) / AW_KV;
  const int per = (tiles_total + args.nsplit - 1) / args.nsplit;
  const int tile0 = split * per, tile1 = min(tiles_total, tile0 + per);
  const int T = tile1 - tile0;  // may be <= 0 for a trailing split: then this CTA writes neutral partials

  if (t == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&k_full[i], 1), tc::mbar_init(&k_empty[i], 1), tc::mbar_init(&v_full[i], 1), tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1), tc::mbar_init(&s_full[2 + i], 1), tc::mbar_init(&p_full[i], 128), tc::mbar_init(&o_full[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  bool ok = true;

  if (warp == 8) {
    if (lane == 0 && T > 0) {
      // ===== TMA producer =====
      tc::mbar_expect_tx(q_full, 4 * AW_Q_BYTES);
      for (int q = 0; q < 2; ++q) {
        tc::tma_load_2d(smem0 + AW_OFF_Q + (2 * q) * AW_Q_BYTES, &maps.qh[z], q_full, 0, h * Nq + q0 + q * AW_Q);
        tc::tma_load_2d(smem0 + AW_OFF_Q + (2 * q + 1) * AW_Q_BYTES, &maps.ql[z], q_full, 0, h * Nq + q0 + q * AW_Q);
      }
      for (int i = 0; i < T; ++i) {
        const int s = i & 1;
        const int row = h * Nk + (tile0 + i) * AW_KV;
        if (i >= 2) ok = tc::mbar_wait(&k_empty[s], ((i >> 1) - 1) & 1) && ok;
        tc::mbar_expect_tx(&k_full[s], 2 * AW_KV_BYTES);
        tc::tma_load_2d(smem0 + AW_OFF_K + (2 * s) * AW_KV_BYTES, &maps.kh[z], &k_full[s], 0, row);
        tc::tma_load_2d(smem0 + AW_OFF_K + (2 * s + 1) * AW_KV_BYTES, &maps.kl[z], &k_full[s], 0, row);
        if (i >= 2) ok = tc::mbar_wait(&v_empty[s], ((i >> 1) - 1) & 1) && ok;
        tc::mbar_expect_tx(&v_full[s], 2 * AW_KV_BYTES);
        tc::tma_load_2d(smem0 + AW_OFF_V + (2 * s) * AW_KV_BYTES, &maps.vh[z], &v_full[s], 0, row);
        tc::tma_load_2d(smem0 + AW_OFF_V + (2 * s + 1) * AW_KV_BYTES, &maps.vl[z], &v_full[s], 0, row);
      }
    }
  } else if (warp == 9) {
    if (lane == 0 && T > 0) {
      // ===== MMA issuer =====
      const uint32_t idS = tc::idesc_f16(AW_Q, AW_KV);                        // Q K^T: both K-major
      const uint32_t idO = tc::idesc_f16(AW_Q, AW_D) | tc::IDESC_B_MN_MAJOR;  // P V: V as stored = MN-major
      auto issue_S = [&](int q, int s, int buf) {
        const uint64_t dQh = tc::smem_desc_sw128(smem0 + AW_OFF_Q + (2 * q) * AW_Q_BYTES);
        const uint64_t dQl = tc::smem_desc_sw128(smem0 + AW_OFF_Q + (2 * q + 1) * AW_Q_BYTES);
        const uint64_t dKh = tc::smem_desc_sw128(smem0 + AW_OFF_K + (2 * s) * AW_KV_BYTES);
        const uint64_t dKl = tc::smem_desc_sw128(smem0 + AW_OFF_K + (2 * s + 1) * AW_KV_BYTES);
        const uint32_t tS = tmem + q * 256 + buf * 64;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);  // 16 dims = 32 bytes
          tc::umma_f16(tS, dQh + adv, dKh + adv, idS, ks ? 1u : 0u);
          tc::umma_f16(tS, dQh + adv, dKl + adv, idS, 1u);  // unscaled lo planes: all three products share the accumulator
          tc::umma_f16(tS, dQl + adv, dKh + adv, idS, 1u);
        }
        tc::umma_commit(&s_full[q * 2 + buf]);
      };
      auto issue_PV = [&](int q, int s) {
        const uint64_t dPh = tc::smem_desc_sw128(smem0 + AW_OFF_P + (2 * q) * AW_Q_BYTES);
        const uint64_t dPl = tc::smem_desc_sw128(smem0 + AW_OFF_P + (2 * q + 1) * AW_Q_BYTES);
        const uint64_t dVh = tc::smem_desc_sw128_mn(smem0 + AW_OFF_V + (2 * s) * AW_KV_BYTES);
        const uint64_t dVl = tc::smem_desc_sw128_mn(smem0 + AW_OFF_V + (2 * s + 1) * AW_KV_BYTES);
        const uint32_t tO0 = tmem + q * 256 + 128, tO1 = tO0 + 64;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t advP = (uint64_t)(ks * 2);    // 16 keys along P's K dimension = 32 bytes
          const uint64_t advV = (uint64_t)(ks * 128);  // 16 keys = two 8-row groups of V = 2048 bytes
          tc::umma_f16(tO0, dPh + advP, dVh + advV, idO, ks ? 1u : 0u);
          tc::umma_f16(tO1, dPh + advP, dVl + advV, idO, ks ? 1u : 0u);
          tc::umma_f16(tO1, dPl + advP, dVh + advV, idO, 1u);
        }
        tc::umma_commit(&o_full[q]);
      };
      ok = tc::mbar_wait(q_full, 0) && ok;
      ok = tc::mbar_wait(&k_full[0], 0) && ok;
      tc::fence_after_sync();
      issue_S(0, 0, 0);
      issue_S(1, 0, 0);
      tc::umma_commit(&k_empty[0]);
      if (T > 1) {
        ok = tc::mbar_wait(&k_full[1], 0) && ok;
        tc::fence_after_sync();
        issue_S(0, 1, 1);
        issue_S(1, 1, 1);
        tc::umma_commit(&k_empty[1]);
      }
      for (int i = 0; i < T; ++i) {
        const int sv = i & 1;           // V stage of tile i; also the K stage and logits buffer of tile i + 2
        const bool more = i + 2 < T;
        ok = tc::mbar_wait(&v_full[sv], (i >> 1) & 1) && ok;
        if (more) ok = tc::mbar_wait(&k_full[sv], ((i + 2) >> 1) & 1) && ok;
        for (int q = 0; q < 2; ++q) {
          ok = tc::mbar_wait(&p_full[q], i & 1) && ok;  // P_q(i) in smem; logits buffer i & 1 and O'_q(i-1) consumed
          tc::fence_after_sync();
          issue_PV(q, sv);
          if (more) issue_S(q, sv, sv);
        }
        tc::umma_commit(&v_empty[sv]);
        if (more) tc::umma_commit(&k_empty[sv]);
      }
    }
  } else {
    // ===== softmax warpgroups: q = 0 (warps 0-3), q = 1 (warps 4-7); thread = query row = TMEM lane =====
    const int q = warp >> 2;
    const int r = t & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tSb = tmem + q * 256 + lane_off, tO0 = tSb + 128, tO1 = tSb + 192;
    unsigned char* sPh = sm + AW_OFF_P + (2 * q) * AW_Q_BYTES;
    unsigned char* sPl = sPh + AW_Q_BYTES;
    const float c2 = args.scale * 1.4426950408889634f;
    float m_i = -INFINITY, l_i = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;

    for (int i = 0; i < T; ++i) {
      ok = tc::mbar_wait(&s_full[q * 2 + (i & 1)], (i >> 1) & 1) && ok;
      tc::fence_after_sync();
      const uint32_t tS = tSb + (i & 1) * 64;
      const int k0 = (tile0 + i) * AW_KV;
      const bool ragged = k0 + AW_KV > Nk;
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        float a0[32];
        tc::tmem_ld32(tS + cc * 32, a0);
        if (!ragged) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, a0[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (k0 + cc * 32 + j < Nk) ? a0[j] : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_i, mx * c2);
      const float corr = tc::ex2(m_i - m_new);
      uint32_t ph[32], pl[32];  // P row, packed half2 (hi / lo planes), kept in registers until PV(i-1) has drained
      float rs = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float a0[32];
        tc::tmem_ld32(tS + cc * 32, a0);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int j = 2 * jj;
          float pa = tc::ex2(fmaf(a0[j], c2, -m_new));
          float pb = tc::ex2(fmaf(a0[j + 1], c2, -m_new));
          if (ragged) {
            if (k0 + cc * 32 + j >= Nk) pa = 0.f;
            if (k0 + cc * 32 + j + 1 >= Nk) pb = 0.f;
          }
          rs += pa + pb;
          tc::split2(pa, pb, ph[cc * 16 + jj], pl[cc * 16 + jj]);
        }
      }
      l_i = l_i * corr + rs;
      m_i = m_new;
      // fold in the previous tile's PV product (this also guarantees the tensor core is done reading P_{i-1})
      if (i > 0) {
        ok = tc::mbar_wait(&o_full[q], (i - 1) & 1) && ok;
        tc::fence_after_sync();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float a0[32], a1[32];
          tc::tmem_ld32(tO0 + cc * 32, a0);
          tc::tmem_ld32(tO1 + cc * 32, a1);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[cc * 32 + j] = (o[cc * 32 + j] + fmaf(a1[j], tc::LO_INV, a0[j])) * corr;
        }
      }
      // store P_i (row r, 8 chunks of 16 bytes, 128-byte swizzle: chunk c lands at c ^ (r % 8))
      {
        const uint32_t rowoff = (uint32_t)r * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = rowoff + (uint32_t)((c ^ (r & 7)) * 16);
          *reinterpret_cast<uint4*>(sPh + off) = make_uint4(ph[4 * c], ph[4 * c + 1], ph[4 * c + 2], ph[4 * c + 3]);
          *reinterpret_cast<uint4*>(sPl + off) = make_uint4(pl[4 * c], pl[4 * c + 1], pl[4 * c + 2], pl[4 * c + 3]);
        }
      }
      tc::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core
      tc::fence_before_sync();   // our tcgen05.ld of S_i / O'_{i-1} are ordered before the MMAs the issuer starts next
      tc::mbar_arrive(&p_full[q]);
    }
    if (T > 0) {
      ok = tc::mbar_wait(&o_full[q], (T - 1) & 1) && ok;
      tc::fence_after_sync();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float a0[32], a1[32];
        tc::tmem_ld32(tO0 + cc * 32, a0);
        tc::tmem_ld32(tO1 + cc * 32, a1);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[cc * 32 + j] += fmaf(a1[j], tc::LO_INV, a0[j]);
      }
    }
    const int qrow = q0 + q * AW_Q + r;
    if (qrow < Nq) {
      if (args.nsplit == 1) {
        const float inv = 1.0f / l_i;
        uint4* dh = reinterpret_cast<uint4*>(pr.Oh + (size_t)qrow * 256 + h * 64);
        uint4* dl = reinterpret_cast<uint4*>(pr.Ol + (size_t)qrow * 256 + h * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) tc::split2(o[8 * c + 2 * i] * inv, o[8 * c + 2 * i + 1] * inv, hi[i], lo[i]);
          dh[c] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dl[c] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(pr.Opart + ((size_t)split * Nq + qrow) * 256 + h * 64);
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        float* ml = pr.ml + (((size_t)split * 4 + h) * Nq + qrow) * 2;
        ml[0] = m_i;  // -inf when this split saw no keys
        ml[1] = l_i;
      }
    }
  }
  __syncwarp();
  if (!ok && args.err_flag) *args.err_flag = 1;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc(tmem, 512);
}

// combine the key-range partials of k_flash_ws: O = sum_s O_s 2^(m_s - m) / sum_s l_s 2^(m_s - m), written as split planes
static __global__ void __launch_bounds__(256) k_attn_merge(const float* __restrict__ Opart, const float* __restrict__ ml, int Nq,
                                                            int nsplit, __half* __restrict__ Oh, __half* __restrict__ Ol) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over Nq * 4 heads * 32 column pairs
  if (idx >= Nq * 128) return;
  const int cp = idx & 31, h = (idx >> 5) & 3, qrow = idx >> 7;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, ml[(((size_t)s * 4 + h) * Nq + qrow) * 2]);
  float l = 0.f, a = 0.f, b = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float* p = ml + (((size_t)s * 4 + h) * Nq + qrow) * 2;
    const float w = tc::ex2(p[0] - m);  // 0 for an empty split (m_s = -inf)
    l = fmaf(p[1], w, l);
    const float2 v = *reinterpret_cast<const float2*>(Opart + ((size_t)s * Nq + qrow) * 256 + h * 64 + 2 * cp);
    a = fmaf(v.x, w, a);
    b = fmaf(v.y, w, b);
  }
  const float inv = 1.0f / l;
  uint32_t hi, lo;
  tc::split2(a * inv, b * inv, hi, lo);
  *reinterpret_cast<uint32_t*>(Oh + (size_t)qrow * 256 + h * 64 + 2 * cp) = hi;
  *reinterpret_cast<uint32_t*>(Ol + (size_t)qrow * 256 + h * 64 + 2 * cp) = lo;
}
