// Warp-specialised tcgen05 + TMA flash attention, A operands in TMEM ("TS" MMAs), split-fp16 operands (~fp32 accuracy).
//
//   O[Nq][256] = softmax(scale * Q K^T) V per head; q / k / v arrive as fp16 hi / lo planes with UNSCALED lo
//   (x ~= hi + lo, lo = fp16(x - hi)), the output leaves as hi / lo planes with the usual 2^11-scaled lo.
//
// Why TMEM operands: a 128 x 64 x 16 SS-mode MMA reads 4 KB of A and 2 KB of B from shared memory per 32 tensor cycles
// (192 B/clk > the 128 B/clk the SM's shared memory delivers), so the smem-operand version of this kernel (attn_ws.cuh)
// is shared-memory-bound at ~2x the tensor time (ncu: tensor pipe 24 % active).  Here Q lives in TMEM for the whole CTA
// and P is written back over its own logits in TMEM, so shared memory only carries the K / V tiles (B operands).
//
// One CTA (320 threads, one per SM) owns TWO 128-query tiles of one head and streams a range of 64-key tiles.
// TMEM (512 columns; per query tile q at q * 256):
//     [  0, 64) logits buffer 0: S (128 x 64 fp32), overwritten in place by P: hi = columns [0, 32), lo = [32, 64)
//     [ 64,128) logits buffer 1                                                 (column c = keys 2c, 2c + 1 as half2)
//     [128,192) O accumulator (128 x 64 fp32), accumulated by the tensor core across ALL key tiles
//     [192,224) Q hi, [224,256) Q lo  (column c = dims 2c, 2c + 1)
//   warp 8 lane 0 : TMA producer - K and V tiles through two 4-stage rings (128-byte swizzled, zero OOB fill)
//   warps 9, 10   : MMA issuers of query tile 0 / 1 (warp-uniform issue, one elected lane; warp 9 also owns the TMEM
//                   allocation).  Two issuers because a barrier wait costs the issuing warp 150-250 cycles and the
//                   tcgen05 queue is shallow: while one warp waits for its P tile the other's MMAs keep the pipe busy.
//                   Per key tile i, issuer q:
//                     PV_q(i): O_q += Ph Vh + Ph Vl + Pl Vh   (A = P from TMEM, B = V MN-major)      12 tcgen05.mma
//                     S_q(i+2) = Qh Kh^T + Qh Kl^T + Ql Kh^T  (A = Q from TMEM) into buffer i & 1    12 tcgen05.mma
//                   the logits run two key tiles ahead of the softmax.
//   warps 0-3 / 4-7: softmax warpgroup of query tile 0 / 1, one thread per query row (= TMEM lane): tcgen05.ld S, base-2
//                   online softmax with a LAZY reference maximum (O and l are only rescaled when the row maximum grew by
//                   more than 2^8 - P then stays <= 256, exact in the hi / lo split - so O normally never leaves TMEM),
//                   P = 2^(s - m) split to fp16 hi / lo and stored over S with tcgen05.st.
// Hand-offs are mbarriers: TMA complete_tx (k_full, v_full), tcgen05.commit (s_full, o_full, k_empty, v_empty) and
// 128-thread arrivals (q_ready, p_full).  With `nsplit` > 1 a CTA covers a slice of the key range and writes un-normalised
// partials (O, m, l) that k_attn_merge combines.
#pragma once
#include "attn_ws.cuh"

constexpr int AS_NS = 4;                     // ring depth; entry e holds K tile e and V tile e - 2 (consumed together)
constexpr int AS_HALF = 2 * AW_KV_BYTES;     // hi + lo plane of one 64 x 64 tile = 16 KB
constexpr int AS_STAGE = 2 * AS_HALF;        // K part at +0, V part at +AS_HALF
constexpr int AS_TILE_BYTES = AS_NS * AS_STAGE;
constexpr size_t AS_SMEM = AS_TILE_BYTES + 1024 + 512;
constexpr uint32_t AS_COL_O = 128, AS_COL_Q = 192;
constexpr int AS_THREADS = 352;     // 8 softmax warps + TMA producer warp + 2 MMA issuer warps
constexpr float AS_RESCALE = 8.0f;  // log2 of the largest P allowed before the reference maximum is refreshed

struct AttnTsMaps {
  CUtensorMap kh[2], kl[2], vh[2], vl[2];  // per problem; 2-D views [4 * N rows][64] of the head-major planes
};
struct AttnTsProblem {
  const __half *Qh, *Ql;  // head-major planes [4][Nq][64]
  __half *Oh, *Ol;        // final output planes [Nq][256] (column h * 64 + d)            (nsplit == 1)
  float* Opart;           // partial O   [nsplit][Nq][256] fp32, un-normalised            (nsplit > 1)
  float* ml;              // partial m,l [nsplit][4][Nq][2]
  int Nq, Nk;
};
struct AttnTsArgs {
  AttnTsProblem p[2];
  float scale;
  int nsplit;
  int* err_flag;
#ifdef B2_ATTN_TIMING
  long long* timing;  // scratch/attn_bench.cu: per-phase clock64 sums of CTA 0
#endif
};

#ifdef B2_ATTN_TIMING
#define AS_T0() long long t_ = clock64()
#define AS_ACC(slot) do { if (stamp) { const long long n_ = clock64(); tacc[slot] += n_ - t_; t_ = n_; } } while (0)
#else
#define AS_T0() do { } while (0)
#define AS_ACC(slot) do { } while (0)
#endif

static __global__ void __launch_bounds__(AS_THREADS, 1) k_flash_ts(const __grid_constant__ AttnTsMaps maps, AttnTsArgs args) {
  extern __shared__ unsigned char as_raw[];
  const uint32_t raw = tc::smem_u32(as_raw);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  unsigned char* sm = as_raw + (smem0 - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AS_TILE_BYTES);
  uint64_t* kv_full = bars;                 // [AS_NS]
  uint64_t* kv_empty = kv_full + AS_NS;     // [AS_NS]
  uint64_t* s_full = kv_empty + AS_NS;      // [4] query tile x logits buffer
  uint64_t* p_full = s_full + 4;            // [2]
  uint64_t* o_full = p_full + 2;            // [2]
  uint64_t* q_ready = o_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 2);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int z = blockIdx.z / args.nsplit, split = blockIdx.z % args.nsplit;
  const AttnTsProblem& pr = args.p[z];
  const int Nq = pr.Nq, Nk = pr.Nk;
  const int h = blockIdx.y, q0 = blockIdx.x * (2 * AW_Q);
  if (q0 >= Nq) return;  // uniform
  const int tiles_total = (Nk + AW_KV - 1) / AW_KV;
  const int per = (tiles_total + args.nsplit - 1) / args.nsplit;
  const int tile0 = split * per, tile1 = min(tiles_total, tile0 + per);
  const int T = tile1 - tile0;  // may be <= 0 for a trailing split: then this CTA writes neutral partials

  if (t == 0) {
    for (int i = 0; i < AS_NS; ++i) tc::mbar_init(&kv_full[i], 1), tc::mbar_init(&kv_empty[i], 2);  // released by both issuers
    for (int i = 0; i < 4; ++i) tc::mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&p_full[i], 128), tc::mbar_init(&o_full[i], 1), tc::mbar_init(&q_ready[i], 128);
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
      // ===== TMA producer: ring entry e = K tile e (if any) + V tile e - 2 (if any), one barrier pair per entry =====
      for (int e = 0; e < T + 2; ++e) {
        const int s = e % AS_NS;
        if (e >= AS_NS) ok = tc::mbar_wait(&kv_empty[s], ((e / AS_NS) - 1) & 1) && ok;
        const bool hk = e < T, hv = e >= 2;
        tc::mbar_expect_tx(&kv_full[s], (hk ? AS_HALF : 0) + (hv ? AS_HALF : 0));
        const uint32_t dst = smem0 + s * AS_STAGE;
        if (hk) {
          const int row = h * Nk + (tile0 + e) * AW_KV;
          tc::tma_load_2d(dst, &maps.kh[z], &kv_full[s], 0, row);
          tc::tma_load_2d(dst + AW_KV_BYTES, &maps.kl[z], &kv_full[s], 0, row);
        }
        if (hv) {
          const int row = h * Nk + (tile0 + e - 2) * AW_KV;
          tc::tma_load_2d(dst + AS_HALF, &maps.vh[z], &kv_full[s], 0, row);
          tc::tma_load_2d(dst + AS_HALF + AW_KV_BYTES, &maps.vl[z], &kv_full[s], 0, row);
        }
      }
    }
  } else if (warp >= 9) {
    if (T > 0) {
      const int q = warp - 9;  // each issuer warp owns one query tile: its barrier waits overlap the other's MMAs
      // ===== MMA issuer: the whole warp runs this with uniform operands, one elected lane issues (tc.cuh) =====
      const uint32_t idS = tc::idesc_f16(AW_Q, AW_KV);                        // Q K^T: A (TMEM) and B K-major
      const uint32_t idO = tc::idesc_f16(AW_Q, AW_D) | tc::IDESC_B_MN_MAJOR;  // P V: V as stored = MN-major
      auto issue_S = [&](int q, int j) {  // logits of key tile j into buffer j & 1
        const int s = j % AS_NS;  // ring entry j
        const uint64_t dKh = tc::smem_desc_sw128(smem0 + s * AS_STAGE);
        const uint64_t dKl = tc::smem_desc_sw128(smem0 + s * AS_STAGE + AW_KV_BYTES);
        const uint32_t tS = tmem + q * 256 + (j & 1) * 64;
        const uint32_t tQh = tmem + q * 256 + AS_COL_Q, tQl = tQh + 32;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);  // 16 dims = 32 bytes of K's rows
          const uint32_t ac = (uint32_t)(ks * 8);   // 16 dims = 8 TMEM columns of Q
          tc::umma_f16_ts_w(tS, tQh + ac, dKh + adv, idS, ks ? 1u : 0u);
          tc::umma_f16_ts_w(tS, tQh + ac, dKl + adv, idS, 1u);
          tc::umma_f16_ts_w(tS, tQl + ac, dKh + adv, idS, 1u);
        }
        tc::umma_commit_w(&s_full[q * 2 + (j & 1)]);
      };
      auto issue_PV = [&](int q, int j) {
        const int s = (j + 2) % AS_NS;  // ring entry j + 2
        const uint64_t dVh = tc::smem_desc_sw128_mn(smem0 + s * AS_STAGE + AS_HALF);
        const uint64_t dVl = tc::smem_desc_sw128_mn(smem0 + s * AS_STAGE + AS_HALF + AW_KV_BYTES);
        const uint32_t tPh = tmem + q * 256 + (j & 1) * 64, tPl = tPh + 32;
        const uint32_t tO = tmem + q * 256 + AS_COL_O;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ac = (uint32_t)(ks * 8);      // 16 keys = 8 TMEM columns of P
          const uint64_t advV = (uint64_t)(ks * 128);  // 16 keys = two 8-row groups of V = 2048 bytes
          tc::umma_f16_ts_w(tO, tPh + ac, dVh + advV, idO, (j | ks) ? 1u : 0u);
          tc::umma_f16_ts_w(tO, tPh + ac, dVl + advV, idO, 1u);
          tc::umma_f16_ts_w(tO, tPl + ac, dVh + advV, idO, 1u);
        }
        tc::umma_commit_w(&o_full[q]);
      };
      ok = tc::mbar_wait(&q_ready[q], 0) && ok;
      for (int j = 0; j < 2 && j < T; ++j) {
        ok = tc::mbar_wait(&kv_full[j], 0) && ok;
        __syncwarp();
        tc::fence_after_sync();
        issue_S(q, j);
        tc::umma_commit_w(&kv_empty[j]);
      }
#ifdef B2_ATTN_TIMING
      const bool stamp = args.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
      AS_T0();
      bool h_kv = false;  // ring entry already seen complete by a pre-poll issued before the previous batch of MMAs
      for (int i = 0; i < T; ++i) {
        const int e = i + 2;  // ring entry: V tile i and (if any) K tile i + 2
        const bool more = e < T;
        if (!h_kv) ok = tc::mbar_wait(&kv_full[e % AS_NS], (e / AS_NS) & 1) && ok;
        AS_ACC(0);
        ok = tc::mbar_wait(&p_full[q], i & 1) && ok;  // P_q(i) stored over S_q(i); O_q rescaled if it had to be
        __syncwarp();
        tc::fence_after_sync();
        AS_ACC(1);
        h_kv = (i + 1 < T) && tc::mbar_test(&kv_full[(e + 1) % AS_NS], ((e + 1) / AS_NS) & 1);
        issue_PV(q, i);
        if (more) issue_S(q, e);  // overwrites buffer i & 1 = P_q(i): in issue order after PV_q(i) has read it
        tc::umma_commit_w(&kv_empty[e % AS_NS]);
        AS_ACC(2);
      }
#ifdef B2_ATTN_TIMING
      if (stamp) for (int i = 0; i < 3; ++i) args.timing[16 + 3 * q + i] = tacc[i];
#endif
    }
  } else {
    // ===== softmax warpgroups: q = 0 (warps 0-3), q = 1 (warps 4-7); thread = query row = TMEM lane =====
    const int q = warp >> 2;
    const int r = t & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tB = tmem + q * 256 + lane_off, tO = tB + AS_COL_O;
    const int qrow = q0 + q * AW_Q + r;
    const float c2 = args.scale * 1.4426950408889634f;
    float m_ref = -INFINITY, l_i = 0.f;
    float o[64];

    if (T > 0) {
      // this thread's query row -> TMEM (zero rows past the end)
      uint32_t w[32];
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const __half* src = (pl ? pr.Ql : pr.Qh) + ((size_t)h * Nq + qrow) * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (qrow < Nq) v = __ldg(reinterpret_cast<const uint4*>(src) + c);
          w[4 * c] = v.x, w[4 * c + 1] = v.y, w[4 * c + 2] = v.z, w[4 * c + 3] = v.w;
        }
        tc::tmem_st32(tB + AS_COL_Q + pl * 32, w);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(&q_ready[q]);
    }

#ifdef B2_ATTN_TIMING
    const bool stamp = args.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (r == 0);
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    AS_T0();
    bool h_s = false;  // pre-polled: logits of the next tile already complete
    for (int i = 0; i < T; ++i) {
      if (!h_s) ok = tc::mbar_wait(&s_full[q * 2 + (i & 1)], (i >> 1) & 1) && ok;
      tc::fence_after_sync();
      AS_ACC(0);
      const uint32_t tS = tB + (i & 1) * 64;
      const int k0 = (tile0 + i) * AW_KV;
      float a[64];
      tc::tmem_ld64(tS, a);
      if (k0 + AW_KV > Nk) {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (k0 + j >= Nk) a[j] = -INFINITY;  // 2^(-inf) = 0
      }
      float mx = a[0];
#pragma unroll
      for (int j = 1; j < 64; ++j) mx = fmaxf(mx, a[j]);
      const float m_new = fmaxf(m_ref, mx * c2);
      bool waited = i == 0 || tc::mbar_test(&o_full[q], (i - 1) & 1);  // PV_q(i-1) landed? (pre-poll, consumed below)
      AS_ACC(1);
      if (__any_sync(0xffffffffu, m_new - m_ref > AS_RESCALE)) {  // also true on the first tile (m_ref = -inf)
        const float corr = tc::ex2(m_ref - m_new);
        l_i *= corr;
        if (i > 0) {  // bring O (in TMEM) to the new reference; PV_q(i-1) must have landed
          if (!waited) ok = tc::mbar_wait(&o_full[q], (i - 1) & 1) && ok;
          tc::fence_after_sync();
          waited = true;
          tc::tmem_ld64(tO, o);
          uint32_t ow[64];
#pragma unroll
          for (int j = 0; j < 64; ++j) ow[j] = __float_as_uint(o[j] * corr);
          tc::tmem_st32(tO, ow);
          tc::tmem_st32(tO + 32, ow + 32);
        }
        m_ref = m_new;
      }
      AS_ACC(2);
      uint32_t ph[32], pl[32];
      float rs = 0.f;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const float pa = tc::ex2(fmaf(a[2 * jj], c2, -m_ref));
        const float pb = tc::ex2(fmaf(a[2 * jj + 1], c2, -m_ref));
        rs += pa + pb;
        tc::split2_unscaled(pa, pb, ph[jj], pl[jj]);
      }
      l_i += rs;
      AS_ACC(3);
      h_s = (i + 1 < T) && tc::mbar_test(&s_full[q * 2 + ((i + 1) & 1)], ((i + 1) >> 1) & 1);
      tc::tmem_st32(tS, ph);  // P_i over S_i (this thread's own row; every column of it is already in registers)
      tc::tmem_st32(tS + 32, pl);
      tc::tmem_st_wait();
      AS_ACC(4);
      if (!waited) ok = tc::mbar_wait(&o_full[q], (i - 1) & 1) && ok;  // every phase is observed once: parities stay unambiguous
      AS_ACC(5);
      tc::fence_before_sync();
      tc::mbar_arrive(&p_full[q]);
      AS_ACC(6);
    }
#ifdef B2_ATTN_TIMING
    if (stamp) for (int i = 0; i < 8; ++i) args.timing[q * 8 + i] = tacc[i];
    if (stamp) args.timing[24] = T;
#endif
#pragma unroll
    for (int j = 0; j < 64; ++j) o[j] = 0.f;
    if (T > 0) {
      ok = tc::mbar_wait(&o_full[q], (T - 1) & 1) && ok;
      tc::fence_after_sync();
      tc::tmem_ld64(tO, o);
    }
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
        ml[0] = m_ref;  // -inf when this split saw no keys
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
