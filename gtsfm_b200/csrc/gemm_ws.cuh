// Persistent, warp-specialised tcgen05 + TMA split-fp16 "NT" GEMM over a BATCH of problems that share one weight matrix:
//     C_z[M_z, N] = [A1_z | A2_z][M_z, K] * B[N, K]^T        z = 0 .. nprob-1  (images of a batch of pairs)
// or, with `b_per_problem`, one (A_z, B_z, N_z) triple per problem (the assignment similarity of every pair of a batch).
//
// Arithmetic: every fp32 value travels as two fp16 planes, x ~= hi + lo * 2^-11 (the producing kernel's epilogue writes
// them); three MMAs per K step,  acc0 += Ah Bh ; acc1 += Ah Bl + Al Bh ; result = acc0 + acc1 * 2^-11, both fp32
// accumulators in TMEM.  The dropped Al * Bl term is 2^-22 relative.
//
// Schedule: one CTA per SM walks output tiles `blockIdx.x + i * gridDim.x` of the whole batch (problem-major, then row
// tile, then column tile - concurrently running CTAs share the A row tile through L2):
//   warp 0 (lane 0)  TMA producer: K chunks of consecutive tiles flow through one 2-stage ring (2 x 64 KB) without draining
//   warp 1           TMEM allocator (512 columns = two accumulator sets of 2 x 128) + MMA issuer (warp-uniform issue, tc.cuh)
//   warps 2-17       epilogue, ONE 32 x 32 block of the tile each (TMEM lane quarter w % 4 = rows, column block (w - 2) / 4):
//                    operands that do not depend on the accumulator (bias) are fetched first, then wait acc_full[set] ->
//                    tcgen05.ld -> release the set (acc_empty) -> 128-bit stores into a padded transpose pad -> read back as
//                    (row, 4 consecutive columns) per lane -> bias / scale / ReLU / residual -> 16-byte fp32 and 8-byte
//                    plane stores (8 lanes cover 128 contiguous bytes of a row).  Tile t's epilogue overlaps tile t + 1's MMAs.
// Tile = 128 x 128: per 64-wide K chunk the tensor pipe needs 768 cycles for the three products.  With K = 256 / 512 a
// tile is only 3072 / 6144 tensor cycles for 16 384 outputs, so the EPILOGUE, not the operand traffic, paces the linears
// (ncu, profiles/r02_gemm_ws.txt: round-2's first version with 8 epilogue warps and scalar stores spent 12.4 k cycles per
// tile, tensor pipe 27 %); hence one block per warp, vector accesses and no exposed dependent global load.
#pragma once
#include "tma.cuh"

constexpr int GW_MAXP = 16;  // problems per launch (2 images x 8 pairs)
constexpr int GW_M = 128, GW_N = 128, GW_K = 64;
constexpr int GW_A_BYTES = GW_M * GW_K * 2;  // 16 KB per plane
constexpr int GW_B_BYTES = GW_N * GW_K * 2;  // 16 KB per plane
constexpr int GW_STAGE_BYTES = 2 * GW_A_BYTES + 2 * GW_B_BYTES;  // 64 KB
constexpr int GW_STAGES = 2;
constexpr int GW_EPI_WARPS = 16;
constexpr int GW_THREADS = 64 + 32 * GW_EPI_WARPS;  // producer warp + MMA warp + 16 epilogue warps
constexpr int GW_PAD = 36;  // fp32 row pitch of the transpose pad: 16-byte aligned rows, conflict-free for 128-bit accesses
constexpr int GW_SCRATCH = GW_EPI_WARPS * 32 * GW_PAD * 4;  // one 32 x 36 fp32 transpose pad per epilogue warp
constexpr size_t GW_SMEM = GW_STAGES * GW_STAGE_BYTES + GW_SCRATCH + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmProblem {
  const float* resid;  // [M][ldr] fp32 or null, added after bias / scale
  float* C;            // optional fp32 output, row-major [M][ldc]
  __half *Ch, *Cl;     // optional split output planes
  int M, N, ldc;
  int vec4;      // outputs / residual may be accessed 16 bytes at a time (leading dimensions and bases 16-byte aligned)
  int tiles_n;   // ceil(N / 128)
  int tile_end;  // running total of tiles up to and including this problem
};
struct GemmWsMaps {  // 128-byte TMA descriptors, passed as a __grid_constant__ kernel parameter
  CUtensorMap a1h[GW_MAXP], a1l[GW_MAXP];  // per problem: first K segment
  CUtensorMap a2h[GW_MAXP], a2l[GW_MAXP];  // optional second K segment (torch.cat([x, msg], -1) without the concat)
  CUtensorMap bh[GW_MAXP], bl[GW_MAXP];    // [0] when every problem shares the weight matrix
};
struct GemmWsArgs {
  GemmProblem p[GW_MAXP];
  int nprob, tiles;
  int K1, K2;
  int b_per_problem;
  const float* bias;  // [N] or null
  int ldr;
  float scale;
  int ldch;        // row-major leading dimension of Ch / Cl (ignored when head_major)
  int head_major;  // 1: Ch / Cl (and C) are written as [N/64][M][64] (attention head layout)
  int relu;        // max(., 0) after bias / scale, before the residual
  int lo_unscaled; // split outputs keep lo = fp16(x - hi) (attention operands)
  int* err_flag;   // set to 1 if an mbarrier wait timed out (pipeline bug): results are then invalid
};

static __global__ void __launch_bounds__(GW_THREADS, 1) k_gemm_ws(const __grid_constant__ GemmWsMaps maps, const __grid_constant__ GemmWsArgs g) {
  extern __shared__ unsigned char gw_raw[];
  const uint32_t raw = tc::smem_u32(gw_raw);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  unsigned char* sm = gw_raw + (smem0 - raw);
  float* scratch_all = reinterpret_cast<float*>(sm + GW_STAGES * GW_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + GW_STAGES * GW_STAGE_BYTES + GW_SCRATCH);
  uint64_t* empty = full + GW_STAGES;
  uint64_t* acc_full = empty + GW_STAGES;  // [2] all MMAs of the tile in accumulator set b have completed
  uint64_t* acc_empty = acc_full + 2;      // [2] the 8 epilogue warps have read set b out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nk = (g.K1 + g.K2) / GW_K;

  if (t == 0) {
    for (int s = 0; s < GW_STAGES; ++s) tc::mbar_init(&full[s], 1), tc::mbar_init(&empty[s], 1);
    for (int b = 0; b < 2; ++b) tc::mbar_init(&acc_full[b], 1), tc::mbar_init(&acc_empty[b], GW_EPI_WARPS);
    tc::fence_mbar_init();
    tc::tma_prefetch_desc(&maps.bh[0]);
    tc::tma_prefetch_desc(&maps.bl[0]);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  bool ok = true;

  auto decode = [&](int tile, int& z, int& m0, int& n0) {
    z = 0;
    while (z + 1 < g.nprob && tile >= g.p[z].tile_end) ++z;
    const int local = tile - (z ? g.p[z - 1].tile_end : 0);
    const int mt = local / g.p[z].tiles_n;
    n0 = (local - mt * g.p[z].tiles_n) * GW_N;
    m0 = mt * GW_M;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int gk = 0;
      for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x) {
        int z, m0, n0;
        decode(tile, z, m0, n0);
        const int zb = g.b_per_problem ? z : 0;
        for (int kc = 0; kc < nk; ++kc, ++gk) {
          const int s = gk % GW_STAGES;
          if (gk >= GW_STAGES) ok = tc::mbar_wait(&empty[s], ((gk / GW_STAGES) - 1) & 1) && ok;
          const int k0 = kc * GW_K;
          const bool seg2 = k0 >= g.K1;
          const CUtensorMap* ah = seg2 ? &maps.a2h[z] : &maps.a1h[z];
          const CUtensorMap* al = seg2 ? &maps.a2l[z] : &maps.a1l[z];
          const int ka = seg2 ? k0 - g.K1 : k0;
          const uint32_t sA = smem0 + s * GW_STAGE_BYTES, sB = sA + 2 * GW_A_BYTES;
          tc::mbar_expect_tx(&full[s], GW_STAGE_BYTES);
          tc::tma_load_2d(sA, ah, &full[s], ka, m0);
          tc::tma_load_2d(sA + GW_A_BYTES, al, &full[s], ka, m0);
          tc::tma_load_2d(sB, &maps.bh[zb], &full[s], k0, n0);
          tc::tma_load_2d(sB + GW_B_BYTES, &maps.bl[zb], &full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp, one elected lane issues) =====
    const uint32_t idesc = tc::idesc_f16(GW_M, GW_N);
    int gk = 0, it = 0;
    bool hint = false;  // next stage already seen full by a pre-poll
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      if (it >= 2) ok = tc::mbar_wait(&acc_empty[b], ((it >> 1) - 1) & 1) && ok;
      const uint32_t acc = tmem + b * (2 * GW_N);
      for (int kc = 0; kc < nk; ++kc, ++gk) {
        const int s = gk % GW_STAGES;
        if (!hint) ok = tc::mbar_wait(&full[s], (gk / GW_STAGES) & 1) && ok;
        __syncwarp();
        tc::fence_after_sync();
        hint = tc::mbar_test(&full[(gk + 1) % GW_STAGES], ((gk + 1) / GW_STAGES) & 1);
        const uint32_t aH = smem0 + s * GW_STAGE_BYTES, aL = aH + GW_A_BYTES, bH = aH + 2 * GW_A_BYTES, bL = bH + GW_B_BYTES;
        const uint64_t dAh = tc::smem_desc_sw128(aH), dAl = tc::smem_desc_sw128(aL), dBh = tc::smem_desc_sw128(bH), dBl = tc::smem_desc_sw128(bL);
#pragma unroll
        for (int ks = 0; ks < GW_K / 16; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K step, in 16-byte units of the start-address field
          const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
          tc::umma_f16_w(acc, dAh + adv, dBh + adv, idesc, first);         // acc0 (+)= Ah Bh
          tc::umma_f16_w(acc + GW_N, dAh + adv, dBl + adv, idesc, first);  // acc1 (+)= Ah Bl
          tc::umma_f16_w(acc + GW_N, dAl + adv, dBh + adv, idesc, 1u);     // acc1  += Al Bh
        }
        tc::umma_commit_w(&empty[s]);
      }
      tc::umma_commit_w(&acc_full[b]);
    }
  } else {
    // ===== epilogue warps: warp e owns ONE 32 x 32 block of every tile (rows = its TMEM lane quarter, column block e / 4) =====
    const int ew = warp - 2;                 // 0..15
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int cb = ew >> 2;                  // 32-column block of the tile
    float* scratch = scratch_all + ew * (32 * GW_PAD);
    const float scale = g.scale;
    const int relu = g.relu;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // read-back: 4 rows per pass, lane -> (row rsub, columns c4 .. c4 + 3)
    int it = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
      int z, m0, n0;
      decode(tile, z, m0, n0);
      const GemmProblem& pb = g.p[z];
      const int M = pb.M, N = pb.N;
      const int b = it & 1;
      const int mw = m0 + quarter * 32;
      const int n = n0 + cb * 32 + c4;  // first of this lane's 4 columns
      // operands that do not depend on the accumulator are fetched before the wait
      float bias4[4] = {0.f, 0.f, 0.f, 0.f};
      if (g.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) bias4[j] = __ldg(g.bias + n + j);
      }
      const bool vec = pb.vec4 && n + 3 < N;  // 16-byte accesses allowed for this lane's chunk
      ok = tc::mbar_wait(&acc_full[b], (it >> 1) & 1) && ok;
      tc::fence_after_sync();
      const uint32_t lane_base = tmem + b * (2 * GW_N) + ((uint32_t)(quarter * 32) << 16) + cb * 32;
      {
        float a0[32], a1[32];
        tc::tmem_ld32(lane_base, a0);
        tc::tmem_ld32(lane_base + GW_N, a1);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[b]);  // this warp's share of the set is in registers
        float4* dst = reinterpret_cast<float4*>(scratch + lane * GW_PAD);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[c] = make_float4(fmaf(a1[4 * c], tc::LO_INV, a0[4 * c]), fmaf(a1[4 * c + 1], tc::LO_INV, a0[4 * c + 1]),
                               fmaf(a1[4 * c + 2], tc::LO_INV, a0[4 * c + 2]), fmaf(a1[4 * c + 3], tc::LO_INV, a0[4 * c + 3]));
      }
      __syncwarp();
      if (n < N) {
        const size_t hm_col = (size_t)(n >> 6) * M * 64 + (n & 63);  // head-major: [N / 64][M][64]
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          const int r = ps * 4 + rsub;
          const int row = mw + r;
          if (row >= M) break;  // rows ascend with ps: nothing further for this lane
          const float4 x4 = *reinterpret_cast<const float4*>(scratch + r * GW_PAD + c4);
          float v[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = (v[j] + bias4[j]) * scale;
            if (relu) v[j] = fmaxf(v[j], 0.f);
          }
          if (pb.resid) {
            const float* rp = pb.resid + (size_t)row * g.ldr + n;
            if (vec) {
              const float4 rr = *reinterpret_cast<const float4*>(rp);
              v[0] += rr.x, v[1] += rr.y, v[2] += rr.z, v[3] += rr.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j < N) v[j] += rp[j];
            }
          }
          const size_t off_c = g.head_major ? hm_col + (size_t)row * 64 : (size_t)row * pb.ldc + n;
          const size_t off_s = g.head_major ? hm_col + (size_t)row * 64 : (size_t)row * g.ldch + n;
          if (pb.C) {
            if (vec) {
              *reinterpret_cast<float4*>(pb.C + off_c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j < N) pb.C[off_c + j] = v[j];
            }
          }
          if (pb.Ch) {
            if (vec) {
              uint32_t h01, l01, h23, l23;
              if (g.lo_unscaled) {
                tc::split2_unscaled_clamped(v[0], v[1], h01, l01);
                tc::split2_unscaled_clamped(v[2], v[3], h23, l23);
              } else {
                tc::split2(v[0], v[1], h01, l01);
                tc::split2(v[2], v[3], h23, l23);
              }
              *reinterpret_cast<uint2*>(pb.Ch + off_s) = make_uint2(h01, h23);
              *reinterpret_cast<uint2*>(pb.Cl + off_s) = make_uint2(l01, l23);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (n + j < N) {
                  __half hh, ll;
                  if (g.lo_unscaled) tc::split_h_unscaled(v[j], hh, ll);
                  else tc::split_h(v[j], hh, ll);
                  pb.Ch[off_s + j] = hh;
                  pb.Cl[off_s + j] = ll;
                }
              }
            }
          }
        }
      }
      __syncwarp();  // scratch is reused by the next tile
    }
  }
  __syncwarp();
  if (!ok && g.err_flag) *g.err_flag = 1;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}
