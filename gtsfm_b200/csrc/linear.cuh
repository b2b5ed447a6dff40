// Shared host-side dispatch for the matcher networks: one "linear" / attention call site expressed for both execution
// paths (tcgen05 split-fp16 planes, or the exact-fp32 SIMT kernels when B2_FORCE_SIMT=1).
#pragma once
#include "attn_tc.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"
#include "gemm_ws.cuh"
#include "attn_ws.cuh"
#include "attn_ts.cuh"
#include "attn_ps.cuh"
#include <cstdlib>

struct TcWeights {  // a model's weight blob in fp32 and as split-fp16 planes (same element offsets in all three)
  const float* f;
  const __half* h;
  const __half* l;
  int* err;     // device flag raised by a timed-out mbarrier wait
  bool use_tc;
  bool use_tma = true;  // TMA-fed kernels (default) vs the cp.async kernels (B2_NO_TMA=1)
  DevBuf* attn_part = nullptr;  // [2] scratch for key-split attention partials (O) ...
  DevBuf* attn_ml = nullptr;    // [2] ... and (m, l)
  int sm_count = 148;
};

// q / k planes feed the warp-specialised attention (single logits accumulator) with an UNSCALED lo plane
static inline bool attn_qk_unscaled(const TcWeights& tw) { return tw.use_tc && tw.use_tma && tw.attn_part && tma_encoder() != nullptr; }
// the TMEM-operand attention kernel (k_flash_ts, default) also takes v with an unscaled lo plane; B2_ATTN_WS=1 selects
// the shared-memory-operand kernel (k_flash_ws) instead
static inline bool attn_use_ts() {
  static const bool ts = [] {
    const char* e = getenv("B2_ATTN_WS");
    return !(e && e[0] == '1');
  }();
  return ts;
}
// persistent warp-specialised GEMM (k_gemm_ws, default); B2_GEMM_WS=0 selects the one-tile-per-CTA k_gemm_tma
static inline bool gemm_use_ws() {
  static const bool ws = [] {
    const char* e = getenv("B2_GEMM_WS");
    return !(e && e[0] == '0');
  }();
  return ws;
}
// ... and its persistent stream-K schedule (k_flash_ps) is what runs unless B2_ATTN_PS=0
static inline bool attn_use_ps() {
  static const bool ps = [] {
    const char* e = getenv("B2_ATTN_PS");
    return !(e && e[0] == '0');
  }();
  return ps;
}
static inline bool attn_v_unscaled(const TcWeights& tw) { return attn_qk_unscaled(tw) && attn_use_ts(); }

struct Pl {  // split-fp16 planes of an activation
  __half* hi;
  __half* lo;
};
static inline Pl planes_of(const DevBuf& b, size_t elems) { return {b.as<__half>(), b.as<__half>() + elems}; }

// One linear / GEMM call site, expressed for both execution paths: fp32 views feed the exact-fp32 SIMT kernel,
// split-fp16 plane views feed the tcgen05 kernel.
struct LinArgs {
  const float* a1f = nullptr;
  Pl a1p{nullptr, nullptr};
  int lda1 = 0, K1 = 0;
  const float* a2f = nullptr;
  Pl a2p{nullptr, nullptr};
  int lda2 = 0, K2 = 0;
  const float* w = nullptr;   // weight inside the blob (B operand), or
  const float* bf = nullptr;  // an fp32 activation B operand with
  Pl bp{nullptr, nullptr};    // its planes
  int ldb = 0;
  const float* bias = nullptr;
  float scale = 1.f;
  const float* resid = nullptr;
  int ldr = 0;
  float* cf = nullptr;  // fp32 output (always written on the SIMT path; on the tcgen05 path only if tc_want_f32)
  int ldc = 0;
  Pl cp{nullptr, nullptr};  // plane output (tcgen05 path)
  int ldch = 0;
  int head_major = 0;
  bool tc_want_f32 = false;
  int relu = 0;  // max(., 0) after bias / scale, before the residual
  int lo_unscaled = 0;  // plane output with an unscaled lo plane (q / k operands of the warp-specialised attention)
  int M = 0, N = 0;
};

// `b` (optional) is the same linear applied to the other image of the pair: same weights, shapes and epilogue, so the
// tcgen05 path runs both as one launch (blockIdx.z).
static int run_linear(b2_context* ctx, cudaStream_t st, const TcWeights& tw, const LinArgs& a, const LinArgs* b = nullptr) {
  if (!tw.use_tc) {
    const LinArgs* both[2] = {&a, b};
    for (int i = 0; i < 2; ++i) {
      if (!both[i] || both[i]->M <= 0 || both[i]->N <= 0) continue;
      const LinArgs& x = *both[i];
      GemmArgs g{};
      g.A1 = x.a1f, g.lda1 = x.lda1, g.K1 = x.K1, g.A2 = x.a2f, g.lda2 = x.lda2, g.K2 = x.K2;
      g.B = x.w ? x.w : x.bf, g.ldb = x.ldb, g.C = x.cf, g.ldc = x.ldc, g.M = x.M, g.N = x.N;
      g.bias = x.bias, g.resid = x.resid, g.ldr = x.ldr, g.scale = x.scale, g.head_major = x.head_major, g.relu = x.relu;
      int rc = launch_gemm(ctx, st, g);
      if (rc) return rc;
    }
    return B2_OK;
  }
  if (a.N <= 0 || (a.M <= 0 && (!b || b->M <= 0))) return B2_OK;
  if (tw.use_tma && tma_encoder()) {
    GemmTmaMaps maps;
    GemmTmaArgs q{};
    const LinArgs* both[2] = {&a, b};
    int maxM = 0;
    double work = 0.0;
    bool ok = true;
    for (int i = 0; i < 2; ++i) {
      const LinArgs& x = both[i] ? *both[i] : a;  // unused second slot mirrors the first (never launched: grid.z = 1)
      ok = ok && tma_map_2d(&maps.a1h[i], x.a1p.hi, x.M, x.K1, x.lda1, TM_M) && tma_map_2d(&maps.a1l[i], x.a1p.lo, x.M, x.K1, x.lda1, TM_M);
      if (x.K2 > 0)
        ok = ok && tma_map_2d(&maps.a2h[i], x.a2p.hi, x.M, x.K2, x.lda2, TM_M) && tma_map_2d(&maps.a2l[i], x.a2p.lo, x.M, x.K2, x.lda2, TM_M);
      else
        maps.a2h[i] = maps.a1h[i], maps.a2l[i] = maps.a1l[i];
      if (!both[i]) continue;
      GemmTcProblem& pr = q.p[i];
      pr.resid = x.resid, pr.C = x.tc_want_f32 ? x.cf : nullptr, pr.Ch = x.cp.hi, pr.Cl = x.cp.lo, pr.M = x.M;
      maxM = x.M > maxM ? x.M : maxM;
      work += 2.0 * x.M * x.N * (x.K1 + x.K2);
    }
    const __half *bh, *bl;
    if (a.w) {
      const size_t off = (size_t)(a.w - tw.f);
      bh = tw.h + off, bl = tw.l + off;
    } else {
      bh = a.bp.hi, bl = a.bp.lo;
    }
    ok = ok && tma_map_2d(&maps.bh, bh, a.N, a.K1 + a.K2, a.ldb, TM_N) && tma_map_2d(&maps.bl, bl, a.N, a.K1 + a.K2, a.ldb, TM_N);
    if (!ok) return b2_fail(ctx, B2_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    q.K1 = a.K1, q.K2 = a.K2, q.N = a.N, q.bias = a.bias, q.ldr = a.ldr, q.scale = a.scale, q.ldc = a.ldc, q.ldch = a.ldch;
    q.head_major = a.head_major, q.relu = a.relu, q.lo_unscaled = a.lo_unscaled, q.err_flag = tw.err;
    if (gemm_use_ws()) {  // persistent, epilogue overlapped with the next tile's MMAs
      GemmWsArgs wq{};
      wq.g = q;
      wq.tiles_n = cdiv(a.N, TM_N), wq.tiles_m0 = cdiv(a.M, TM_M);
      wq.tiles = wq.tiles_n * (wq.tiles_m0 + (b ? cdiv(b->M, TM_M) : 0));
      if (wq.tiles <= 0) return B2_OK;
      b2_prof_work(ctx, "k_gemm_ws", work);
      B2_LAUNCH(ctx, k_gemm_ws, wq.tiles < tw.sm_count ? wq.tiles : tw.sm_count, GW_THREADS, GW_SMEM, st, maps, wq);
      B2_CHECK_LAUNCH(ctx);
      return B2_OK;
    }
    dim3 grid(cdiv(a.N, TM_N), cdiv(maxM, TM_M), b ? 2 : 1);
    b2_prof_work(ctx, "k_gemm_tma", work);
    B2_LAUNCH(ctx, k_gemm_tma, grid, 128, TM_GEMM_SMEM, st, maps, q);
    B2_CHECK_LAUNCH(ctx);
    return B2_OK;
  }
  GemmTcArgs t{};
  const LinArgs* both[2] = {&a, b};
  int maxM = 0;
  double work = 0.0;
  for (int i = 0; i < 2; ++i) {
    if (!both[i]) continue;
    const LinArgs& x = *both[i];
    GemmTcProblem& q = t.p[i];
    q.A1h = x.a1p.hi, q.A1l = x.a1p.lo, q.A2h = x.a2p.hi, q.A2l = x.a2p.lo, q.resid = x.resid;
    q.C = x.tc_want_f32 ? x.cf : nullptr, q.Ch = x.cp.hi, q.Cl = x.cp.lo, q.M = x.M;
    maxM = x.M > maxM ? x.M : maxM;
    work += 2.0 * x.M * x.N * (x.K1 + x.K2);
  }
  t.lda1 = a.lda1, t.K1 = a.K1, t.lda2 = a.lda2, t.K2 = a.K2;
  if (a.w) {
    const size_t off = (size_t)(a.w - tw.f);
    t.Bh = tw.h + off, t.Bl = tw.l + off;
  } else {
    t.Bh = a.bp.hi, t.Bl = a.bp.lo;
  }
  t.ldb = a.ldb, t.N = a.N, t.bias = a.bias, t.ldr = a.ldr, t.scale = a.scale, t.ldc = a.ldc, t.ldch = a.ldch;
  t.head_major = a.head_major, t.relu = a.relu, t.lo_unscaled = a.lo_unscaled, t.err_flag = tw.err;
  dim3 grid(cdiv(a.N, TC_N), cdiv(maxM, TC_M), b ? 2 : 1);
  b2_prof_work(ctx, "k_gemm_tc", work);
  B2_LAUNCH(ctx, k_gemm_tc, grid, 128, TC_GEMM_SMEM, st, t);
  B2_CHECK_LAUNCH(ctx);
  return B2_OK;
}

// One launch, two attention problems.  tcgen05 path: q / k / v / o buffers hold split fp16 planes (hi, then lo at + N * 256).
struct FlashJob {
  const DevBuf *q, *k, *v, *o;
  int nq, nk;
  int capq, capk;  // allocated rows of the query-side / key-side buffers (lo plane offset = cap * 256 halves)
};
static int run_flash2(b2_context* ctx, cudaStream_t st, const TcWeights& tw, const FlashJob& a, const FlashJob& b, float scale) {
  if (!tw.use_tc) {
    int rc;
    if ((rc = launch_flash(ctx, st, a.q->as<float>(), a.k->as<float>(), a.v->as<float>(), a.o->as<float>(), a.nq, a.nk, scale))) return rc;
    return launch_flash(ctx, st, b.q->as<float>(), b.k->as<float>(), b.v->as<float>(), b.o->as<float>(), b.nq, b.nk, scale);
  }
  const FlashJob* jobs[2] = {&a, &b};
  if (tw.use_tma && tma_encoder() && tw.attn_part && attn_use_ts() && attn_use_ps()) {
    AttnTsMaps tmaps;
    AttnPsArgs pa{};
    bool okm = true;
    int items = 0;
    for (int i = 0; i < 2; ++i) {
      const FlashJob& j = *jobs[i];
      const Pl q = planes_of(*j.q, (size_t)j.capq * 256), k = planes_of(*j.k, (size_t)j.capk * 256), v = planes_of(*j.v, (size_t)j.capk * 256),
               o = planes_of(*j.o, (size_t)j.capq * 256);
      AttnPsProblem& p = pa.p[i];
      p.Qh = q.hi, p.Ql = q.lo, p.Oh = o.hi, p.Ol = o.lo, p.Nq = j.nq, p.Nk = j.nk;
      p.qt = (j.nq > 0 && j.nk > 0) ? cdiv(j.nq, 2 * AW_Q) : 0, p.tiles = j.nk > 0 ? cdiv(j.nk, AW_KV) : 1;
      items += p.qt * 4;
      if (p.qt == 0) {  // nothing to do for this problem: mirror the other one's maps so the struct is fully initialised
        continue;
      }
      okm = okm && tma_map_2d(&tmaps.kh[i], k.hi, (uint64_t)4 * j.nk, 64, 64, AW_KV) && tma_map_2d(&tmaps.kl[i], k.lo, (uint64_t)4 * j.nk, 64, 64, AW_KV);
      okm = okm && tma_map_2d(&tmaps.vh[i], v.hi, (uint64_t)4 * j.nk, 64, 64, AW_KV) && tma_map_2d(&tmaps.vl[i], v.lo, (uint64_t)4 * j.nk, 64, 64, AW_KV);
    }
    if (!okm) return b2_fail(ctx, B2_ERR_CUDA, "cuTensorMapEncodeTiled failed (attention)");
    for (int i = 0; i < 2; ++i)
      if (pa.p[i].qt == 0) tmaps.kh[i] = tmaps.kh[1 - i], tmaps.kl[i] = tmaps.kl[1 - i], tmaps.vh[i] = tmaps.vh[1 - i], tmaps.vl[i] = tmaps.vl[1 - i];
    pa.W0 = pa.p[0].qt * 4 * pa.p[0].tiles;
    pa.W = pa.W0 + pa.p[1].qt * 4 * pa.p[1].tiles;
    if (pa.W <= 0) return B2_OK;
    int ncta = pa.W < tw.sm_count ? pa.W : tw.sm_count;
    pa.quota = cdiv(pa.W, ncta);
    ncta = cdiv(pa.W, pa.quota);
    const int tmax = pa.p[0].tiles > pa.p[1].tiles ? pa.p[0].tiles : pa.p[1].tiles;
    pa.max_splits = cdiv(tmax, pa.quota) + 1;
    B2_CUDA(ctx, tw.attn_part[0].ensure((size_t)items * pa.max_splits * 256 * 64 * 4));
    B2_CUDA(ctx, tw.attn_ml[0].ensure((size_t)items * pa.max_splits * 256 * 2 * 4));
    {  // arrival counters: zero once per allocation, the kernel leaves them zero
      DevBuf& cnt = tw.attn_ml[1];
      const size_t need = (size_t)items * 2 * sizeof(int);
      if (cnt.cap < need) {
        B2_CUDA(ctx, cnt.ensure(need * 4));
        B2_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, cnt.cap, st));
      }
      pa.arrivals = cnt.as<int>();
    }
    pa.Opart = tw.attn_part[0].as<float>(), pa.ml = tw.attn_ml[0].as<float>();
    pa.scale = scale, pa.err_flag = tw.err;
    b2_prof_work(ctx, "k_flash_ps", 4.0 * 2.0 * 2.0 * 64 * ((double)a.nq * a.nk + (double)b.nq * b.nk));
    B2_LAUNCH(ctx, k_flash_ps, ncta, AS_THREADS, AS_SMEM, st, tmaps, pa);
    B2_CHECK_LAUNCH(ctx);
    return B2_OK;
  }
  if (tw.use_tma && tma_encoder() && tw.attn_part) {
    AttnWsMaps maps;
    AttnWsArgs wa{};
    const int qt = cdiv(a.nq > b.nq ? a.nq : b.nq, 2 * AW_Q);
    if (qt <= 0) return B2_OK;
    // key-range split factor: fewest (rounds of CTAs over the SMs) / split
    const int items = qt * 4 * 2;
    int nsplit = 1;
    double best = 1e30;
    const int max_tiles = cdiv(a.nk > b.nk ? a.nk : b.nk, AW_KV);
    for (int sp = 1; sp <= 4 && sp <= max_tiles; ++sp) {
      const double cost = (double)cdiv(items * sp, tw.sm_count) / sp + 0.03 * (sp - 1);  // small penalty for the merge pass
      if (cost < best - 1e-9) best = cost, nsplit = sp;
    }
    bool okm = true;
    const bool ts = attn_use_ts();
    AttnTsMaps tmaps;
    AttnTsArgs ta{};
    for (int i = 0; i < 2; ++i) {
      const FlashJob& j = *jobs[i];
      const Pl q = planes_of(*j.q, (size_t)j.capq * 256), k = planes_of(*j.k, (size_t)j.capk * 256), v = planes_of(*j.v, (size_t)j.capk * 256),
               o = planes_of(*j.o, (size_t)j.capq * 256);
      CUtensorMap *kh = ts ? &tmaps.kh[i] : &maps.kh[i], *kl = ts ? &tmaps.kl[i] : &maps.kl[i];
      CUtensorMap *vh = ts ? &tmaps.vh[i] : &maps.vh[i], *vl = ts ? &tmaps.vl[i] : &maps.vl[i];
      if (!ts)
        okm = okm && tma_map_2d(&maps.qh[i], q.hi, (uint64_t)4 * j.nq, 64, 64, AW_Q) && tma_map_2d(&maps.ql[i], q.lo, (uint64_t)4 * j.nq, 64, 64, AW_Q);
      okm = okm && tma_map_2d(kh, k.hi, (uint64_t)4 * j.nk, 64, 64, AW_KV) && tma_map_2d(kl, k.lo, (uint64_t)4 * j.nk, 64, 64, AW_KV);
      okm = okm && tma_map_2d(vh, v.hi, (uint64_t)4 * j.nk, 64, 64, AW_KV) && tma_map_2d(vl, v.lo, (uint64_t)4 * j.nk, 64, 64, AW_KV);
      AttnWsProblem& p = wa.p[i];
      p.Oh = o.hi, p.Ol = o.lo, p.Nq = j.nq, p.Nk = j.nk;
      if (nsplit > 1) {
        B2_CUDA(ctx, tw.attn_part[i].ensure((size_t)nsplit * j.nq * 256 * 4));
        B2_CUDA(ctx, tw.attn_ml[i].ensure((size_t)nsplit * 4 * j.nq * 2 * 4));
        p.Opart = tw.attn_part[i].as<float>(), p.ml = tw.attn_ml[i].as<float>();
      }
      AttnTsProblem& tp = ta.p[i];
      tp.Qh = q.hi, tp.Ql = q.lo, tp.Oh = p.Oh, tp.Ol = p.Ol, tp.Opart = p.Opart, tp.ml = p.ml, tp.Nq = j.nq, tp.Nk = j.nk;
    }
    if (!okm) return b2_fail(ctx, B2_ERR_CUDA, "cuTensorMapEncodeTiled failed (attention)");
    wa.scale = scale, wa.nsplit = nsplit, wa.err_flag = tw.err;
    ta.scale = scale, ta.nsplit = nsplit, ta.err_flag = tw.err;
    const double flops = 4.0 * 2.0 * 2.0 * 64 * ((double)a.nq * a.nk + (double)b.nq * b.nk);
    if (ts) {
      b2_prof_work(ctx, "k_flash_ts", flops);
      B2_LAUNCH(ctx, k_flash_ts, dim3(qt, 4, 2 * nsplit), AS_THREADS, AS_SMEM, st, tmaps, ta);
    } else {
      b2_prof_work(ctx, "k_flash_ws", flops);
      B2_LAUNCH(ctx, k_flash_ws, dim3(qt, 4, 2 * nsplit), AW_THREADS, AW_SMEM, st, maps, wa);
    }
    B2_CHECK_LAUNCH(ctx);
    if (nsplit > 1) {
      for (int i = 0; i < 2; ++i) {
        const FlashJob& j = *jobs[i];
        const Pl o = planes_of(*j.o, (size_t)j.capq * 256);
        B2_LAUNCH(ctx, k_attn_merge, cdiv(j.nq * 128, 256), 256, 0, st, tw.attn_part[i].as<float>(), tw.attn_ml[i].as<float>(), j.nq, nsplit,
                  o.hi, o.lo);
        B2_CHECK_LAUNCH(ctx);
      }
    }
    return B2_OK;
  }
  AttnArgs args{};
  for (int i = 0; i < 2; ++i) {
    const FlashJob& j = *jobs[i];
    AttnProblem& p = args.p[i];
    const Pl q = planes_of(*j.q, (size_t)j.capq * 256), k = planes_of(*j.k, (size_t)j.capk * 256), v = planes_of(*j.v, (size_t)j.capk * 256),
             o = planes_of(*j.o, (size_t)j.capq * 256);
    p.Qh = q.hi, p.Ql = q.lo, p.Kh = k.hi, p.Kl = k.lo, p.Vh = v.hi, p.Vl = v.lo, p.Oh = o.hi, p.Ol = o.lo;
    p.Nq = j.nq, p.Nk = j.nk;
  }
  args.scale = scale, args.err_flag = tw.err;
  const int qt = cdiv(a.nq > b.nq ? a.nq : b.nq, AT_Q);
  if (qt <= 0) return B2_OK;
  b2_prof_work(ctx, "k_flash_tc", 4.0 * 2.0 * 2.0 * 64 * ((double)a.nq * a.nk + (double)b.nq * b.nk));  // 4 heads x (QK^T + PV) x 2 FLOP/MAC
  B2_LAUNCH(ctx, k_flash_tc, dim3(qt, 4, 2), 128, AT_SMEM, st, args);
  B2_CHECK_LAUNCH(ctx);
  return B2_OK;
}

